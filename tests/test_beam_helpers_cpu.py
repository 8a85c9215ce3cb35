"""Host-side pieces of the beam search (vlp_b200/beam.py, semantics of the reference's modeling.py:1326-1350, 1390-1406) that need
no GPU: beam expansion / re-ordering by back pointers and the duplicate-n-gram candidate rule."""
import torch

from vlp_b200 import beam


def test_expand_beams_repeats_each_item_consecutively():
    x = torch.arange(6).view(3, 2)
    y = beam._expand_beams(x, 2)
    assert y.tolist() == [[0, 1], [0, 1], [2, 3], [2, 3], [4, 5], [4, 5]]


def test_reorder_follows_back_pointers_per_batch_item():
    B, K = 2, 3
    x = torch.arange(B * K * 4, dtype=torch.float32).view(B * K, 2, 2)
    back = torch.tensor([[2, 0, 0], [1, 1, 2]])
    y = beam._reorder(x, back, B, K)
    xs = x.view(B, K, 2, 2)
    for b in range(B):
        for k in range(K):
            assert torch.equal(y.view(B, K, 2, 2)[b, k], xs[b, back[b, k]])


def test_dup_ngram_candidates_match_reference_rule():
    # trigram blocking: the last two words (7, 8) occurred before followed by 9 and by 4
    seq = [7, 8, 9, 1, 7, 8, 4, 7, 8]
    assert beam._dup_ngram_candidates(seq, 3, None) == [4, 9]
    assert beam._dup_ngram_candidates([1, 2], 3, None) == []
    assert beam._dup_ngram_candidates(seq, 3, {8}) == []          # tail contains an ignored word
    assert beam._dup_ngram_candidates(seq, 3, {9}) == [4]         # completion word ignored


def test_floor_division_back_pointers():
    """The reference's torch.div(k_ids, K) (modeling.py:1317) yields floats on torch >= 1.6; the rewrite must produce integer parents."""
    K = 4
    flat = torch.tensor([[0, 5, 7, 15]])
    back = torch.div(flat, K, rounding_mode="floor")
    assert back.dtype == torch.int64 and back.tolist() == [[0, 1, 1, 3]]
