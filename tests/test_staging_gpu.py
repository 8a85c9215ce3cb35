"""GPU: input staging (vlp_b200/staging.py).  (1) `vlpk_mask_synth` is bit-identical to packing the mask the
reference loader builds (seq2seq_loader.py:291-301, restated in synth.attention_mask) for every text length and both modes;
(2) a model step fed through BatchStager with bf16 features + (len_b, mode) equals the step fed with the loader's int64 matrix;
(3) the staged batch moves less than half the bytes."""
import pytest
import torch

from vlp_b200 import ops, staging, synth
from vlp_b200 import vlp_modules as vm

pytestmark = pytest.mark.gpu


def test_mask_synthesis_is_bit_identical_to_the_loader_mask():
    for d in (synth.BERT_BASE, synth.TINY):
        L_, R = d.seq_len, d.regions
        cases = [(tl, mode) for tl in range(0, d.text + 1) for mode in ("s2s", "bi")]
        ref = torch.stack([synth.attention_mask(d, tl, mode) for tl, mode in cases]).cuda()
        want = ops.pack_mask(ref, "zero_one")
        lb, md = staging.mask_descriptor([tl for tl, _ in cases], [m for _, m in cases])
        got = staging.PackedAttentionMask.synthesize(lb.cuda(), md.cuda(), R, L_)
        assert torch.equal(got.bits, want)
        lb2, md2 = staging.describe_mask(ref.cpu(), R)             # and the descriptor can be recovered from a loader-built matrix
        ok = [(int(a) == int(b)) for a, b in zip(lb2, lb)]
        assert all(ok)
        for (tl, mode), m in zip(cases, md2):
            if tl > 0:
                assert int(m) == (1 if mode == "s2s" else 0), (tl, mode)


def test_staged_step_matches_matrix_mask_step():
    d = synth.SMALL_L123
    cfg = vm.BertConfig(d.vocab, hidden_size=d.hidden, num_hidden_layers=d.layers, num_attention_heads=d.heads, intermediate_size=d.inter,
                        type_vocab_size=d.type_vocab, max_position_embeddings=d.max_pos, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = vm.BertForPreTrainingLossMask(cfg, enable_butd=True, len_vis_input=d.regions)
    model.load_state_dict(synth.make_state_dict(d, 0))
    model = model.cuda().bfloat16().eval()
    host = synth.make_batch(d, 6, seed=3, mode="mix", ragged=True)
    lb, md = staging.describe_mask(host["input_mask"], d.regions)

    def step(b):
        model.zero_grad(set_to_none=True)
        loss = model(b["img"], b["vis_pe"], b["input_ids"], b["segment_ids"], b["input_mask"], b["masked_ids"], None, b["is_next"],
                     masked_pos=b["masked_pos"], masked_weights=b["masked_weights"], task_idx=b["task_idx"], drop_worst_ratio=0.0)[0]
        loss.backward()
        return float(loss), model.bert.encoder.layer[0].attention.self.query.weight.grad.clone()

    stager = staging.BatchStager("cuda", len_vis_input=d.regions, max_len=d.seq_len)
    compact = {k: v for k, v in host.items() if k != "input_mask"}
    compact["len_b"], compact["mode"] = lb, md
    stager.put(compact)
    compact_bytes = stager.h2d_bytes
    b = stager.get()
    assert isinstance(b["input_mask"], staging.PackedAttentionMask) and b["img"].dtype == torch.bfloat16
    l1, g1 = step(b)
    b.done()
    stager.put(host)                                              # the loader's int64 matrix through the same stager
    full_bytes = stager.h2d_bytes
    b2 = stager.get()
    l2, g2 = step(b2)
    b2.done()
    assert l1 == l2 and torch.equal(g1, g2)
    assert compact_bytes < full_bytes
    fp32_bytes = sum(v.numel() * v.element_size() for v in host.values())
    assert compact_bytes < 0.5 * fp32_bytes
