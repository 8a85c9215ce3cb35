"""GPU: BertEmbeddings backward (vlpk_embed_bwd + the table-gradient kernels of csrc/tables.cu, vlpk_embed_tables_bwd) against fp32
PyTorch autograd of the same forward (modeling.py:217-241: gather / region splice / sum / LayerNorm) on identical inputs.  The
kernels carry the pre-LN gradient in bf16 and accumulate table rows in fp32; the word gradient is rounded to bf16 once -> rel-L2
within bf16 resolution (1e-2), heavily duplicated ids ([CLS]-like) included."""
import pytest
import torch
import torch.nn.functional as F

from vlp_b200 import ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,L,R,H,V,vis", [(3, 15, 4, 128, 300, True), (64, 123, 100, 768, 28996, True), (2, 9, 0, 128, 50, False)])
def test_table_grads_match_fp32_autograd(B, L, R, H, V, vis):
    gen = torch.Generator().manual_seed(9)
    P, T = 512, 6
    tabs = [(torch.randn(n, H, generator=gen) * 0.05).cuda().bfloat16() for n in (V, P, T)]
    ln_g, ln_b = (1 + 0.1 * torch.randn(H, generator=gen)).cuda().bfloat16(), (0.1 * torch.randn(H, generator=gen)).cuda().bfloat16()
    ids = torch.randint(0, V, (B, L), generator=gen).cuda()
    ids[:, 0] = 1                                                 # a heavily duplicated id ([CLS]-like)
    tt = torch.randint(0, T, (B, L), generator=gen).cuda()
    visf = (torch.randn(B, max(R, 1), H, generator=gen)).cuda().bfloat16()
    vpef = (torch.randn(B, max(R, 1), H, generator=gen)).cuda().bfloat16()
    dy = (torch.randn(B, L, H, generator=gen) * 0.1).cuda().bfloat16()

    leaves = [t.clone().requires_grad_(True) for t in tabs]
    y = ops.EmbedFn.apply(visf if vis else None, vpef if vis else None, leaves[0], leaves[1], leaves[2], ln_g, ln_b, ids, tt, None, vis, R, 0.0,
                          False)
    y.backward(dy)
    torch.cuda.synchronize()
    got = [t.grad.float().cpu() for t in leaves]

    ref = [t.float().clone().requires_grad_(True) for t in tabs]
    w = ref[0][ids]
    p = ref[1][torch.arange(L, device="cuda")].unsqueeze(0).expand(B, -1, -1)
    if vis:
        w = torch.cat((w[:, :1], visf.float(), w[:, R + 1:]), dim=1)
        p = torch.cat((p[:, :1], vpef.float(), p[:, R + 1:]), dim=1)
    z = w + p + ref[2][tt]
    yr = F.layer_norm(z, (H,), ln_g.float(), ln_b.float(), 1e-5)
    yr.backward(dy.float())
    assert float((y.float() - yr).norm() / yr.norm()) < 1e-2
    for name, a, b in zip(("word", "pos", "type"), got, [t.grad.cpu() for t in ref]):
        assert float(b.norm()) > 0 and float((a - b).norm() / b.norm()) < 1e-2, name
        untouched = (b.abs().sum(-1) == 0)
        assert float(a[untouched].abs().sum()) == 0.0, name       # rows never looked up get an exactly-zero gradient


def test_table_rows_add_matches_index_add():
    """vlpk_table_rows_add (data-parallel path: all ranks' looked-up rows added into an existing bf16 table gradient): duplicates summed in
    fp32, scaled, added once per id; position rows into an fp32 table — vs torch index_add_ in fp32."""
    from vlp_b200 import _lib as L
    gen = torch.Generator().manual_seed(4)
    n, H, V, P = 4 * 64 * 23, 768, 28996, 512
    ids = torch.randint(0, V, (n,), generator=gen)
    ids[::7] = 101                                               # heavy duplication ([CLS])
    ids[3::11] = 102
    pos = torch.randint(0, 123, (n,), generator=gen)
    rows = (torch.randn(n, H, generator=gen) * 0.05).bfloat16()
    base = (torch.randn(V, H, generator=gen) * 0.01).bfloat16()
    d_word = base.clone().cuda()
    scratch = torch.empty(V, H, device="cuda", dtype=torch.float32)
    owner = torch.empty(V, device="cuda", dtype=torch.int32)
    d_pos = torch.zeros(P, H, device="cuda", dtype=torch.float32)
    scale = 0.25
    ids_d, pos_d, rows_d = ids.cuda(), pos.cuda(), rows.cuda()
    L.call("vlpk_table_rows_add", n, ids_d.data_ptr(), pos_d.data_ptr(), rows_d.data_ptr(), H, V, P, scale, d_word.data_ptr(),
           scratch.data_ptr(), owner.data_ptr(), d_pos.data_ptr(), L.stream())
    torch.cuda.synchronize()
    add = torch.zeros(V, H).index_add_(0, ids, rows.float() * scale)
    want = base.float() + add
    touched = add.abs().sum(-1) > 0
    assert torch.equal(d_word.cpu()[~touched], base[~touched])                      # untouched rows are bit-identical
    err = (d_word.float().cpu()[touched] - want[touched]).abs().max()
    assert float(err) <= 2 ** -8 * float(want[touched].abs().max()) + 1e-6              # one bf16 rounding of the sum
    want_pos = torch.zeros(P, H).index_add_(0, pos, rows.float() * scale)
    assert float((d_pos.cpu() - want_pos).abs().max()) < 1e-4 * float(want_pos.abs().max())
