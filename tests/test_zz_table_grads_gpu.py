"""GPU: the opt-in embedding-table gradient kernels (csrc/tables.cu, vlpk_embed_tables_bwd) against the default torch scatter
(index_add_ in fp32) inside the same autograd Function, on identical inputs.  Both accumulate in fp32; the word gradient is
rounded to bf16 once on either path -> agreement to fp32 summation-order noise (1e-5 of each tensor's scale)."""
import pytest
import torch

from tools.gating import unverified_on_gpu
from vlp_b200 import ops

pytestmark = [pytest.mark.gpu, unverified_on_gpu]


@pytest.mark.parametrize("B,L,R,H,V,vis", [(3, 15, 4, 128, 300, True), (64, 123, 100, 768, 28996, True), (2, 9, 0, 128, 50, False)])
def test_table_grads_match_torch_scatter(B, L, R, H, V, vis, monkeypatch):
    gen = torch.Generator().manual_seed(9)
    P, T = 512, 6
    tabs = [(torch.randn(n, H, generator=gen) * 0.05).cuda().bfloat16() for n in (V, P, T)]
    ln_g, ln_b = torch.ones(H).cuda().bfloat16(), torch.zeros(H).cuda().bfloat16()
    ids = torch.randint(0, V, (B, L), generator=gen).cuda()
    ids[:, 0] = 1                                                 # a heavily duplicated id ([CLS]-like)
    tt = torch.randint(0, T, (B, L), generator=gen).cuda()
    visf = (torch.randn(B, max(R, 1), H, generator=gen)).cuda().bfloat16()
    dy = (torch.randn(B, L, H, generator=gen) * 0.1).cuda().bfloat16()
    res = []
    for fused in (False, True):
        monkeypatch.setattr(ops, "FUSED_TABLE_GRADS", fused)
        leaves = [t.clone().requires_grad_(True) for t in tabs]
        y = ops.EmbedFn.apply(visf if vis else None, visf if vis else None, leaves[0], leaves[1], leaves[2], ln_g, ln_b, ids, tt, None, vis,
                              R, 0.0, False)
        y.backward(dy)
        torch.cuda.synchronize()
        res.append([t.grad.float().cpu() for t in leaves])
    for name, a, b in zip(("word", "pos", "type"), res[1], res[0]):
        scale = float(b.abs().max())
        assert scale > 0 and float((a - b).abs().max()) <= 1e-2 * scale, name      # bf16 table gradients: one ulp of the scale
        assert float((a - b).norm() / b.norm()) < 4e-3, name
