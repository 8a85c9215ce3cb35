"""GPU: vlpk_mask_pack (warp-per-row, __ballot_sync) against a bit-pack computed with torch on the host — bit-exact, for the three
mask dtypes the module surface passes (int64 0/1 `input_mask`, fp32 / bf16 additive extended masks), 3-D and broadcast masks,
ragged kv (reference semantics: get_extended_attention_mask, modeling.py:807-833)."""
import pytest
import torch

from vlp_b200 import ops

pytestmark = pytest.mark.gpu


def _pack_host(m01):
    B, R, KV = m01.shape
    out = torch.zeros(B, R, 4, dtype=torch.int64)
    for j in range(KV):
        out[:, :, j >> 5] |= m01[:, :, j].to(torch.int64) << (j & 31)
    return out.to(torch.int32)          # wraps bit 31 into the sign, like the kernel's uint32 words viewed as int32


@pytest.mark.parametrize("dtype", [torch.int64, torch.float32, torch.bfloat16])
def test_mask_pack_is_bit_exact(dtype):
    g = torch.Generator().manual_seed(2)
    for (B, R, KV) in ((64, 123, 123), (3, 1, 77), (2, 2, 128), (5, 15, 15), (1, 1, 1)):
        m01 = (torch.rand(B, R, KV, generator=g) < 0.6).to(torch.int64)
        mask = m01.cuda() if dtype == torch.int64 else ((1 - m01).to(torch.float32) * -10000.0).to(dtype).cuda()
        got = ops.pack_mask(mask, mode="zero_one" if dtype == torch.int64 else "additive").cpu()
        want = _pack_host(m01)
        assert got.shape == want.shape and torch.equal(got.view(torch.int32), want)
        assert int(got.abs().sum()) != 0 or KV == 1
