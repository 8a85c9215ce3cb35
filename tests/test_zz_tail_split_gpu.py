"""GPU: the wave-remainder split experiment (VLPK_GEMM_TAIL_SPLIT / vlpk_debug_set_option) must not change any output bit: the
two launches use the same kernels and the same k-order per output element, only the tile width of the trailing rows differs."""
import pytest
import torch

from tools.gating import unverified_on_gpu
from vlp_b200 import _lib as L

pytestmark = [pytest.mark.gpu, unverified_on_gpu]
BF = torch.bfloat16


def _gemm(M, N, K, A, B, b_mn, epi, aux, bias, D0, D1=None, colsum=None):
    L.call("vlpk_gemm", M, N, K, 0, A.data_ptr(), A.stride(0), b_mn, B.data_ptr(), B.stride(0), L.ptr(bias), D0.data_ptr(), D0.stride(0),
           L.ptr(D1), D1.stride(0) if D1 is not None else 0, L.ptr(aux), aux.stride(0) if aux is not None else 0, epi, 1, 0, L.stream())


@pytest.mark.parametrize("N,K,b_mn,epi", [(768, 3072, 1, 3), (768, 768, 1, 0), (3072, 768, 0, 1), (3072, 768, 1, 4)])
def test_split_and_unsplit_gemm_agree_bitwise(N, K, b_mn, epi):
    M = 64 * 123
    assert L.lib().vlpk_debug_plan_tail_split(M, N, K, 0, b_mn, epi, 256, 2, 1) > 0
    g = torch.Generator().manual_seed(0)
    A = (torch.randn(M, K, generator=g)).cuda().to(BF)
    B = (torch.randn(K, N, generator=g) if b_mn else torch.randn(N, K, generator=g)).cuda().to(BF) * 0.05
    bias = None if b_mn else torch.randn(N, generator=g).cuda().to(BF)
    aux = torch.randn(M, N, generator=g).cuda().to(BF) if epi in (3, 4) else None
    outs = []
    try:
        for on in (0, 1):
            L.lib().vlpk_debug_set_option(b"tail_split", on)
            D0 = torch.zeros(M, N, device="cuda", dtype=BF)
            D1 = torch.zeros(M, N, device="cuda", dtype=BF) if epi == 1 else None
            _gemm(M, N, K, A, B, b_mn, epi, aux, bias, D0, D1)
            torch.cuda.synchronize()
            outs.append((D0, D1))
    finally:
        L.lib().vlpk_debug_set_option(b"tail_split", 0)
    assert torch.equal(outs[0][0], outs[1][0]) and float(outs[0][0].float().abs().sum()) > 0
    if epi == 1:
        assert torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("dtype", [torch.int64, torch.float32, torch.bfloat16])
def test_warp_per_row_mask_pack_matches_the_row_walk(dtype):
    """Option "mask_pack_warp": same bitmask as the validated one-thread-per-row kernel, 3-D and broadcast 2-D masks, ragged kv."""
    from vlp_b200 import ops
    g = torch.Generator().manual_seed(2)
    for (B, R, KV) in ((64, 123, 123), (3, 1, 77), (2, 2, 128), (5, 15, 15)):
        m01 = (torch.rand(B, R, KV, generator=g) < 0.6).to(torch.int64)
        mask = m01.cuda() if dtype == torch.int64 else ((1 - m01).to(torch.float32) * -10000.0).to(dtype).cuda()
        mode = "zero_one" if dtype == torch.int64 else "additive"
        outs = []
        try:
            for on in (0, 1):
                L.lib().vlpk_debug_set_option(b"mask_pack_warp", on)
                outs.append(ops.pack_mask(mask, mode=mode).clone())
                torch.cuda.synchronize()
        finally:
            L.lib().vlpk_debug_set_option(b"mask_pack_warp", 0)
        assert torch.equal(outs[0], outs[1]) and int(outs[0].abs().sum()) != 0
