"""Tile-shape sweep for the hot GEMM shapes (B = 64 -> M = 7872), L2 flushed between iterations so that operands come from
HBM as they do inside a training step.  Used to calibrate the cost model in csrc/gemm.cu.  python tools/tune_tiles.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vlp_b200 import _lib as L

DEV, BF = "cuda", torch.bfloat16
M = 7872
flush = torch.empty(256 * 1024 * 1024 // 4, device=DEV, dtype=torch.float32)


def timeit(fn, iters=12):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def gemm(Mg, N, K, A, B, a_mn=0, b_mn=0, bias=None, epi=0, aux=None, splits=1, bn=0, D0=None, D1=None, colsum=None):
    L.call("vlpk_gemm", Mg, N, K, a_mn, A.data_ptr(), A.stride(0), b_mn, B.data_ptr(), B.stride(0), L.ptr(bias), D0.data_ptr(), D0.stride(0),
           L.ptr(D1), D1.stride(0) if D1 is not None else 0, L.ptr(aux), aux.stride(0) if aux is not None else 0, epi, splits, bn, L.stream())


def main():
    torch.manual_seed(0)
    print("shape / epilogue                      " + "".join(f"{'cg%d bn%d' % (cg, bn):>12s}" for cg in (1, 2) for bn in (128, 192, 256)) + f"{'auto':>12s}")
    cases = [("fwd QKV   N2304 K768  store", 2304, 768, 0, 0), ("fwd out   N768  K768  store", 768, 768, 0, 0),
             ("fwd FFNup N3072 K768  gelu", 3072, 768, 0, 1), ("fwd FFNdn N768  K3072 store", 768, 3072, 0, 0),
             ("dgrad dU  N3072 K768  mul", 3072, 768, 1, 4), ("dgrad dy1 N768  K3072 add", 768, 3072, 1, 3),
             ("dgrad dx  N768  K2304 add", 768, 2304, 1, 3), ("dgrad dctx N768 K768  store", 768, 768, 1, 0)]
    for name, N, K, b_mn, epi in cases:
        A = torch.randn(M, K, device=DEV).to(BF)
        Bm = (torch.randn(K, N, device=DEV) if b_mn else torch.randn(N, K, device=DEV)).to(BF)
        bias = None if b_mn else torch.randn(N, device=DEV).to(BF)
        D0 = torch.zeros(M, N, device=DEV, dtype=BF)
        D1 = torch.zeros(M, N, device=DEV, dtype=BF) if epi == 1 else None
        aux = torch.randn(M, N, device=DEV).to(BF) if epi in (3, 4) else None
        row = f"{name:38s}"
        for cg in (1, 2):
            for bn in (128, 192, 256):
                if bn == 192 and b_mn:
                    row += f"{'-':>12s}"
                    continue
                L.lib().vlpk_debug_set_cta_group(cg)
                try:
                    t = timeit(lambda: gemm(M, N, K, A, Bm, b_mn=b_mn, bias=bias, epi=epi, aux=aux, bn=bn, D0=D0, D1=D1))
                    row += f"{t:12.1f}"
                except RuntimeError:
                    row += f"{'n/a':>12s}"
        L.lib().vlpk_debug_set_cta_group(0)
        t = timeit(lambda: gemm(M, N, K, A, Bm, b_mn=b_mn, bias=bias, epi=epi, aux=aux, bn=0, D0=D0, D1=D1))
        row += f"{t:12.1f}"
        print(row)
    print("wgrad (auto split-K)")
    for Nf, Kf in [(768, 768), (768, 3072), (3072, 768), (2304, 768)]:
        dY = torch.randn(M, Nf, device=DEV).to(BF)
        X = torch.randn(M, Kf, device=DEV).to(BF)
        D = torch.zeros(Nf, Kf, device=DEV)
        row = f"  dW [{Nf}x{Kf}]".ljust(38)
        for cg in (1, 2):
            for bn in (128, 256):
                L.lib().vlpk_debug_set_cta_group(cg)
                t = timeit(lambda: gemm(Nf, Kf, M, dY, X, a_mn=1, b_mn=1, epi=6, splits=0, bn=bn, D0=D))
                row += f"  cg{cg}bn{bn}:{t:6.1f}"
        L.lib().vlpk_debug_set_cta_group(0)
        t = timeit(lambda: gemm(Nf, Kf, M, dY, X, a_mn=1, b_mn=1, epi=6, splits=0, bn=0, D0=D))
        print(row + f"  auto:{t:6.1f}")


if __name__ == "__main__":
    main()
