"""Per-shape comparison of vlpk::gemm_kernel with cuBLAS (torch.matmul / F.linear) on the 12 hot GEMM shapes of one BertLayer at
B = 64 (M = 7872 token rows): 4 forward, 4 dgrad, 4 wgrad.  Same box, same clocks, CUDA events on the launch stream, median of
`iters` runs; "cold" = a 512 MB buffer is rewritten between iterations (operands come from HBM, as inside a training step),
"warm" = back-to-back (operands L2-resident).  Writes a markdown table (default gpurun_out/r02_gemm_vs_cublas.md).

    python tools/gemm_vs_cublas.py [out.md]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from vlp_b200 import _lib as L

DEV, BF = "cuda", torch.bfloat16
M = 7872
H, I = 768, 3072


def timeit(fn, flush, iters=15):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def vgemm(Mg, N, K, A, B, a_mn=0, b_mn=0, bias=None, epi=0, aux=None, splits=1, bn=0, D0=None, D1=None):
    L.call("vlpk_gemm", Mg, N, K, a_mn, A.data_ptr(), A.stride(0), b_mn, B.data_ptr(), B.stride(0), L.ptr(bias), D0.data_ptr(), D0.stride(0),
           L.ptr(D1), D1.stride(0) if D1 is not None else 0, L.ptr(aux), aux.stride(0) if aux is not None else 0, epi, splits, bn, L.stream())


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02_gemm_vs_cublas.md"
    torch.manual_seed(0)
    flush = torch.empty(512 * 1024 * 1024 // 4, device=DEV, dtype=torch.float32)
    rows = []
    # (name, kind, N_out, K_in, epilogue)  forward: y[M,N] = x[M,K] W[N,K]^T + b
    fwd = [("fwd QKV", 3 * H, H, 0), ("fwd attn-out", H, H, 0), ("fwd FFN-up+GELU", I, H, 1), ("fwd FFN-down", H, I, 0)]
    for name, N, K, epi in fwd:
        x = torch.randn(M, K, device=DEV).to(BF)
        w = (torch.randn(N, K, device=DEV) * 0.05).to(BF)
        b = torch.randn(N, device=DEV).to(BF)
        y = torch.empty(M, N, device=DEV, dtype=BF)
        y1 = torch.empty(M, N, device=DEV, dtype=BF) if epi == 1 else None
        ours = lambda: vgemm(M, N, K, x, w, bias=b, epi=epi, D0=y, D1=y1)
        ref = lambda: F.linear(x, w, b)
        rows.append((name, M, N, K, 2.0 * M * N * K, [timeit(ours, flush), timeit(ref, flush), timeit(ours, None), timeit(ref, None)],
                     "cuBLAS: linear+bias" + (" (no GELU, one output)" if epi == 1 else "")))
    # dgrad: dx[M,K] = dy[M,N] W[N,K]  (ours: contraction over N with W read MN-major)
    dg = [("dgrad dU (x gelu')", H, I, 4), ("dgrad dy1 (+res)", I, H, 3), ("dgrad dx (+res)", 3 * H, H, 3), ("dgrad dctx", H, H, 0)]
    for name, N, K, epi in dg:
        dy = torch.randn(M, N, device=DEV).to(BF)
        w = (torch.randn(N, K, device=DEV) * 0.05).to(BF)
        aux = torch.randn(M, K, device=DEV).to(BF) if epi in (3, 4) else None
        dx = torch.empty(M, K, device=DEV, dtype=BF)
        ours = lambda: vgemm(M, K, N, dy, w, b_mn=1, epi=epi, aux=aux, D0=dx)
        ref = lambda: torch.matmul(dy, w)
        rows.append((name, M, K, N, 2.0 * M * N * K, [timeit(ours, flush), timeit(ref, flush), timeit(ours, None), timeit(ref, None)],
                     "cuBLAS: plain matmul (no fused add / mul)"))
    # wgrad: dw[N,K] = dy[M,N]^T x[M,K]   (ours: split-K, fp32 TMA reduce-add into a pre-zeroed buffer)
    wg = [("wgrad Wqkv", 3 * H, H), ("wgrad Wo", H, H), ("wgrad W1", I, H), ("wgrad W2", H, I)]
    for name, N, K in wg:
        dy = torch.randn(M, N, device=DEV).to(BF)
        x = torch.randn(M, K, device=DEV).to(BF)
        dw = torch.zeros(N, K, device=DEV, dtype=torch.float32)
        ours = lambda: vgemm(N, K, M, dy, x, a_mn=1, b_mn=1, epi=6, splits=0, D0=dw)
        ref = lambda: torch.matmul(dy.t(), x)
        rows.append((name, N, K, M, 2.0 * M * N * K, [timeit(ours, flush), timeit(ref, flush), timeit(ours, None), timeit(ref, None)],
                     "ours: fp32 reduce-add output; cuBLAS: bf16 output"))
    lines = ["# r02 — vlpk::gemm_kernel vs cuBLAS on the hot shapes of one BertLayer (B = 64, M = 7872)", "",
             f"device: {torch.cuda.get_device_name(0)}; torch {torch.__version__}; median of 15, CUDA events; cold = 512 MB rewritten between iterations", "",
             "| GEMM | M x N x K | ours cold us | cuBLAS cold us | ours cold TF/s | cuBLAS cold TF/s | ours/cuBLAS | ours warm us | cuBLAS warm us | note |",
             "|---|---|---|---|---|---|---|---|---|---|"]
    tot = [0.0, 0.0, 0.0, 0.0]
    for name, m, n, k, fl, t, note in rows:
        for i in range(4):
            tot[i] += t[i]
        lines.append(f"| {name} | {m} x {n} x {k} | {t[0]:.1f} | {t[1]:.1f} | {fl / t[0] / 1e6:.0f} | {fl / t[1] / 1e6:.0f} | {t[1] / t[0]:.2f} | {t[2]:.1f} | {t[3]:.1f} | {note} |")
    flops = sum(r[4] for r in rows)
    lines.append(f"| **sum (one layer fwd+bwd)** | | {tot[0]:.1f} | {tot[1]:.1f} | {flops / tot[0] / 1e6:.0f} | {flops / tot[1] / 1e6:.0f} | {tot[1] / tot[0]:.2f} | {tot[2]:.1f} | {tot[3]:.1f} | |")
    text = "\n".join(lines) + "\n"
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    open(out_path, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
