"""Launches a few hot GEMMs once each (after one warm-up each) for an ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -o gpurun_out/prof python tools/gemm_probe.py [names...]
names: ffn_down, dy1, w2, ffn_up, out, qkv (default: ffn_down dy1 w2)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tools.gemm_vs_cublas import vgemm

DEV, BF = "cuda", torch.bfloat16
M, H, I = 7872, 768, 3072


def main():
    names = sys.argv[1:] or ["ffn_down", "dy1", "w2"]
    torch.manual_seed(0)
    flush = torch.empty(512 * 1024 * 1024 // 4, device=DEV, dtype=torch.float32)
    for name in names:
        if name in ("ffn_down", "ffn_up", "out", "qkv"):
            N, K, epi = {"ffn_down": (H, I, 0), "ffn_up": (I, H, 1), "out": (H, H, 0), "qkv": (3 * H, H, 0)}[name]
            x = torch.randn(M, K, device=DEV).to(BF)
            w = (torch.randn(N, K, device=DEV) * 0.05).to(BF)
            b = torch.randn(N, device=DEV).to(BF)
            y = torch.empty(M, N, device=DEV, dtype=BF)
            y1 = torch.empty(M, N, device=DEV, dtype=BF) if epi == 1 else None
            fn = lambda: vgemm(M, N, K, x, w, bias=b, epi=epi, D0=y, D1=y1)
        elif name == "dy1":
            dy = torch.randn(M, I, device=DEV).to(BF)
            w = (torch.randn(I, H, device=DEV) * 0.05).to(BF)
            aux = torch.randn(M, H, device=DEV).to(BF)
            dx = torch.empty(M, H, device=DEV, dtype=BF)
            fn = lambda: vgemm(M, H, I, dy, w, b_mn=1, epi=3, aux=aux, D0=dx)
        elif name == "w2":
            dy = torch.randn(M, H, device=DEV).to(BF)
            x = torch.randn(M, I, device=DEV).to(BF)
            dw = torch.zeros(H, I, device=DEV, dtype=torch.float32)
            fn = lambda: vgemm(H, I, M, dy, x, a_mn=1, b_mn=1, epi=6, splits=0, D0=dw)
        else:
            raise SystemExit(f"unknown {name}")
        for _ in range(2):
            flush.zero_()
            fn()
        torch.cuda.synchronize()


if __name__ == "__main__":
    main()
