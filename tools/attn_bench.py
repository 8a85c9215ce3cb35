"""Times vlpk_attn_core_fwd / bwd alone at the production shape (B = 64, 12 heads, L = 123, dropout 0.1): median CUDA-event time of a
loop of `reps` back-to-back launches over rotating buffers.  python tools/attn_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from vlp_b200 import _lib as L

DEV, BF = "cuda", torch.bfloat16


def main():
    B, heads, Lq, H = 64, 12, 123, 768
    nset = 6
    torch.manual_seed(0)
    qkvs = [torch.randn(B, Lq, 3 * H, device=DEV).to(BF) for _ in range(nset)]
    dctxs = [torch.randn(B, Lq, H, device=DEV).to(BF) for _ in range(nset)]
    ctx = [torch.zeros(B, Lq, H, device=DEV, dtype=BF) for _ in range(nset)]
    dqkv = [torch.zeros(B, Lq, 3 * H, device=DEV, dtype=BF) for _ in range(nset)]
    lse = [torch.zeros(B, heads, Lq, device=DEV) for _ in range(nset)]
    mask = torch.zeros(B, Lq, Lq, device=DEV, dtype=torch.int64)
    mask[:, :, :102] = 1
    mask[:, 102:, 102:] = torch.tril(torch.ones(21, 21, device=DEV, dtype=torch.int64))
    bits = torch.zeros(B, Lq, 4, device=DEV, dtype=torch.int32)
    L.call("vlpk_mask_pack", mask.data_ptr(), 2, 1, B, Lq, Lq, Lq * Lq, Lq, bits.data_ptr(), L.stream())
    drop = L.VlpkDropout(0.1, 99, None)

    def fwd(i):
        q = qkvs[i]
        L.call("vlpk_attn_core_fwd", B, heads, Lq, Lq, q.data_ptr(), 3 * H, q[..., H:].data_ptr(), q[..., 2 * H:].data_ptr(), 3 * H, bits.data_ptr(),
               Lq, ctx[i].data_ptr(), H, lse[i].data_ptr(), drop, 3, L.stream())

    def bwd(i):
        q = qkvs[i]
        L.call("vlpk_attn_core_bwd", B, heads, Lq, q.data_ptr(), q[..., H:].data_ptr(), q[..., 2 * H:].data_ptr(), 3 * H, bits.data_ptr(), Lq,
               ctx[i].data_ptr(), dctxs[i].data_ptr(), H, lse[i].data_ptr(), dqkv[i].data_ptr(), dqkv[i][..., H:].data_ptr(),
               dqkv[i][..., 2 * H:].data_ptr(), 3 * H, drop, 3, L.stream())

    for name, fn in (("attn fwd", fwd), ("attn bwd", bwd)):
        for i in range(nset):
            fn(i)
        torch.cuda.synchronize()
        ts = []
        reps = 24
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for r in range(reps):
                fn(r % nset)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / reps)
        ts.sort()
        print(f"{name}: {ts[len(ts) // 2]:.1f} us per launch (median of 7 x {reps} back-to-back)")


if __name__ == "__main__":
    main()
