"""BertAdam.step(): host enqueue time and device time per step on the BERT-base parameter set."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vlp_b200.optimization import BertAdam

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
model, d = bench.build_model(dev, "img2txt")
named = [(n, p) for n, p in model.named_parameters()]
no_decay = ("bias", "LayerNorm.bias", "LayerNorm.weight")
groups = [{"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": 0.01},
          {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0}]
opt = BertAdam(groups, lr=3e-5, warmup=0.1, t_total=100000)
for n, p in named:
    p.grad = torch.randn_like(p) * 0.01
for _ in range(3):
    opt.step()
torch.cuda.synchronize()
hs, ds = [], []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record(); opt.step(); e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    hs.append((t1 - t0) * 1e3); ds.append(e0.elapsed_time(e1))
print("BertAdam.step host enqueue ms:", [round(x, 2) for x in hs])
print("BertAdam.step device ms      :", [round(x, 2) for x in ds])
nparam = sum(p.numel() for _, p in named)
print(f"{nparam/1e6:.1f} M parameters; 30 B/param algorithmic -> {nparam*30/1e9:.2f} GB; at {min(ds):.3f} ms = {nparam*30/min(ds)/1e6:.0f} GB/s")
