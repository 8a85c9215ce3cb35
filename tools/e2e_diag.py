"""Host-side time breakdown of one staged training step (where does the e2e loop spend host time?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vlp_b200 import staging, synth

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
model, d = bench.build_model(dev, "img2txt")
B = 64
host = synth.make_batch(d, B, seed=1234)
lb, md = staging.describe_mask(host["input_mask"], d.regions)
compact = {k: v for k, v in host.items() if k != "input_mask"}
compact["img"] = compact["img"].bfloat16(); compact["vis_pe"] = compact["vis_pe"].bfloat16()
compact["len_b"], compact["mode"] = lb, md
stager = staging.BatchStager(dev, len_vis_input=d.regions, max_len=d.seq_len)
fields = {k: (tuple(v.shape), v.dtype) for k, v in compact.items()}
for _ in range(2):
    slot = stager.slot(fields)
    for k, v in compact.items():
        slot[k].copy_(v)
    print("pinned?", {k: v.is_pinned() for k, v in slot.items()})
    stager.put(slot); stager.get().done()
dev_batch = {k: v.to(dev) for k, v in host.items()}
dev_batch["img"] = dev_batch["img"].bfloat16(); dev_batch["vis_pe"] = dev_batch["vis_pe"].bfloat16()

def step(b):
    model.zero_grad(set_to_none=True)
    return bench.step_fn(model, b, "img2txt")

for _ in range(5):
    step(dev_batch)
torch.cuda.synchronize()
# host time of a device-resident step (enqueue only)
t = []
for _ in range(10):
    t0 = time.perf_counter(); step(dev_batch); t.append(time.perf_counter() - t0)
torch.cuda.synchronize()
print("host enqueue time, device-resident step: ms", [round(x * 1e3, 2) for x in t])
stager.put(stager.slot(fields))
tp, tg, ts = [], [], []
for i in range(12):
    t0 = time.perf_counter(); b = stager.get(); t1 = time.perf_counter()
    stager.put(stager.slot(fields)); t2 = time.perf_counter()
    loss = step(b); b.done(); t3 = time.perf_counter()
    tg.append(t1 - t0); tp.append(t2 - t1); ts.append(t3 - t2)
torch.cuda.synchronize()
print("get  ms", [round(x * 1e3, 2) for x in tg])
print("put  ms", [round(x * 1e3, 2) for x in tp])
print("step ms", [round(x * 1e3, 2) for x in ts])
# raw copy cost
x = slot["img"]
torch.cuda.synchronize(); t0 = time.perf_counter(); y = x.to(dev, non_blocking=True); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"img H2D ({x.numel()*2/1e6:.1f} MB): enqueue {1e3*(t1-t0):.2f} ms, complete {1e3*(t2-t0):.2f} ms")
