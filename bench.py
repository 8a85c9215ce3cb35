"""bench.py — VLP hot path throughput on B200 (driver contract: one JSON line on stdout from rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload = BASELINE.json configs[1]: BERT-base 12L/768H, 100 regions x 2048 + 20 tokens (L = 123), batch 64 per GPU,
bf16, seq2seq mask, 3 masked positions, dropout 0.1 (the reference's training setting), forward + backward through
BertForPreTrainingLossMask (+ NCCL gradient all-reduce when N > 1; weak scaling).  One step = one batch.

  value : samples/s, inputs resident in HBM, device-timed (CUDA events), max over ranks.
  e2e   : same metric through the public module API with HOST buffers: every step's 12-tensor batch is copied from pinned
          host memory (double-buffered on a copy stream) and the loss is read back to the host, all inside the timed region.
  roofline : dominant kernel family (tcgen05 GEMM), algorithmic FLOPs / CUDA-event time of its launches inside one real step.
  cpu_baseline : the fp32 oracle port (same unfused eager op sequence as the reference) on the host cores, bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "image_text_samples_per_sec"
WORKLOAD = ("BASELINE.json configs[1]: BERT-base 12L/768H/12 heads/3072, 100 regions x 2048-d + 20 text tokens (L=123), "
            "batch 64 per GPU, seq2seq mask, 3 masked positions, dropout 0.1, fwd+bwd")
PER_GPU_BATCH = 64


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port timed on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(batch=8, warmup=1, steps=2):
    """fp32, eager, train mode, dropout 0.1 — the reference's own op sequence (oracle/vlp_oracle.py restates it op for op)."""
    from oracle import vlp_oracle as O
    from vlp_b200 import synth
    avail = os.cpu_count() or 1
    d = synth.BERT_BASE
    sd = synth.make_state_dict(d, 0)
    for k, v in sd.items():
        if k != "cls.predictions.decoder.weight":
            v.requires_grad_(True)
    b = synth.make_batch(d, batch, seed=1234)
    # "all the host threads it can use": eager PyTorch on small per-op tensors slows down when oversubscribed, so probe a few
    # thread counts on one forward pass and keep the fastest (the count actually used is reported as `cores`).
    cands = sorted({c for c in (8, 16, 32, 64, avail) if c <= avail})
    probe = {}
    for c in cands:
        torch.set_num_threads(c)
        with torch.no_grad():
            O.pretraining_loss(sd, d, b)
            t0 = time.perf_counter()
            O.pretraining_loss(sd, d, b)
            probe[c] = time.perf_counter() - t0
    cores = min(probe, key=probe.get)
    torch.set_num_threads(cores)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        for v in sd.values():
            v.grad = None
        loss = O.pretraining_loss(sd, d, b, p_hidden=0.1, p_attn=0.1, training=True)[0]
        loss.backward()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    per_step = sum(times) / len(times)
    return {"value": batch / per_step, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"oracle port (fp32 eager PyTorch, train mode, dropout 0.1), BERT-base L=123, batch {batch}, {warmup} warm-up + {steps} timed fwd+bwd steps; "
                      f"{cores} of {avail} host threads (fastest of {cands} on a forward probe)",
            "s_per_step": per_step}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = cpu_reference_run(batch=8, warmup=min(args.warmup, 1), steps=max(1, min(args.steps, 3)))
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": res["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "note": "CPU arm: bounded sample, batch 8 per step"},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def build_model(device):
    from vlp_b200 import synth
    from vlp_b200 import vlp_modules as vm
    d = synth.BERT_BASE
    cfg = vm.BertConfig(d.vocab, hidden_size=d.hidden, num_hidden_layers=d.layers, num_attention_heads=d.heads, intermediate_size=d.inter,
                        type_vocab_size=d.type_vocab, max_position_embeddings=d.max_pos, hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1)
    torch.manual_seed(0)
    model = vm.BertForPreTrainingLossMask(cfg, enable_butd=True, len_vis_input=d.regions, tasks="img2txt")
    model = model.to(device=device, dtype=torch.bfloat16).train()
    # bert.pooler.* never receives a gradient in img2txt training (SURVEY.md §7 "DDP unused parameters"); freezing it replaces the
    # reference's find_unused_parameters=True graph walk.
    for p in model.bert.pooler.parameters():
        p.requires_grad_(False)
    return model, d


BATCH_ORDER = ["input_ids", "segment_ids", "input_mask", "masked_ids", "masked_pos", "masked_weights", "is_next", "task_idx", "img",
               "vis_masked_pos", "vis_pe", "ans_labels"]


def step_fn(model, b):
    """The reference's training-loop body, run_img2txt_dist.py:479-483 + :575 (loss.backward())."""
    loss_tuple = model(b["img"], b["vis_pe"], b["input_ids"], b["segment_ids"], b["input_mask"], b["masked_ids"], None, b["is_next"],
                       masked_pos=b["masked_pos"], masked_weights=b["masked_weights"], task_idx=b["task_idx"],
                       vis_masked_pos=b["vis_masked_pos"], mask_image_regions=False, drop_worst_ratio=0.0)
    loss = loss_tuple[0] + loss_tuple[1] + loss_tuple[2]
    loss.backward()
    return loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="vlp_b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch.distributed as dist
    from vlp_b200 import _lib as L
    from vlp_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    warmup = max(args.warmup, 3)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    model, d = build_model(device)
    net = model
    reducer = None
    if world > 1:
        # backward runs in 4 groups of 3 layers; each group's gradients are one contiguous bf16 arena (42.5 MB) that is handed to
        # NCCL (all-reduce AVG over NVLink) as soon as the group finishes, while the next group is still computing.
        if os.environ.get("VLP_BENCH_DP", "arena") == "torch_ddp":
            model.bert.encoder.layers_per_call = 3
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], gradient_as_bucket_view=True, bucket_cap_mb=45,
                                                            broadcast_buffers=False)
        else:
            from vlp_b200.dp import GradientAllReducer
            reducer = GradientAllReducer(model, layers_per_call=3)
            reducer.broadcast_parameters(0)
    B = PER_GPU_BATCH
    host = synth.make_batch(d, B, seed=1234 + rank, mode="s2s")

    def to_dev(hb, non_blocking=False):
        out = {}
        for k in BATCH_ORDER:
            out[k] = hb[k].to(device, non_blocking=non_blocking)
        return out

    dev_batch = to_dev(host)
    dev_batch["img"] = dev_batch["img"].bfloat16()
    dev_batch["vis_pe"] = dev_batch["vis_pe"].bfloat16()

    def barrier():
        if world > 1:
            dist.barrier()

    def run_steps(n, get_batch, read_loss=None):
        for i in range(n):
            net.zero_grad(set_to_none=True)
            loss = step_fn(net, get_batch(i))
            if reducer is not None:
                reducer.finish()
            if read_loss is not None:
                read_loss(i, loss)

    # ---------------- device-resident timing ("value") ----------------
    run_steps(warmup, lambda i: dev_batch)
    if world > 1:
        # after the all-reduce every rank must hold the same averaged gradients (ranks see different data shards)
        enc = model.bert.encoder.layer
        chk = torch.stack([enc[0].attention.self.query.weight.grad.float().sum(), enc[11].output.dense.weight.grad.float().sum(),
                           model.vis_embed[0].weight.grad.float().sum(), model.bert.embeddings.word_embeddings.weight.grad.float().sum()])
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        assert all(torch.equal(allc[0], c) for c in allc), f"gradients differ across ranks after all-reduce: {allc}"

    torch.cuda.synchronize()
    barrier()
    launches0 = L.lib().vlpk_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        torch.cuda.synchronize()
        e0.record()
        run_steps(args.steps, lambda i: dev_batch)
        e1.record()
        torch.cuda.synchronize()
    barrier()
    launches = L.lib().vlpk_launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    ms_per_step = ms_total / args.steps
    value = B * world * args.steps / (ms_total / 1e3)

    # ---------------- end-to-end timing with host buffers ("e2e") ----------------
    pinned = [{k: host[k].clone().pin_memory() for k in BATCH_ORDER} for _ in range(2)]
    h2d_bytes = sum(pinned[0][k].numel() * pinned[0][k].element_size() for k in BATCH_ORDER)
    copy_stream = torch.cuda.Stream()
    loss_host = torch.zeros(args.steps + warmup, dtype=torch.float32).pin_memory()
    slots = [None, None]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        s = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[s])
            slots[s] = to_dev(pinned[s], non_blocking=True)
            ready[s].record(copy_stream)

    def get_e2e_batch(i):
        s = i & 1
        torch.cuda.current_stream().wait_event(ready[s])
        b = slots[s]
        prefetch(i + 1)
        return b

    def read_loss(i, loss):
        s = i & 1
        consumed[s].record(torch.cuda.current_stream())
        loss_host[i].copy_(loss.detach().float().reshape(()), non_blocking=True)

    for ev in consumed:
        ev.record(torch.cuda.current_stream())
    prefetch(0)
    run_steps(warmup, get_e2e_batch, read_loss)
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(warmup, warmup + args.steps):
        net.zero_grad(set_to_none=True)
        loss_i = step_fn(net, get_e2e_batch(i))
        if reducer is not None:
            reducer.finish()
        read_loss(i, loss_i)
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    barrier()
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = B * world * args.steps / float(t_e2e)
    final_loss = float(loss_host[warmup + args.steps - 1])

    # ---------------- per-kernel-family profile of one real step (roofline) ----------------
    names = ["gemm_fwd", "gemm_dgrad", "gemm_wgrad", "attn_fwd", "attn_bwd", "ln_fwd", "ln_bwd", "embed", "misc"]
    prof = {}
    if rank == 0:
        L.lib().vlpk_profile_reset()
        L.lib().vlpk_profile_enable(1)
        run_steps(2, lambda i: dev_batch)
        torch.cuda.synchronize()
        L.lib().vlpk_profile_enable(0)
        for i, n in enumerate(names):
            a, w, c = C.c_double(), C.c_double(), C.c_int64()
            L.lib().vlpk_profile_get(i, C.byref(a), C.byref(w), C.byref(c))
            prof[n] = {"ms_per_step": a.value / 2, "work_per_step": w.value / 2, "launches_per_step": c.value // 2}
        L.lib().vlpk_profile_reset()
    else:
        run_steps(2, lambda i: dev_batch)
        torch.cuda.synchronize()
    barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    gemm_ms = sum(prof[n]["ms_per_step"] for n in ("gemm_fwd", "gemm_dgrad", "gemm_wgrad"))
    gemm_fl = sum(prof[n]["work_per_step"] for n in ("gemm_fwd", "gemm_dgrad", "gemm_wgrad"))
    gemm_n = sum(prof[n]["launches_per_step"] for n in ("gemm_fwd", "gemm_dgrad", "gemm_wgrad"))
    achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    peak_tf = peaks["bf16_tflops_sustained"]
    fl = synth.flops_per_sample()
    step_tf = value / world * fl["total"] / 1e12
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic, traffic_src = tj["gemm_dram_bytes_per_launch_avg"], tj["source"]
    roofline = {"bound": "tensor", "kernel": "vlpk::gemm_kernel (tcgen05 GEMM family: fwd/dgrad/wgrad)", "achieved": achieved, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']}); kernel timed inside a long step",
                "launches_per_step": gemm_n, "flops_per_launch_avg": gemm_fl / max(gemm_n, 1), "avg_launch_us": gemm_ms * 1e3 / max(gemm_n, 1),
                "gemm_share_of_step": gemm_ms / ms_per_step,
                "whole_step": {"achieved": step_tf, "frac": step_tf / peak_tf, "flops_per_sample": fl["total"]},
                "families_ms_per_step": {n: round(prof[n]["ms_per_step"], 4) for n in names},
                "hbm_kernels": {n: {"GBps": (prof[n]["work_per_step"] / (prof[n]["ms_per_step"] * 1e-3) / 1e9 if prof[n]["ms_per_step"] > 0 else 0.0),
                                    "frac_of_hbm_peak": (prof[n]["work_per_step"] / (prof[n]["ms_per_step"] * 1e-3) / 1e9 / peaks["hbm_gbs"]
                                                         if prof[n]["ms_per_step"] > 0 else 0.0)} for n in ("ln_fwd", "ln_bwd")}}
    line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": B * world, "seq_len": d.seq_len, "parallelism": f"dp{world}" + ("" if world == 1 else (" torch-DDP" if reducer is None else " NCCL all-reduce of flat bf16 gradient arenas overlapped with backward")),
                       "l2": "per-step working set (2.3 GB saved activations + 0.23 GB weights) is far larger than the 126 MB L2; no explicit flush",
                       "timing": "CUDA events on the launch stream, barrier + synchronize both sides, max over ranks"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "how": "pinned host batch (fp32 features + int64 ids/mask as the reference loader emits) -> double-buffered H2D on a copy stream -> "
                           "BertForPreTrainingLossMask fwd+bwd -> loss copied to pinned host memory, host wall clock around the loop"},
            "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks.summary(), "final_loss": final_loss}
    if world == 1 and not args.no_cpu_baseline:
        res = cpu_reference_run(batch=8, warmup=1, steps=2)
        line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
