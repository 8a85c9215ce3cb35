"""Turns ncu output (brought back in gpurun_out/) into the tracked summaries under profiles/.

    python tools/ncu_summary.py launches gpurun_out/r02_launches.csv profiles/r02_launches.md "<title / command>"
    python tools/ncu_summary.py full gpurun_out/r02_full.ncu-rep profiles/r02_ncu_full_summary.json [profiles/r02_ncu_traffic.json]

`launches`: the --metrics gpu__time_duration.sum --csv launch list, aggregated by kernel (cold-cache, serialised: compare SHARES).
`full`    : a --set full report; per launch the duration, DRAM bytes, tensor-pipe / SM / DRAM / L2 throughput, occupancy, IPC and the
            top warp-stall reasons; optionally the average DRAM bytes per GEMM launch (bench.py's roofline.traffic, labelled static).
"""
import csv
import io
import json
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*$", "", name)
    return name.replace("vlpk::", "").replace("void ", "").strip()


def launches(src, dst, title):
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        us = float(r["Metric Value"].replace(",", "")) / (1e3 if r["Metric Unit"] == "ns" else 1.0)
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0.0, 0])
        a[0] += us
        a[1] += 1
        total += us
    n = sum(a[1] for a in agg.values())
    out = [f"# {title}", "",
           "`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none` — cold-cache, serialised replays: compare SHARES, "
           "not absolutes.", "", f"{n} launches, total {total / 1e3:.2f} ms.", "", "| share | total us | launches | avg us | kernel |", "|---|---|---|---|---|"]
    for k, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        out.append(f"| {100 * us / total:.1f}% | {us:.0f} | {c} | {us / c:.1f} | `{k[:90]}` |")
    open(dst, "w").write("\n".join(out) + "\n")
    print(f"{dst}: {n} launches, {total / 1e3:.2f} ms")


KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__inst_executed.avg.per_cycle_elapsed"]


def full(src, dst, traffic_dst=None):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    out = []
    for r in data:
        e = OrderedDict(kernel=short(r[ki]))
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                try:
                    e[k] = float(r[i].replace(",", ""))
                except ValueError:
                    e[k] = r[i]
                e.setdefault("_units", {})[k] = units[i]
        st = []
        for i, h in stall:
            try:
                st.append((float(r[i].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
        st.sort(reverse=True)
        e["top_stalls_per_issue"] = {n: round(v, 3) for v, n in st[:5] if n != "selected"}
        out.append(e)
    u = out[0].get("_units", {}) if out else {}
    for e in out:
        e.pop("_units", None)
    json.dump({"source": src, "units": u, "launches": out}, open(dst, "w"), indent=1)
    print(f"{dst}: {len(out)} launches")
    if traffic_dst:
        g = [e for e in out if e["kernel"].startswith("gemm_kernel")]
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
        tot = sum(e["dram__bytes_read.sum"] * scale[u["dram__bytes_read.sum"]] + e["dram__bytes_write.sum"] * scale[u["dram__bytes_write.sum"]] for e in g)
        json.dump({"gemm_dram_bytes_per_launch_avg": tot / max(len(g), 1), "gemm_launches": len(g),
                   "source": f"ncu --set full of one 2-layer training step (tools/prof_step.py --layers 2), {len(g)} gemm_kernel launches (fwd, dgrad, wgrad, "
                             f"region projections, MLM head); {src}"}, open(traffic_dst, "w"), indent=1)
        print(f"{traffic_dst}: {tot / max(len(g), 1) / 1e6:.1f} MB per GEMM launch over {len(g)} launches")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "ncu launch list")
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
