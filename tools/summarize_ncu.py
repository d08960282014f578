"""Turns the ncu artefacts a gpurun call left in gpurun_out/ into the small text
summaries committed under profiles/ (the .ncu-rep files themselves stay in gpurun_out/).

  python tools/summarize_ncu.py r1     # reads gpurun_out/r1_*  -> profiles/r1_*.md|json
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__cluster_size", "sm__inst_executed.avg.per_cycle_elapsed",
]


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def raw_table(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return hdr, units, data


def launch_list():
    path = os.path.join(G, f"{tag}_launches.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("cvx::<unnamed>::", "").replace("cvx::(anonymous namespace)::", "")
        v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(r[ui], 1)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, f"{tag}_launch_list.md"), "w") as f:
        f.write(f"# {tag}: kernel launch list (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n"
                "Command: `ncu ... -s 200 -c 400 python bench.py --steps 2 --warmup 3 --utts-per-gpu 120 --no-cpu-baseline`\n"
                "(per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's\n"
                "`kernel_ms_per_step`, not absolutes)\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / tot:.3f} |\n")
    print("launch list:", {k: round(v[1] / tot, 3) for k, v in agg.items()})


def kernel_report(name, title, note=""):
    rep = os.path.join(G, f"{tag}_prof_{name}.ncu-rep")
    if not os.path.exists(rep):
        return None
    hdr, units, data = raw_table(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [m for m in METRICS if m in idx]
    lines = [f"# {tag}: {title} (ncu --set full --clock-control none)\n", note, "",
             "| launch | " + " | ".join(c.replace("sm__", "").replace(".avg", "").replace("pct_of_peak_sustained_", "%") for c in cols) + " |",
             "|---|" + "---|" * len(cols)]
    traffic = []
    kn = idx.get("Kernel Name")
    for i, r in enumerate(data):
        vals = []
        for c in cols:
            v, u = r[idx[c]], units[idx[c]]
            vals.append(f"{v} {u}".strip())
        tagname = (r[kn].split("<")[0].split("::")[-1] if kn is not None else "")
        lines.append(f"| {i} {tagname} | " + " | ".join(vals) + " |")
        traffic.append(to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) +
                       to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]]))
    open(os.path.join(P, f"{tag}_{name}.md"), "w").write("\n".join(lines) + "\n")
    return traffic


os.makedirs(P, exist_ok=True)
launch_list()
t_gemm = kernel_report("gemm", "TDNN GEMM kernels, 10 consecutive launches = the 10 frame layers of one 524k-row batch "
                       "(v5 c-vector, plan order: tdnn1_share, tdnn2-4_xvec, AM tdnn1-5, tdnn5_xvec = pooled layer with the "
                       "fused-pooling epilogue; multi-tap layers run tdnn_gemm_halo_kernel, the others tdnn_gemm_pair_kernel)",
                       "Command: `ncu --set full --clock-control none --import-source on -k regex:tdnn_gemm -s 20 -c 10 python bench.py "
                       "--steps 1 --warmup 3 --utts-per-gpu 60 --no-cpu-baseline`")
t_pool = kernel_report("pool", "stats_pool_partial_kernel<__half> (stand-alone pooling, CVX_FUSED_POOL=0)",
                       "Command: `CVX_FUSED_POOL=0 ncu --set full ... -k regex:stats_pool_partial -s 2 -c 2 ...`")
kernel_report("stage", "stage_input_kernel<__half>", "Command: `ncu --set full ... -k regex:stage_input -s 2 -c 1 ...`")
out = {}
if t_gemm:
    out["tdnn_gemm_dram_bytes_per_launch"] = sum(t_gemm) / len(t_gemm)
    out["tdnn_gemm_dram_bytes_by_launch"] = t_gemm
if t_pool:
    out["stats_pool_dram_bytes_per_launch"] = sum(t_pool) / len(t_pool)
    out["stats_pool_dram_bytes_by_launch"] = t_pool
out["source"] = f"ncu --set full captures of round {tag} (see {tag}_gemm.md, {tag}_pool.md); bench.py reads this file"
json.dump(out, open(os.path.join(P, "traffic.json"), "w"), indent=1)
print(json.dumps(out)[:600])
