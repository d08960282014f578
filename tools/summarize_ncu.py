"""Turns the ncu artefacts a gpurun call left in gpurun_out/ (tools/profile_r2.sh) into the small text
summaries committed under profiles/ (the .ncu-rep files themselves stay in gpurun_out/).

  python tools/summarize_ncu.py r2     # reads gpurun_out/r2_* , r2x_*  -> profiles/r2_*.md, tensor_pipe.json, sass_summary.md
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__cluster_size", "sm__inst_executed.avg.per_cycle_elapsed",
]
SHORT = {"gpu__time_duration.sum": "time", "sm__cycles_elapsed.avg.per_second": "SM clock",
         "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor pipe %",
         "sm__throughput.avg.pct_of_peak_sustained_elapsed": "SM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "DRAM %",
         "dram__bytes_read.sum": "DRAM read", "dram__bytes_write.sum": "DRAM write", "lts__t_sector_hit_rate.pct": "L2 hit %",
         "sm__warps_active.avg.pct_of_peak_sustained_active": "warps active %", "launch__registers_per_thread": "regs",
         "launch__grid_size": "grid", "launch__block_size": "block", "launch__cluster_size": "cluster",
         "sm__inst_executed.avg.per_cycle_elapsed": "IPC"}


def num(v):
    return float(v.replace(",", ""))


def to_bytes(v, unit):
    return num(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def to_us(v, unit):
    return num(v) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


def raw_table(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def short_kernel(name):
    n = name.replace("void ", "").replace("cvx::<unnamed>::", "").replace("cvx::(anonymous namespace)::", "").replace("unnamed>::", "")
    return n.split("(")[0]


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
        a = agg.setdefault(short_kernel(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += to_us(r[vi], r[ui])
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(P, f"{tag}_launch_list.md"), "w") as f:
        f.write(f"# {tag}: kernel launch list of bench.py's timed steps (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n"
                "Command (tools/profile_r2.sh): `ncu --metrics gpu__time_duration.sum --clock-control none -s 165 -c 110 --csv "
                "python bench.py --quick --steps 2 --warmup 3 --utts-per-gpu 120` = the two timed device-resident calls "
                "(4 device batches each: pool_meta, stage_input, 10 frame layers, finalize per batch; split + segment GEMM + "
                "utt_average per call).\n(per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's "
                "`kernel_ms_per_step`, not absolutes)\n\n| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / tot:.3f} |\n")
    print("launch list:", {k: round(v[1] / tot, 3) for k, v in agg.items()})


def gemm_report(prefix, title, cmd):
    """Per-layer table of the frame-layer GEMMs of one batch; returns the json summary."""
    rep = os.path.join(G, f"{prefix}_prof_gemm.ncu-rep")
    plan_path = os.path.join(G, f"{prefix}_plan.json")
    if not (os.path.exists(rep) and os.path.exists(plan_path)):
        return None
    plan = json.load(open(plan_path))
    hdr, units, data = raw_table(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    layers = plan["layers"]
    assert len(data) == len(layers), (len(data), len(layers))
    lines = [f"# {prefix}: {title} (ncu --set full --clock-control none)\n", cmd, "",
             f"One device batch of {plan['rows']} rows ({plan['utts']} x {plan['frames']} frames), {plan['topology']}; launches in plan "
             "order.  `alg bytes` = every source buffer read once + output written once (stored pitch, fp16) + weights once "
             "(the pooled layer writes per-slot partial sums instead of its activation).",
             "",
             "| # | layer | K -> N | kernel | time us | SM clock GHz | tensor pipe % | DRAM % | DRAM read MB | DRAM write MB | alg MB | DRAM / alg | L2 hit % | TFLOP/s (this launch) | regs | grid x block (cluster) |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    per_layer, tot_flops, tot_time, wt_flop, wt_time, tot_dram, tot_alg = [], 0.0, 0.0, 0.0, 0.0, 0.0, 0.0
    for i, (r, L) in enumerate(zip(data, layers)):
        g = lambda m: r[idx[m]]
        us = to_us(g("gpu__time_duration.sum"), units[idx["gpu__time_duration.sum"]])
        tp = num(g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))
        rd = to_bytes(g("dram__bytes_read.sum"), units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(g("dram__bytes_write.sum"), units[idx["dram__bytes_write.sum"]])
        alg = L["in_bytes"] + L["out_bytes"] + L["w_bytes"]
        kern = short_kernel(g("Kernel Name"))
        tfl = L["flops"] / (us * 1e-6) / 1e12
        lines.append(f"| {i} | {L['name']}{' (pooled)' if L.get('pooled') else ''} | {L['K']} -> {L['N']} | `{kern}` | {us:.1f} | "
                     f"{num(g('sm__cycles_elapsed.avg.per_second')):.3f} | {tp:.1f} | "
                     f"{num(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | "
                     f"{alg / 1e6:.1f} | {(rd + wr) / alg:.2f} | {num(g('lts__t_sector_hit_rate.pct')):.1f} | {tfl:.0f} | "
                     f"{g('launch__registers_per_thread')} | {g('launch__grid_size')} x {g('launch__block_size')} ({g('launch__cluster_size')}) |")
        per_layer.append({"layer": L["name"], "K": L["K"], "N": L["N"], "kernel": kern, "time_us": round(us, 1),
                          "tensor_pipe_pct": round(tp, 2), "dram_bytes": rd + wr, "algorithmic_bytes": alg,
                          "dram_over_algorithmic": round((rd + wr) / alg, 3)})
        tot_flops += L["flops"]; tot_time += us; wt_flop += tp * L["flops"]; wt_time += tp * us
        tot_dram += rd + wr; tot_alg += alg
    summ = {"tensor_pipe_pct_flop_weighted": round(wt_flop / tot_flops, 2), "tensor_pipe_pct_time_weighted": round(wt_time / tot_time, 2),
            "dram_over_algorithmic": round(tot_dram / tot_alg, 3), "dram_bytes_per_launch_avg": tot_dram / len(layers),
            "algorithmic_bytes_per_launch_avg": tot_alg / len(layers), "per_layer": per_layer}
    lines += ["", f"FLOP-weighted tensor-pipe utilisation over the {len(layers)} frame layers: **{summ['tensor_pipe_pct_flop_weighted']} %** "
              f"(time-weighted {summ['tensor_pipe_pct_time_weighted']} %).  DRAM traffic {tot_dram / 1e9:.2f} GB vs {tot_alg / 1e9:.2f} GB algorithmic "
              f"= **{summ['dram_over_algorithmic']}x**.  (Isolated launches under ncu run at burst clocks; bench.py's `layer_ms_per_step` has the "
              "sustained, power-capped times.)"]
    open(os.path.join(P, f"{prefix}_gemm.md"), "w").write("\n".join(lines) + "\n")
    return summ


def kernel_report(prefix, name, title, note=""):
    rep = os.path.join(G, f"{prefix}_prof_{name}.ncu-rep")
    if not os.path.exists(rep):
        return
    hdr, units, data = raw_table(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [m for m in METRICS if m in idx]
    lines = [f"# {prefix}: {title} (ncu --set full --clock-control none)\n", note, "",
             "| launch | " + " | ".join(SHORT[c] for c in cols) + " | DRAM GB/s |", "|---|" + "---|" * (len(cols) + 1)]
    for i, r in enumerate(data):
        vals = [f"{r[idx[c]]} {units[idx[c]]}".strip() for c in cols]
        rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        us = to_us(r[idx["gpu__time_duration.sum"]], units[idx["gpu__time_duration.sum"]])
        lines.append(f"| {i} `{short_kernel(r[idx['Kernel Name']])}` | " + " | ".join(vals) + f" | {(rd + wr) / us / 1e3:.0f} |")
    open(os.path.join(P, f"{prefix}_{name}.md"), "w").write("\n".join(lines) + "\n")


def sass_summary():
    lib = os.path.join(ROOT, "speaker-embedding-with-phonetic-information_b200", "lib", "libcvx.so")
    if not os.path.exists(lib):
        return
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    pats = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "FHADD", "FHFMA", "HMMA", "SYNCS"]
    per, cur = collections.OrderedDict(), None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = short_kernel(cur)
            per[cur] = collections.Counter()
            continue
        if cur:
            for p in pats:
                if re.search(r"\b" + p + r"[A-Z0-9_.]*", ln):
                    per[cur][p] += 1
    with open(os.path.join(P, "sass_summary.md"), "w") as f:
        f.write("# SASS evidence per kernel (`cuobjdump -sass lib/libcvx.so`, sm_100a)\n\n"
                "`UTCHMMA` = tcgen05.mma kind::f16 (`.2CTA` forms in the pair/halo kernels), `LDTM` = tcgen05.ld, `UTMALDG`/`UTMASTG` = TMA "
                "tile load/store, `UTCBAR` = tcgen05.commit, `FHADD`/`FHFMA` = mixed-precision fp32 += fp16 (pooling), `HMMA` = legacy mma.sync "
                "(must be 0).\n\n| kernel | " + " | ".join(pats) + " |\n|---|" + "---|" * len(pats) + "\n")
        for k, c in per.items():
            if sum(c.values()) == 0 and "tdnn" not in k and "pool" not in k:
                continue
            f.write(f"| `{k}` | " + " | ".join(str(c.get(p, 0)) for p in pats) + " |\n")
    print("sass summary:", len(per), "kernels")


def sparse_vs_dense():
    """same batch, same box, back to back: offset forwarding on (default) vs CVX_FORWARD_OFFSETS=0"""
    ra, rb = os.path.join(G, f"{tag}_prof_gemm.ncu-rep"), os.path.join(G, f"{tag}dense_prof_gemm.ncu-rep")
    plan_path = os.path.join(G, f"{tag}_plan.json")
    if not (os.path.exists(ra) and os.path.exists(rb) and os.path.exists(plan_path)):
        return
    layers = json.load(open(plan_path))["layers"]
    rows = []
    for rep in (ra, rb):
        hdr, units, data = raw_table(rep)
        idx = {h: i for i, h in enumerate(hdr)}
        rows.append([(to_us(r[idx["gpu__time_duration.sum"]], units[idx["gpu__time_duration.sum"]]),
                      num(r[idx["sm__cycles_elapsed.avg.per_second"]]),
                      num(r[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])) for r in data])
    lines = [f"# {tag}: offset forwarding (sparse hidden activations) vs the dense form, same batch, same box (ncu --set full --clock-control none)\n",
             "Commands: `ncu … -k regex:tdnn_gemm -s 33 -c 10 python tools/profile_step.py` and the same with `CVX_FORWARD_OFFSETS=0` "
             "(tools/profile_r2.sh + one extra capture).  The kernels, tiles and tensor-pipe utilisation are identical; what differs is the "
             "SM clock the 1 kW cap allows, because about half of the activation operand is exact zeros in the sparse form.", "",
             "| # | layer | K -> N | sparse: time us | SM clock GHz | tensor pipe % | dense: time us | SM clock GHz | tensor pipe % | time ratio |",
             "|---|---|---|---|---|---|---|---|---|---|"]
    for i, (L, a, b) in enumerate(zip(layers, rows[0], rows[1])):
        lines.append(f"| {i} | {L['name']} | {L['K']} -> {L['N']} | {a[0]:.1f} | {a[1]:.3f} | {a[2]:.1f} | {b[0]:.1f} | {b[1]:.3f} | {b[2]:.1f} | {a[0] / b[0]:.3f} |")
    ta, tb = sum(a[0] for a in rows[0]), sum(b[0] for b in rows[1])
    lines += ["", f"Sum over the 10 frame layers: {ta:.0f} us vs {tb:.0f} us = **{ta / tb:.3f}** (the first layers read raw features and do not change)."]
    open(os.path.join(P, f"{tag}_sparse_vs_dense.md"), "w").write("\n".join(lines) + "\n")


os.makedirs(P, exist_ok=True)
sparse_vs_dense()
launch_list()
s = gemm_report(tag, "the 10 frame layers of one 500 000-row batch of the headline net (v5 c-vector), incl. the pooled layer's kPool launch",
                "Command (tools/profile_r2.sh): `ncu --set full --clock-control none --import-source on -k regex:tdnn_gemm -s 33 -c 10 "
                "python tools/profile_step.py` (4th call; 11 matches per call = 10 frame layers + the segment GEMM)")
sx = gemm_report(tag + "x", "short segments: the 5 frame layers of one 500 000-row batch (2 500 x 200-frame segments) of the x-vector net",
                 "Command: `ncu --set full --clock-control none -k regex:tdnn_gemm -s 18 -c 5 python tools/profile_step.py --topology v2_xvector "
                 "--utts 2500 --frames 200 --chunk-size -1 --tag r2x`")
kernel_report(tag, "aux", "the other kernels of the same call: pool_meta, stage_input_tiled, stats_pool_finalize_fused, split_f32, utt_average",
              "Command: `ncu --set full --clock-control none -k 'regex:pool_meta|stage_input|stats_pool_finalize|split_f32|utt_average' -s 15 -c 5 "
              "python tools/profile_step.py` (4th call)")
kernel_report(tag, "pool", "stats_pool_partial_kernel<__half> (stand-alone HBM-bound pooling, CVX_FUSED_POOL=0)",
              "Command: `CVX_FUSED_POOL=0 ncu --set full --clock-control none -k regex:stats_pool_partial -s 3 -c 1 python tools/profile_step.py`")
if s:
    s["source"] = (f"ncu --set full capture of round {tag}: profiles/{tag}_gemm.md (tools/profile_r2.sh, tools/summarize_ncu.py); "
                   "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active per frame layer of one 500 000-row batch of the headline net")
    if sx:
        s["short_segments_xvector"] = {k: sx[k] for k in ("tensor_pipe_pct_flop_weighted", "tensor_pipe_pct_time_weighted", "dram_over_algorithmic")}
    json.dump(s, open(os.path.join(P, "tensor_pipe.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in s.items() if k != "per_layer"}))
sass_summary()
