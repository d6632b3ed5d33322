"""Summarise ncu outputs brought back in gpurun_out/ into profiles/ (tracked).

    python tools/ncu_summary.py <tag> [--launches gpurun_out/launches.csv] [--rep gpurun_out/prof.ncu-rep]
"""
import argparse
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    # the shared-memory data pipe: LSU wavefronts (LDS / STS) and UMMA operand reads; their sum is the utilisation of
    # the pipe that bounds the two tcgen05 kernels
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def short(name):
    return name.split("(")[0].replace("void ", "").replace("gs::", "").replace("(anonymous namespace)::", "")


def launches_table(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        d = agg.setdefault(short(r[ki]), [0, 0.0])
        d[0] += 1
        d[1] += v
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / tot:.3f} |")
    out.append(f"\n{len(rows) - 1} launches captured, {tot / 1e3:.2f} ms of kernel time (ncu-serialised, cold cache: compare shares).")
    return "\n".join(out)


def rep_table(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    ki = hdr.index("Kernel Name")
    out = ["| # | kernel | " + " | ".join(f"{m} [{units[i]}]" for m, i in cols) + " |",
           "|---|---|" + "---:|" * len(cols)]
    for n, r in enumerate(rows[2:]):
        out.append(f"| {n} | `{short(r[ki])}` | " + " | ".join(r[i] for _, i in cols) + " |")
    return "\n".join(out)


# ncu kernel name -> the name bench.py's profiler uses for it
BENCH_NAMES = {"gcn_fused_kernel": "bf16_gemm_gcn", "tcn_fused_kernel": "bf16_tconv", "head_stream_kernel": "head",
               "front_mma_kernel": "bf16_front", "dtw_ws_kernel": "dtw_wavefront", "dtw_pipeline2_kernel": "dtw_wavefront", "stj_tc_kernel": "stj_gate",
               "se_kernel": "se_gate", "dtw_backtrack_kernel": "dtw_backtrack"}


def traffic_json(reps, out_path, source):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of every kernel in the captures -> the JSON bench.py
    reads for `roofline.traffic` (regenerated from the capture of THIS code state, never edited by hand)."""
    import json
    agg = collections.OrderedDict()
    pipes = {}
    for path in reps:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        pipe = [hdr.index(m) if m in hdr else -1 for m in (
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            name = short(r[ki]).split("<")[0].split("::")[-1]
            b = float(r[ri].replace(",", "")) * scale[units[ri]] + float(r[wi].replace(",", "")) * scale[units[wi]]
            agg.setdefault(BENCH_NAMES.get(name, name), []).append(b)
            pipes.setdefault(BENCH_NAMES.get(name, name), []).append(
                [float(r[i].replace(",", "")) if i >= 0 and r[i] else 0.0 for i in pipe])
    out = {"source": source}
    total = 0.0
    for k, v in agg.items():
        out[k] = {"bytes_per_launch": sum(v) / len(v), "launches": len(v),
                  "per_launch_GB": [round(x / 1e9, 3) for x in v],
                  # shared-memory data pipe of each launch: LDS/STS wavefronts + UMMA operand wavefronts, % of peak;
                  # their sum near 100 = the kernel is bound by that pipe
                  "smem_pipe_lsu_pct": [round(p[0], 1) for p in pipes[k]],
                  "smem_pipe_tc_pct": [round(p[1], 1) for p in pipes[k]],
                  "tensor_pipe_pct": [round(p[2], 1) for p in pipes[k]]}
        total += sum(v)
    out["total_bytes_all_captured_launches"] = total
    json.dump(out, open(out_path, "w"), indent=1)
    print(out_path, f"{total / 1e9:.2f} GB over the captured launches")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--traffic", nargs="*", help="ncu-rep files to fold into profiles/<tag>_traffic.json")
    ap.add_argument("--launches")
    ap.add_argument("--rep")
    ap.add_argument("--cmd", default="")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    out = [f"# ncu summary — {a.tag}", ""]
    if a.cmd:
        out += [f"Command profiled: `{a.cmd}`", ""]
    if a.note:
        out += [a.note, ""]
    if a.launches:
        out += ["## Launch list (`--metrics gpu__time_duration.sum --clock-control none`)", "",
                launches_table(a.launches), ""]
    if a.rep:
        out += ["## Full capture (`--set full --clock-control none --import-source on`)", "", rep_table(a.rep), ""]
    if a.traffic:
        traffic_json(a.traffic, os.path.join(ROOT, "profiles", f"{a.tag}_traffic.json"),
                     f"ncu --set full --clock-control none captures {[os.path.basename(t) for t in a.traffic]} ({a.cmd})")
        return
    path = os.path.join(ROOT, "profiles", f"{a.tag}.md")
    open(path, "w").write("\n".join(out))
    print(path)


if __name__ == "__main__":
    main()
