"""Turns the ncu artefacts a gpurun call left under gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py <tag> --launches gpurun_out/x_launches.csv --rep name=gpurun_out/x.ncu-rep ...

Writes profiles/<tag>_launches.md (per-kernel totals / shares of the step from the gpu__time_duration launch list),
profiles/<tag>_<name>.csv (selected raw metrics of an `ncu --set full` capture) and updates
profiles/roofline_traffic.json (dram bytes read+written per launch, which bench.py copies into `roofline.traffic`)."""
import argparse
import collections
import csv
import json
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PROF = ROOT / "profiles"

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]


def launches(tag, path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    mi = hdr.index("Metric Name")
    t, n = collections.defaultdict(float), collections.Counter()
    for r in rows[h + 1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        name = r[ki].split("(")[0].replace("void ", "")
        t[name] += v
        n[name] += 1
    tot = sum(t.values())
    out = [f"# {tag}: ncu launch list ({Path(path).name}; gpu__time_duration.sum, --clock-control none)", "",
           "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
           "| kernel | launches | total ms | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(t.items(), key=lambda x: -x[1]):
        out.append(f"| `{k}` | {n[k]} | {v / 1e3:.3f} | {100 * v / tot:.1f}% | {v / n[k]:.2f} |")
    out.append(f"| total | {sum(n.values())} | {tot / 1e3:.3f} | 100% | |")
    (PROF / f"{tag}_launches.md").write_text("\n".join(out) + "\n")
    print("\n".join(out))


def report(tag, name, path):
    res = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(res.stdout.splitlines()))
    hdr, units = rows[0], rows[1]
    traffic = {}
    with open(PROF / f"{tag}_{name}.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch", "kernel", "metric", "unit", "value"])
        for li, vals in enumerate(rows[2:]):
            d = dict(zip(hdr, vals))
            kern = d.get("Kernel Name", "").split("(")[0].replace("void ", "")
            for k in KEEP:
                if k in d and d[k] != "":
                    w.writerow([li, kern, k, units[hdr.index(k)], d[k]])
            try:
                mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                rd = float(d["dram__bytes_read.sum"]) * mult[units[hdr.index("dram__bytes_read.sum")]]
                wr = float(d["dram__bytes_write.sum"]) * mult[units[hdr.index("dram__bytes_write.sum")]]
                traffic.setdefault(kern.split("<")[0], []).append(rd + wr)
            except (KeyError, ValueError):
                pass
    tf = PROF / "roofline_traffic.json"
    cur = json.loads(tf.read_text()) if tf.exists() else {}
    for k, v in traffic.items():
        cur[k] = sum(v) / len(v)
    cur.setdefault("_source", {})[name] = (f"{tag}: ncu --set full, {Path(path).name}, "
                                           "dram__bytes_read.sum + dram__bytes_write.sum per launch")
    tf.write_text(json.dumps(cur, indent=1) + "\n")
    print(name, {k: sum(v) / len(v) for k, v in traffic.items()})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches")
    ap.add_argument("--rep", action="append", default=[])
    a = ap.parse_args()
    PROF.mkdir(exist_ok=True)
    if a.launches:
        launches(a.tag, a.launches)
    for r in a.rep:
        nm, p = r.split("=", 1)
        report(a.tag, nm, p)
