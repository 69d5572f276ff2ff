"""Summaries for profiles/: (1) launch list -> per-kernel share / DRAM bytes + traffic.json, (2) `ncu --set full` raw page -> a metric table.
usage: python scripts/ncu_summary.py launches <launches.csv> <out.md> <traffic.json>
       python scripts/ncu_summary.py full <raw.csv> <out.md>"""
import collections
import csv
import json
import sys


def launches(path, out_md, traffic_json):
    rows = list(csv.reader(open(path)))
    hdr = None
    per = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            unit = d.get("Metric Unit", "")
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e6, "us": 1e3, "ns": 1.0, "s": 1e9, "msecond": 1e6, "usecond": 1e3, "nsecond": 1.0,
                    "second": 1e9}.get(unit, 1.0)
            per[d["Kernel Name"].split("(")[0]][d["Metric Name"]].append(v * mult)
    tot = sum(sum(m["gpu__time_duration.sum"]) for m in per.values())
    lines = ["| kernel | launches | total ms | share | avg us | DRAM read MB / launch | DRAM written MB / launch |", "|---|---|---|---|---|---|---|"]
    traffic = {}
    for k, m in sorted(per.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
        t = m["gpu__time_duration.sum"]
        rd, wr = m.get("dram__bytes_read.sum", [0]), m.get("dram__bytes_write.sum", [0])
        lines.append(f"| `{k}` | {len(t)} | {sum(t) / 1e6:.2f} | {100 * sum(t) / tot:.1f} % | {sum(t) / len(t) / 1e3:.1f} | {sum(rd) / len(rd) / 1e6:.1f} | {sum(wr) / len(wr) / 1e6:.1f} |")
        traffic[k] = (sum(rd) / len(rd) + sum(wr) / len(wr))
    open(out_md, "w").write("\n".join(lines) + "\n")
    # one colour phase of the 8192^2 world = 3 part-launches of pass 1, pass 2 and pass 3 — of the plain instantiation <PASS, 0> (what the
    # skip gate picks on the mixed world) or of the row-skipping one <PASS, 1> with its classification (the gate's probe ticks)
    def pick(tag):
        return {k: v for k, v in traffic.items() if f", {tag}>" in k or "tick_pass3" in k or (tag == 1 and "classify" in k)}
    plain, skip = pick(0), pick(1)
    phase = 3 * sum(plain.values())
    json.dump({"tick_phase_bytes_per_launch": phase, "per_kernel_bytes_per_launch": {k: 3 * v for k, v in plain.items()},
               "row_skipping_phase_bytes_per_launch": 3 * sum(skip.values()),
               "source": f"{path} (ncu dram__bytes_read.sum + dram__bytes_write.sum of `bench.py --steps 2 --warmup 1 --no-cpu-baseline`, 8192x8192 mixed; a colour phase = 961 chunks is launched as 3 parts, so one phase = 3 launches of each kernel; <PASS, 0> = the instantiation the skip gate uses on this world)",
               "algorithmic_bytes_per_launch": 36 * 7936 * 7936 // 4}, open(traffic_json, "w"))
    print("\n".join(lines))
    print("phase MB", phase / 1e6)


def full(path, out_md):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = [("duration (us)", "gpu__time_duration.sum"), ("registers / thread", "launch__registers_per_thread"), ("warps active (% of peak)", "sm__warps_active.avg.pct_of_peak_sustained_active"),
            ("issue slots used per scheduler (IPC)", "smsp__issue_active.avg.per_cycle_active"), ("active threads per instruction", "smsp__thread_inst_executed_per_inst_executed.ratio"),
            ("warp instructions executed", "smsp__inst_executed.sum"), ("stall / issue: barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
            ("stall / issue: fixed-latency dependency (wait)", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
            ("stall / issue: instruction fetch (no_instruction)", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
            ("stall / issue: shared memory (short_scoreboard)", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
            ("stall / issue: global memory (long_scoreboard)", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
            ("stall / issue: branch resolving", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
            ("stall / issue: not selected", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
            ("instruction-cache hit rate (%)", "sm__icc_request_hit_rate.pct"), ("DRAM read", "dram__bytes_read.sum"), ("DRAM written", "dram__bytes_write.sum"),
            ("DRAM throughput (% of peak)", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("SM throughput (% of peak)", "sm__throughput.avg.pct_of_peak_sustained_elapsed")]
    kernels = rows[2:]
    lines = ["| metric | " + " | ".join("`" + r[idx["Kernel Name"]].split("(")[0] + "`" for r in kernels) + " |", "|---|" + "---|" * len(kernels)]
    for label, key in want:
        if key not in idx:
            continue
        u = units[idx[key]]
        lines.append(f"| {label} | " + " | ".join(f"{r[idx[key]]} {u}".strip() for r in kernels) + " |")
    open(out_md, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3])
