"""Write profiles/<tag>_summary.md from the artefacts of tools/ncu_round.sh (run here, no GPU)."""
import csv, json, subprocess, sys
from collections import defaultdict
from pathlib import Path

tag = sys.argv[1]
G = Path("gpurun_out"); P = Path("profiles"); P.mkdir(exist_ok=True)
METRICS = ["gpu__time_duration.sum", "gcc__cache_requests_type_instruction.sum", "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[h.index("Kernel Name")]}
        for m in METRICS:
            if m in h:
                d[m] = (r[h.index(m)], units[h.index(m)])
        res.append(d)
    return res


def to_bytes(v, unit):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(v) * f


lines = [f"# ncu evidence, {tag}", "", f"Command: `{sys.argv[2] if len(sys.argv) > 2 else 'python bench.py --gpus 1 --steps 20 --warmup 5'}` (4096 envs x 128 agents, 1 x B200),",
         "`--clock-control none`.  Per-launch times under ncu are cold-cache and serialised: compare shares.", ""]
# launch list
rows = [r for r in csv.reader(open(G / f"launches_{tag}.csv")) if len(r) > 5]
hi = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
h = rows[hi]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
d = defaultdict(list)
for r in rows[hi + 1:]:
    d[r[ki]].append(float(r[vi].replace(",", "")) / 1e3)
tot = sum(sum(v) for k, v in d.items() if k.startswith("nmmo"))
lines += ["## Launch list (gpu__time_duration.sum, first 260 launches: spin-up, warm-up, the timed window, the dense-writer leg)", "", "| kernel | launches | mean us | share of nmmo time |", "|---|---|---|---|"]
for k, v in d.items():
    lines.append(f"| {k[:48]} | {len(v)} | {sum(v)/len(v):.1f} | {100*sum(v)/tot if k.startswith('nmmo') else 0:.1f} % |")
summary = {}
for name, title in ((f"prof_{tag}_alive", "Full capture at tick 12 of the driver's window (ticks 5-25: every agent alive, incremental observation writer)"),
                    (f"prof_{tag}_tick40", "Full capture at tick ~40 (incremental observation writer)"),
                    (f"prof_{tag}_c5", "Full capture, config 5 (1024 agents per env), tick 12"),
                    (f"prof_{tag}_dense", "Full capture, dense observation writer (obs_full=1)")):
    rep = G / f"{name}.ncu-rep"
    if not rep.exists():
        continue
    lines += ["", f"## {title}", "", "| kernel | " + " | ".join(m.split(".")[0].replace("__", " ") for m in METRICS) + " |", "|---|" + "---|" * len(METRICS)]
    for k in raw(rep):
        lines.append(f"| {k['kernel']} | " + " | ".join(f"{k[m][0]} {k[m][1]}" if m in k else "-" for m in METRICS) + " |")
        rd = to_bytes(*k["dram__bytes_read.sum"]); wr = to_bytes(*k["dram__bytes_write.sum"])
        dur = float(k["gpu__time_duration.sum"][0]) * {"ms": 1e-3, "us": 1e-6, "s": 1}[k["gpu__time_duration.sum"][1]]
        summary[f"{name.split('_')[-1]}:{k['kernel']}"] = {"dram_bytes": rd + wr, "seconds": dur, "dram_gbs": (rd + wr) / dur / 1e9}
        lines.append(f"|  | DRAM traffic {(rd+wr)/1e6:.1f} MB per launch = {(rd+wr)/dur/1e9:.0f} GB/s | | | | | | | | | | | | |")
(P / f"{tag}_summary.md").write_text("\n".join(lines) + "\n")
(P / f"{tag}_traffic.json").write_text(json.dumps(summary, indent=1))
print("\n".join(lines))
