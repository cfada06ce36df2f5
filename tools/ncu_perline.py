"""Per-source-line warp instructions (per agent record) and stall samples of one kernel of an .ncu-rep (development aid).
usage: ncu_perline.py rep kernel n_agents [min_inst_per_agent]"""
import csv, subprocess, sys
rep, kernel, n_agents = sys.argv[1], sys.argv[2], float(sys.argv[3])
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 2.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; data = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 8 and r[0] not in ('', 'Line No'):
        try: data.append((cur, int(r[0]), int(r[7]), int(r[6]), r[1].strip()[:100]))
        except Exception: pass
tot = sum(d[2] for d in data); ts = sum(d[3] for d in data)
print("warp-inst per agent", tot / n_agents, "samples", ts)
for d in sorted(data):
    if d[2] / n_agents >= thr or d[3] > ts * 0.008:
        print(f"{d[0][:14]:14s} {d[1]:4d} {d[2]/n_agents:7.1f} smp={100*d[3]/ts:5.1f}%  {d[4]}")
