"""Warp-instruction and stall-sample shares of source-line ranges of one kernel of an .ncu-rep (development aid).
usage: ncu_regions.py rep kernel file.cu name:lo-hi [name:lo-hi ...]   (lines of other files are listed as 'inlined:<file>')"""
import csv, subprocess, sys, collections
rep, kernel, fname = sys.argv[1:4]
regions = []
for a in sys.argv[4:]:
    n, r = a.split(":"); lo, hi = r.split("-"); regions.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kernel],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; S = collections.Counter(); I = collections.Counter()
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 8 and r[0] not in ('', 'Line No'):
        try: ln, smp, ins = int(r[0]), int(r[6]), int(r[7])
        except Exception: continue
        key = f"inlined:{cur}"
        if cur == fname:
            key = "other"
            for n, lo, hi in regions:
                if lo <= ln <= hi: key = n; break
        S[key] += smp; I[key] += ins
ts, ti = sum(S.values()), sum(I.values())
print(f"total samples {ts}  total warp-inst {ti}")
for k, v in sorted(I.items(), key=lambda x: -x[1]):
    print(f"{k:28s} inst {100*v/ti:5.1f}%  samples {100*S[k]/ts:5.1f}%")
