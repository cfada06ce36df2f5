"""Top source lines by warp-stall samples for one kernel of an .ncu-rep (development aid)."""
import csv, subprocess, sys
rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kernel],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; data = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 8 and r[0] not in ('', 'Line No'):
        try: data.append((int(r[6]), int(r[7]), cur, int(r[0]), r[1].strip()[:105]))
        except Exception: pass
tot = sum(d[0] for d in data); toti = sum(d[1] for d in data)
print('total samples', tot, 'total warp-inst', toti)
for d in sorted(data, reverse=True)[:top]:
    print(f"{d[0]:7d} {100*d[0]/tot:5.1f}% ins={100*d[1]/max(1,toti):5.1f}% {d[2]}:{d[3]} {d[4]}")
