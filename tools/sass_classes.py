"""SASS instruction-class counts per kernel of the built library (cuobjdump -sass): the evidence that the tables move
with 1-D TMA bulk copies (UBLKCP) completed on mbarriers (SYNCS), and how big each kernel is.  Writes markdown."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "nmmo_b200/_build/libnmmo_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
CLASSES = [("UBLKCP", r"UBLKCP"), ("UTMALDG/UTMASTG", r"UTMA(LDG|STG)"), ("SYNCS (mbarrier)", r"SYNCS"), ("BAR", r"\bBAR\b"),
           ("ATOMS (smem atomics)", r"ATOMS"), ("ATOMG/RED (global atomics)", r"\b(ATOMG|RED)\b"), ("LDS", r"\bLDS"), ("STS", r"\bSTS"),
           ("LDG", r"\bLDG"), ("STG", r"\bSTG"), ("LDL/STL (local)", r"\b(LDL|STL)"), ("SHFL/VOTE/MATCH", r"\b(SHFL|VOTE|MATCH|REDUX)"),
           ("CALL", r"\bCALL\b"), ("BRA", r"\bBRA\b")]
cnt = collections.defaultdict(collections.Counter)
arch = set(re.findall(r"arch = (sm_\w+)", out))
cur = None
for l in out.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", l)
    if cur and m:
        ins = m.group(1)
        cnt[cur]["total"] += 1
        for name, rx in CLASSES:
            if re.search(rx, ins):
                cnt[cur][name] += 1
print(f"# SASS instruction classes per kernel ({lib}, arch {', '.join(sorted(arch))})\n")
print("| kernel | total | " + " | ".join(n for n, _ in CLASSES) + " |")
print("|---|---|" + "---|" * len(CLASSES))
for k, c in sorted(cnt.items(), key=lambda x: -x[1]["total"]):
    if not k.startswith("nmmo") and "k_" not in k:
        continue
    print(f"| {k[:60]} | {c['total']} | " + " | ".join(str(c[n]) for n, _ in CLASSES) + " |")
