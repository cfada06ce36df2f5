"""SASS instruction count per kernel of the built library (code size is what the step kernel's time follows)."""
import re, subprocess, sys, collections
lib = sys.argv[1] if len(sys.argv) > 1 else "nmmo_b200/_build/libnmmo_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cnt, cur = collections.Counter(), None
for l in out.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m: cur = m.group(1); continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+\S", l): cnt[cur] += 1
for k, v in sorted(cnt.items(), key=lambda x: -x[1])[:6]: print(v, k)
