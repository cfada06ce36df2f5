"""Build the CUDA library in-tree: nmmo_b200/_build/libnmmo_b200.so (sm_100a only)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
SRC = ROOT / "csrc"
OUT = ROOT / "_build" / "libnmmo_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v"]


def sources():
    return sorted(SRC.glob("*.cu")) + sorted(SRC.glob("*.cuh")) + sorted((ROOT.parent / "include").glob("*.h"))


def needs_build() -> bool:
    if not OUT.exists():
        return True
    t = OUT.stat().st_mtime
    return any(s.stat().st_mtime > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return OUT
    OUT.parent.mkdir(exist_ok=True)
    cmd = [NVCC, *FLAGS, "-o", str(OUT), str(SRC / "nmmo_api.cu")]
    env = dict(os.environ)
    env.pop("CC", None); env.pop("CXX", None)
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    log = r.stdout + r.stderr
    (OUT.parent / "build.log").write_text(" ".join(cmd) + "\n" + log)
    if verbose or r.returncode:
        print(log, file=sys.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed; see nmmo_b200/_build/build.log")
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(OUT)
