"""BASELINE.json configs[2] / SURVEY.md 8(d) config 3 as a driver-run test: the env-sharded recorded-RNG check
(tests/check_sharded_parity.py) on the one GPU the driver's test box has -- same code path as the 8-GPU run whose result
is kept under profiles/ (full-size config, envs sharded by global index, sampled envs against the oracle every tick with
draws injected into both sides, whole episodes under the forager policy)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
@pytest.mark.parametrize("policy,ticks", [("forage", 300), ("random", 120)])
def test_config3_check_on_one_gpu(policy, ticks):
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "check_sharded_parity.py"), "--envs", "96", "--check", "4",
                        "--ticks", str(ticks), "--policy", policy], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["mismatches"] == 0 and line["envs_checked"] == 4 and line["ticks"] == ticks
    if policy == "forage":
        assert line["mean_alive_fraction_checked"] > 0.3
