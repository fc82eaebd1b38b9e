"""bench.py's reference arm runs on the CPU (the unmodified reference package from baseline/_ref; the oracle port of its
forward only where that copy is missing), so its JSON contract can be checked here: one line on rank 0 with the base keys, `impl`, `cpu_baseline` and a zero-copy `e2e`; other ranks print
nothing and exit 0."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--small", "--steps", "2",
                           "--warmup", "1", "--gpus", "2"], capture_output=True, text=True, timeout=600, env=env)


def test_reference_arm_prints_one_contract_line_on_rank_zero():
    r = _run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d)
    assert d["impl"] == "reference" and d["metric"] == "frames_per_sec" and d["unit"] == "frames/s"
    assert d["n_gpus"] == 2 and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["value"] > 0 and abs(d["value"] - 1e3 * d["config"]["frames_per_step"] / d["ms_per_step"]) < 1e-6 * d["value"]
    assert d["vs_baseline"] is None and d["config"]["valid"] is False          # --small is a debug size, flagged as such
    cb = d["cpu_baseline"]
    installed = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "dprt", "models", "__init__.py"))
    assert cb["kind"] == ("reference" if installed else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
