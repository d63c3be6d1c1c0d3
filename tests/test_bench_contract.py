"""The driver's bench contract, checked on CPU through the reference arm (`bench.py --impl reference` times the reference's own CPU
path — the oracle port — and must print ONE JSON line with the contract's keys), and statically for the GPU arm's keys."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    # OMP_NUM_THREADS=1 is what torchrun exports to its workers: the arm must still use every core it may run on
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env={**os.environ, "OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "ViT-B/16 train images/sec" and d["unit"] == "images/sec"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_gpu_arm_emits_every_contract_key():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "clocks", "e2e", "h2d_bytes_per_step", "d2h_bytes_per_step", "gpu_launches", "roofline", "bound",
                "achieved", "peak", "frac", "traffic", "cpu_baseline", "cores", "kind", "sample"):
        assert re.search(r'["\']%s["\']' % re.escape(key), src), key
    assert "/root/reference" not in src          # nothing at run time may read the reference tree (absent on the GPU box)
