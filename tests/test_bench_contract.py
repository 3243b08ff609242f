"""CPU test of bench.py's reference arm (the one leg of bench.py that runs without a GPU): one JSON
line with the keys the driver reads, on a miniature of the default workload."""
import json
import os
import subprocess
import sys

import pytest

import refapi as R


@pytest.mark.skipif(not R.have_reference(), reason="needs oracle/_ref (the compiled reference)")
def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(R.ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "lap2d_48", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "numeric_qr_factorization_fp64_gflops"
    assert d["unit"] == "GFLOP/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"] == "lap2d_48"


def test_b200_arm_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    out = subprocess.run([sys.executable, os.path.join(R.ROOT, "bench.py"), "--workload", "lap2d_48", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)
