"""bench.py contract on a GPU-less box: the reference arm must print ONE JSON line with the keys the
driver reads.  Without a CUDA device `--impl reference` cannot run the reference CUDA build and falls back
to the CPU oracle port (the reference has no CPU path of its own); `--impl reference-cpu` asks for it."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("impl", ["reference", "reference-cpu"])
def test_reference_arm_json_line(impl):
    if impl == "reference" and torch.cuda.is_available():
        pytest.skip("with a GPU this arm runs the reference CUDA build (covered by the GPU bench runs)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", impl, "--steps", "1", "--warmup", "1", "--gpus", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gvoxel/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]
