"""CPU checks of bench.py's contract pieces that do not need a GPU: the reference arm prints one well-formed JSON line,
and the child mode used for the N-device BKZ tour fails loudly (no CPU fallback) with one JSON line."""
import json
import os
import subprocess
import sys

import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _last_json(out):
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out[-500:]
    return json.loads(lines[0])


@pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built")
def test_reference_arm_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-500:]
    j = _last_json(p.stdout)
    assert j["impl"] == "reference" and j["metric"] == "gso_update_row_GBps" and j["unit"] == "GB/s"
    assert j["value"] > 0 and j["higher_is_better"] is True and j["cpu_baseline"]["kind"] == "reference"
    assert j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0


def test_bkz_child_without_gpu_reports_the_error():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the child would run the real tour")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--bkz-child", "2"], capture_output=True,
                       text=True, timeout=120)
    assert p.returncode == 0
    j = _last_json(p.stdout)
    assert "error" in j and "no CUDA device" in j["error"]
