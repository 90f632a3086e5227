"""tcgen05 Conv2D kernel vs oracle / generic kernel, bit-exact.  Each group runs in its own subprocess with a timeout."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
HERE = Path(__file__).resolve().parent


def _run(which):
    p = subprocess.run([sys.executable, str(HERE / "tc_check.py"), which], capture_output=True, text=True, timeout=300)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert lines, f"no result (rc={p.returncode})\nstdout:\n{p.stdout[-2000:]}\nstderr:\n{p.stderr[-3000:]}"
    res = json.loads(lines[-1])
    assert res["ok"], res.get("error")


def test_tc_pointwise_packed_gemm():
    _run("pointwise")


def test_tc_conv3x3_implicit_gemm():
    _run("conv3x3")
