import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
MODELS = GOLDEN / "models"
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def kats():
    return json.loads((GOLDEN / "kats.json").read_text())


@pytest.fixture(scope="session")
def samples():
    return dict(np.load(GOLDEN / "samples.npz"))


def f32(x):
    return np.asarray(x, dtype=np.float32)


def splitmix_bytes(seed, n, offset=0):
    """BASELINE.md synthetic inputs: byte k = low 8 bits of splitmix64 mix(seed + k), as int8."""
    k = (np.arange(offset, offset + n, dtype=np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        z = k
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z & np.uint64(0xFF)).astype(np.uint8).view(np.int8)
