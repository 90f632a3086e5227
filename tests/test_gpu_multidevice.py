"""mf_options.devices (ABI 3): predict_many sharded over the GPUs of the box INSIDE the C-ABI call (north_star).
On a one-GPU box the routing object, the per-replica host threads, the contiguous sharding and the one-time weight broadcast
are exercised with two replicas on the same GPU (MF_ALLOW_DUPLICATE_DEVICES, a test-only switch, in a subprocess so the
environment variable does not leak); with >= 2 GPUs the same checks run over every visible device."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import microflow_rs_b200 as mf
import oracle
from conftest import MODELS, splitmix_bytes

pytestmark = pytest.mark.gpu
HERE = Path(__file__).resolve().parent

CHILD = r"""
import json, sys
import numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import microflow_rs_b200 as mf
import oracle
from conftest import MODELS, splitmix_bytes
devices = json.loads(sys.argv[1])
res = {"ok": False}
o = oracle.Model(MODELS / "person_detect.tflite", fast=True)
g = mf.Model(MODELS / "person_detect.tflite", devices=devices)
one = mf.Model(MODELS / "person_detect.tflite", device=devices[0])
res["devices"] = g.devices
res["bcast"] = g.weight_broadcast
n = 1000 + len(devices) + 1                                  # ragged: the last shard is shorter
xs = splitmix_bytes(0x5EED0004, n * o.in_elems).reshape(n, -1)
q1, l1 = one.predict_many_logits(xs)
qg, lg = g.predict_many_logits(xs)
res["rows_equal_1gpu"] = bool(np.array_equal(q1, qg) and np.array_equal(l1, lg))
want_f, want_q = o.predict_many_quantized(xs[-40:], threads=oracle.max_threads())     # rows of the LAST shard vs the oracle
res["last_shard_equals_oracle"] = bool(np.array_equal(qg[-40:], want_q) and np.array_equal(g.predict_many_quantized(xs)[-40:], want_f))
res["one_sample"] = bool(np.array_equal(g.predict_quantized(xs[3]), one.predict_quantized(xs[3])))
res["tiny_batches"] = all(bool(np.array_equal(g.predict_many_quantized(xs[:k]), one.predict_many_quantized(xs[:k]))) for k in (1, 2, 3, 65))
pin = mf.PinnedBuffer((n, o.in_elems), np.int8); pin.array[:] = xs
outs = [mf.PinnedBuffer((n, o.out_elems), np.float32) for _ in range(2)]
for k in range(2):
    g.predict_many_quantized_async(pin.array, outs[k].array)
g.synchronize()
ref = one.predict_many_quantized(xs)
res["async"] = bool(np.array_equal(outs[0].array, ref) and np.array_equal(outs[1].array, ref))
# large host-resident calls: contiguous 2048-sample chunks claimed dynamically by the devices' host threads (same rows, same order)
nb = len(devices) * 2 * 2048 + 37
big = np.concatenate([xs] * (nb // n + 1))[:nb]
res["dynamic_chunks"] = bool(np.array_equal(g.predict_many_quantized(big), one.predict_many_quantized(big)))
qb, lb = g.predict_many_logits(big)
q1b, l1b = one.predict_many_logits(big)
res["dynamic_chunks_logits"] = bool(np.array_equal(qb, q1b) and np.array_equal(lb, l1b))
res["launches"] = g.launch_count()
res["ok"] = all(res[k] for k in ("rows_equal_1gpu", "last_shard_equals_oracle", "one_sample", "tiny_batches", "async", "dynamic_chunks", "dynamic_chunks_logits"))
g.close(); one.close()
print(json.dumps(res))
"""


def _child(devices, env_extra):
    code = CHILD % {"root": str(HERE.parent), "tests": str(HERE)}
    env = dict(os.environ, **env_extra)
    p = subprocess.run([sys.executable, "-c", code, json.dumps(devices)], capture_output=True, text=True, timeout=600, env=env)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert lines, f"no result (rc={p.returncode})\nstdout:\n{p.stdout[-2000:]}\nstderr:\n{p.stderr[-3000:]}"
    return json.loads(lines[-1])


def test_single_entry_device_list_is_a_plain_model():
    m = mf.Model(MODELS / "sine.tflite", devices=[0])
    try:
        assert m.devices == [0] and m.weight_broadcast == "none"
        np.testing.assert_array_equal(m.predict(np.full((1, 1), 0.5, np.float32)).reshape(-1), np.float32([0.41348344]))   # tests/sine.rs:9-11
    finally:
        m.close()


def test_two_replicas_shard_predict_many_inside_the_call():
    res = _child([0, 0], {"MF_ALLOW_DUPLICATE_DEVICES": "1"})
    assert res["ok"], res
    assert res["devices"] == [0, 0] and res["bcast"] == "memcpy_peer"


def test_all_gpus_of_the_box():
    n = mf.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    res = _child(list(range(n)), {})
    assert res["ok"], res
    assert res["devices"] == list(range(n)) and res["bcast"] in ("nccl", "memcpy_peer")


def test_bad_device_lists_are_refused():
    with pytest.raises(mf.MicroflowError) as e:
        mf.Model(MODELS / "sine.tflite", devices=[0, 0])
    assert e.value.status == 9
    with pytest.raises(mf.MicroflowError) as e:
        mf.Model(MODELS / "sine.tflite", devices=[0, 99])
    assert e.value.status == 9
