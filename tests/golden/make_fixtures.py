"""Regenerates the data fixtures under tests/golden/ from the reference checkout.

Run HERE (the build container, where /root/reference exists); the GPU box only ever sees the committed
outputs.  Nothing in tests/, bench.py or smoke() reads /root/reference at run time.

Outputs (all DATA, no reference source code):
  models/{sine,speech,person_detect}.tflite   the three workloads (reference models/*.tflite, byte-for-byte)
  samples.npz                                 int8 inputs YES, NO (1x1960) and PERSON, NO_PERSON (96x96x1),
                                              parsed from reference samples/features/{speech,person_detect}.rs
  sine_microflow.csv                          500 (x, y_microflow) rows, reference analysis/accuracy/data/sine-microflow.csv
The hand-transcribed known-answer tests (kats.json) are NOT generated here; they cite reference file:line.
"""
import re
import shutil
import sys
from pathlib import Path

import numpy as np

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent


def parse_rs_consts(path):
    """`pub const NAME: BufferND<i8, ...> = [matrix![ [a],[b]; ... ]];` -> {NAME: flat int list (row-major)}."""
    text = path.read_text()
    out = {}
    for m in re.finditer(r"pub const (\w+):\s*Buffer\dD<i8,([^>]*)>\s*=\s*(.*?);\s*(?=pub const|\Z)", text, re.S):
        name, dims, body = m.group(1), m.group(2), m.group(3)
        dims = [int(d) for d in re.findall(r"\d+", dims)]
        vals = [int(v) for v in re.findall(r"-?\d+", body)]
        n = int(np.prod(dims))
        assert len(vals) == n, (name, len(vals), n)
        out[name] = np.array(vals, dtype=np.int8)
    return out


def main():
    (OUT / "models").mkdir(exist_ok=True)
    for name in ("sine", "speech", "person_detect"):
        shutil.copyfile(REF / "models" / f"{name}.tflite", OUT / "models" / f"{name}.tflite")
    shutil.copyfile(REF / "analysis/accuracy/data/sine-microflow.csv", OUT / "sine_microflow.csv")
    sp = parse_rs_consts(REF / "samples/features/speech.rs")
    pd_ = parse_rs_consts(REF / "samples/features/person_detect.rs")
    np.savez_compressed(
        OUT / "samples.npz",
        YES=sp["YES"].reshape(1, 1960), NO=sp["NO"].reshape(1, 1960),
        PERSON=pd_["PERSON"].reshape(1, 96, 96, 1), NO_PERSON=pd_["NO_PERSON"].reshape(1, 96, 96, 1),
    )
    print("fixtures written to", OUT)


if __name__ == "__main__":
    main()
