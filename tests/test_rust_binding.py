"""The Rust side of the boundary is source only (no rustc in the image).  What can be verified without a compiler: the -sys crate
binds exactly the functions the C header declares (and the library exports), its mf_options mirrors the C struct, the safe wrapper
covers every input/output rank and element type the macro accepts, and -- where the reference checkout is present (this
container, not the GPU box) -- the macro patch applies cleanly to it."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

import microflow_rs_b200 as mf
from conftest import ROOT

RUST = ROOT / "rust"


def _header_functions():
    header = (ROOT / "include" / "microflow_cuda.h").read_text()
    return sorted(set(re.findall(r"\b(mf_[a-z0-9_]+)\s*\(", header)))


def test_sys_crate_declares_every_exported_function():
    src = (RUST / "microflow-cuda-sys" / "src" / "lib.rs").read_text()
    block = src[src.index('extern "C" {'):]
    declared = sorted(set(re.findall(r"pub fn (mf_[a-z0-9_]+)\s*\(", block)))
    assert declared == _header_functions() == sorted(mf.ABI_SYMBOLS)


def test_sys_options_struct_matches_the_c_struct():
    src = (RUST / "microflow-cuda-sys" / "src" / "lib.rs").read_text()
    body = re.search(r"pub struct mf_options \{(.*?)\n\}", src, re.S).group(1)
    fields = re.findall(r"pub (\w+): ([^,]+),", body)
    py = [(n, t) for n, t in mf._Options._fields_]
    assert [f[0] for f in fields] == [n for n, _ in py]
    sizes = {"u32": 4, "i32": 4, "[i32; MF_MAX_DEVICES]": 4 * mf.MAX_DEVICES}
    import ctypes as C
    assert sum(sizes[t.strip()] for _, t in fields) == C.sizeof(mf._Options)
    assert f"pub const MF_ABI_VERSION: c_int = {mf.lib().mf_abi_version()};" in src


def test_wrapper_covers_every_buffer_rank_and_element_type():
    src = (RUST / "microflow-cuda" / "src" / "lib.rs").read_text()
    assert "HostBuffer for SMatrix<T, R, C>" in src                                     # Buffer2D (sine, speech)
    assert "HostBuffer for [SMatrix<[T; CH], R, C>; B]" in src                          # Buffer4D (person_detect)
    for t in ("i8", "u8", "f32"):
        assert f"impl Element for {t}" in src
    for fn in ("pub fn predict<", "pub fn predict_quantized<", "pub fn predict_many_quantized<", "pub fn predict_many<", "pub fn cached("):
        assert fn in src, fn
    assert "MF_LAYOUT_NALGEBRA" in src and "n_devices" in src


def test_macro_patch_names_only_functions_the_wrapper_defines():
    patch = (RUST / "microflow-macros-cuda.patch").read_text()
    src = (RUST / "microflow-cuda" / "src" / "lib.rs").read_text()
    used = set(re.findall(r"microflow::cuda::(\w+)", patch))
    assert used >= {"Handle", "cached", "predict", "predict_quantized", "predict_many_quantized", "predict_many"}
    for name in used:
        assert re.search(rf"pub (fn|struct) {name}\b", src), name


def test_macro_patch_applies_to_the_reference_checkout(tmp_path):
    ref = Path("/root/reference")
    if not (ref / "microflow-macros" / "src" / "lib.rs").exists() or not shutil.which("git"):
        pytest.skip("reference checkout not present (GPU box)")
    for f in ("Cargo.toml", "microflow-macros/Cargo.toml", "src/lib.rs", "microflow-macros/src/lib.rs"):
        (tmp_path / f).parent.mkdir(parents=True, exist_ok=True)
        shutil.copy(ref / f, tmp_path / f)
    subprocess.run(["git", "init", "-q", "."], cwd=tmp_path, check=True)
    r = subprocess.run(["git", "apply", "--check", str(RUST / "microflow-macros-cuda.patch")], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    subprocess.run(["git", "apply", str(RUST / "microflow-macros-cuda.patch")], cwd=tmp_path, check=True)
    out = (tmp_path / "microflow-macros" / "src" / "lib.rs").read_text()
    assert '#[cfg(feature = "cuda")]' in out and "microflow::cuda::predict_quantized" in out
    assert out.count("pub fn predict_quantized(") == 2      # CPU emission kept, CUDA emission added
