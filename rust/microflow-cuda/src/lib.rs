//! Drop-in for the functions `#[model("x.tflite")]` generates (microflow-macros/src/lib.rs:188-196), backed by the
//! CUDA library.  UNBUILT here (no Rust toolchain).  Buffer types are the reference's: `Buffer2D = SMatrix<T, R, C>`
//! (column-major) and `Buffer4D = [SMatrix<[T; CH], R, C>; B]` (src/buffer.rs:5-16); handles are created with
//! `MF_LAYOUT_NALGEBRA`, so those buffers cross the C ABI as they lie in memory (the transposition to NHWC runs on the GPU).
//! `nhwc_from_buffer4d` / `rowmajor_from_buffer2d` remain for callers that want NHWC handles.
use core::ffi::c_void;
use std::ffi::CStr;
use std::sync::OnceLock;

use microflow_cuda_sys as sys;
use nalgebra::SMatrix;

pub type Buffer2D<T, const R: usize, const C: usize> = SMatrix<T, R, C>;
pub type Buffer4D<T, const B: usize, const R: usize, const C: usize, const CH: usize> = [SMatrix<[T; CH], R, C>; B];

pub struct Handle(*mut sys::mf_model);
unsafe impl Send for Handle {}
unsafe impl Sync for Handle {}

fn check(rc: i32) {
    if rc != sys::MF_OK {
        let msg = unsafe { CStr::from_ptr(sys::mf_last_error()) }.to_string_lossy().into_owned();
        // the reference reports these at compile time (abort_call_site!); at run time the closest equivalent is a panic
        panic!("microflow_cuda status {rc}: {msg}");
    }
}

impl Handle {
    /// `include_bytes!("model.tflite")` -> parsed, pre-processed, uploaded once.
    pub fn from_bytes(bytes: &[u8]) -> Self {
        // MF_LAYOUT_NALGEBRA: the library takes and returns the reference's column-major buffers as they lie in memory,
        // so `predict*` below hands over `input.as_ptr()` without a host-side transposition.
        let opt = sys::mf_options {
            struct_size: core::mem::size_of::<sys::mf_options>() as u32,
            device: -1,
            chunk: 0,
            flags: 0,
            layout: sys::MF_LAYOUT_NALGEBRA,
        };
        let mut h = core::ptr::null_mut();
        check(unsafe { sys::mf_model_create_from_tflite(bytes.as_ptr() as *const c_void, bytes.len(), &opt, &mut h) });
        Handle(h)
    }
}
impl Drop for Handle {
    fn drop(&mut self) {
        unsafe { sys::mf_model_destroy(self.0) }
    }
}

/// NHWC row-major bytes of a reference 4-D buffer (batch, row, col, channel)
pub fn nhwc_from_buffer4d<T: Copy, const B: usize, const R: usize, const C: usize, const CH: usize>(x: &Buffer4D<T, B, R, C, CH>) -> Vec<T> {
    let mut v = Vec::with_capacity(B * R * C * CH);
    for b in 0..B {
        for i in 0..R {
            for j in 0..C {
                v.extend_from_slice(&x[b][(i, j)]);
            }
        }
    }
    v
}
/// row-major elements of a reference 2-D buffer
pub fn rowmajor_from_buffer2d<T: Copy, const R: usize, const C: usize>(x: &Buffer2D<T, R, C>) -> Vec<T> {
    let mut v = Vec::with_capacity(R * C);
    for i in 0..R {
        for j in 0..C {
            v.push(x[(i, j)]);
        }
    }
    v
}

/// `predict_quantized` for a model with a 4-D int8 input and a 2-D output (person_detect shape).
pub fn predict_quantized_4d<const B: usize, const R: usize, const C: usize, const CH: usize, const OR: usize, const OC: usize>(
    model: &Handle,
    input: &Buffer4D<i8, B, R, C, CH>,
) -> Buffer2D<f32, OR, OC> {
    // `[SMatrix<[i8; CH], R, C>; B]` is B contiguous column-major matrices of CH-byte cells: exactly MF_LAYOUT_NALGEBRA
    let mut out = Buffer2D::<f32, OR, OC>::zeros();
    check(unsafe { sys::mf_predict_quantized(model.0, input.as_ptr() as *const c_void, out.as_mut_ptr()) });
    out
}

/// New entry point: n independent samples in the model's host layout (column-major per sample for handles made by
/// `Handle::from_bytes`), host buffers (pinned via mf_host_alloc for full PCIe speed).
pub fn predict_many_quantized(model: &Handle, samples: &[i8], n: usize, out: &mut [f32]) {
    check(unsafe { sys::mf_predict_many_quantized(model.0, samples.as_ptr() as *const c_void, n, out.as_mut_ptr()) });
}

/// What the patched macro keeps in a `static`: one handle per `#[model]` struct.
pub fn cached(cell: &'static OnceLock<Handle>, bytes: &'static [u8]) -> &'static Handle {
    cell.get_or_init(|| Handle::from_bytes(bytes))
}
