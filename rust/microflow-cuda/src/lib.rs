//! Drop-in for the functions `#[model("x.tflite")]` generates (microflow-macros/src/lib.rs:188-196), backed by
//! libmicroflow_cuda.so.  UNBUILT in this repository (no Rust toolchain in the build image); written against nalgebra 0.32 as
//! the reference pins it (Cargo.toml:25).
//!
//! The generated API takes and returns the reference's own buffer types: `Buffer2D<T, R, C> = SMatrix<T, R, C>` and
//! `Buffer4D<T, B, R, C, CH> = [SMatrix<[T; CH], R, C>; B]` (src/buffer.rs:5-16), both column-major.  Handles are created
//! with `MF_LAYOUT_NALGEBRA`, so those buffers cross the C ABI exactly as they lie in memory (`as_ptr()`, no host copy); the
//! transposition to NHWC runs on the GPU.  Every input rank the macro accepts is covered (rank 2: sine, speech; rank 4:
//! person_detect -- lib.rs:79-96), for `f32` (`predict`), `i8` and `u8` (`predict_quantized`).
use core::ffi::c_void;
use std::ffi::CStr;
use std::sync::OnceLock;

pub use microflow_cuda_sys as sys;
use nalgebra::SMatrix;

pub type Buffer2D<T, const R: usize, const C: usize> = SMatrix<T, R, C>;
pub type Buffer4D<T, const B: usize, const R: usize, const C: usize, const CH: usize> = [SMatrix<[T; CH], R, C>; B];

mod sealed {
    pub trait Sealed {}
    impl Sealed for i8 {}
    impl Sealed for u8 {}
    impl Sealed for f32 {}
}

/// Element types that cross the ABI: plain numbers whose all-zero bit pattern is a valid value.
pub trait Element: sealed::Sealed + Copy + core::fmt::Debug + PartialEq + 'static {
    /// tflite.fbs TensorType of a quantized element (0 for f32)
    const DTYPE: i32;
}
impl Element for i8 {
    const DTYPE: i32 = sys::MF_DTYPE_I8;
}
impl Element for u8 {
    const DTYPE: i32 = sys::MF_DTYPE_U8;
}
impl Element for f32 {
    const DTYPE: i32 = 0;
}

/// A reference buffer type as one contiguous run of `ELEMS` elements in nalgebra (column-major) order.
///
/// # Safety
/// `as_ptr` / `as_mut_ptr` must address exactly `ELEMS` initialised elements of `Elem`, laid out as MF_LAYOUT_NALGEBRA expects:
/// per batch entry a column-major `R x C` matrix of `CH`-element cells.
pub unsafe trait HostBuffer: Sized {
    type Elem: Element;
    const ELEMS: usize;
    fn zeroed() -> Self;
    fn as_ptr(&self) -> *const Self::Elem;
    fn as_mut_ptr(&mut self) -> *mut Self::Elem;
}

unsafe impl<T: Element, const R: usize, const C: usize> HostBuffer for SMatrix<T, R, C> {
    type Elem = T;
    const ELEMS: usize = R * C;
    fn zeroed() -> Self {
        // SMatrix<T, R, C> is `ArrayStorage<T, R, C>` = `[[T; R]; C]` of plain numbers
        unsafe { core::mem::zeroed() }
    }
    fn as_ptr(&self) -> *const T {
        self.as_slice().as_ptr()
    }
    fn as_mut_ptr(&mut self) -> *mut T {
        self.as_mut_slice().as_mut_ptr()
    }
}

unsafe impl<T: Element, const B: usize, const R: usize, const C: usize, const CH: usize> HostBuffer for [SMatrix<[T; CH], R, C>; B] {
    type Elem = T;
    const ELEMS: usize = B * R * C * CH;
    fn zeroed() -> Self {
        unsafe { core::mem::zeroed() }
    }
    fn as_ptr(&self) -> *const T {
        // `[SMatrix<[T; CH], R, C>; B]` is B contiguous column-major matrices of CH-element cells: [batch][col][row][chan]
        self.as_slice().as_ptr() as *const T
    }
    fn as_mut_ptr(&mut self) -> *mut T {
        self.as_mut_slice().as_mut_ptr() as *mut T
    }
}

/// One model on the GPU(s): parsed, pre-processed (the macro's `preprocess()` constants) and uploaded once.
pub struct Handle {
    raw: *mut sys::mf_model,
    in_elems: usize,
    out_elems: usize,
    in_dtype: i32,
}
unsafe impl Send for Handle {}
unsafe impl Sync for Handle {}

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::mf_last_error()) }.to_string_lossy().into_owned()
}

fn check(rc: i32) {
    if rc != sys::MF_OK {
        // the reference reports these at compile time (abort_call_site!) and is infallible at run time; the closest run-time
        // equivalent is a panic carrying the same text
        panic!("microflow_cuda status {rc}: {}", last_error());
    }
}

/// Which GPUs a handle runs on.  `MICROFLOW_CUDA_DEVICES=all` or `=0,1,2,3` selects a multi-device model (predict_many shards
/// the samples over them inside the library); unset = the current device.
fn devices_from_env() -> (i32, [i32; sys::MF_MAX_DEVICES]) {
    let mut list = [0i32; sys::MF_MAX_DEVICES];
    match std::env::var("MICROFLOW_CUDA_DEVICES") {
        Ok(v) if v.trim() == "all" => (-1, list),
        Ok(v) => {
            let mut n = 0usize;
            for tok in v.split(',').filter(|t| !t.trim().is_empty()) {
                if n == sys::MF_MAX_DEVICES {
                    panic!("MICROFLOW_CUDA_DEVICES lists more than {} devices", sys::MF_MAX_DEVICES);
                }
                list[n] = tok.trim().parse().unwrap_or_else(|_| panic!("MICROFLOW_CUDA_DEVICES: bad ordinal '{tok}'"));
                n += 1;
            }
            (n as i32, list)
        }
        Err(_) => (0, list),
    }
}

impl Handle {
    /// `include_bytes!("model.tflite")` -> a model on the device(s) named by `MICROFLOW_CUDA_DEVICES`.
    pub fn from_bytes(bytes: &[u8]) -> Self {
        let (n_devices, devices) = devices_from_env();
        Self::with_devices(bytes, n_devices, devices)
    }

    /// `n_devices`: 0 = current device, -1 = every visible GPU, n = `devices[..n]`.
    pub fn with_devices(bytes: &[u8], n_devices: i32, devices: [i32; sys::MF_MAX_DEVICES]) -> Self {
        let opt = sys::mf_options {
            struct_size: core::mem::size_of::<sys::mf_options>() as u32,
            device: -1,
            chunk: 0,
            flags: 0,
            layout: sys::MF_LAYOUT_NALGEBRA,
            n_devices,
            devices,
        };
        let mut raw = core::ptr::null_mut();
        check(unsafe { sys::mf_model_create_from_tflite(bytes.as_ptr() as *const c_void, bytes.len(), &opt, &mut raw) });
        let (mut ti, mut to) = (sys::mf_tensor_info::default(), sys::mf_tensor_info::default());
        check(unsafe { sys::mf_model_io_info(raw, &mut ti, &mut to) });
        Handle { raw, in_elems: ti.elems as usize, out_elems: to.elems as usize, in_dtype: ti.dtype }
    }

    pub fn raw(&self) -> *mut sys::mf_model {
        self.raw
    }

    fn check_shapes<I: HostBuffer, O: HostBuffer>(&self, quantized: bool) {
        assert_eq!(I::ELEMS, self.in_elems, "input buffer does not match the model's input tensor");
        assert_eq!(O::ELEMS, self.out_elems, "output buffer does not match the model's output tensor");
        if quantized {
            assert_eq!(I::Elem::DTYPE, self.in_dtype, "input element type does not match the model's input tensor type");
        }
    }
}
impl Drop for Handle {
    fn drop(&mut self) {
        unsafe { sys::mf_model_destroy(self.raw) }
    }
}

/// `M::predict(input)` (lib.rs:188-191): quantize with the input tensor's (scale, zero point), run, dequantize.
pub fn predict<I: HostBuffer<Elem = f32>, O: HostBuffer<Elem = f32>>(model: &Handle, input: &I) -> O {
    model.check_shapes::<I, O>(false);
    let mut out = O::zeroed();
    check(unsafe { sys::mf_predict(model.raw, input.as_ptr(), out.as_mut_ptr()) });
    out
}

/// `M::predict_quantized(input)` (lib.rs:193-196) for `i8` and `u8` models.
pub fn predict_quantized<I: HostBuffer, O: HostBuffer<Elem = f32>>(model: &Handle, input: &I) -> O {
    model.check_shapes::<I, O>(true);
    let mut out = O::zeroed();
    check(unsafe { sys::mf_predict_quantized(model.raw, input.as_ptr() as *const c_void, out.as_mut_ptr()) });
    out
}

/// New entry point: `n` independent samples.  A slice of the reference's buffers is `n` contiguous samples in MF_LAYOUT_NALGEBRA
/// order, so it crosses the ABI as it is; the library pipelines H2D / compute / D2H in chunks and, for a multi-device handle,
/// shards the samples into contiguous ranges, one per GPU (no collective on the inference path).
pub fn predict_many_quantized<I: HostBuffer, O: HostBuffer<Elem = f32>>(model: &Handle, inputs: &[I]) -> Vec<O> {
    model.check_shapes::<I, O>(true);
    let n = inputs.len();
    let mut out: Vec<O> = (0..n).map(|_| O::zeroed()).collect();
    if n > 0 {
        check(unsafe { sys::mf_predict_many_quantized(model.raw, inputs.as_ptr() as *const c_void, n, out.as_mut_ptr() as *mut f32) });
    }
    out
}

/// `predict_many` for `f32` inputs (quantized on the device, src/tensor.rs:80-86).
pub fn predict_many<I: HostBuffer<Elem = f32>, O: HostBuffer<Elem = f32>>(model: &Handle, inputs: &[I]) -> Vec<O> {
    model.check_shapes::<I, O>(false);
    let n = inputs.len();
    let mut out: Vec<O> = (0..n).map(|_| O::zeroed()).collect();
    if n > 0 {
        check(unsafe { sys::mf_predict_many(model.raw, inputs.as_ptr() as *const f32, n, out.as_mut_ptr() as *mut f32) });
    }
    out
}

/// Raw-slice form of `predict_many_quantized` for callers that keep their samples in pinned memory (`PinnedVec`): `samples` holds
/// `n` samples in the model's host layout, `out` receives `n * out_elems` floats.
pub fn predict_many_quantized_raw<T: Element>(model: &Handle, samples: &[T], n: usize, out: &mut [f32]) {
    assert_eq!(T::DTYPE, model.in_dtype, "element type does not match the model's input tensor type");
    assert!(samples.len() >= n * model.in_elems && out.len() >= n * model.out_elems, "buffers too small for {n} samples");
    check(unsafe { sys::mf_predict_many_quantized(model.raw, samples.as_ptr() as *const c_void, n, out.as_mut_ptr()) });
}

/// Page-locked host memory (`mf_host_alloc`): what the H2D / D2H legs want for full PCIe speed.
pub struct PinnedVec<T: Element> {
    ptr: *mut T,
    len: usize,
}
impl<T: Element> PinnedVec<T> {
    pub fn zeroed(len: usize) -> Self {
        let mut p: *mut c_void = core::ptr::null_mut();
        check(unsafe { sys::mf_host_alloc(&mut p, len * core::mem::size_of::<T>()) });
        unsafe { core::ptr::write_bytes(p as *mut u8, 0, len * core::mem::size_of::<T>()) };
        PinnedVec { ptr: p as *mut T, len }
    }
    pub fn as_slice(&self) -> &[T] {
        unsafe { core::slice::from_raw_parts(self.ptr, self.len) }
    }
    pub fn as_mut_slice(&mut self) -> &mut [T] {
        unsafe { core::slice::from_raw_parts_mut(self.ptr, self.len) }
    }
}
impl<T: Element> Drop for PinnedVec<T> {
    fn drop(&mut self) {
        unsafe { sys::mf_host_free(self.ptr as *mut c_void) };
    }
}

/// What the patched macro keeps in a `static`: one handle per `#[model]` struct, created on first use.
pub fn cached(cell: &'static OnceLock<Handle>, bytes: &'static [u8]) -> &'static Handle {
    cell.get_or_init(|| Handle::from_bytes(bytes))
}

/// NHWC row-major elements of a reference 4-D buffer (for callers that create MF_LAYOUT_NHWC handles themselves)
pub fn nhwc_from_buffer4d<T: Copy, const B: usize, const R: usize, const C: usize, const CH: usize>(x: &Buffer4D<T, B, R, C, CH>) -> Vec<T> {
    let mut v = Vec::with_capacity(B * R * C * CH);
    for m in x.iter() {
        for i in 0..R {
            for j in 0..C {
                v.extend_from_slice(&m[(i, j)]);
            }
        }
    }
    v
}
