//! Raw bindings to `include/microflow_cuda.h` (ABI version 2).  UNBUILT here: no Rust toolchain in the image.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

pub const MF_OK: c_int = 0;
pub const MF_DTYPE_U8: i32 = 3;
pub const MF_DTYPE_I8: i32 = 9;
pub const MF_FLAG_HOST_ONLY: u32 = 1;
pub const MF_FLAG_FORCE_GENERIC: u32 = 2;
pub const MF_FLAG_NO_TENSOR_CORE: u32 = 4;
pub const MF_LAYOUT_NHWC: u32 = 0;
/// host buffers in nalgebra's column-major order: `Buffer4D` / `Buffer2D` memory is passed as it is
pub const MF_LAYOUT_NALGEBRA: u32 = 1;

#[repr(C)]
pub struct mf_model {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct mf_options {
    pub struct_size: u32,
    pub device: i32,
    pub chunk: u32,
    pub flags: u32,
    pub layout: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct mf_tensor_info {
    pub rank: i32,
    pub dims: [i32; 4],
    pub dtype: i32,
    pub scale: f32,
    pub zero_point: i32,
    pub elems: u64,
}

extern "C" {
    pub fn mf_abi_version() -> c_int;
    pub fn mf_last_error() -> *const c_char;
    pub fn mf_device_count(count: *mut c_int) -> c_int;
    pub fn mf_model_create_from_tflite(buf: *const c_void, len: usize, opt: *const mf_options, out: *mut *mut mf_model) -> c_int;
    pub fn mf_model_create_from_file(path: *const c_char, opt: *const mf_options, out: *mut *mut mf_model) -> c_int;
    pub fn mf_model_destroy(m: *mut mf_model);
    pub fn mf_model_io_info(m: *const mf_model, input: *mut mf_tensor_info, output: *mut mf_tensor_info) -> c_int;
    pub fn mf_predict(m: *mut mf_model, in_f32: *const f32, out_f32: *mut f32) -> c_int;
    pub fn mf_predict_quantized(m: *mut mf_model, in_q: *const c_void, out_f32: *mut f32) -> c_int;
    pub fn mf_predict_many(m: *mut mf_model, in_f32: *const f32, n: usize, out_f32: *mut f32) -> c_int;
    pub fn mf_predict_many_quantized(m: *mut mf_model, in_q: *const c_void, n: usize, out_f32: *mut f32) -> c_int;
    pub fn mf_predict_many_logits(m: *mut mf_model, in_q: *const c_void, n: usize, out_q: *mut c_void, logits_q: *mut c_void) -> c_int;
    pub fn mf_host_alloc(p: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn mf_host_free(p: *mut c_void) -> c_int;
}
