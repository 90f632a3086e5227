//! Raw bindings to `include/microflow_cuda.h` (ABI version 3): every exported function and every struct, field for field.
//! UNBUILT in this repository (no Rust toolchain in the build image); `tests/test_rust_binding.py` checks on every run that the
//! set of functions declared here equals the set the header declares and the library exports.
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

pub const MF_ABI_VERSION: c_int = 3;
pub const MF_MAX_DEVICES: usize = 16;

pub const MF_OK: c_int = 0;
pub const MF_ERR_FILE: c_int = 1;
pub const MF_ERR_INVALID_MODEL: c_int = 2;
pub const MF_ERR_UNSUPPORTED_TYPE: c_int = 3;
pub const MF_ERR_UNSUPPORTED_RANK: c_int = 4;
pub const MF_ERR_UNSUPPORTED_OP: c_int = 5;
pub const MF_ERR_UNSUPPORTED_ACTIVATION: c_int = 6;
pub const MF_ERR_UNSUPPORTED_SHAPE: c_int = 7;
pub const MF_ERR_VIEW_OUT_OF_BOUNDS: c_int = 8;
pub const MF_ERR_INVALID_ARG: c_int = 9;
pub const MF_ERR_NO_DEVICE: c_int = 10;
pub const MF_ERR_CUDA: c_int = 11;
pub const MF_ERR_NONFINITE_CONSTANT: c_int = 12;

pub const MF_DTYPE_U8: i32 = 3;
pub const MF_DTYPE_I8: i32 = 9;
pub const MF_PAD_SAME: i32 = 0;
pub const MF_PAD_VALID: i32 = 1;
pub const MF_ACT_NONE: i32 = 0;
pub const MF_ACT_RELU: i32 = 1;
pub const MF_ACT_RELU6: i32 = 3;
pub const MF_OP_AVERAGE_POOL_2D: i32 = 1;
pub const MF_OP_CONV_2D: i32 = 3;
pub const MF_OP_DEPTHWISE_CONV_2D: i32 = 4;
pub const MF_OP_FULLY_CONNECTED: i32 = 9;
pub const MF_OP_RESHAPE: i32 = 22;
pub const MF_OP_SOFTMAX: i32 = 25;

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
pub struct mf_op {
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
    /// 0: the single `device`; -1: every visible GPU; n >= 1: `devices[..n]` (predict_many* shard over them inside the call)
    pub n_devices: i32,
    pub devices: [i32; MF_MAX_DEVICES],
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

#[repr(C)]
#[derive(Clone, Copy)]
pub struct mf_layer_info {
    pub op: i32,
    pub in_dims: [i32; 4],
    pub out_dims: [i32; 4],
    pub in_rank: i32,
    pub out_rank: i32,
    pub kh: i32,
    pub kw: i32,
    pub stride_h: i32,
    pub stride_w: i32,
    pub padding: i32,
    pub activation: i32,
    pub in_zero_point: i32,
    pub out_zero_point: i32,
    pub in_scale: f32,
    pub out_scale: f32,
    pub act_lo: i32,
    pub act_hi: i32,
    pub n_c0: i32,
    pub n_c1: i32,
    pub macs: u64,
    pub bytes: u64,
    pub weight_bytes: u64,
    pub kernel: [c_char; 48],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct mf_conv_desc {
    pub dtype: i32,
    pub depthwise: i32,
    pub in_h: i32,
    pub in_w: i32,
    pub in_c: i32,
    pub out_h: i32,
    pub out_w: i32,
    pub out_c: i32,
    pub kh: i32,
    pub kw: i32,
    pub stride_h: i32,
    pub stride_w: i32,
    pub padding: i32,
    pub activation: i32,
    pub in_zero_point: i32,
    pub out_scale: f32,
    pub out_zero_point: i32,
    pub filters: *const c_void,
    pub filter_zero_points: *const i32,
    pub n_filter_zero_points: i32,
    pub c0: *const f32,
    pub c1: *const f32,
    pub n_c1: i32,
    pub impl_: i32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct mf_fc_desc {
    pub dtype: i32,
    pub in_features: i32,
    pub out_features: i32,
    pub weights_nk: *const c_void,
    pub weight_zero_point: i32,
    pub out_scale: f32,
    pub out_zero_point: i32,
    pub activation: i32,
    pub c0: *const f32,
    pub c1: f32,
    pub c2: *const i32,
    pub c3: i32,
    pub impl_: i32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct mf_pool_desc {
    pub dtype: i32,
    pub in_h: i32,
    pub in_w: i32,
    pub chans: i32,
    pub out_h: i32,
    pub out_w: i32,
    pub filter_h: i32,
    pub filter_w: i32,
    pub stride_h: i32,
    pub stride_w: i32,
    pub padding: i32,
    pub activation: i32,
    pub out_scale: f32,
    pub out_zero_point: i32,
    pub c0: f32,
    pub c1: f32,
    pub impl_: i32,
}

extern "C" {
    // ---- library
    pub fn mf_abi_version() -> c_int;
    pub fn mf_last_error() -> *const c_char;
    pub fn mf_status_string(status: c_int) -> *const c_char;
    pub fn mf_device_count(count: *mut c_int) -> c_int;
    // ---- model = what `#[model("path")]` builds at Rust compile time (microflow-macros/src/lib.rs:46-208)
    pub fn mf_model_create_from_tflite(buf: *const c_void, len: usize, opt: *const mf_options, out: *mut *mut mf_model) -> c_int;
    pub fn mf_model_create_from_file(path: *const c_char, opt: *const mf_options, out: *mut *mut mf_model) -> c_int;
    pub fn mf_model_destroy(m: *mut mf_model);
    pub fn mf_model_io_info(m: *const mf_model, input: *mut mf_tensor_info, output: *mut mf_tensor_info) -> c_int;
    pub fn mf_model_num_layers(m: *const mf_model) -> c_int;
    pub fn mf_model_layer_info(m: *const mf_model, layer: c_int, out: *mut mf_layer_info) -> c_int;
    pub fn mf_model_layer_constants(m: *const mf_model, layer: c_int, c0: *mut f32, c1: *mut f32, c2: *mut i32, c3: *mut i32, cap: c_int) -> c_int;
    pub fn mf_model_dump(m: *const mf_model, path: *const c_char) -> c_int;
    // ---- generated API (lib.rs:188-196) and its batched forms
    pub fn mf_predict(m: *mut mf_model, in_f32: *const f32, out_f32: *mut f32) -> c_int;
    pub fn mf_predict_quantized(m: *mut mf_model, in_q: *const c_void, out_f32: *mut f32) -> c_int;
    pub fn mf_predict_many(m: *mut mf_model, in_f32: *const f32, n: usize, out_f32: *mut f32) -> c_int;
    pub fn mf_predict_many_quantized(m: *mut mf_model, in_q: *const c_void, n: usize, out_f32: *mut f32) -> c_int;
    pub fn mf_predict_many_quantized_async(m: *mut mf_model, in_q: *const c_void, n: usize, out_f32: *mut f32) -> c_int;
    pub fn mf_predict_many_logits(m: *mut mf_model, in_q: *const c_void, n: usize, out_q: *mut c_void, logits_q: *mut c_void) -> c_int;
    pub fn mf_predict_many_device(m: *mut mf_model, d_in_q: *const c_void, n: usize, d_out_f32: *mut f32, d_out_q: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn mf_predict_many_device_on(m: *mut mf_model, index: c_int, d_in_q: *const c_void, n: usize, d_out_f32: *mut f32, d_out_q: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn mf_predict_trace(m: *mut mf_model, in_q: *const c_void, n: usize, layer_outs: *const *mut c_void) -> c_int;
    pub fn mf_model_synchronize(m: *mut mf_model) -> c_int;
    pub fn mf_model_set_profiling(m: *mut mf_model, enabled: c_int) -> c_int;
    pub fn mf_model_layer_times_ms(m: *mut mf_model, ms: *mut f32, cap: c_int) -> c_int;
    pub fn mf_model_launch_count(m: *const mf_model, count: *mut u64) -> c_int;
    pub fn mf_model_layer_launched(m: *const mf_model, layer: c_int) -> *const c_char;
    pub fn mf_model_devices(m: *const mf_model, devices: *mut i32, cap: c_int) -> c_int;
    pub fn mf_model_weight_broadcast(m: *const mf_model) -> *const c_char;
    pub fn mf_model_blob(m: *const mf_model, d_ptr: *mut *mut c_void, bytes: *mut usize) -> c_int;
    pub fn mf_features_from_bmp_gray8(bmp: *const c_void, len: usize, out: *mut c_void, cap: usize, height: *mut i32, width: *mut i32) -> c_int;
    pub fn mf_predict_many_bmp(m: *mut mf_model, bmps: *const *const c_void, lens: *const usize, n: usize, out_f32: *mut f32) -> c_int;
    pub fn mf_host_alloc(p: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn mf_host_free(p: *mut c_void) -> c_int;
    // ---- per-operator hooks (microflow::ops::*, src/ops/mod.rs:8-13)
    pub fn mf_op_conv_2d(d: *const mf_conv_desc, input: *const c_void, out: *mut c_void, batch: usize) -> c_int;
    pub fn mf_op_conv_chain(ops: *const mf_conv_desc, n_ops: c_int, input: *const c_void, out: *mut c_void, batch: usize, fuse: c_int) -> c_int;
    pub fn mf_op_conv_2d_create(d: *const mf_conv_desc, out: *mut *mut mf_op) -> c_int;
    pub fn mf_op_run_device(op: *mut mf_op, d_in: *const c_void, d_out: *mut c_void, batch: usize, stream: *mut c_void) -> c_int;
    pub fn mf_op_kernel_name(op: *const mf_op) -> *const c_char;
    pub fn mf_op_destroy(op: *mut mf_op);
    pub fn mf_op_fully_connected(d: *const mf_fc_desc, input: *const c_void, out: *mut c_void, batch: usize) -> c_int;
    pub fn mf_op_average_pool_2d(d: *const mf_pool_desc, input: *const c_void, out: *mut c_void, batch: usize) -> c_int;
    pub fn mf_op_softmax(dtype: i32, rows: i32, cols: i32, in_scale: f32, out_scale: f32, out_zero_point: i32, input: *const c_void, out: *mut c_void, batch: usize) -> c_int;
    pub fn mf_op_layout_transpose(input: *const c_void, out: *mut c_void, batch: usize, rows: i32, cols: i32, elem_bytes: i32, to_nalgebra: i32) -> c_int;
    pub fn mf_op_quantize(dtype: i32, scale: f32, zero_point: i32, input: *const f32, out: *mut c_void, n: usize) -> c_int;
    pub fn mf_op_dequantize(dtype: i32, scale: f32, zero_point: i32, input: *const c_void, out: *mut f32, n: usize) -> c_int;
}
