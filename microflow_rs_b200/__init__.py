"""microflow_rs_b200 -- B200 (sm_100a) backend for MicroFlow's quantized op-kernel hot path.

Host-side mirror, in Python over ctypes, of the reference's user-visible surface:

    reference (Rust)                                   here
    -----------------------------------------------    ------------------------------------------
    #[model("models/x.tflite")] struct M;              M = microflow_rs_b200.model("models/x.tflite")
    M::predict(buf_f32) -> buf_f32                     M.predict(x_f32)            (lib.rs:188-191)
    M::predict_quantized(buf_i8) -> buf_f32            M.predict_quantized(x_i8)   (lib.rs:193-196)
    (new) predict_many over independent samples        M.predict_many(xs) / M.predict_many_quantized(xs)
    microflow::ops::{conv_2d, depthwise_conv_2d,       microflow_rs_b200.ops.{conv_2d, depthwise_conv_2d,
      fully_connected, average_pool_2d, softmax}         fully_connected, average_pool_2d, softmax}

Everything goes through the C ABI in include/microflow_cuda.h (libmicroflow_cuda.so, built in-tree by
_build.py).  There is no CPU fallback: without the CUDA extension or without a B200 the calls raise.
torch is not needed by this module; bench.py / tests use it only for device buffers and torch.distributed.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

from . import _build

__all__ = ["model", "Model", "ConvOp", "ops", "MicroflowError", "lib", "build", "device_count", "PinnedBuffer", "features_from_bmp"]

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libmicroflow_cuda.so"

DTYPE_I8, DTYPE_U8 = 9, 3
PAD = {"same": 0, "valid": 1, 0: 0, 1: 1}
ACT = {"none": 0, "relu": 1, "relu6": 3, 0: 0, 1: 1, 3: 3}
OP_NAMES = {1: "average_pool_2d", 3: "conv_2d", 4: "depthwise_conv_2d", 9: "fully_connected", 22: "reshape", 25: "softmax"}
FLAG_HOST_ONLY, FLAG_FORCE_GENERIC, FLAG_NO_TENSOR_CORE = 1, 2, 4


class MicroflowError(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"microflow_cuda status {status}: {text}")
        self.status = status
        self.text = text


LAYOUT_NHWC, LAYOUT_NALGEBRA = 0, 1   # mf_options.layout: host buffers row-major NHWC, or the reference's column-major nalgebra buffers


MAX_DEVICES = 16


class _Options(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("chunk", C.c_uint32), ("flags", C.c_uint32), ("layout", C.c_uint32),
                ("n_devices", C.c_int32), ("devices", C.c_int32 * MAX_DEVICES)]


class _TensorInfo(C.Structure):
    _fields_ = [("rank", C.c_int32), ("dims", C.c_int32 * 4), ("dtype", C.c_int32), ("scale", C.c_float), ("zero_point", C.c_int32),
                ("elems", C.c_uint64)]


class _LayerInfo(C.Structure):
    _fields_ = [("op", C.c_int32), ("in_dims", C.c_int32 * 4), ("out_dims", C.c_int32 * 4), ("in_rank", C.c_int32), ("out_rank", C.c_int32),
                ("kh", C.c_int32), ("kw", C.c_int32), ("stride_h", C.c_int32), ("stride_w", C.c_int32), ("padding", C.c_int32),
                ("activation", C.c_int32), ("in_zero_point", C.c_int32), ("out_zero_point", C.c_int32), ("in_scale", C.c_float),
                ("out_scale", C.c_float), ("act_lo", C.c_int32), ("act_hi", C.c_int32), ("n_c0", C.c_int32), ("n_c1", C.c_int32),
                ("macs", C.c_uint64), ("bytes", C.c_uint64), ("weight_bytes", C.c_uint64), ("kernel", C.c_char * 48)]


class _ConvDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("depthwise", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32), ("in_c", C.c_int32),
                ("out_h", C.c_int32), ("out_w", C.c_int32), ("out_c", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
                ("stride_h", C.c_int32), ("stride_w", C.c_int32), ("padding", C.c_int32), ("activation", C.c_int32),
                ("in_zero_point", C.c_int32), ("out_scale", C.c_float), ("out_zero_point", C.c_int32), ("filters", C.c_void_p),
                ("filter_zero_points", C.c_void_p), ("n_filter_zero_points", C.c_int32), ("c0", C.c_void_p), ("c1", C.c_void_p),
                ("n_c1", C.c_int32), ("impl", C.c_int32)]


class _FcDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("in_features", C.c_int32), ("out_features", C.c_int32), ("weights_nk", C.c_void_p),
                ("weight_zero_point", C.c_int32), ("out_scale", C.c_float), ("out_zero_point", C.c_int32), ("activation", C.c_int32),
                ("c0", C.c_void_p), ("c1", C.c_float), ("c2", C.c_void_p), ("c3", C.c_int32), ("impl", C.c_int32)]


class _PoolDesc(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32), ("chans", C.c_int32), ("out_h", C.c_int32),
                ("out_w", C.c_int32), ("filter_h", C.c_int32), ("filter_w", C.c_int32), ("stride_h", C.c_int32), ("stride_w", C.c_int32),
                ("padding", C.c_int32), ("activation", C.c_int32), ("out_scale", C.c_float), ("out_zero_point", C.c_int32),
                ("c0", C.c_float), ("c1", C.c_float), ("impl", C.c_int32)]


# every symbol include/microflow_cuda.h declares (tests check the .so exports each one)
ABI_SYMBOLS = [
    "mf_abi_version", "mf_last_error", "mf_status_string", "mf_device_count", "mf_model_create_from_tflite", "mf_model_create_from_file",
    "mf_model_destroy", "mf_model_io_info", "mf_model_num_layers", "mf_model_layer_info", "mf_model_layer_constants", "mf_model_dump",
    "mf_predict", "mf_predict_quantized", "mf_predict_many", "mf_predict_many_quantized", "mf_predict_many_quantized_async", "mf_predict_many_logits", "mf_predict_many_device",
    "mf_predict_trace", "mf_model_synchronize", "mf_model_set_profiling", "mf_model_layer_times_ms", "mf_model_launch_count", "mf_model_blob",
    "mf_host_alloc", "mf_host_free", "mf_op_conv_2d", "mf_op_conv_2d_create", "mf_op_run_device", "mf_op_kernel_name", "mf_op_destroy", "mf_op_fully_connected", "mf_op_average_pool_2d", "mf_op_softmax", "mf_op_quantize",
    "mf_op_dequantize", "mf_op_layout_transpose", "mf_model_layer_launched", "mf_model_devices", "mf_model_weight_broadcast", "mf_predict_many_device_on", "mf_op_conv_chain", "mf_features_from_bmp_gray8", "mf_predict_many_bmp",
]

_lib = None


def build(force=False, verbose=False):
    """Compile libmicroflow_cuda.so for sm_100a (nvcc; works without a GPU)."""
    return _build.build(force=force, verbose=verbose)


def lib():
    """Loads the CUDA extension; fails loudly if it is missing (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            if os.environ.get("MICROFLOW_NO_AUTOBUILD"):
                raise ImportError(f"{LIB_PATH} is missing: run `python -m microflow_rs_b200._build` (nvcc, sm_100a). There is no CPU fallback.")
            build()
        L = C.CDLL(str(LIB_PATH))
        L.mf_last_error.restype = C.c_char_p
        L.mf_status_string.restype = C.c_char_p
        L.mf_status_string.argtypes = [C.c_int]
        L.mf_model_create_from_tflite.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(_Options), C.POINTER(C.c_void_p)]
        L.mf_model_create_from_file.argtypes = [C.c_char_p, C.POINTER(_Options), C.POINTER(C.c_void_p)]
        L.mf_model_destroy.argtypes = [C.c_void_p]
        L.mf_model_destroy.restype = None
        L.mf_model_io_info.argtypes = [C.c_void_p, C.POINTER(_TensorInfo), C.POINTER(_TensorInfo)]
        L.mf_model_num_layers.argtypes = [C.c_void_p]
        L.mf_model_layer_info.argtypes = [C.c_void_p, C.c_int, C.POINTER(_LayerInfo)]
        L.mf_model_layer_constants.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_int]
        L.mf_model_dump.argtypes = [C.c_void_p, C.c_char_p]
        L.mf_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mf_predict_quantized.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.mf_predict_many.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.mf_predict_many_quantized.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.mf_predict_many_quantized_async.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.mf_predict_many_logits.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.mf_predict_many_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mf_predict_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.mf_op_layout_transpose.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        L.mf_model_synchronize.argtypes = [C.c_void_p]
        L.mf_model_set_profiling.argtypes = [C.c_void_p, C.c_int]
        L.mf_model_layer_times_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.mf_model_launch_count.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.mf_model_blob.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.mf_model_layer_launched.argtypes = [C.c_void_p, C.c_int]
        L.mf_model_layer_launched.restype = C.c_char_p
        L.mf_model_devices.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_int]
        L.mf_model_weight_broadcast.argtypes = [C.c_void_p]
        L.mf_model_weight_broadcast.restype = C.c_char_p
        L.mf_predict_many_device_on.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mf_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        L.mf_host_free.argtypes = [C.c_void_p]
        L.mf_device_count.argtypes = [C.POINTER(C.c_int)]
        L.mf_op_conv_2d.argtypes = [C.POINTER(_ConvDesc), C.c_void_p, C.c_void_p, C.c_size_t]
        L.mf_features_from_bmp_gray8.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        L.mf_predict_many_bmp.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_size_t, C.c_void_p]
        L.mf_op_conv_chain.argtypes = [C.POINTER(_ConvDesc), C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.mf_op_conv_2d_create.argtypes = [C.POINTER(_ConvDesc), C.POINTER(C.c_void_p)]
        L.mf_op_run_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.mf_op_kernel_name.argtypes = [C.c_void_p]
        L.mf_op_kernel_name.restype = C.c_char_p
        L.mf_op_destroy.argtypes = [C.c_void_p]
        L.mf_op_destroy.restype = None
        L.mf_op_fully_connected.argtypes = [C.POINTER(_FcDesc), C.c_void_p, C.c_void_p, C.c_size_t]
        L.mf_op_average_pool_2d.argtypes = [C.POINTER(_PoolDesc), C.c_void_p, C.c_void_p, C.c_size_t]
        L.mf_op_softmax.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]
        L.mf_op_quantize.argtypes = [C.c_int32, C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]
        L.mf_op_dequantize.argtypes = [C.c_int32, C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]
        _lib = L
    return _lib


def _check(status):
    if status != 0:
        raise MicroflowError(status, lib().mf_last_error().decode(errors="replace"))


def device_count():
    n = C.c_int(0)
    _check(lib().mf_device_count(C.byref(n)))
    return n.value


class PinnedBuffer:
    """Page-locked host memory (mf_host_alloc) exposed as a numpy array; used for the H2D/D2H legs."""

    def __init__(self, shape, dtype):
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        self.dtype = np.dtype(dtype)
        nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self._p = C.c_void_p()
        _check(lib().mf_host_alloc(C.byref(self._p), nbytes))
        buf = (C.c_uint8 * max(nbytes, 1)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self._p:
            self.array = None
            lib().mf_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _dtype_code(a):
    if a.dtype == np.int8:
        return DTYPE_I8
    if a.dtype == np.uint8:
        return DTYPE_U8
    raise TypeError(f"quantized tensors must be int8 or uint8, got {a.dtype}")


class Model:
    """What `#[model("path.tflite")]` generates in the reference: predict / predict_quantized (+ predict_many)."""

    def __init__(self, path_or_bytes, device=-1, chunk=0, flags=0, layout=LAYOUT_NHWC, devices=None):
        """devices: None = the single `device`; "all" = every visible GPU; a list of ordinals = those GPUs.  With more than one
        device predict_many* shard the samples into contiguous ranges, one per GPU, inside the library (mf_options.devices)."""
        self._h = C.c_void_p()
        opt = _Options(C.sizeof(_Options), device, chunk, flags, layout)
        if isinstance(devices, str) and devices == "all":
            opt.n_devices = -1
        elif devices is not None:
            devices = [int(d) for d in devices]
            if len(devices) > MAX_DEVICES:
                raise ValueError(f"at most {MAX_DEVICES} devices")
            opt.n_devices = len(devices)
            for i, d in enumerate(devices):
                opt.devices[i] = d
        if isinstance(path_or_bytes, (str, os.PathLike)):
            _check(lib().mf_model_create_from_file(str(path_or_bytes).encode(), C.byref(opt), C.byref(self._h)))
        else:
            data = bytes(path_or_bytes)
            _check(lib().mf_model_create_from_tflite(data, len(data), C.byref(opt), C.byref(self._h)))
        ti, to = _TensorInfo(), _TensorInfo()
        _check(lib().mf_model_io_info(self._h, C.byref(ti), C.byref(to)))
        self.in_shape = tuple(ti.dims[: ti.rank])
        self.out_shape = tuple(to.dims[: to.rank])
        self.in_scale, self.in_zp, self.in_elems = np.float32(ti.scale), ti.zero_point, int(ti.elems)
        self.out_scale, self.out_zp, self.out_elems = np.float32(to.scale), to.zero_point, int(to.elems)
        self.dtype = np.uint8 if ti.dtype == DTYPE_U8 else np.int8
        self.out_dtype = np.uint8 if to.dtype == DTYPE_U8 else np.int8
        self.flags = flags
        self.layers = []
        for i in range(lib().mf_model_num_layers(self._h)):
            li = _LayerInfo()
            _check(lib().mf_model_layer_info(self._h, i, C.byref(li)))
            self.layers.append(dict(
                op=OP_NAMES.get(li.op, str(li.op)), in_shape=tuple(li.in_dims[: li.in_rank]), out_shape=tuple(li.out_dims[: li.out_rank]),
                kernel_hw=(li.kh, li.kw), strides=(li.stride_h, li.stride_w), padding=li.padding, activation=li.activation,
                in_zp=li.in_zero_point, out_zp=li.out_zero_point, in_scale=np.float32(li.in_scale), out_scale=np.float32(li.out_scale),
                clamp=(li.act_lo, li.act_hi), n_c0=li.n_c0, n_c1=li.n_c1, macs=int(li.macs), bytes=int(li.bytes), weight_bytes=int(li.weight_bytes),
                kernel=li.kernel.decode(), out_elems=int(np.prod(li.out_dims[: li.out_rank])) if li.out_rank else 0))

    def close(self):
        if getattr(self, "_h", None):
            lib().mf_model_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the macro's pre-processing, for parity tests ------------------------------------------------
    def layer_constants(self, i):
        n = max(self.layers[i]["n_c0"], self.layers[i]["n_c1"], 1)
        c0, c1, c2 = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.int32)
        c3 = C.c_int32(0)
        _check(lib().mf_model_layer_constants(self._h, i, c0.ctypes.data, c1.ctypes.data, c2.ctypes.data, C.byref(c3), n))
        return c0[: self.layers[i]["n_c0"]], c1[: self.layers[i]["n_c1"]], c2, c3.value

    def dump(self, path):
        _check(lib().mf_model_dump(self._h, str(path).encode()))

    # ---- generated API (lib.rs:188-196) ------------------------------------------------------------
    def predict(self, x):
        x = np.ascontiguousarray(np.asarray(x, np.float32).reshape(-1))
        if x.size != self.in_elems:
            raise ValueError(f"predict expects {self.in_shape}, got {x.size} elements")
        out = np.zeros(self.out_elems, np.float32)
        _check(lib().mf_predict(self._h, x.ctypes.data, out.ctypes.data))
        return out.reshape(self.out_shape)

    def predict_quantized(self, x):
        x = np.ascontiguousarray(np.asarray(x).reshape(-1))
        if x.dtype != self.dtype or x.size != self.in_elems:
            raise ValueError(f"predict_quantized expects {self.dtype} {self.in_shape}")
        out = np.zeros(self.out_elems, np.float32)
        _check(lib().mf_predict_quantized(self._h, x.ctypes.data, out.ctypes.data))
        return out.reshape(self.out_shape)

    # ---- batched: n independent samples ----------------------------------------------------------------
    def _n(self, xs, dtype):
        xs = np.asarray(xs)
        if xs.dtype != dtype:
            raise ValueError(f"expected dtype {dtype}, got {xs.dtype}")
        if not xs.flags["C_CONTIGUOUS"]:
            xs = np.ascontiguousarray(xs)
        if xs.size % self.in_elems:
            raise ValueError(f"input size {xs.size} is not a multiple of {self.in_elems}")
        return xs, xs.size // self.in_elems

    def _out(self, out, n):
        """A caller-supplied result array is written by the library through its raw pointer: it must be exactly what the C side
        expects (float32, C-contiguous, writable, n * out_elems elements)."""
        if out is None:
            return np.zeros((n, self.out_elems), np.float32)
        if not isinstance(out, np.ndarray) or out.dtype != np.float32 or not out.flags["C_CONTIGUOUS"] or not out.flags["WRITEABLE"] or \
                out.size != n * self.out_elems:
            raise ValueError(f"out must be a writable C-contiguous float32 array of {n} x {self.out_elems} elements")
        return out

    def predict_many(self, xs, out=None):
        xs, n = self._n(xs, np.float32)
        out = self._out(out, n)
        _check(lib().mf_predict_many(self._h, xs.ctypes.data, n, out.ctypes.data))
        return out

    def predict_many_quantized(self, xs, out=None):
        xs, n = self._n(xs, self.dtype)
        out = self._out(out, n)
        _check(lib().mf_predict_many_quantized(self._h, xs.ctypes.data, n, out.ctypes.data))
        return out

    def predict_many_quantized_async(self, xs_pinned, out_pinned):
        """Enqueue only (pinned numpy views from PinnedBuffer); call synchronize() before reading `out_pinned`.
        The copies land after this call returns, so both arrays are used in place: no silent conversion copies."""
        if not isinstance(xs_pinned, np.ndarray) or xs_pinned.dtype != self.dtype or not xs_pinned.flags["C_CONTIGUOUS"] or \
                xs_pinned.size % self.in_elems:
            raise ValueError(f"xs_pinned must be a C-contiguous {np.dtype(self.dtype)} array of n x {self.in_elems} elements")
        n = xs_pinned.size // self.in_elems
        if out_pinned is None:
            raise ValueError("out_pinned is required (the result arrives asynchronously)")
        out_pinned = self._out(out_pinned, n)
        _check(lib().mf_predict_many_quantized_async(self._h, xs_pinned.ctypes.data, n, out_pinned.ctypes.data))

    def predict_many_logits(self, xs, want_logits=True):
        """Returns (final quantized output [n, out_elems], pre-softmax int8 logits or None)."""
        xs, n = self._n(xs, self.dtype)
        outq = np.zeros((n, self.out_elems), self.out_dtype)
        logits = None
        lp = None
        if want_logits:
            tail = [L for L in self.layers if L["op"] == "softmax"]
            if tail:
                le = int(np.prod(tail[-1]["in_shape"]))
                logits = np.zeros((n, le), self.out_dtype)
                lp = logits.ctypes.data
        _check(lib().mf_predict_many_logits(self._h, xs.ctypes.data, n, outq.ctypes.data, lp))
        return outq, logits

    def predict_many_bmp(self, images):
        """images: a list of BMP files as bytes (8-bit gray, the model's input size) -> staged and run as one batch."""
        n = len(images)
        arr = (C.c_char_p * n)(*images)
        lens = (C.c_size_t * n)(*[len(b) for b in images])
        out = np.zeros((n, self.out_elems), np.float32)
        _check(lib().mf_predict_many_bmp(self._h, arr, lens, n, out.ctypes.data))
        return out

    def predict_many_device(self, d_in_ptr, n, d_out_f32_ptr=None, d_out_q_ptr=None, stream=None):
        """Device-resident buffers given as raw pointers (e.g. torch.Tensor.data_ptr()); asynchronous."""
        _check(lib().mf_predict_many_device(self._h, C.c_void_p(d_in_ptr), n, C.c_void_p(d_out_f32_ptr or 0), C.c_void_p(d_out_q_ptr or 0),
                                            C.c_void_p(stream or 0)))

    def predict_trace(self, xs):
        """Quantized output of every layer for the given samples (parity debugging)."""
        xs, n = self._n(xs, self.dtype)
        outs = [np.zeros((n,) + tuple(L["out_shape"][1:] if len(L["out_shape"]) > 1 else L["out_shape"]), self.dtype) for L in self.layers]
        outs = [np.zeros((n, L["out_elems"]), self.dtype) for L in self.layers]
        ptrs = (C.c_void_p * len(outs))(*[o.ctypes.data for o in outs])
        _check(lib().mf_predict_trace(self._h, xs.ctypes.data, n, ptrs))
        return outs

    def synchronize(self):
        _check(lib().mf_model_synchronize(self._h))

    def set_profiling(self, on):
        _check(lib().mf_model_set_profiling(self._h, int(bool(on))))

    def layer_times_ms(self):
        ms = np.zeros(len(self.layers), np.float32)
        _check(lib().mf_model_layer_times_ms(self._h, ms.ctypes.data, len(ms)))
        return ms

    def launched_kernels(self):
        """Per layer: the kernel the most recent predict*/trace call actually launched ('' = none / inside the previous launch)."""
        return [(lib().mf_model_layer_launched(self._h, i) or b"").decode() for i in range(len(self.layers))]

    @property
    def devices(self):
        buf = (C.c_int32 * MAX_DEVICES)()
        n = lib().mf_model_devices(self._h, buf, MAX_DEVICES)
        return [int(buf[i]) for i in range(min(n, MAX_DEVICES))]

    @property
    def weight_broadcast(self):
        return lib().mf_model_weight_broadcast(self._h).decode()

    def predict_many_device_on(self, index, d_in_ptr, n, d_out_f32_ptr=None, d_out_q_ptr=None, stream=None):
        """predict_many_device on the replica of devices[index] of a multi-device model."""
        _check(lib().mf_predict_many_device_on(self._h, index, C.c_void_p(d_in_ptr), n, C.c_void_p(d_out_f32_ptr or 0), C.c_void_p(d_out_q_ptr or 0),
                                               C.c_void_p(stream or 0)))

    def launch_count(self):
        c = C.c_uint64(0)
        _check(lib().mf_model_launch_count(self._h, C.byref(c)))
        return c.value

    def blob(self):
        """(device pointer, bytes) of the static weights/constants blob (multi-GPU init broadcast)."""
        p, n = C.c_void_p(), C.c_size_t()
        _check(lib().mf_model_blob(self._h, C.byref(p), C.byref(n)))
        return p.value or 0, n.value


class ConvOp:
    """Persistent Conv2D / DepthwiseConv2D operator (mf_op_conv_2d_create): plan once, run on device-resident buffers."""

    def __init__(self, in_hwc, in_zp, filters, filter_zp, out_scale, out_zp, act, pad, strides, c0, c1, out_hw, depthwise=False, impl=0, dtype=np.int8):
        filters = np.ascontiguousarray(filters)
        H, W, Cin = in_hwc
        if depthwise:
            _, KH, KW, Cout = filters.shape
        else:
            Cout, KH, KW, _ = filters.shape
        fz = np.ascontiguousarray(np.atleast_1d(filter_zp), np.int32)
        c0 = np.ascontiguousarray(c0, np.float32)
        c1 = np.ascontiguousarray(np.atleast_1d(c1), np.float32)
        d = _ConvDesc(DTYPE_U8 if np.dtype(dtype) == np.uint8 else DTYPE_I8, int(depthwise), H, W, Cin, out_hw[0], out_hw[1], Cout, KH, KW, strides[0],
                      strides[1], PAD[pad], ACT[act], int(in_zp), np.float32(out_scale), int(out_zp), filters.ctypes.data, fz.ctypes.data, len(fz),
                      c0.ctypes.data, c1.ctypes.data, len(c1), impl)
        self._h = C.c_void_p()
        _check(lib().mf_op_conv_2d_create(C.byref(d), C.byref(self._h)))
        self.kernel = lib().mf_op_kernel_name(self._h).decode()
        self.in_elems = H * W * Cin
        self.out_elems = out_hw[0] * out_hw[1] * Cout
        self.macs = out_hw[0] * out_hw[1] * Cout * KH * KW * (1 if depthwise else Cin)
        self.weight_bytes = filters.size + 8 * Cout

    def run_device(self, d_in_ptr, d_out_ptr, batch, stream=None):
        _check(lib().mf_op_run_device(self._h, C.c_void_p(d_in_ptr), C.c_void_p(d_out_ptr), batch, C.c_void_p(stream or 0)))

    def close(self):
        if getattr(self, "_h", None):
            lib().mf_op_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def features_from_bmp(data):
    """samples/person.bmp -> the reference's `features::PERSON` tensor (samples/features/person_detect.rs): int8 [H, W, 1], top row first."""
    h, w = C.c_int32(0), C.c_int32(0)
    _check(lib().mf_features_from_bmp_gray8(data, len(data), None, 0, C.byref(h), C.byref(w)))
    out = np.zeros((h.value, w.value, 1), np.int8)
    _check(lib().mf_features_from_bmp_gray8(data, len(data), out.ctypes.data, out.size, C.byref(h), C.byref(w)))
    return out


def model(path, **kw):
    """`#[model("path")]` equivalent."""
    return Model(path, **kw)


class _Ops:
    """Per-operator hooks mirroring `microflow::ops::*` (src/ops/mod.rs:8-13); arrays are [batch, ...] NHWC."""

    last_kernel = ""

    def _done(self, status):
        _check(status)
        _Ops.last_kernel = lib().mf_last_error().decode()

    def conv_2d(self, x, in_zp, filters, filter_zp, out_scale, out_zp, act, pad, strides, c0, c1, out_hw, depthwise=False, impl=0):
        x = np.ascontiguousarray(x)
        filters = np.ascontiguousarray(filters)
        B, H, W, Cin = x.shape
        if depthwise:
            _, KH, KW, Cout = filters.shape
        else:
            Cout, KH, KW, _ = filters.shape
        fz = np.ascontiguousarray(np.atleast_1d(filter_zp), np.int32)
        c0 = np.ascontiguousarray(c0, np.float32)
        c1 = np.ascontiguousarray(np.atleast_1d(c1), np.float32)
        out = np.zeros((B, out_hw[0], out_hw[1], Cout), x.dtype)
        d = _ConvDesc(_dtype_code(x), int(depthwise), H, W, Cin, out_hw[0], out_hw[1], Cout, KH, KW, strides[0], strides[1], PAD[pad], ACT[act],
                      int(in_zp), np.float32(out_scale), int(out_zp), filters.ctypes.data, fz.ctypes.data, len(fz), c0.ctypes.data, c1.ctypes.data,
                      len(c1), impl)
        self._done(lib().mf_op_conv_2d(C.byref(d), x.ctypes.data, out.ctypes.data, B))
        return out

    def depthwise_conv_2d(self, *a, **k):
        return self.conv_2d(*a, depthwise=True, **k)

    def conv_chain(self, x, layers, fuse):
        """layers: list of dicts with the keyword arguments of conv_2d (in_zp, filters, filter_zp, out_scale, out_zp, act, pad, strides,
        c0, c1, out_hw, depthwise); op k consumes op k-1's output.  fuse=True: one fused_chain_kernel launch (or MicroflowError 7)."""
        x = np.ascontiguousarray(x)
        B, H, W, Cc = x.shape
        descs = (_ConvDesc * len(layers))()
        keep = []
        for i, L in enumerate(layers):
            filters = np.ascontiguousarray(L["filters"])
            dw = bool(L.get("depthwise", False))
            if dw:
                _, KH, KW, Cout = filters.shape
            else:
                Cout, KH, KW, _ = filters.shape
            fz = np.ascontiguousarray(np.atleast_1d(L["filter_zp"]), np.int32)
            c0 = np.ascontiguousarray(L["c0"], np.float32)
            c1 = np.ascontiguousarray(np.atleast_1d(L["c1"]), np.float32)
            keep += [filters, fz, c0, c1]
            oh, ow = L["out_hw"]
            descs[i] = _ConvDesc(_dtype_code(x), int(dw), H, W, Cc, oh, ow, Cout, KH, KW, L["strides"][0], L["strides"][1], PAD[L["pad"]], ACT[L["act"]],
                                 int(L["in_zp"]), np.float32(L["out_scale"]), int(L["out_zp"]), filters.ctypes.data, fz.ctypes.data, len(fz),
                                 c0.ctypes.data, c1.ctypes.data, len(c1), 0)
            H, W, Cc = oh, ow, Cout
        out = np.zeros((B, H, W, Cc), x.dtype)
        self._done(lib().mf_op_conv_chain(descs, len(layers), x.ctypes.data, out.ctypes.data, B, int(bool(fuse))))
        return out

    def fully_connected(self, x, w_nk, w_zp, out_scale, out_zp, act, c0, c1, c2, c3, impl=0):
        x = np.ascontiguousarray(x)
        w_nk = np.ascontiguousarray(w_nk)
        B, K = x.shape
        N = w_nk.shape[0]
        c0 = np.ascontiguousarray(c0, np.float32)
        c2 = np.ascontiguousarray(c2, np.int32)
        out = np.zeros((B, N), x.dtype)
        d = _FcDesc(_dtype_code(x), K, N, w_nk.ctypes.data, int(w_zp), np.float32(out_scale), int(out_zp), ACT[act], c0.ctypes.data, np.float32(c1),
                    c2.ctypes.data, int(c3), impl)
        self._done(lib().mf_op_fully_connected(C.byref(d), x.ctypes.data, out.ctypes.data, B))
        return out

    def average_pool_2d(self, x, filter_hw, out_scale, out_zp, act, pad, strides, c0, c1, out_hw, impl=0):
        x = np.ascontiguousarray(x)
        B, H, W, Cc = x.shape
        out = np.zeros((B, out_hw[0], out_hw[1], Cc), x.dtype)
        d = _PoolDesc(_dtype_code(x), H, W, Cc, out_hw[0], out_hw[1], filter_hw[0], filter_hw[1], strides[0], strides[1], PAD[pad], ACT[act],
                      np.float32(out_scale), int(out_zp), np.float32(c0), np.float32(c1), impl)
        self._done(lib().mf_op_average_pool_2d(C.byref(d), x.ctypes.data, out.ctypes.data, B))
        return out

    def softmax(self, x, in_scale, out_scale, out_zp):
        x = np.ascontiguousarray(x)
        B, rows, cols = x.shape
        out = np.zeros_like(x)
        self._done(lib().mf_op_softmax(_dtype_code(x), rows, cols, np.float32(in_scale), np.float32(out_scale), int(out_zp), x.ctypes.data,
                                       out.ctypes.data, B))
        return out

    def quantize(self, x, scale, zp, dtype=np.int8):
        x = np.ascontiguousarray(x, np.float32)
        out = np.zeros(x.shape, dtype)
        _check(lib().mf_op_quantize(DTYPE_U8 if np.dtype(dtype) == np.uint8 else DTYPE_I8, np.float32(scale), int(zp), x.ctypes.data, out.ctypes.data,
                                    x.size))
        return out

    def layout_transpose(self, x, to_nalgebra):
        """x: [batch, rows, cols, cell...] (to_nalgebra) or [batch, cols, rows, cell...]; returns the other memory order."""
        x = np.ascontiguousarray(x)
        b, d1, d2 = x.shape[:3]
        elem = int(np.prod(x.shape[3:], dtype=np.int64)) * x.itemsize
        rows, cols = (d1, d2) if to_nalgebra else (d2, d1)
        out = np.zeros((b, d2, d1) + x.shape[3:], x.dtype)
        _check(lib().mf_op_layout_transpose(x.ctypes.data, out.ctypes.data, b, rows, cols, elem, 1 if to_nalgebra else 0))
        return out

    def dequantize(self, q, scale, zp):
        q = np.ascontiguousarray(q)
        out = np.zeros(q.shape, np.float32)
        _check(lib().mf_op_dequantize(_dtype_code(q), np.float32(scale), int(zp), q.ctypes.data, out.ctypes.data, q.size))
        return out


ops = _Ops()
