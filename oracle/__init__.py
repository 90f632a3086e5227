"""ctypes wrapper around oracle/microflow_oracle.c -- the CPU parity oracle.

TEST INFRASTRUCTURE ONLY (see the header of microflow_oracle.c).  Importable only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg.  The product package
(microflow_rs_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
ACT = {"none": 0, "relu": 1, "relu6": 3}
PAD = {"same": 0, "valid": 1}
OP_NAMES = {1: "average_pool_2d", 3: "conv_2d", 4: "depthwise_conv_2d", 9: "fully_connected", 22: "reshape", 25: "softmax"}

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")


def build(force=False):
    """Compile the oracle with oracle/Makefile (gcc).  Building the checker is not using it."""
    targets = [_DIR / "libmicroflow_oracle.so", _DIR / "libmicroflow_oracle_fast.so"]
    src = _DIR / "microflow_oracle.c"
    if force or any((not t.exists()) or t.stat().st_mtime < src.stat().st_mtime for t in targets):
        env = dict(os.environ, CC="gcc")
        subprocess.run(["make", "-C", str(_DIR), "-B" if force else "-s", "all"], check=True, env=env,
                       stdout=subprocess.DEVNULL)
    return targets


def _u8(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.int8:
        return a.view(np.uint8)
    if a.dtype == np.uint8:
        return a
    raise TypeError(f"expected int8/uint8, got {a.dtype}")


class _Lib:
    def __init__(self, fast=False):
        build()
        self.lib = C.CDLL(str(_DIR / ("libmicroflow_oracle_fast.so" if fast else "libmicroflow_oracle.so")))
        L = self.lib
        L.mfo_quantize.restype = C.c_int32
        L.mfo_quantize.argtypes = [C.c_float, C.c_float, C.c_int32, C.c_int]
        L.mfo_dequantize.restype = C.c_float
        L.mfo_dequantize.argtypes = [C.c_int32, C.c_float, C.c_int32]
        L.mfo_relu.restype = C.c_int32
        L.mfo_relu.argtypes = [C.c_int32, C.c_int32]
        L.mfo_relu6.restype = C.c_int32
        L.mfo_relu6.argtypes = [C.c_int32, C.c_float, C.c_int32, C.c_int]
        L.mfo_expf.restype = C.c_float
        L.mfo_expf.argtypes = [C.c_float]
        L.mfo_softmax_scalar.restype = C.c_int32
        L.mfo_softmax_scalar.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int32, C.c_int]
        L.mfo_conv_2d.restype = C.c_int
        L.mfo_conv_2d.argtypes = [C.c_int, _u8p, C.c_int, C.c_int, C.c_int, C.c_int32, _u8p, C.c_int, C.c_int, C.c_int, _i32p, C.c_int,
                                  C.c_float, C.c_int32, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_int, _u8p, C.c_int, C.c_int]
        L.mfo_depthwise_conv_2d.restype = C.c_int
        L.mfo_depthwise_conv_2d.argtypes = L.mfo_conv_2d.argtypes
        L.mfo_fully_connected.restype = C.c_int
        L.mfo_fully_connected.argtypes = [C.c_int, _u8p, C.c_int, C.c_int, _u8p, C.c_int, C.c_int32, C.c_float, C.c_int32, C.c_int, _f32p,
                                          C.c_float, _i32p, C.c_int32, _u8p]
        L.mfo_average_pool_2d.restype = C.c_int
        L.mfo_average_pool_2d.argtypes = [C.c_int, _u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int32, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.c_float, C.c_float, _u8p, C.c_int, C.c_int]
        L.mfo_softmax.restype = C.c_int
        L.mfo_softmax.argtypes = [C.c_int, _u8p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int32, _u8p]
        L.mfo_conv_preprocess.restype = None
        L.mfo_conv_preprocess.argtypes = [C.c_float, _f32p, C.c_int, _f32p, C.c_int, _i32p, _i32p, C.c_int, C.c_float, C.c_int, _f32p, _f32p]
        L.mfo_fc_preprocess.restype = None
        L.mfo_fc_preprocess.argtypes = [C.c_int, C.c_float, C.c_int32, C.c_int, _u8p, C.c_int, C.c_int, C.c_float, C.c_int32, C.c_float, _i32p,
                                        C.c_int32, C.c_float, _f32p, C.POINTER(C.c_float), _i32p, C.POINTER(C.c_int32)]
        L.mfo_pool_preprocess.restype = None
        L.mfo_pool_preprocess.argtypes = [C.c_float, C.c_int32, C.c_float, C.c_int32, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.mfo_model_load.restype = C.c_int
        L.mfo_model_load.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.mfo_model_free.restype = None
        L.mfo_model_free.argtypes = [C.c_void_p]
        L.mfo_model_io.restype = C.c_int
        L.mfo_model_io.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int * 4), C.POINTER(C.c_float), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_int), C.POINTER(C.c_int * 4), C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_int)]
        L.mfo_model_num_layers.restype = C.c_int
        L.mfo_model_num_layers.argtypes = [C.c_void_p]
        L.mfo_model_layer_info.restype = C.c_int
        L.mfo_model_layer_info.argtypes = [C.c_void_p, C.c_int, _i32p, _f32p]
        L.mfo_model_layer_consts.restype = C.c_int
        L.mfo_model_layer_consts.argtypes = [C.c_void_p, C.c_int, _f32p, _f32p, _i32p, C.POINTER(C.c_int32), C.c_int]
        L.mfo_predict_quantized.restype = C.c_int
        L.mfo_predict_quantized.argtypes = [C.c_void_p, _u8p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.mfo_predict.restype = C.c_int
        L.mfo_predict.argtypes = [C.c_void_p, _f32p, _f32p, C.c_void_p]
        L.mfo_predict_many_quantized.restype = C.c_int
        L.mfo_predict_many_quantized.argtypes = [C.c_void_p, _u8p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int]
        L.mfo_max_threads.restype = C.c_int


_LIBS = {}


def lib(fast=False):
    if fast not in _LIBS:
        _LIBS[fast] = _Lib(fast).lib
    return _LIBS[fast]


# ---- scalar functions (src/quantize.rs, src/activation.rs) ------------------------------------
def quantize(x, scale, zp, is_u8=False):
    return lib().mfo_quantize(np.float32(x), np.float32(scale), int(zp), int(is_u8))


def dequantize(q, scale, zp):
    return np.float32(lib().mfo_dequantize(int(q), np.float32(scale), int(zp)))


def relu(x, zp):
    return lib().mfo_relu(int(x), int(zp))


def relu6(x, scale, zp, is_u8=False):
    return lib().mfo_relu6(int(x), np.float32(scale), int(zp), int(is_u8))


def expf(x):
    return np.float32(lib().mfo_expf(np.float32(x)))


def softmax_scalar(x, s, scale, zp, is_u8=False):
    return lib().mfo_softmax_scalar(np.float32(x), np.float32(s), np.float32(scale), int(zp), int(is_u8))


# ---- operators (src/ops/*.rs); arrays are NHWC int8/uint8, one sample ---------------------------
def _sign(a, like):
    return a.view(like.dtype)


def conv_2d(x, in_zp, filt, filt_zp, out_scale, out_zp, act, pad, strides, c0, c1, out_hw, depthwise=False):
    """x [H,W,Cin]; filt OHWI [Cout,KH,KW,Cin] (depthwise: [1,KH,KW,Cout]); returns [OH,OW,Cout]."""
    is_u8 = x.dtype == np.uint8
    H, W, Cin = x.shape
    if depthwise:
        _, KH, KW, Cout = filt.shape
    else:
        Cout, KH, KW, _ = filt.shape
    OH, OW = out_hw
    out = np.zeros((OH, OW, Cout), np.uint8)
    fz = np.ascontiguousarray(np.atleast_1d(filt_zp), np.int32)
    c0 = np.ascontiguousarray(c0, np.float32)
    c1 = np.ascontiguousarray(np.atleast_1d(c1), np.float32)
    fn = lib().mfo_depthwise_conv_2d if depthwise else lib().mfo_conv_2d
    rc = fn(int(is_u8), _u8(x), H, W, Cin, int(in_zp), _u8(filt), Cout, KH, KW, fz, len(fz), np.float32(out_scale), int(out_zp),
            ACT[act], PAD[pad], strides[0], strides[1], c0, c1, len(c1), out, OH, OW)
    if rc:
        raise RuntimeError(f"oracle conv rc={rc}")
    return _sign(out, x)


def depthwise_conv_2d(*a, **k):
    return conv_2d(*a, depthwise=True, **k)


def fully_connected(x, w_nk, w_zp, out_scale, out_zp, act, c0, c1, c2, c3):
    """x [R,K]; w_nk = TFLite layout [N,K]; returns [R,N]."""
    is_u8 = x.dtype == np.uint8
    R, K = x.shape
    N = w_nk.shape[0]
    out = np.zeros((R, N), np.uint8)
    rc = lib().mfo_fully_connected(int(is_u8), _u8(x), R, K, _u8(w_nk), N, int(w_zp), np.float32(out_scale), int(out_zp), ACT[act],
                                   np.ascontiguousarray(c0, np.float32), np.float32(c1), np.ascontiguousarray(c2, np.int32), int(c3), out)
    if rc:
        raise RuntimeError(f"oracle fc rc={rc}")
    return _sign(out, x)


def average_pool_2d(x, filter_hw, out_scale, out_zp, act, pad, strides, c0, c1, out_hw):
    is_u8 = x.dtype == np.uint8
    H, W, Cc = x.shape
    OH, OW = out_hw
    out = np.zeros((OH, OW, Cc), np.uint8)
    rc = lib().mfo_average_pool_2d(int(is_u8), _u8(x), H, W, Cc, filter_hw[0], filter_hw[1], np.float32(out_scale), int(out_zp), ACT[act],
                                   PAD[pad], strides[0], strides[1], np.float32(c0), np.float32(c1), out, OH, OW)
    if rc:
        raise RuntimeError(f"oracle pool rc={rc}")
    return _sign(out, x)


def softmax(x, in_scale, out_scale, out_zp):
    is_u8 = x.dtype == np.uint8
    rows, cols = x.shape
    out = np.zeros((rows, cols), np.uint8)
    lib().mfo_softmax(int(is_u8), _u8(x), rows, cols, np.float32(in_scale), np.float32(out_scale), int(out_zp), out)
    return _sign(out, x)


# ---- pre-processing (microflow-macros/src/ops/*.rs `preprocess`) ------------------------------
def conv_preprocess(in_scale, w_scale, b_scale, bias, b_zp, out_scale, n_out):
    w_scale = np.ascontiguousarray(np.atleast_1d(w_scale), np.float32)
    b_scale = np.ascontiguousarray(np.atleast_1d(b_scale), np.float32)
    bias = np.ascontiguousarray(bias, np.int32)
    b_zp = np.ascontiguousarray(np.atleast_1d(b_zp), np.int32)
    c0 = np.zeros(n_out, np.float32)
    c1 = np.zeros(len(w_scale), np.float32)
    lib().mfo_conv_preprocess(np.float32(in_scale), w_scale, len(w_scale), b_scale, len(b_scale), bias, b_zp, len(b_zp), np.float32(out_scale),
                              n_out, c0, c1)
    return c0, c1


def fc_preprocess(in_scale, in_zp, shape1, w_nk, w_scale, w_zp, b_scale, bias, b_zp, out_scale):
    is_u8 = w_nk.dtype == np.uint8
    N, K = w_nk.shape
    c0 = np.zeros(N, np.float32)
    c2 = np.zeros(N, np.int32)
    c1 = C.c_float()
    c3 = C.c_int32()
    lib().mfo_fc_preprocess(int(is_u8), np.float32(in_scale), int(in_zp), int(shape1), _u8(w_nk), N, K, np.float32(w_scale), int(w_zp),
                            np.float32(b_scale), np.ascontiguousarray(bias, np.int32), int(b_zp), np.float32(out_scale), c0, C.byref(c1), c2,
                            C.byref(c3))
    return c0, np.float32(c1.value), c2, c3.value


def pool_preprocess(in_scale, in_zp, out_scale, out_zp):
    c0, c1 = C.c_float(), C.c_float()
    lib().mfo_pool_preprocess(np.float32(in_scale), int(in_zp), np.float32(out_scale), int(out_zp), C.byref(c0), C.byref(c1))
    return np.float32(c0.value), np.float32(c1.value)


# ---- whole models (microflow-macros/src/lib.rs generated predict*) ------------------------------
class Model:
    """The oracle's equivalent of `#[model("x.tflite")] struct M;`."""

    def __init__(self, path_or_bytes, fast=False):
        data = Path(path_or_bytes).read_bytes() if isinstance(path_or_bytes, (str, os.PathLike)) else bytes(path_or_bytes)
        self._lib = lib(fast)
        self._h = C.c_void_p()
        rc = self._lib.mfo_model_load(data, len(data), C.byref(self._h))
        if rc:
            raise RuntimeError(f"oracle: mfo_model_load rc={rc}")
        ir, orr, isu8 = C.c_int(), C.c_int(), C.c_int()
        idims, odims = (C.c_int * 4)(), (C.c_int * 4)()
        isc, osc = C.c_float(), C.c_float()
        izp, ozp = C.c_int32(), C.c_int32()
        self._lib.mfo_model_io(self._h, C.byref(ir), C.byref(idims), C.byref(isc), C.byref(izp), C.byref(orr), C.byref(odims), C.byref(osc),
                               C.byref(ozp), C.byref(isu8))
        self.in_shape = tuple(idims[: ir.value])
        self.out_shape = tuple(odims[: orr.value])
        self.in_scale, self.in_zp = np.float32(isc.value), izp.value
        self.out_scale, self.out_zp = np.float32(osc.value), ozp.value
        self.dtype = np.uint8 if isu8.value else np.int8
        self.in_elems = int(np.prod(self.in_shape))
        self.out_elems = int(np.prod(self.out_shape))
        self.layers = []
        for i in range(self._lib.mfo_model_num_layers(self._h)):
            info = np.zeros(24, np.int32)
            sc = np.zeros(2, np.float32)
            self._lib.mfo_model_layer_info(self._h, i, info, sc)
            d = dict(op=OP_NAMES.get(int(info[0]), str(info[0])), out_elems=int(info[1]), in_elems=int(info[2]), act=int(info[3]),
                     pad=int(info[4]), strides=(int(info[5]), int(info[6])), KH=int(info[7]), KW=int(info[8]), Cout=int(info[9]),
                     in_zp=int(info[10]), out_zp=int(info[11]), n_c1=int(info[12]), n_wq=int(info[13]),
                     out_shape=tuple(int(v) for v in info[15:15 + info[14]]), in_shape=tuple(int(v) for v in info[20:20 + info[19]]),
                     in_scale=sc[0], out_scale=sc[1])
            self.layers.append(d)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.mfo_model_free(self._h)
            self._h = None

    def layer_consts(self, i):
        n = max(self.layers[i]["Cout"], 1)
        c0, c1, c2 = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.int32)
        c3 = C.c_int32()
        self._lib.mfo_model_layer_consts(self._h, i, c0, c1, c2, C.byref(c3), n)
        return c0, c1, c2, c3.value

    def predict_quantized(self, x, return_q=False, trace=False):
        """x: int8/uint8 array of in_shape.  Returns f32 out (and the quantized output / per-layer outputs)."""
        xq = _u8(np.asarray(x).reshape(-1))
        assert xq.size == self.in_elems
        out = np.zeros(self.out_elems, np.float32)
        outq = np.zeros(self.out_elems, np.uint8)
        lay = None
        ptrs = None
        if trace:
            lay = [np.zeros(L["out_elems"], np.uint8) for L in self.layers]
            ptrs = (C.c_void_p * len(lay))(*[a.ctypes.data for a in lay])
        rc = self._lib.mfo_predict_quantized(self._h, xq, out.ctypes.data, outq.ctypes.data, ptrs)
        if rc:
            raise RuntimeError(f"oracle predict rc={rc}")
        res = [out.reshape(self.out_shape)]
        if return_q:
            res.append(outq.view(self.dtype).reshape(self.out_shape))
        if trace:
            res.append([a.view(self.dtype).reshape(L["out_shape"]) for a, L in zip(lay, self.layers)])
        return res[0] if len(res) == 1 else tuple(res)

    def predict(self, x, return_q=False):
        xf = np.ascontiguousarray(np.asarray(x, np.float32).reshape(-1))
        assert xf.size == self.in_elems
        out = np.zeros(self.out_elems, np.float32)
        outq = np.zeros(self.out_elems, np.uint8)
        rc = self._lib.mfo_predict(self._h, xf, out, outq.ctypes.data)
        if rc:
            raise RuntimeError(f"oracle predict rc={rc}")
        if return_q:
            return out.reshape(self.out_shape), outq.view(self.dtype).reshape(self.out_shape)
        return out.reshape(self.out_shape)

    def predict_many_quantized(self, xs, threads=1):
        """xs [n, *in_shape[1:]] (in_shape[0] == 1 per sample).  Returns (f32 [n, out_elems], q [n, out_elems])."""
        xs = _u8(np.asarray(xs))
        n = xs.size // self.in_elems
        out = np.zeros((n, self.out_elems), np.float32)
        outq = np.zeros((n, self.out_elems), np.uint8)
        rc = self._lib.mfo_predict_many_quantized(self._h, xs.reshape(-1), n, out.ctypes.data, outq.ctypes.data, int(threads))
        if rc:
            raise RuntimeError(f"oracle predict_many rc={rc}")
        return out, outq.view(self.dtype)


def max_threads():
    return lib().mfo_max_threads()
