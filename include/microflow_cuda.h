/*
 * microflow_cuda.h -- C ABI of libmicroflow_cuda.so, the B200 (sm_100a) backend for MicroFlow's
 * quantized op-kernel hot path.
 *
 * The reference (matteocarnelos/microflow-rs) has NO FFI boundary: its "operator API" is the set of
 * generic Rust functions that the `#[model("x.tflite")]` proc-macro emits calls to, plus the generated
 * `predict` / `predict_quantized` associated functions.  This header is what a `microflow-cuda-sys`
 * crate would bind (see INTEGRATION.md for the Rust-side stub); each entry point cites the reference
 * interface it replaces.  Plain pointers and sizes only -- no torch / CUDA types in the signatures
 * (streams are passed as `void*` = cudaStream_t).
 *
 * Tensor layout: NHWC, row-major, one byte per element (int8 or uint8 -- `T: Quantized`,
 * src/quantize.rs:6-7).  Sample s of a batched call starts at byte s * in_elems.  The reference's
 * nalgebra buffers are column-major ([batch][col][row][chan], src/buffer.rs:10-16): a model created
 * with mf_options.layout = MF_LAYOUT_NALGEBRA takes and returns HOST buffers in exactly that memory
 * order (the Rust shim then passes `input.as_ptr()` with no copy); the transposition runs on the
 * device.  Device-resident entry points (mf_predict_many_device, mf_op_run_device) are NHWC only.
 *
 * Errors: the reference reports every error at Rust compile time (abort_call_site!) and is infallible
 * at run time.  Here every function returns an mf_status; mf_last_error() gives the text a shim would
 * put in the compile error / panic message.
 */
#ifndef MICROFLOW_CUDA_H
#define MICROFLOW_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library itself is built with -fvisibility=hidden */
#endif

#define MF_ABI_VERSION 3

typedef enum mf_status {
    MF_OK = 0,
    MF_ERR_FILE = 1,                   /* microflow-macros/src/lib.rs:50-55 "couldn't find '{}'" */
    MF_ERR_INVALID_MODEL = 2,          /* lib.rs:56-58 "invalid model" */
    MF_ERR_UNSUPPORTED_TYPE = 3,       /* lib.rs:71-78, ops/conv_2d.rs:39-46 (INT8/UINT8 only) */
    MF_ERR_UNSUPPORTED_RANK = 4,       /* lib.rs:79-96 (ranks 2 and 4 only) */
    MF_ERR_UNSUPPORTED_OP = 5,         /* lib.rs:148 */
    MF_ERR_UNSUPPORTED_ACTIVATION = 6, /* microflow-macros/src/activation.rs:31-34 */
    MF_ERR_UNSUPPORTED_SHAPE = 7,      /* ops/reshape.rs:47-55 and shape mismatches rustc would reject */
    MF_ERR_VIEW_OUT_OF_BOUNDS = 8,     /* src/tensor.rs:222 (VALID view indexing out of bounds panics) */
    MF_ERR_INVALID_ARG = 9,
    MF_ERR_NO_DEVICE = 10,             /* CUDA extension / device missing: the product has NO CPU fallback */
    MF_ERR_CUDA = 11,
    MF_ERR_NONFINITE_CONSTANT = 12     /* a requantization constant is NaN/Inf (scale 0): rejected at load */
} mf_status;

/* tflite.fbs TensorType values the reference accepts */
#define MF_DTYPE_U8 3
#define MF_DTYPE_I8 9

/* tflite.fbs:548-560 */
#define MF_PAD_SAME 0
#define MF_PAD_VALID 1
#define MF_ACT_NONE 0
#define MF_ACT_RELU 1
#define MF_ACT_RELU6 3

/* BuiltinOperator codes the reference supports (lib.rs:138-148) */
#define MF_OP_AVERAGE_POOL_2D 1
#define MF_OP_CONV_2D 3
#define MF_OP_DEPTHWISE_CONV_2D 4
#define MF_OP_FULLY_CONNECTED 9
#define MF_OP_RESHAPE 22
#define MF_OP_SOFTMAX 25

/* mf_options.flags */
#define MF_FLAG_HOST_ONLY 1u      /* parse + preprocess only (what the proc-macro does); no device needed */
#define MF_FLAG_FORCE_GENERIC 2u  /* run every layer on the generic direct kernels (cross-check path) */
#define MF_FLAG_NO_TENSOR_CORE 4u /* keep SIMT fast kernels but never pick a tcgen05 kernel */

/* mf_options.layout: memory order of the HOST input / output buffers of predict* */
#define MF_LAYOUT_NHWC 0u     /* [sample][row][col][chan], row-major (default) */
#define MF_LAYOUT_NALGEBRA 1u /* the reference's own buffers: Buffer4D = [SMatrix<[T; CH], R, C>; B] and Buffer2D = SMatrix<T, R, C>,
                                 both column-major (src/buffer.rs:5-16): [sample][col][row][chan] */

#define MF_MAX_DEVICES 16

typedef struct mf_options {
    uint32_t struct_size; /* = sizeof(mf_options) of the caller; a shorter (older) struct leaves the missing fields at 0 */
    int32_t device;       /* CUDA device ordinal; -1 = current device (ignored when n_devices > 0) */
    uint32_t chunk;       /* samples per internal chunk of predict_many (0 = default) */
    uint32_t flags;       /* MF_FLAG_* */
    uint32_t layout;      /* MF_LAYOUT_* (ABI version 2) */
    /* ABI version 3: the GPUs of the box this model runs on.  n_devices == 0: the single `device` above.  n_devices == -1: every
     * visible device.  n_devices >= 1: devices[0 .. n_devices).  With more than one device the model is replicated: the static
     * weight blob is uploaded to devices[0] and copied to the others by ONE broadcast at create (NCCL when libnccl.so.2 can be
     * loaded, else cudaMemcpyPeer), and mf_predict_many* shard the n independent samples into contiguous ranges
     * [r*ceil(n/G), min(n, (r+1)*ceil(n/G))), one per device, each driven by its own host thread -- no collective and no
     * exchange of any kind on the inference path (the reference handles one sample per call, src/ops/conv_2d.rs:40). */
    int32_t n_devices;
    int32_t devices[MF_MAX_DEVICES];
} mf_options;

typedef struct mf_tensor_info {
    int32_t rank;    /* 2 or 4 (1-D shapes are reported as [1,n], lib.rs:68-70) */
    int32_t dims[4]; /* unused trailing dims = 1 */
    int32_t dtype;   /* MF_DTYPE_I8 / MF_DTYPE_U8 */
    float scale;
    int32_t zero_point;
    uint64_t elems;  /* per sample */
} mf_tensor_info;

typedef struct mf_layer_info {
    int32_t op;                 /* MF_OP_* */
    int32_t in_dims[4], out_dims[4];
    int32_t in_rank, out_rank;
    int32_t kh, kw, stride_h, stride_w, padding, activation;
    int32_t in_zero_point, out_zero_point;
    float in_scale, out_scale;
    int32_t act_lo, act_hi;     /* fused clamp after saturation */
    int32_t n_c0, n_c1;         /* lengths of the constant vectors (per-channel vs per-tensor) */
    uint64_t macs;              /* multiply-accumulates per sample */
    uint64_t bytes;             /* algorithmic bytes per sample: input + output (+ weights/constants once) */
    uint64_t weight_bytes;
    char kernel[48];            /* name of the CUDA kernel the engine selected ("" in host-only mode) */
} mf_layer_info;

typedef struct mf_model mf_model;

/* ---- library ------------------------------------------------------------------------------- */
int mf_abi_version(void);
const char *mf_last_error(void);       /* thread-local text of the last failure */
const char *mf_status_string(int status);
int mf_device_count(int *count);       /* MF_ERR_NO_DEVICE when there is no usable GPU */

/* ---- model = what `#[model("path")]` builds at Rust compile time (lib.rs:46-208) ----------- */
int mf_model_create_from_tflite(const void *buf, size_t len, const mf_options *opt, mf_model **out);
int mf_model_create_from_file(const char *path, const mf_options *opt, mf_model **out);
void mf_model_destroy(mf_model *m);
int mf_model_io_info(const mf_model *m, mf_tensor_info *in, mf_tensor_info *out);
int mf_model_num_layers(const mf_model *m);
int mf_model_layer_info(const mf_model *m, int layer, mf_layer_info *out);
/* pre-processed constants of a layer (the tuples the macro emits, ops/<op>.rs `preprocess`):
 * conv/dw: c0[n_c0], c1[n_c1]; fc: c0[n], c1[0], c2[n], *c3; pool: c0[0], c1[0].  `cap` = capacity of each array. */
int mf_model_layer_constants(const mf_model *m, int layer, float *c0, float *c1, int32_t *c2, int32_t *c3, int cap);
/* text dump of the graph + constants: the equivalent of target/microflow-expansion.rs (lib.rs:205) */
int mf_model_dump(const mf_model *m, const char *path);

/* ---- generated API: predict / predict_quantized (lib.rs:188-196); host buffers, one sample -- */
int mf_predict(mf_model *m, const float *in_f32, float *out_f32);
int mf_predict_quantized(mf_model *m, const void *in_q, float *out_f32);

/* ---- batched: n independent samples (new; the reference handles one sample per call) -------- */
/* host buffers (pinned or pageable); H2D / compute / D2H are pipelined in chunks inside the call */
int mf_predict_many(mf_model *m, const float *in_f32, size_t n, float *out_f32);
int mf_predict_many_quantized(mf_model *m, const void *in_q, size_t n, float *out_f32);
/* same, but returns as soon as the copies and kernels are enqueued; the host buffers must be PINNED (mf_host_alloc) and stay
 * valid until mf_model_synchronize().  Back-to-back calls pipeline: the H2D of call k+1 overlaps the kernels of call k.
 * (The blocking calls above serve requests of <= 64 samples by replaying a captured CUDA graph) */
int mf_predict_many_quantized_async(mf_model *m, const void *in_q, size_t n, float *out_f32);
/* strict-parity variant: final quantized output (out_q, out_elems bytes/sample) and, optionally, the input of
 * the trailing softmax ("logits", may be NULL) */
int mf_predict_many_logits(mf_model *m, const void *in_q, size_t n, void *out_q, void *logits_q);
/* device-resident buffers (NHWC); asynchronous on `stream` (cudaStream_t, NULL = the model's own stream): everything the call
 * enqueues is ordered after the work already on `stream` and before work enqueued on it afterwards (large chunks run on two
 * internal streams that are forked from / joined to `stream` by events).  d_out_f32 and d_out_q may each be NULL.  The model's
 * workspaces are shared with the host-path entry points: synchronize (mf_model_synchronize) between un-waited
 * mf_predict_many_quantized_async calls and a device-path call on a different stream. */
int mf_predict_many_device(mf_model *m, const void *d_in_q, size_t n, float *d_out_f32, void *d_out_q, void *stream);
/* multi-device models: the same call on the replica of devices[index] (buffers and stream must belong to that device);
 * mf_predict_many_device itself addresses devices[0] */
int mf_predict_many_device_on(mf_model *m, int index, const void *d_in_q, size_t n, float *d_out_f32, void *d_out_q, void *stream);
/* every layer's quantized output for n samples, into host buffers layer_outs[i] (n * out_elems(i) bytes; NULL = skip) */
int mf_predict_trace(mf_model *m, const void *in_q, size_t n, void *const *layer_outs);
int mf_model_synchronize(mf_model *m);

/* per-layer device timing (CUDA events on the launching stream), summed over every mf_predict_many_device call since
 * profiling was enabled or mf_model_layer_times_ms() was last read (reading synchronizes and resets) */
int mf_model_set_profiling(mf_model *m, int enabled);
int mf_model_layer_times_ms(mf_model *m, float *ms, int cap);
/* number of kernels launched by this model since creation (for bench.py's gpu_launches) */
int mf_model_launch_count(const mf_model *m, uint64_t *count);
/* name of the kernel the most recent predict* / trace call actually launched for `layer` ("" = none yet, or the layer ran inside
 * the previous layer's fused launch); mf_layer_info.kernel is the engine's choice for large batches, this is what ran */
const char *mf_model_layer_launched(const mf_model *m, int layer);
/* devices this model runs on (ABI 3): returns the count, fills devices[0 .. min(count, cap)) */
int mf_model_devices(const mf_model *m, int32_t *devices, int cap);
/* how the weight blob reached devices[1..]: "none" (one device), "nccl" or "memcpy_peer" */
const char *mf_model_weight_broadcast(const mf_model *m);

/* ---- static weights blob (multi-GPU init: one broadcast of d_ptr[0..bytes) from rank 0) ------ */
int mf_model_blob(const mf_model *m, void **d_ptr, size_t *bytes);

/* ---- host staging of the reference's sample formats: the step before the path (samples/person.bmp -> samples/features/person_detect.rs) */
/* An uncompressed 8-bit BMP with the identity gray palette -> `height * width` int8 features, top row first (the pixel bytes read as
 * int8; BMP rows are stored bottom-up).  out == NULL only reports the dimensions.  The speech frontend (the .wav samples ->
 * samples/features/speech.rs) is TensorFlow Lite Micro's audio frontend, which the reference does not contain: not provided. */
int mf_features_from_bmp_gray8(const void *bmp, size_t len, void *out, size_t cap, int32_t *height, int32_t *width);
/* n BMP images of the model's input size -> staged into one NHWC batch -> mf_predict_many_quantized */
int mf_predict_many_bmp(mf_model *m, const void *const *bmps, const size_t *lens, size_t n, float *out_f32);

/* ---- pinned host memory for the H2D/D2H legs ------------------------------------------------- */
int mf_host_alloc(void **p, size_t bytes);
int mf_host_free(void *p);

/* ---- per-operator hooks: one call = one reference op on `batch` independent samples ---------- *
 * They mirror the calls the macro emits so the reference's op KATs can drive a single
 * CUDA kernel.  Host pointers.  `impl`: 0 = auto (fast kernel if eligible), 1 = generic kernel, 2 = require fast. */
typedef struct mf_conv_desc {          /* microflow::ops::conv_2d / depthwise_conv_2d (src/ops/conv_2d.rs:28-49) */
    int32_t dtype;                     /* MF_DTYPE_I8 / MF_DTYPE_U8 */
    int32_t depthwise;                 /* 0: filters OHWI [cout][kh][kw][cin]; 1: weights [1][kh][kw][cout] */
    int32_t in_h, in_w, in_c;
    int32_t out_h, out_w, out_c;
    int32_t kh, kw, stride_h, stride_w, padding, activation;
    int32_t in_zero_point;
    float out_scale;
    int32_t out_zero_point;
    const void *filters;
    const int32_t *filter_zero_points; int32_t n_filter_zero_points;
    const float *c0;                   /* constants.0, length out_c */
    const float *c1; int32_t n_c1;     /* constants.1, per-channel or single */
    int32_t impl;
} mf_conv_desc;
int mf_op_conv_2d(const mf_conv_desc *d, const void *in, void *out, size_t batch);

/* a straight-line sequence of conv_2d / depthwise_conv_2d operators (the chain the macro emits, microflow-macros/src/lib.rs:198-201,
 * restricted to convolutions): op k consumes op k-1's output.  fuse = 0: one kernel per operator.  fuse = 1: the sequence must be
 * n x [depthwise 3x3 / stride 1 / SAME, 128 channels] -> [1x1 conv, 128 -> 128] on a feature map of <= 128 pixels and runs as ONE
 * launch of the fused low-resolution stage kernel (activations stay in shared memory); anything else: MF_ERR_UNSUPPORTED_SHAPE. */
int mf_op_conv_chain(const mf_conv_desc *ops, int n_ops, const void *in, void *out, size_t batch, int fuse);

/* persistent form of the same operator: plan once (weights/constants uploaded, kernel selected), then run on
 * DEVICE-resident NHWC buffers (32-byte aligned, as cudaMalloc returns them), asynchronously on `stream` (cudaStream_t).  Used by pipelines and by bench.py's
 * BASELINE config 5 (synthetic 224x224x128->128 Conv2D roofline). */
typedef struct mf_op mf_op;
int mf_op_conv_2d_create(const mf_conv_desc *d, mf_op **out);
int mf_op_run_device(mf_op *op, const void *d_in, void *d_out, size_t batch, void *stream);
const char *mf_op_kernel_name(const mf_op *op);
void mf_op_destroy(mf_op *op);

typedef struct mf_fc_desc {            /* microflow::ops::fully_connected (src/ops/fully_connected.rs:24-41) */
    int32_t dtype;
    int32_t in_features, out_features; /* K, N; input is [batch][K] */
    const void *weights_nk;            /* TFLite byte layout [N][K] (the reference's W[k][j] = bytes[j*K+k]) */
    int32_t weight_zero_point;
    float out_scale; int32_t out_zero_point; int32_t activation;
    const float *c0; float c1; const int32_t *c2; int32_t c3;
    int32_t impl;
} mf_fc_desc;
int mf_op_fully_connected(const mf_fc_desc *d, const void *in, void *out, size_t batch);

typedef struct mf_pool_desc {          /* microflow::ops::average_pool_2d (src/ops/average_pool_2d.rs:29-45) */
    int32_t dtype;
    int32_t in_h, in_w, chans, out_h, out_w;
    int32_t filter_h, filter_w, stride_h, stride_w, padding, activation;
    float out_scale; int32_t out_zero_point;
    float c0, c1;
    int32_t impl;
} mf_pool_desc;
int mf_op_average_pool_2d(const mf_pool_desc *d, const void *in, void *out, size_t batch);

/* microflow::ops::softmax (src/ops/softmax.rs:15-27): over the whole rows x cols buffer of each sample */
int mf_op_softmax(int32_t dtype, int32_t rows, int32_t cols, float in_scale, float out_scale, int32_t out_zero_point,
                  const void *in, void *out, size_t batch);
/* MF_LAYOUT_NALGEBRA <-> MF_LAYOUT_NHWC conversion of `batch` matrices of rows x cols cells of elem_bytes each (host buffers; the
 * device kernel predict* runs when a model is created with mf_options.layout = MF_LAYOUT_NALGEBRA).  Column-major storage is
 * nalgebra's (src/buffer.rs:5-16).  to_nalgebra = 0: [cols][rows][cell] -> [rows][cols][cell]; 1: the opposite direction. */
int mf_op_layout_transpose(const void *in, void *out, size_t batch, int32_t rows, int32_t cols, int32_t elem_bytes, int32_t to_nalgebra);
/* Tensor::quantize / dequantize (src/tensor.rs:80-92, :246-262; src/quantize.rs:16-29) */
int mf_op_quantize(int32_t dtype, float scale, int32_t zero_point, const float *in, void *out, size_t n);
int mf_op_dequantize(int32_t dtype, float scale, int32_t zero_point, const void *in, float *out, size_t n);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* MICROFLOW_CUDA_H */
