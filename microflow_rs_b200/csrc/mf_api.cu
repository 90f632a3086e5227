// mf_api.cu -- extern "C" ABI of libmicroflow_cuda.so (include/microflow_cuda.h).
//
// The product path has NO CPU fallback: without a usable CUDA device every compute entry point fails with
// MF_ERR_NO_DEVICE (only MF_FLAG_HOST_ONLY models -- parse + preprocess, i.e. the proc-macro's job -- work).
#include <cuda_runtime.h>

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/microflow_cuda.h"
#include "mf_engine.h"
#include "mf_fused.h"

using namespace mf;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
#define MF_CUDA(expr)                                                                                         \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess) return fail(MF_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

struct Slot {
    cudaStream_t stream = nullptr;
    uint8_t *act[2] = {nullptr, nullptr};
    uint8_t *in_q = nullptr;
    float *in_f32 = nullptr;
    float *out_f32 = nullptr;
    uint8_t *out_q = nullptr;
    uint8_t *logits = nullptr;
    uint8_t *in_t = nullptr;      // MF_LAYOUT_NALGEBRA only: transposed copies of the input / outputs
    uint8_t *out_t = nullptr;
};

// One host thread per device of a multi-device model: it owns every CUDA call made for its replica (enqueueing a step is ~60
// driver calls; eight devices fed from one thread would be launch-bound).  Tasks run in FIFO order.
class Worker {
  public:
    Worker() : th_([this] { loop(); }) {}
    ~Worker() {
        {
            std::lock_guard<std::mutex> l(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        th_.join();
    }
    void post(std::function<void()> f) {
        {
            std::lock_guard<std::mutex> l(mu_);
            q_.push_back(std::move(f));
            ++pending_;
        }
        cv_.notify_all();
    }
    void drain() {
        std::unique_lock<std::mutex> l(mu_);
        done_.wait(l, [this] { return pending_ == 0; });
    }

  private:
    void loop() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [this] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                f = std::move(q_.front());
                q_.pop_front();
            }
            f();
            {
                std::lock_guard<std::mutex> l(mu_);
                --pending_;
            }
            done_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::deque<std::function<void()>> q_;
    size_t pending_ = 0;
    bool stop_ = false;
    std::thread th_;
};

// a fused low-resolution chain (mf_fused.h): layers [first .. last] = n_pairs x (depthwise 3x3 s1, pointwise 1x1) in one launch
struct Chain {
    int first = -1, last = -1;
    FusedChainPlan plan;
    size_t o_wimg = SIZE_MAX, o_consts = SIZE_MAX;
};

}  // namespace

struct mf_model {
    // ---- multi-device model (mf_options.n_devices > 1): this object only routes; replicas[r] is a complete single-device model on
    // devices[r], workers[r] the host thread that drives it.  Everything below `replicas` is unused in a routing object.
    std::vector<mf_model *> replicas;
    std::vector<std::unique_ptr<Worker>> workers;
    const char *bcast = "none";             // how the weight blob reached replicas[1..]
    std::mutex group_mu;                    // serialises group calls and guards the deferred error of asynchronous ones
    int deferred_rc = 0;
    std::string deferred_err;
    std::vector<const char *> launched;     // per layer: kernel the most recent predict*/trace call launched ("" = none)
    ModelSpec spec;
    std::vector<LayerExec> layers;
    uint32_t flags = 0;
    uint32_t layout = MF_LAYOUT_NHWC;
    // MF_LAYOUT_NALGEBRA: per sample the host tensors are `mats` column-major R x C matrices of `ch`-element cells
    struct Geo { int mats = 1, R = 1, C = 1, ch = 1; bool transpose = false; } geo_in, geo_out;
    bool host_only = false;
    int device = 0, num_sms = 148;
    uint8_t *d_blob = nullptr;
    size_t blob_bytes = 0;
    size_t chunk = 0;
    Slot slot[2];
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;   // (layers + 1) per profiled chunk
    size_t prof_chunks = 0;
    uint64_t launches = 0;
    int softmax_tail = -1;                  // index of a trailing softmax layer (its input = "logits")
    // classifier tail run as one launch (tail_fused_kernel): layers [tail_first .. tail_last] = global average pool, 1x1 conv,
    // reshapes, softmax.  Per-layer tracing runs the layers one by one instead.
    int tail_first = -1, tail_conv = -1, tail_last = -1;
    TailArgs tail;
    // fully_connected (fc_warp_kernel) -> reshape* -> softmax (last layer) as one launch: layers [fc_tail_first .. fc_tail_last]
    int fc_tail_first = -1, fc_tail_last = -1;
    // fused low-resolution chains (mf_fused.h): layers [first .. last] = n_pairs x (depthwise 3x3 s1, pointwise 1x1) in one launch
    std::vector<Chain> chains;
    size_t slot_rr = 0;                     // round-robin position of the host-path stream slots
    cudaEvent_t split_ev[3] = {nullptr, nullptr, nullptr};   // MF_SPLIT=2: fork / join events of the two half-batch streams
    // Small host-path calls (n <= kGraphMaxN, the reference's one-sample predict() above all) replay a captured CUDA graph:
    // H2D from a pinned staging buffer, every layer, D2H into pinned staging -- one graph launch instead of ~30 stream
    // operations.  One executable graph per (n, input kind, outputs wanted); capture happens on first use.
    struct SmallGraph {
        size_t n = 0;
        int kind = 0;                       // bit 0: f32 input, bit 1: f32 output, bit 2: quantized output, bit 3: logits
        cudaGraphExec_t exec = nullptr;
        uint64_t kernels = 0;               // kernel nodes per replay (for mf_model_launch_count)
    };
    std::vector<SmallGraph> graphs;
    bool graphs_disabled = false;           // set when capture / instantiation fails once: the stream path is used from then on
    uint8_t *h_stage_in = nullptr;          // pinned staging, kGraphMaxN samples each
    float *h_stage_out_f32 = nullptr;
    uint8_t *h_stage_out_q = nullptr, *h_stage_logits = nullptr;
    std::mutex mu;
};

namespace {

int check_device(int *count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        if (count) *count = 0;
        return fail(MF_ERR_NO_DEVICE, std::string("no usable CUDA device (") + cudaGetErrorString(e) +
                                          "): libmicroflow_cuda has no CPU fallback; only MF_FLAG_HOST_ONLY models can be created");
    }
    if (count) *count = n;
    return MF_OK;
}

void free_slot(Slot &s) {
    if (s.stream) cudaStreamDestroy(s.stream);
    for (auto *p : {(void *)s.act[0], (void *)s.act[1], (void *)s.in_q, (void *)s.in_f32, (void *)s.out_f32, (void *)s.out_q, (void *)s.logits, (void *)s.in_t,
                    (void *)s.out_t})
        if (p) cudaFree(p);
    s = Slot{};
}

int alloc_slot(mf_model *m, Slot &s) {
    const size_t c = m->chunk;
    MF_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    MF_CUDA(cudaMalloc(&s.act[0], c * m->spec.max_elems));
    MF_CUDA(cudaMalloc(&s.act[1], c * m->spec.max_elems));
    MF_CUDA(cudaMalloc(&s.in_q, c * m->spec.in_elems));
    // in_f32 (4 bytes per input element: 1.2 GB per slot for person_detect at the default chunk) is only needed by the f32
    // predict() entry points: allocated on their first use (ensure_f32_staging)
    MF_CUDA(cudaMalloc(&s.out_f32, c * m->spec.out_elems * sizeof(float)));
    MF_CUDA(cudaMalloc(&s.out_q, c * m->spec.out_elems));
    {   // "logits" = the input of the trailing softmax (a few bytes per sample), not a whole activation tensor
        const size_t le = m->softmax_tail >= 0 ? m->layers[(size_t)m->softmax_tail].spec.in_elems : 0;
        MF_CUDA(cudaMalloc(&s.logits, c * (le ? le : 1)));
    }
    if (m->geo_in.transpose) MF_CUDA(cudaMalloc(&s.in_t, c * m->spec.in_elems));
    if (m->geo_out.transpose) MF_CUDA(cudaMalloc(&s.out_t, c * m->spec.out_elems * sizeof(float)));
    return MF_OK;
}

int ensure_f32_staging(mf_model *m, Slot &s) {
    if (!s.in_f32) MF_CUDA(cudaMalloc(&s.in_f32, m->chunk * m->spec.in_elems * sizeof(float)));
    return MF_OK;
}

// MF_LAYOUT_NALGEBRA, input side: s.in_q holds the quantized input in the reference's column-major order -> NHWC in s.in_t
int stage_input_layout(mf_model *m, Slot &s, size_t n, cudaStream_t st, const uint8_t **d_in) {
    *d_in = s.in_q;
    if (!m->geo_in.transpose) return MF_OK;
    const auto &g = m->geo_in;
    cudaError_t e = launch_layout_transpose(s.in_q, s.in_t, (long long)n * g.mats, g.R, g.C, g.ch, st);
    if (e != cudaSuccess) return fail(MF_ERR_CUDA, std::string("layout_transpose_kernel launch failed: ") + cudaGetErrorString(e));
    m->launches += 1;
    *d_in = s.in_t;
    return MF_OK;
}
// output side: NHWC result in `src` (elements of `esz` bytes) -> column-major copy in s.out_t; *d_out is what the D2H copy reads
int stage_output_layout(mf_model *m, Slot &s, size_t n, const void *src, size_t esz, cudaStream_t st, const void **d_out) {
    *d_out = src;
    if (!m->geo_out.transpose) return MF_OK;
    const auto &g = m->geo_out;
    cudaError_t e = launch_layout_transpose((const uint8_t *)src, s.out_t, (long long)n * g.mats, g.C, g.R, g.ch * (int)esz, st);   // roles of R and C swapped
    if (e != cudaSuccess) return fail(MF_ERR_CUDA, std::string("layout_transpose_kernel launch failed: ") + cudaGetErrorString(e));
    m->launches += 1;
    *d_out = s.out_t;
    return MF_OK;
}

// Runs all layers for `n` <= chunk samples.  d_in: quantized input.  Outputs (each optional): dequantized f32, final
// quantized bytes, and the input of the trailing softmax.  layer_outs (optional): device->host copies of every layer.
int run_chunk(mf_model *m, Slot &s, const uint8_t *d_in, size_t n, float *d_out_f32, uint8_t *d_out_q, uint8_t *d_logits, cudaStream_t st,
              cudaEvent_t *prof, void *const *layer_outs_host, size_t sample_offset) {
    const uint8_t *cur = d_in;
    int flip = 0;
    bool f32_done = false;   // a fused classifier tail also wrote the dequantized output
    const bool i8_out = !m->spec.is_u8_out;
    // Programmatic dependent launch between consecutive layers (mf_kernels.h): off while per-layer events or trace copies sit
    // between the kernels, and for the first kernel of a chunk (its predecessor in the stream is a copy or another call).
    static const bool env_pdl = [] { const char *e = std::getenv("MF_PDL"); return !e || std::atoi(e) != 0; }();
    const bool use_pdl = env_pdl && !prof && !layer_outs_host;
    int pdl = 0;
    if (prof) MF_CUDA(cudaEventRecord(prof[0], st));
    for (size_t i = 0; i < m->layers.size(); ++i) {
        const LayerExec &L = m->layers[i];
        const Chain *chain = nullptr;
        if (!layer_outs_host)
            for (const auto &c : m->chains)
                if (c.first == (int)i) chain = &c;
        if (chain) {                                            // n x (depthwise 3x3 + pointwise 1x1) on a small feature map in one launch
            uint8_t *dst = s.act[flip];
            cudaError_t e = fused_chain_launch(chain->plan, cur, dst, (long long)n, m->num_sms, st, pdl);
            if (e != cudaSuccess) return fail(MF_ERR_CUDA, std::string("fused_chain_kernel launch failed: ") + cudaGetErrorString(e));
            m->launches += 1;
            m->launched[i] = "fused_chain_kernel";
            for (size_t k = i + 1; k <= (size_t)chain->last; ++k) m->launched[k] = "";
            pdl = use_pdl;
            cur = dst;
            flip ^= 1;
            if (prof)
                for (size_t k = i; k <= (size_t)chain->last; ++k) MF_CUDA(cudaEventRecord(prof[k + 1], st));
            i = (size_t)chain->last;
            continue;
        }
        if ((int)i == m->tail_first && !layer_outs_host) {      // pool + conv + softmax in one launch
            TailArgs t = m->tail;
            t.in = cur; t.out = s.act[flip]; t.logits = d_logits; t.batch = (long long)n; t.pdl = pdl;
            if (d_out_f32 && i8_out) { t.out_f32 = d_out_f32; t.dq_scale = m->spec.out_scale; t.dq_zp = (float)m->spec.out_zp; f32_done = true; }
            cudaError_t e = launch_tail_fused(t, st);
            if (e != cudaSuccess) return fail(MF_ERR_CUDA, std::string("tail_fused_kernel launch failed: ") + cudaGetErrorString(e));
            m->launches += 1;
            m->launched[i] = "tail_fused_kernel";
            for (size_t k = i + 1; k <= (size_t)m->tail_last; ++k) m->launched[k] = "";
            pdl = use_pdl;
            cur = t.out;
            flip ^= 1;
            if (prof)
                for (size_t k = i; k <= (size_t)m->tail_last; ++k) MF_CUDA(cudaEventRecord(prof[k + 1], st));
            i = (size_t)m->tail_last;
            continue;
        }
        if ((int)i == m->fc_tail_first && !layer_outs_host) {   // fully_connected + softmax in one launch (speech's classifier)
            FcArgs a = L.fc;
            const SoftmaxArgs &sa = m->layers[(size_t)m->fc_tail_last].sm;
            a.in = cur; a.out = d_logits; a.batch = (long long)n; a.pdl = pdl;
            a.sm_out = s.act[flip];
            a.exp_lut = sa.exp_lut; a.sm_rows = sa.rows; a.sm_cols = sa.cols;
            a.sm_out_scale = sa.out_scale; a.sm_out_zp = sa.out_zp; a.sm_lo = sa.lo; a.sm_hi = sa.hi;
            if (d_out_f32 && i8_out) { a.out_f32 = d_out_f32; a.dq_scale = m->spec.out_scale; a.dq_zp = (float)m->spec.out_zp; f32_done = true; }
            cudaError_t e = launch_fc_warp(a, st);
            if (e != cudaSuccess) return fail(MF_ERR_CUDA, std::string("fc_warp_kernel(+softmax) launch failed: ") + cudaGetErrorString(e));
            m->launches += 1;
            m->launched[i] = "fc_warp_kernel";
            for (size_t k = i + 1; k <= (size_t)m->fc_tail_last; ++k) m->launched[k] = "";
            pdl = use_pdl;
            cur = a.sm_out;
            flip ^= 1;
            if (prof)
                for (size_t k = i; k <= (size_t)m->fc_tail_last; ++k) MF_CUDA(cudaEventRecord(prof[k + 1], st));
            i = (size_t)m->fc_tail_last;
            continue;
        }
        if (d_logits && (int)i == m->softmax_tail) {
            MF_CUDA(cudaMemcpyAsync(d_logits, cur, n * L.spec.in_elems, cudaMemcpyDeviceToDevice, st));
            pdl = 0;
        }
        if (L.kernel != Kernel::None) {
            uint8_t *dst = s.act[flip];
            std::string err;
            cudaError_t e = L.run(cur, dst, (long long)n, m->num_sms, st, &err, pdl);
            if (e != cudaSuccess) return fail(MF_ERR_CUDA, std::string(kernel_name(L.kernel)) + " launch failed: " + cudaGetErrorString(e) + " " + err);
            m->launches += 1;
            m->launched[i] = L.launched_name(cur, dst, (long long)n);
            pdl = use_pdl;
            cur = dst;
            flip ^= 1;
        }
        if (layer_outs_host && layer_outs_host[i])
            MF_CUDA(cudaMemcpyAsync((uint8_t *)layer_outs_host[i] + sample_offset * L.spec.out_elems, cur, n * L.spec.out_elems, cudaMemcpyDeviceToHost, st));
        if (prof) MF_CUDA(cudaEventRecord(prof[i + 1], st));
    }
    const size_t total = n * m->spec.out_elems;
    if (d_out_f32 && !f32_done) {
        cudaError_t e = launch_dequantize(cur, d_out_f32, total, m->spec.out_scale, (float)m->spec.out_zp, m->spec.is_u8_out, st, pdl);   // src/tensor.rs:89-92
        if (e != cudaSuccess) return fail(MF_ERR_CUDA, std::string("dequantize launch failed: ") + cudaGetErrorString(e));
        m->launches += 1;
    }
    if (d_out_q) MF_CUDA(cudaMemcpyAsync(d_out_q, cur, total, cudaMemcpyDeviceToDevice, st));
    return MF_OK;
}

// Recognises [average_pool_2d over the whole image] -> [1x1 conv_2d, <= 8 outputs, dp4a kernel] -> reshape* -> softmax (last
// layer) and prepares the single-launch tail.  Anything else keeps the layer-by-layer path.
void plan_tail(mf_model *m) {
    const int n = (int)m->layers.size();
    if (m->softmax_tail < 0) return;
    for (int i = 0; i + 2 < n; ++i) {
        const LayerExec &P = m->layers[(size_t)i], &C = m->layers[(size_t)i + 1];
        if (P.kernel != Kernel::PoolGeneric || C.kernel != Kernel::PwConvDp4a) continue;
        const LayerSpec &p = P.spec, &c = C.spec;
        const PoolArgs &pa = P.pool;
        // the pool window must cover the whole image from a single output position
        if (p.is_u8 || p.OH != 1 || p.OW != 1 || -pa.off_r > 0 || -pa.off_c > 0 || -pa.off_r + p.KH < p.H || -pa.off_c + p.KW < p.W) continue;
        if (p.Cin % 128 != 0 || p.Cin > 512 || c.H != 1 || c.W != 1 || c.Cin != p.Cin || c.Cout > 8 || c.is_u8) continue;
        int j = i + 2;
        while (j < n && m->layers[(size_t)j].kernel == Kernel::None) ++j;
        if (j != m->softmax_tail || j != n - 1 || m->layers[(size_t)j].kernel != Kernel::Softmax) continue;
        const SoftmaxArgs &sa = m->layers[(size_t)j].sm;
        if (sa.rows * sa.cols != c.Cout) continue;
        TailArgs &t = m->tail;
        t = TailArgs{};
        t.HW = p.H * p.W; t.C = p.Cin; t.N = c.Cout;
        t.inv_len = 1.0f / (float)(p.H * p.W);      // pool_generic_kernel: __fdiv_rn(1.0f, float(len)), len = every pixel
        t.pool_c0 = pa.c0; t.pool_c1 = pa.c1; t.pool_lo = pa.lo; t.pool_hi = pa.hi;
        t.w = C.conv.w; t.c0z = C.conv.c0z; t.c1 = C.conv.c1; t.kcorr = C.conv.kcorr; t.conv_lo = C.conv.lo; t.conv_hi = C.conv.hi;
        t.exp_lut = sa.exp_lut; t.sm_rows = sa.rows; t.sm_cols = sa.cols; t.out_scale = sa.out_scale; t.out_zp = sa.out_zp; t.sm_lo = sa.lo; t.sm_hi = sa.hi;
        m->tail_first = i; m->tail_conv = i + 1; m->tail_last = j;
        return;
    }
}

// Recognises fully_connected (<= 8 outputs, fc_warp_kernel) -> reshape* -> softmax as the last layers.
void plan_fc_tail(mf_model *m) {
    const int n = (int)m->layers.size();
    if (m->softmax_tail != n - 1 || m->tail_first >= 0) return;
    int j = n - 2;
    while (j >= 0 && m->layers[(size_t)j].kernel == Kernel::None) --j;
    if (j < 0 || m->layers[(size_t)j].kernel != Kernel::FcWarp) return;
    const SoftmaxArgs &sa = m->layers[(size_t)n - 1].sm;
    if (m->layers[(size_t)n - 1].kernel != Kernel::Softmax || sa.rows * sa.cols != m->layers[(size_t)j].fc.N) return;
    m->fc_tail_first = j;
    m->fc_tail_last = n - 1;
}

// Recognises runs of [depthwise_conv_2d 3x3 / s1 / SAME, 128 ch] -> [conv_2d 1x1, 128 -> 128] on one small feature map (both on
// their fast kernels: int8, weight zero-points 0, accumulators within 2^22) and packs their static data for fused_chain_kernel.
void plan_chains(const std::vector<LayerExec> &layers, BlobBuilder &bb, std::vector<Chain> &chains) {
    const int n = (int)layers.size();
    auto is_dw = [&](int i, int H, int W) {
        const LayerExec &E = layers[(size_t)i];
        const LayerSpec &L = E.spec;
        return E.kernel == Kernel::DwConv3x3Rows && !E.big_acc && !L.is_u8 && L.Cin == kFusedC && L.Cout == kFusedC && L.KH == 3 && L.KW == 3 && L.sh == 1 &&
               L.sw == 1 && L.pad == MF_PAD_SAME && L.H == H && L.W == W && L.OH == H && L.OW == W;
    };
    auto is_pw = [&](int i, int H, int W) {
        const LayerExec &E = layers[(size_t)i];
        const LayerSpec &L = E.spec;
        return E.kernel == Kernel::ConvTcPointwise && !E.big_acc && !L.is_u8 && L.Cin == kFusedC && L.Cout == kFusedC && L.H == H && L.W == W;
    };
    for (int i = 0; i + 1 < n;) {
        const LayerSpec &L0 = layers[(size_t)i].spec;
        const int H = L0.H, W = L0.W;
        int pairs = 0;
        while (i + 2 * pairs + 1 < n && pairs < kFusedMaxPairs && is_dw(i + 2 * pairs, H, W) && is_pw(i + 2 * pairs + 1, H, W) &&
               fused_chain_smem(pairs + 1, H, W) != 0)
            ++pairs;
        if (pairs < 2) { ++i; continue; }          // a single pair gains nothing over its two kernels
        Chain c;
        c.first = i; c.last = i + 2 * pairs - 1;
        c.plan.n_pairs = pairs; c.plan.H = H; c.plan.W = W;
        std::vector<uint8_t> wimg((size_t)pairs * 128 * 128), consts((size_t)pairs * kFusedConstBytes);
        for (int l = 0; l < pairs; ++l) {
            const LayerSpec &D = layers[(size_t)(i + 2 * l)].spec, &P = layers[(size_t)(i + 2 * l + 1)].spec;
            auto consts_of = [](const LayerSpec &S, std::vector<float> &c0z, std::vector<float> &c1, std::vector<int32_t> &kcorr, bool dw) {
                c0z.resize(128); c1.resize(128); kcorr.assign(128, 0);
                for (int b = 0; b < 128; ++b) {
                    c0z[(size_t)b] = (float)S.out_zp + S.c0[(size_t)b];                     // conv_2d.rs:94-95: the same single f32 add
                    c1[(size_t)b] = S.c1[(size_t)b < S.c1.size() ? (size_t)b : 0];
                    int32_t sum = 0;
                    if (dw) for (int t = 0; t < 9; ++t) sum += (int8_t)S.w[(size_t)t * 128 + b];
                    else for (int k = 0; k < 128; ++k) sum += (int8_t)S.w[(size_t)b * 128 + k];
                    kcorr[(size_t)b] = S.in_zp * sum;
                }
            };
            std::vector<float> dz, d1, pz, p1;
            std::vector<int32_t> dk, pk;
            consts_of(D, dz, d1, dk, true);
            consts_of(P, pz, p1, pk, false);
            fused_pack_pointwise_image(P.w.data(), wimg.data() + (size_t)l * 128 * 128);
            fused_pack_consts(D.w.data(), dz.data(), d1.data(), pz.data(), p1.data(), pk.data(), consts.data() + (size_t)l * kFusedConstBytes);
            c.plan.dw_zp[l] = D.in_zp;
            c.plan.dw_lo[l] = (float)D.act_lo; c.plan.dw_hi[l] = (float)D.act_hi;
            c.plan.pw_lo[l] = (float)P.act_lo; c.plan.pw_hi[l] = (float)P.act_hi;
        }
        std::string why;
        if (!fused_chain_finalize(c.plan, &why)) { ++i; continue; }
        c.o_wimg = bb.add(wimg.data(), wimg.size());
        c.o_consts = bb.add(consts.data(), consts.size());
        chains.push_back(c);
        i = c.last + 1;
    }
}

// after LayerExec::resolve: a layer that resolve() downgraded breaks its chain (it then runs layer by layer)
void resolve_chains(const std::vector<LayerExec> &layers, std::vector<Chain> &chains, const uint8_t *d_blob) {
    for (auto it = chains.begin(); it != chains.end();) {
        bool ok = true;
        for (int k = it->first; k <= it->last; ++k) {
            const Kernel want = ((k - it->first) & 1) ? Kernel::ConvTcPointwise : Kernel::DwConv3x3Rows;
            ok = ok && layers[(size_t)k].kernel == want;
        }
        if (!ok) { it = chains.erase(it); continue; }
        it->plan.d_wimg = d_blob + it->o_wimg;
        it->plan.d_consts = d_blob + it->o_consts;
        ++it;
    }
}

int need_device(const mf_model *m) {
    if (!m) return fail(MF_ERR_INVALID_ARG, "null model");
    if (m->host_only) return fail(MF_ERR_NO_DEVICE, "model was created with MF_FLAG_HOST_ONLY: it has no device state and cannot predict");
    return MF_OK;
}

constexpr size_t kGraphMaxN = 64;

// Returns MF_OK after serving the call from a graph, or -1 when the caller should take the stream path (never a CPU path).
int predict_small_graph(mf_model *m, const void *in_q, const float *in_f32, size_t n, float *out_f32, void *out_q, void *logits) {
    static const bool env_graph = [] { const char *e = std::getenv("MF_GRAPH"); return !e || std::atoi(e) != 0; }();
    if (!env_graph || m->graphs_disabled || n == 0 || n > kGraphMaxN || n > m->chunk) return -1;
    const size_t ie = m->spec.in_elems, oe = m->spec.out_elems;
    const size_t le = m->softmax_tail >= 0 ? m->layers[(size_t)m->softmax_tail].spec.in_elems : 0;
    const int kind = (in_f32 ? 1 : 0) | (out_f32 ? 2 : 0) | (out_q ? 4 : 0) | (logits ? 8 : 0);
    Slot &s = m->slot[0];
    if (in_f32 && ensure_f32_staging(m, s) != MF_OK) { (void)cudaGetLastError(); return -1; }
    if (!m->h_stage_in) {
        if (cudaMallocHost(&m->h_stage_in, kGraphMaxN * ie * sizeof(float)) != cudaSuccess ||
            cudaMallocHost(&m->h_stage_out_f32, kGraphMaxN * oe * sizeof(float)) != cudaSuccess ||
            cudaMallocHost(&m->h_stage_out_q, kGraphMaxN * oe) != cudaSuccess ||
            cudaMallocHost(&m->h_stage_logits, kGraphMaxN * (le ? le : 1)) != cudaSuccess) {
            (void)cudaGetLastError();
            m->graphs_disabled = true;
            return -1;
        }
    }
    mf_model::SmallGraph *g = nullptr;
    for (auto &e : m->graphs)
        if (e.n == n && e.kind == kind) g = &e;
    if (!g) {
        // both slots' streams may still hold work of earlier asynchronous calls that uses the same device buffers
        if (cudaStreamSynchronize(m->slot[0].stream) != cudaSuccess || cudaStreamSynchronize(m->slot[1].stream) != cudaSuccess) return -1;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        const uint64_t launches0 = m->launches;
        bool ok = cudaStreamBeginCapture(s.stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            if (in_f32) {
                ok = cudaMemcpyAsync(s.in_f32, m->h_stage_in, n * ie * sizeof(float), cudaMemcpyHostToDevice, s.stream) == cudaSuccess &&
                     launch_quantize(s.in_f32, s.in_q, n * ie, m->spec.in_scale, (float)m->spec.in_zp, m->spec.is_u8_in, s.stream) == cudaSuccess;
                m->launches += 1;
            } else {
                ok = cudaMemcpyAsync(s.in_q, m->h_stage_in, n * ie, cudaMemcpyHostToDevice, s.stream) == cudaSuccess;
            }
            const uint8_t *d_in = s.in_q;
            ok = ok && stage_input_layout(m, s, n, s.stream, &d_in) == MF_OK;
            ok = ok && run_chunk(m, s, d_in, n, out_f32 ? s.out_f32 : nullptr, out_q ? s.out_q : nullptr, logits ? s.logits : nullptr, s.stream, nullptr, nullptr, 0) == MF_OK;
            const void *d_o = nullptr;   // the transposed copy (if any) is consumed by the D2H right behind it, so one buffer serves both outputs
            if (ok && out_f32)
                ok = stage_output_layout(m, s, n, s.out_f32, sizeof(float), s.stream, &d_o) == MF_OK &&
                     cudaMemcpyAsync(m->h_stage_out_f32, d_o, n * oe * sizeof(float), cudaMemcpyDeviceToHost, s.stream) == cudaSuccess;
            if (ok && out_q)
                ok = stage_output_layout(m, s, n, s.out_q, 1, s.stream, &d_o) == MF_OK &&
                     cudaMemcpyAsync(m->h_stage_out_q, d_o, n * oe, cudaMemcpyDeviceToHost, s.stream) == cudaSuccess;
            if (ok && logits) ok = cudaMemcpyAsync(m->h_stage_logits, s.logits, n * le, cudaMemcpyDeviceToHost, s.stream) == cudaSuccess;
            const bool ended = cudaStreamEndCapture(s.stream, &graph) == cudaSuccess;   // always end the capture, even after a failure
            ok = ok && ended && graph != nullptr;
        }
        if (ok) ok = cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        const uint64_t per_replay = m->launches - launches0;
        m->launches = launches0;                          // capturing launched nothing; replays are counted below
        if (!ok) {
            (void)cudaGetLastError();
            m->graphs_disabled = true;
            return -1;
        }
        if (m->graphs.size() >= 32) {                     // bounded cache: drop the oldest
            cudaGraphExecDestroy(m->graphs.front().exec);
            m->graphs.erase(m->graphs.begin());
        }
        mf_model::SmallGraph e;
        e.n = n; e.kind = kind; e.exec = exec; e.kernels = per_replay;
        m->graphs.push_back(e);
        g = &m->graphs.back();
    }
    std::memcpy(m->h_stage_in, in_f32 ? (const void *)in_f32 : in_q, n * ie * (in_f32 ? sizeof(float) : 1));
    // slot 0's stream orders the replay after any earlier asynchronous call that used slot 0's device buffers
    MF_CUDA(cudaGraphLaunch(g->exec, s.stream));
    MF_CUDA(cudaStreamSynchronize(s.stream));
    if (out_f32) std::memcpy(out_f32, m->h_stage_out_f32, n * oe * sizeof(float));
    if (out_q) std::memcpy(out_q, m->h_stage_out_q, n * oe);
    if (logits) std::memcpy(logits, m->h_stage_logits, n * le);
    m->launches += g->kernels;
    return MF_OK;
}

// host-buffer batched path: chunks alternate between two streams so H2D(c+1) overlaps compute(c) and D2H(c-1)
// single_piece: the whole call is ONE piece on ONE stream slot (the multi-device chunk loop pipelines its chunks itself)
int predict_many_host(mf_model *m, const void *in_q, const float *in_f32, size_t n, float *out_f32, void *out_q, void *logits, bool wait = true,
                      bool single_piece = false) {
    int rc = need_device(m);
    if (rc) return rc;
    if ((!in_q && !in_f32) || (!out_f32 && !out_q)) return fail(MF_ERR_INVALID_ARG, "null input or output buffer");
    if (logits && m->softmax_tail < 0) return fail(MF_ERR_INVALID_ARG, "model has no softmax layer: no logits to return");
    std::lock_guard<std::mutex> lock(m->mu);
    MF_CUDA(cudaSetDevice(m->device));
    const size_t ie = m->spec.in_elems, oe = m->spec.out_elems;
    const size_t le = m->softmax_tail >= 0 ? m->layers[(size_t)m->softmax_tail].spec.in_elems : 0;
    if (wait && n <= kGraphMaxN) {
        rc = predict_small_graph(m, in_q, in_f32, n, out_f32, out_q, logits);
        if (rc >= 0) return rc;
    }
    // host path: pieces small enough that the H2D of piece c+1 overlaps the compute of piece c (two streams), large enough
    // to keep the per-launch fixed costs amortised
    size_t piece = std::min(m->chunk, std::max<size_t>(1024, (n + 1) / 2));
    // a blocking call cannot hide the compute of its last piece behind a later copy, so it uses smaller pieces than the
    // asynchronous form (whose calls pipeline into each other); MF_HOST_PIECE overrides both for experiments
    static const int env_piece = [] { const char *e = std::getenv("MF_HOST_PIECE"); return e ? std::atoi(e) : 0; }();
    const size_t cap = env_piece > 0 ? (size_t)env_piece : (wait ? 2048 : 4096);
    if (piece > cap) piece = cap;
    if (single_piece) piece = std::min(m->chunk, n);
    size_t ci = m->slot_rr;
    for (size_t off = 0; off < n; off += piece, ++ci) {
        Slot &s = m->slot[ci & 1];
        const size_t cn = std::min(piece, n - off);
        if (in_f32) {   // predict(): quantize the f32 input on the device (src/tensor.rs:80-86, :246-256)
            rc = ensure_f32_staging(m, s);
            if (rc) return rc;
            MF_CUDA(cudaMemcpyAsync(s.in_f32, in_f32 + off * ie, cn * ie * sizeof(float), cudaMemcpyHostToDevice, s.stream));
            cudaError_t e = launch_quantize(s.in_f32, s.in_q, cn * ie, m->spec.in_scale, (float)m->spec.in_zp, m->spec.is_u8_in, s.stream);
            if (e != cudaSuccess) return fail(MF_ERR_CUDA, std::string("quantize launch failed: ") + cudaGetErrorString(e));
            m->launches += 1;
        } else {
            MF_CUDA(cudaMemcpyAsync(s.in_q, (const uint8_t *)in_q + off * ie, cn * ie, cudaMemcpyHostToDevice, s.stream));
        }
        const uint8_t *d_in = s.in_q;
        rc = stage_input_layout(m, s, cn, s.stream, &d_in);
        if (rc) return rc;
        rc = run_chunk(m, s, d_in, cn, out_f32 ? s.out_f32 : nullptr, out_q ? s.out_q : nullptr, logits ? s.logits : nullptr, s.stream, nullptr, nullptr, 0);
        if (rc) return rc;
        const void *d_o = nullptr;
        if (out_f32) {
            rc = stage_output_layout(m, s, cn, s.out_f32, sizeof(float), s.stream, &d_o);
            if (rc) return rc;
            MF_CUDA(cudaMemcpyAsync(out_f32 + off * oe, d_o, cn * oe * sizeof(float), cudaMemcpyDeviceToHost, s.stream));
        }
        if (out_q) {
            rc = stage_output_layout(m, s, cn, s.out_q, 1, s.stream, &d_o);
            if (rc) return rc;
            MF_CUDA(cudaMemcpyAsync((uint8_t *)out_q + off * oe, d_o, cn * oe, cudaMemcpyDeviceToHost, s.stream));
        }
        if (logits) MF_CUDA(cudaMemcpyAsync((uint8_t *)logits + off * le, s.logits, cn * le, cudaMemcpyDeviceToHost, s.stream));
    }
    m->slot_rr = ci;                                   // the next call continues the slot rotation (pipelining across async calls)
    if (wait) {
        MF_CUDA(cudaStreamSynchronize(m->slot[0].stream));
        MF_CUDA(cudaStreamSynchronize(m->slot[1].stream));
    }
    return MF_OK;
}

// One complete single-device model on o.device.  upload_weights == false (replicas 1.. of a multi-device model): the blob is
// allocated and zeroed, its contents arrive by the broadcast in create_model.
int create_single(const uint8_t *buf, size_t len, const mf_options &o, bool upload_weights, mf_model **out) {
    std::unique_ptr<mf_model, void (*)(mf_model *)> m(new mf_model(), mf_model_destroy);
    std::string err;
    int rc = parse_tflite(buf, len, m->spec, err);
    if (rc != MF_OK) return fail(rc, err);
    m->flags = o.flags;
    m->layout = o.layout;
    if (o.layout == MF_LAYOUT_NALGEBRA) {
        auto geo = [](int rank, const int *d) {
            mf_model::Geo g;
            if (rank == 4) { g.mats = d[0]; g.R = d[1]; g.C = d[2]; g.ch = d[3]; }   // Buffer4D = [SMatrix<[T; CH], R, C>; B]
            else { g.mats = 1; g.R = d[0]; g.C = d[1]; g.ch = 1; }                   // Buffer2D = SMatrix<T, R, C>
            g.transpose = g.R > 1 && g.C > 1;                                        // a single row or column is the same bytes either way
            return g;
        };
        m->geo_in = geo(m->spec.in_rank, m->spec.in_dims);
        m->geo_out = geo(m->spec.out_rank, m->spec.out_dims);
    }
    m->host_only = (o.flags & MF_FLAG_HOST_ONLY) != 0;
    bool have_device = false;
    if (!m->host_only) {
        int n = 0;
        rc = check_device(&n);
        if (rc) return rc;
        if (o.device >= n) return fail(MF_ERR_INVALID_ARG, "device ordinal out of range");
        if (o.device >= 0) MF_CUDA(cudaSetDevice(o.device));
        MF_CUDA(cudaGetDevice(&m->device));
        cudaDeviceProp prop{};
        MF_CUDA(cudaGetDeviceProperties(&prop, m->device));
        m->num_sms = prop.multiProcessorCount;
        if (prop.major != 10) return fail(MF_ERR_NO_DEVICE, std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                                              ": this library is built for sm_100a (B200) only");
        have_device = true;
    }
    const int impl = (o.flags & MF_FLAG_FORCE_GENERIC) ? 1 : ((o.flags & MF_FLAG_NO_TENSOR_CORE) ? 2 : 0);
    BlobBuilder bb;
    m->layers.resize(m->spec.layers.size());
    for (size_t i = 0; i < m->layers.size(); ++i) {
        m->layers[i].spec = m->spec.layers[i];
        m->layers[i].plan(bb, impl, have_device);
        if (m->spec.layers[i].op == MF_OP_SOFTMAX && i + 1 == m->layers.size()) m->softmax_tail = (int)i;
    }
    if (m->softmax_tail < 0)
        for (size_t i = m->layers.size(); i-- > 0;) {
            if (m->spec.layers[i].op == MF_OP_RESHAPE) continue;
            if (m->spec.layers[i].op == MF_OP_SOFTMAX) m->softmax_tail = (int)i;
            break;
        }
    if (have_device && impl == 0 && !std::getenv("MF_NO_CHAIN_FUSE")) plan_chains(m->layers, bb, m->chains);
    m->blob_bytes = bb.bytes().size();
    m->launched.assign(m->layers.size(), "");
    if (have_device) {
        m->chunk = o.chunk ? o.chunk : 8192;
        if (m->blob_bytes) {
            MF_CUDA(cudaMalloc(&m->d_blob, m->blob_bytes));
            if (upload_weights) MF_CUDA(cudaMemcpy(m->d_blob, bb.bytes().data(), m->blob_bytes, cudaMemcpyHostToDevice));
            else MF_CUDA(cudaMemset(m->d_blob, 0, m->blob_bytes));
        }
        for (auto &L : m->layers)
            if (!L.resolve(m->d_blob, &err)) return fail(MF_ERR_CUDA, err);
        resolve_chains(m->layers, m->chains, m->d_blob);
        if (impl != 1 && !std::getenv("MF_NO_TAIL_FUSE")) { plan_tail(m.get()); plan_fc_tail(m.get()); }
        for (int k = 0; k < 2; ++k) {
            rc = alloc_slot(m.get(), m->slot[k]);
            if (rc) return rc;
        }
    }
    *out = m.release();
    return MF_OK;
}

// ---- one broadcast of the static weight blob from replicas[0] to the others -------------------------------------------------
// NCCL is loaded at run time (libnccl.so.2: the torch-bundled or the system one) so the library has no link-time dependency on
// it; without it the same bytes travel by cudaMemcpyPeer.  Either way it happens once, at create; nothing is exchanged later.
struct NcclApi {
    typedef int (*InitAllFn)(void **, int, const int *);
    typedef int (*BcastFn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
    typedef int (*VoidFn)(void);
    typedef int (*DestroyFn)(void *);
    InitAllFn comm_init_all = nullptr;
    BcastFn broadcast = nullptr;
    VoidFn group_start = nullptr, group_end = nullptr;
    DestroyFn comm_destroy = nullptr;
    bool ok = false;
};
const NcclApi &nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        if (std::getenv("MF_NO_NCCL")) return a;
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return a;
        a.comm_init_all = (NcclApi::InitAllFn)dlsym(h, "ncclCommInitAll");
        a.broadcast = (NcclApi::BcastFn)dlsym(h, "ncclBroadcast");
        a.group_start = (NcclApi::VoidFn)dlsym(h, "ncclGroupStart");
        a.group_end = (NcclApi::VoidFn)dlsym(h, "ncclGroupEnd");
        a.comm_destroy = (NcclApi::DestroyFn)dlsym(h, "ncclCommDestroy");
        a.ok = a.comm_init_all && a.broadcast && a.group_start && a.group_end && a.comm_destroy;
        return a;
    }();
    return api;
}

int broadcast_blob(mf_model *g) {
    const size_t G = g->replicas.size(), bytes = g->replicas[0]->blob_bytes;
    g->bcast = "none";
    if (G < 2 || !bytes) return MF_OK;
    const NcclApi &nc = nccl_api();
    bool distinct = true;
    for (size_t r = 0; r < G; ++r)
        for (size_t q = 0; q < r; ++q) distinct = distinct && g->replicas[r]->device != g->replicas[q]->device;
    if (nc.ok && distinct) {
        void *comms[MF_MAX_DEVICES] = {};
        int devs[MF_MAX_DEVICES] = {};
        for (size_t r = 0; r < G; ++r) devs[r] = g->replicas[r]->device;
        bool ok = nc.comm_init_all(comms, (int)G, devs) == 0;
        if (ok) {
            ok = nc.group_start() == 0;
            for (size_t r = 0; ok && r < G; ++r) {
                ok = cudaSetDevice(devs[r]) == cudaSuccess &&
                     nc.broadcast(g->replicas[0]->d_blob, g->replicas[r]->d_blob, bytes, /*ncclUint8*/ 1, /*root*/ 0, comms[r], g->replicas[r]->slot[0].stream) == 0;
            }
            ok = (nc.group_end() == 0) && ok;
            for (size_t r = 0; r < G; ++r)
                if (cudaSetDevice(devs[r]) != cudaSuccess || cudaStreamSynchronize(g->replicas[r]->slot[0].stream) != cudaSuccess) ok = false;
        }
        for (void *c : comms)
            if (c) nc.comm_destroy(c);
        if (ok) { g->bcast = "nccl"; return MF_OK; }
        (void)cudaGetLastError();
    }
    for (size_t r = 1; r < G; ++r)
        MF_CUDA(cudaMemcpyPeer(g->replicas[r]->d_blob, g->replicas[r]->device, g->replicas[0]->d_blob, g->replicas[0]->device, bytes));
    g->bcast = "memcpy_peer";
    return MF_OK;
}

int create_model(const uint8_t *buf, size_t len, const mf_options *opt, mf_model **out) {
    if (!out) return fail(MF_ERR_INVALID_ARG, "null output pointer");
    *out = nullptr;
    mf_options o{};
    o.struct_size = sizeof(mf_options);
    o.device = -1;
    if (opt) {   // struct_size versions the struct: an ABI-1 caller (16 bytes, no `layout`) or an ABI-2 caller (20 bytes, no device list) is accepted
        if (opt->struct_size < 16) return fail(MF_ERR_INVALID_ARG, "mf_options.struct_size too small");
        std::memcpy(&o, opt, std::min<size_t>(opt->struct_size, sizeof(mf_options)));
        o.struct_size = sizeof(mf_options);
    }
    if (o.layout != MF_LAYOUT_NHWC && o.layout != MF_LAYOUT_NALGEBRA) return fail(MF_ERR_INVALID_ARG, "mf_options.layout must be MF_LAYOUT_NHWC or MF_LAYOUT_NALGEBRA");
    if (o.n_devices > MF_MAX_DEVICES || o.n_devices < -1) return fail(MF_ERR_INVALID_ARG, "mf_options.n_devices out of range");
    std::vector<int> devs;
    if (o.n_devices != 0 && !(o.flags & MF_FLAG_HOST_ONLY)) {
        int n = 0;
        int rc = check_device(&n);
        if (rc) return rc;
        if (o.n_devices == -1) { for (int d = 0; d < n && d < MF_MAX_DEVICES; ++d) devs.push_back(d); }
        else devs.assign(o.devices, o.devices + o.n_devices);
        for (size_t i = 0; i < devs.size(); ++i) {
            if (devs[i] < 0 || devs[i] >= n) return fail(MF_ERR_INVALID_ARG, "mf_options.devices: ordinal out of range");
            // MF_ALLOW_DUPLICATE_DEVICES=1 (tests only): several replicas on one GPU, so that a one-GPU box exercises the routing,
            // the per-replica host threads and the sharding of a multi-device model
            for (size_t j = 0; j < i; ++j)
                if (devs[j] == devs[i] && !std::getenv("MF_ALLOW_DUPLICATE_DEVICES")) return fail(MF_ERR_INVALID_ARG, "mf_options.devices: duplicate ordinal");
        }
    }
    if (devs.size() <= 1) {
        if (devs.size() == 1) o.device = devs[0];
        return create_single(buf, len, o, true, out);
    }
    int prev = 0;
    (void)cudaGetDevice(&prev);
    std::unique_ptr<mf_model, void (*)(mf_model *)> g(new mf_model(), mf_model_destroy);
    for (size_t r = 0; r < devs.size(); ++r) {
        mf_options od = o;
        od.device = devs[r];
        mf_model *rep = nullptr;
        int rc = create_single(buf, len, od, r == 0, &rep);
        if (rc) { (void)cudaSetDevice(prev); return rc; }
        g->replicas.push_back(rep);
        g->workers.emplace_back(new Worker());
    }
    int rc = broadcast_blob(g.get());
    (void)cudaSetDevice(prev);
    if (rc) return rc;
    *out = g.release();
    return MF_OK;
}

// contiguous shard of replica r: [r * ceil(n / G), min(n, (r + 1) * ceil(n / G)))
inline void shard_range(size_t n, size_t r, size_t G, size_t &lo, size_t &hi) {
    const size_t per = (n + G - 1) / G;
    lo = std::min(n, r * per);
    hi = std::min(n, lo + per);
}

// multi-device predict_many*: every replica runs its shard on its own host thread; nothing is exchanged
int group_predict_many(mf_model *g, const void *in_q, const float *in_f32, size_t n, float *out_f32, void *out_q, void *logits, bool wait) {
    if ((!in_q && !in_f32) || (!out_f32 && !out_q)) return fail(MF_ERR_INVALID_ARG, "null input or output buffer");
    std::lock_guard<std::mutex> lock(g->group_mu);
    const size_t G = g->replicas.size();
    const mf_model *p = g->replicas[0];
    const size_t ie = p->spec.in_elems, oe = p->spec.out_elems;
    const size_t le = p->softmax_tail >= 0 ? p->layers[(size_t)p->softmax_tail].spec.in_elems : 0;
    struct Res { int rc = 0; std::string err; };
    auto res = std::make_shared<std::vector<Res>>(G);
    // Large host-resident batches: the host -> device links of a box are not equally fast once several GPUs copy at the same time
    // (pairs of GPUs share an uplink: profiles/r02c_h2d_sweep_8gpu.txt measured 21-37 GB/s per GPU with eight copying), and with equal
    // shards the slowest link sets the time of the call.  So the range is cut into contiguous 2048-sample chunks that the replicas'
    // host threads claim from a shared counter: a GPU on a faster link takes more of them.  Every chunk still lands at its own offset
    // of the caller's buffers, so the result is the same row for row; with equal links this degenerates to equal contiguous
    // shards.  MF_SHARD_DYNAMIC=0 keeps the static split [r * ceil(n / G), ...) below (always used for smaller calls).
    static const bool env_dyn = [] { const char *e = std::getenv("MF_SHARD_DYNAMIC"); return !e || std::atoi(e) != 0; }();
    constexpr size_t kDynChunk = 2048;
    if (env_dyn && n >= G * 2 * kDynChunk) {
        auto next = std::make_shared<std::atomic<size_t>>(0);
        const size_t nchunks = (n + kDynChunk - 1) / kDynChunk;
        for (size_t r = 0; r < G; ++r) {
            mf_model *rep = g->replicas[r];
            g->workers[r]->post([=] {
                int rc = MF_OK;
                size_t k = 0;
                for (;;) {
                    const size_t c = next->fetch_add(1);
                    if (c >= nchunks) break;
                    const size_t lo = c * kDynChunk, cn = std::min(kDynChunk, n - lo);
                    const size_t slot = rep->slot_rr & 1;            // the stream slot predict_many_host is about to use for this chunk
                    rc = predict_many_host(rep, in_q ? (const uint8_t *)in_q + lo * ie : nullptr, in_f32 ? in_f32 + lo * ie : nullptr, cn,
                                           out_f32 ? out_f32 + lo * oe : nullptr, out_q ? (uint8_t *)out_q + lo * oe : nullptr,
                                           logits ? (uint8_t *)logits + lo * le : nullptr, /*wait=*/false, /*single_piece=*/true);
                    if (rc) break;
                    // two chunks in flight per GPU: before claiming another one, wait for the chunk enqueued before this one
                    if (k >= 1 && (cudaSetDevice(rep->device) != cudaSuccess || cudaStreamSynchronize(rep->slot[slot ^ 1].stream) != cudaSuccess)) {
                        rc = fail(MF_ERR_CUDA, "cudaStreamSynchronize failed in the multi-device chunk loop");
                        break;
                    }
                    ++k;
                }
                if (!rc && wait && (cudaSetDevice(rep->device) != cudaSuccess || cudaStreamSynchronize(rep->slot[0].stream) != cudaSuccess ||
                                    cudaStreamSynchronize(rep->slot[1].stream) != cudaSuccess))
                    rc = fail(MF_ERR_CUDA, "cudaStreamSynchronize failed at the end of a multi-device call");
                if (rc) {
                    (*res)[r].rc = rc;
                    (*res)[r].err = g_err;
                    if (!wait) {
                        std::lock_guard<std::mutex> l(g->group_mu);
                        if (!g->deferred_rc) { g->deferred_rc = rc; g->deferred_err = g_err; }
                    }
                }
            });
        }
        if (!wait) return MF_OK;
        for (auto &w : g->workers) w->drain();
        for (size_t r = 0; r < G; ++r)
            if ((*res)[r].rc) return fail((*res)[r].rc, "device " + std::to_string(g->replicas[r]->device) + ": " + (*res)[r].err);
        return MF_OK;
    }
    for (size_t r = 0; r < G; ++r) {
        size_t lo, hi;
        shard_range(n, r, G, lo, hi);
        if (hi <= lo) continue;
        mf_model *rep = g->replicas[r];
        g->workers[r]->post([=] {
            int rc = predict_many_host(rep, in_q ? (const uint8_t *)in_q + lo * ie : nullptr, in_f32 ? in_f32 + lo * ie : nullptr, hi - lo,
                                       out_f32 ? out_f32 + lo * oe : nullptr, out_q ? (uint8_t *)out_q + lo * oe : nullptr,
                                       logits ? (uint8_t *)logits + lo * le : nullptr, wait);
            if (rc) {
                (*res)[r].rc = rc;
                (*res)[r].err = g_err;
                if (!wait) {   // asynchronous call: the error surfaces at mf_model_synchronize
                    std::lock_guard<std::mutex> l(g->group_mu);
                    if (!g->deferred_rc) { g->deferred_rc = rc; g->deferred_err = g_err; }
                }
            }
        });
    }
    if (!wait) return MF_OK;
    for (auto &w : g->workers) w->drain();
    for (size_t r = 0; r < G; ++r)
        if ((*res)[r].rc) return fail((*res)[r].rc, "device " + std::to_string(g->replicas[r]->device) + ": " + (*res)[r].err);
    return MF_OK;
}

inline mf_model *primary(mf_model *m) { return (m && !m->replicas.empty()) ? m->replicas[0] : m; }
inline const mf_model *primary(const mf_model *m) { return (m && !m->replicas.empty()) ? m->replicas[0] : m; }

}  // namespace

// =================================================================================================
extern "C" {

int mf_abi_version(void) { return MF_ABI_VERSION; }
const char *mf_last_error(void) { return g_err.c_str(); }
const char *mf_status_string(int s) {
    switch (s) {
        case MF_OK: return "ok";
        case MF_ERR_FILE: return "model file not found";
        case MF_ERR_INVALID_MODEL: return "invalid TensorFlow Lite model";
        case MF_ERR_UNSUPPORTED_TYPE: return "unsupported tensor type";
        case MF_ERR_UNSUPPORTED_RANK: return "unsupported tensor rank";
        case MF_ERR_UNSUPPORTED_OP: return "unsupported operator";
        case MF_ERR_UNSUPPORTED_ACTIVATION: return "unsupported fused activation";
        case MF_ERR_UNSUPPORTED_SHAPE: return "unsupported shape";
        case MF_ERR_VIEW_OUT_OF_BOUNDS: return "VALID view out of bounds";
        case MF_ERR_INVALID_ARG: return "invalid argument";
        case MF_ERR_NO_DEVICE: return "no CUDA device (no CPU fallback)";
        case MF_ERR_CUDA: return "CUDA error";
        case MF_ERR_NONFINITE_CONSTANT: return "non-finite requantization constant";
    }
    return "unknown status";
}
int mf_device_count(int *count) { return check_device(count); }

int mf_model_create_from_tflite(const void *buf, size_t len, const mf_options *opt, mf_model **out) {
    if (!buf) return fail(MF_ERR_INVALID_ARG, "null model buffer");
    return create_model((const uint8_t *)buf, len, opt, out);
}
int mf_model_create_from_file(const char *path, const mf_options *opt, mf_model **out) {
    if (!path) return fail(MF_ERR_INVALID_ARG, "null path");
    std::ifstream f(path, std::ios::binary);
    if (!f) return fail(MF_ERR_FILE, std::string("couldn't find '") + path + "', please provide a valid path");   // lib.rs:50-55
    std::vector<char> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return create_model((const uint8_t *)data.data(), data.size(), opt, out);
}
void mf_model_destroy(mf_model *m) {
    if (!m) return;
    if (!m->replicas.empty() || !m->workers.empty()) {   // routing object of a multi-device model
        for (auto &w : m->workers) w->drain();
        m->workers.clear();
        for (mf_model *r : m->replicas) mf_model_destroy(r);
        delete m;
        return;
    }
    if (!m->host_only) {
        cudaSetDevice(m->device);
        cudaDeviceSynchronize();
        for (auto e : m->prof_events) cudaEventDestroy(e);
        for (auto &g : m->graphs) cudaGraphExecDestroy(g.exec);
        for (auto e : m->split_ev)
            if (e) cudaEventDestroy(e);
        for (void *hp : {(void *)m->h_stage_in, (void *)m->h_stage_out_f32, (void *)m->h_stage_out_q, (void *)m->h_stage_logits})
            if (hp) cudaFreeHost(hp);
        free_slot(m->slot[0]);
        free_slot(m->slot[1]);
        if (m->d_blob) cudaFree(m->d_blob);
    }
    delete m;
}

int mf_model_io_info(const mf_model *m, mf_tensor_info *in, mf_tensor_info *out) {
    if (!m) return fail(MF_ERR_INVALID_ARG, "null model");
    m = primary(m);
    if (in) {
        *in = mf_tensor_info{};
        in->rank = m->spec.in_rank;
        for (int i = 0; i < 4; ++i) in->dims[i] = m->spec.in_dims[i];
        in->dtype = m->spec.is_u8_in ? MF_DTYPE_U8 : MF_DTYPE_I8;
        in->scale = m->spec.in_scale; in->zero_point = m->spec.in_zp; in->elems = m->spec.in_elems;
    }
    if (out) {
        *out = mf_tensor_info{};
        out->rank = m->spec.out_rank;
        for (int i = 0; i < 4; ++i) out->dims[i] = m->spec.out_dims[i];
        out->dtype = m->spec.is_u8_out ? MF_DTYPE_U8 : MF_DTYPE_I8;
        out->scale = m->spec.out_scale; out->zero_point = m->spec.out_zp; out->elems = m->spec.out_elems;
    }
    return MF_OK;
}
int mf_model_num_layers(const mf_model *m) { return m ? (int)primary(m)->layers.size() : 0; }
int mf_model_layer_info(const mf_model *m, int i, mf_layer_info *o) {
    m = primary(m);
    if (!m || !o || i < 0 || (size_t)i >= m->layers.size()) return fail(MF_ERR_INVALID_ARG, "bad layer index");
    const LayerExec &E = m->layers[(size_t)i];
    const LayerSpec &L = E.spec;
    *o = mf_layer_info{};
    o->op = L.op;
    for (int k = 0; k < 4; ++k) { o->in_dims[k] = L.in_dims[k]; o->out_dims[k] = L.out_dims[k]; }
    o->in_rank = L.in_rank; o->out_rank = L.out_rank;
    o->kh = L.KH; o->kw = L.KW; o->stride_h = L.sh; o->stride_w = L.sw; o->padding = L.pad; o->activation = L.act;
    o->in_zero_point = L.in_zp; o->out_zero_point = L.out_zp; o->in_scale = L.in_scale; o->out_scale = L.out_scale;
    o->act_lo = L.act_lo; o->act_hi = L.act_hi;
    o->n_c0 = (int)L.c0.size(); o->n_c1 = (int)L.c1.size();
    o->macs = L.macs; o->bytes = E.alg_bytes; o->weight_bytes = E.weight_bytes;
    const char *kn = m->host_only ? "" : kernel_name(E.kernel);
    if (!m->host_only) kn = E.launched_name(m->slot[0].act[0], m->slot[0].act[1], 1 << 20);   // the variant large batches run on
    if (m->tail_first >= 0 && i >= m->tail_first && i <= m->tail_last && E.kernel != Kernel::None)
        kn = i == m->tail_first ? "tail_fused_kernel" : "(in tail_fused_kernel)";
    if (m->fc_tail_first >= 0 && i > m->fc_tail_first && i <= m->fc_tail_last && E.kernel != Kernel::None) kn = "(in fc_warp_kernel)";
    for (const auto &c : m->chains)
        if (i >= c.first && i <= c.last) kn = i == c.first ? "fused_chain_kernel" : "(in fused_chain_kernel)";
    std::snprintf(o->kernel, sizeof o->kernel, "%s", kn);
    return MF_OK;
}
int mf_model_layer_constants(const mf_model *m, int i, float *c0, float *c1, int32_t *c2, int32_t *c3, int cap) {
    m = primary(m);
    if (!m || i < 0 || (size_t)i >= m->layers.size() || cap < 0) return fail(MF_ERR_INVALID_ARG, "bad layer index");
    const LayerSpec &L = m->layers[(size_t)i].spec;
    for (size_t k = 0; c0 && k < L.c0.size() && k < (size_t)cap; ++k) c0[k] = L.c0[k];
    for (size_t k = 0; c1 && k < L.c1.size() && k < (size_t)cap; ++k) c1[k] = L.c1[k];
    for (size_t k = 0; c2 && k < L.c2.size() && k < (size_t)cap; ++k) c2[k] = L.c2[k];
    if (c3) *c3 = L.c3;
    return MF_OK;
}
int mf_model_dump(const mf_model *m, const char *path) {
    if (!m || !path) return fail(MF_ERR_INVALID_ARG, "null argument");
    m = primary(m);
    FILE *f = std::fopen(path, "w");
    if (!f) return fail(MF_ERR_FILE, std::string("cannot open '") + path + "' for writing");
    std::fprintf(f, "# microflow_cuda model dump (equivalent of target/microflow-expansion.rs)\n");
    std::fprintf(f, "input: rank %d dims [%d,%d,%d,%d] scale %.9g zp %d %s\n", m->spec.in_rank, m->spec.in_dims[0], m->spec.in_dims[1], m->spec.in_dims[2],
                 m->spec.in_dims[3], m->spec.in_scale, m->spec.in_zp, m->spec.is_u8_in ? "u8" : "i8");
    for (size_t i = 0; i < m->layers.size(); ++i) {
        const LayerExec &E = m->layers[i];
        const LayerSpec &L = E.spec;
        std::fprintf(f, "layer %zu: op %d in [%d,%d,%d,%d] out [%d,%d,%d,%d] k %dx%d s %dx%d pad %d act %d izp %d ozp %d is %.9g os %.9g clamp [%d,%d] kernel %s%s%s\n", i,
                     L.op, L.in_dims[0], L.in_dims[1], L.in_dims[2], L.in_dims[3], L.out_dims[0], L.out_dims[1], L.out_dims[2], L.out_dims[3], L.KH, L.KW, L.sh,
                     L.sw, L.pad, L.act, L.in_zp, L.out_zp, L.in_scale, L.out_scale, L.act_lo, L.act_hi, m->host_only ? "-" : kernel_name(E.kernel),
                     E.why_not_fast.empty() ? (E.big_acc ? "  # accumulator range beyond 2^22" : "") : "  # ", E.why_not_fast.c_str());
        std::fprintf(f, "  c0:");
        for (float v : L.c0) std::fprintf(f, " %.9g", v);
        std::fprintf(f, "\n  c1:");
        for (float v : L.c1) std::fprintf(f, " %.9g", v);
        if (!L.c2.empty()) {
            std::fprintf(f, "\n  c2:");
            for (int32_t v : L.c2) std::fprintf(f, " %d", v);
            std::fprintf(f, "\n  c3: %d", L.c3);
        }
        std::fprintf(f, "\n");
    }
    std::fprintf(f, "output: rank %d dims [%d,%d,%d,%d] scale %.9g zp %d\n", m->spec.out_rank, m->spec.out_dims[0], m->spec.out_dims[1], m->spec.out_dims[2],
                 m->spec.out_dims[3], m->spec.out_scale, m->spec.out_zp);
    std::fclose(f);
    return MF_OK;
}

// host-buffer entry points: a multi-device model shards the samples over its replicas, a single-device model runs them itself
static int predict_host_any(mf_model *m, const void *in_q, const float *in_f32, size_t n, float *out_f32, void *out_q, void *logits, bool wait = true) {
    if (!m) return fail(MF_ERR_INVALID_ARG, "null model");
    if (!m->replicas.empty()) return group_predict_many(m, in_q, in_f32, n, out_f32, out_q, logits, wait);
    return predict_many_host(m, in_q, in_f32, n, out_f32, out_q, logits, wait);
}
int mf_predict(mf_model *m, const float *in, float *out) { return predict_host_any(m, nullptr, in, 1, out, nullptr, nullptr); }
int mf_predict_quantized(mf_model *m, const void *in_q, float *out) { return predict_host_any(m, in_q, nullptr, 1, out, nullptr, nullptr); }
int mf_predict_many(mf_model *m, const float *in, size_t n, float *out) { return predict_host_any(m, nullptr, in, n, out, nullptr, nullptr); }
int mf_predict_many_quantized(mf_model *m, const void *in_q, size_t n, float *out) { return predict_host_any(m, in_q, nullptr, n, out, nullptr, nullptr); }
int mf_predict_many_quantized_async(mf_model *m, const void *in_q, size_t n, float *out) {
    return predict_host_any(m, in_q, nullptr, n, out, nullptr, nullptr, /*wait=*/false);
}
int mf_predict_many_logits(mf_model *m, const void *in_q, size_t n, void *out_q, void *logits_q) {
    if (!out_q) return fail(MF_ERR_INVALID_ARG, "null out_q");
    if (m && logits_q && primary(m)->softmax_tail < 0) return fail(MF_ERR_INVALID_ARG, "model has no softmax layer: no logits to return");
    return predict_host_any(m, in_q, nullptr, n, nullptr, out_q, logits_q);
}

int mf_predict_many_device_on(mf_model *m, int index, const void *d_in_q, size_t n, float *d_out_f32, void *d_out_q, void *stream) {
    if (!m) return fail(MF_ERR_INVALID_ARG, "null model");
    if (m->replicas.empty()) return index == 0 ? mf_predict_many_device(m, d_in_q, n, d_out_f32, d_out_q, stream) : fail(MF_ERR_INVALID_ARG, "device index out of range");
    if (index < 0 || (size_t)index >= m->replicas.size()) return fail(MF_ERR_INVALID_ARG, "device index out of range");
    return mf_predict_many_device(m->replicas[(size_t)index], d_in_q, n, d_out_f32, d_out_q, stream);
}

int mf_predict_many_device(mf_model *m, const void *d_in_q, size_t n, float *d_out_f32, void *d_out_q, void *stream) {
    m = primary(m);      // device buffers belong to one device: a multi-device model serves this call on devices[0] (see mf_predict_many_device_on)
    int rc = need_device(m);
    if (rc) return rc;
    if (!d_in_q) return fail(MF_ERR_INVALID_ARG, "null device input");
    std::lock_guard<std::mutex> lock(m->mu);
    MF_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : m->slot[0].stream;
    const size_t ie = m->spec.in_elems, oe = m->spec.out_elems;
    size_t nchunks = (n + m->chunk - 1) / m->chunk;
    size_t ci = 0;
    if (m->profiling) {   // event groups accumulate over calls until mf_model_layer_times_ms() reads and resets them
        ci = m->prof_chunks;
        const size_t need = (m->prof_chunks + nchunks) * (m->layers.size() + 1);
        while (m->prof_events.size() < need) {
            cudaEvent_t e;
            MF_CUDA(cudaEventCreate(&e));
            m->prof_events.push_back(e);
        }
        m->prof_chunks += nchunks;
    }
    // Each chunk of >= 4096 samples runs as two half-chunks on the model's two internal streams (fork / join by events on the
    // caller's stream): launch ramp and tail of one half overlap steady-state work of the other.  Measured 1.038 -> 1.011 ms/step
    // at batch 8192 with full-size grids on both streams; restricting the depthwise kernels to 2 CTAs/SM so that they can share
    // an SM with the other half's tcgen05 kernel was SLOWER (1.059), i.e. mixing the two instruction streams on one scheduler does
    // not pay -- the reason dw -> pw fusion is not pursued.  Off while per-layer events are recorded.
    static const int env_split = [] { const char *e = std::getenv("MF_SPLIT"); return e ? std::atoi(e) : 2; }();
    for (size_t off = 0; off < n; off += m->chunk, ++ci) {
        const size_t cn = std::min(m->chunk, n - off);
        if (env_split == 2 && !m->profiling && cn >= 4096) {
            if (!m->split_ev[0])
                for (auto &e : m->split_ev) MF_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            MF_CUDA(cudaEventRecord(m->split_ev[0], st));
            const size_t half = (cn + 1) / 2;
            for (int h = 0; h < 2; ++h) {
                Slot &s = m->slot[h];
                const size_t o2 = off + (h ? half : 0), c2 = h ? cn - half : half;
                MF_CUDA(cudaStreamWaitEvent(s.stream, m->split_ev[0], 0));
                rc = run_chunk(m, s, (const uint8_t *)d_in_q + o2 * ie, c2, d_out_f32 ? d_out_f32 + o2 * oe : nullptr, d_out_q ? (uint8_t *)d_out_q + o2 * oe : nullptr,
                               nullptr, s.stream, nullptr, nullptr, 0);
                if (rc) return rc;
                MF_CUDA(cudaEventRecord(m->split_ev[1 + h], s.stream));
                MF_CUDA(cudaStreamWaitEvent(st, m->split_ev[1 + h], 0));
            }
            continue;
        }
        cudaEvent_t *prof = m->profiling ? &m->prof_events[ci * (m->layers.size() + 1)] : nullptr;
        rc = run_chunk(m, m->slot[0], (const uint8_t *)d_in_q + off * ie, cn, d_out_f32 ? d_out_f32 + off * oe : nullptr,
                       d_out_q ? (uint8_t *)d_out_q + off * oe : nullptr, nullptr, st, prof, nullptr, 0);
        if (rc) return rc;
    }
    return MF_OK;
}

int mf_predict_trace(mf_model *m, const void *in_q, size_t n, void *const *layer_outs) {
    m = primary(m);
    int rc = need_device(m);
    if (rc) return rc;
    if (!in_q || !layer_outs) return fail(MF_ERR_INVALID_ARG, "null argument");
    std::lock_guard<std::mutex> lock(m->mu);
    MF_CUDA(cudaSetDevice(m->device));
    Slot &s = m->slot[0];
    const size_t ie = m->spec.in_elems;
    for (size_t off = 0; off < n; off += m->chunk) {
        const size_t cn = std::min(m->chunk, n - off);
        MF_CUDA(cudaMemcpyAsync(s.in_q, (const uint8_t *)in_q + off * ie, cn * ie, cudaMemcpyHostToDevice, s.stream));
        rc = run_chunk(m, s, s.in_q, cn, nullptr, nullptr, nullptr, s.stream, nullptr, layer_outs, off);
        if (rc) return rc;
        MF_CUDA(cudaStreamSynchronize(s.stream));
    }
    return MF_OK;
}

int mf_model_synchronize(mf_model *m) {
    if (m && !m->replicas.empty()) {
        for (auto &w : m->workers) w->drain();
        int first = MF_OK;
        for (mf_model *r : m->replicas) {
            int rc = mf_model_synchronize(r);
            if (rc && !first) first = rc;
        }
        std::lock_guard<std::mutex> lock(m->group_mu);
        if (m->deferred_rc) {
            const int rc = m->deferred_rc;
            const std::string e = m->deferred_err;
            m->deferred_rc = 0;
            m->deferred_err.clear();
            return fail(rc, e);
        }
        return first;
    }
    int rc = need_device(m);
    if (rc) return rc;
    MF_CUDA(cudaSetDevice(m->device));
    MF_CUDA(cudaDeviceSynchronize());
    return MF_OK;
}
int mf_model_set_profiling(mf_model *m, int enabled) {
    m = primary(m);
    int rc = need_device(m);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(m->mu);
    m->profiling = enabled != 0;
    m->prof_chunks = 0;
    return MF_OK;
}
int mf_model_layer_times_ms(mf_model *m, float *ms, int cap) {
    m = primary(m);
    int rc = need_device(m);
    if (rc) return rc;
    if (!ms || cap < (int)m->layers.size()) return fail(MF_ERR_INVALID_ARG, "ms buffer too small");
    std::lock_guard<std::mutex> lock(m->mu);
    MF_CUDA(cudaSetDevice(m->device));
    MF_CUDA(cudaDeviceSynchronize());
    const size_t L = m->layers.size();
    for (size_t i = 0; i < L; ++i) ms[i] = 0.f;
    for (size_t c = 0; c < m->prof_chunks; ++c)
        for (size_t i = 0; i < L; ++i) {
            float t = 0.f;
            MF_CUDA(cudaEventElapsedTime(&t, m->prof_events[c * (L + 1) + i], m->prof_events[c * (L + 1) + i + 1]));
            ms[i] += t;
        }
    m->prof_chunks = 0;
    return MF_OK;
}
int mf_model_launch_count(const mf_model *m, uint64_t *count) {
    if (!m || !count) return fail(MF_ERR_INVALID_ARG, "null argument");
    *count = m->launches;
    for (const mf_model *r : m->replicas) *count += r->launches;
    return MF_OK;
}
const char *mf_model_layer_launched(const mf_model *m, int layer) {
    m = primary(m);
    if (!m || layer < 0 || (size_t)layer >= m->launched.size()) return "";
    return m->launched[(size_t)layer];
}
int mf_model_devices(const mf_model *m, int32_t *devices, int cap) {
    if (!m) return 0;
    if (m->replicas.empty()) {
        if (m->host_only) return 0;
        if (devices && cap > 0) devices[0] = m->device;
        return 1;
    }
    for (size_t r = 0; r < m->replicas.size() && devices && (int)r < cap; ++r) devices[r] = m->replicas[r]->device;
    return (int)m->replicas.size();
}
const char *mf_model_weight_broadcast(const mf_model *m) { return m ? m->bcast : "none"; }
int mf_model_blob(const mf_model *m, void **d_ptr, size_t *bytes) {
    m = primary(m);
    int rc = need_device(m);
    if (rc) return rc;
    if (d_ptr) *d_ptr = m->d_blob;
    if (bytes) *bytes = m->blob_bytes;
    return MF_OK;
}

// ---- host staging of the reference's sample formats ------------------------------------------------
int mf_features_from_bmp_gray8(const void *bmp, size_t len, void *out, size_t cap, int32_t *height, int32_t *width) {
    std::string err;
    int h = 0, w = 0;
    int rc = features_from_bmp_gray8((const uint8_t *)bmp, len, (uint8_t *)out, out ? cap : 0, &h, &w, err);
    if (height) *height = h;
    if (width) *width = w;
    if (rc != MF_OK && !(rc == MF_ERR_INVALID_ARG && !out && h > 0)) return fail(rc, err);      // out == NULL: a size query
    return MF_OK;
}
int mf_predict_many_bmp(mf_model *m, const void *const *bmps, const size_t *lens, size_t n, float *out) {
    if (!m || !bmps || !lens || !out) return fail(MF_ERR_INVALID_ARG, "null argument");
    const mf_model *p = primary(m);
    if (p->spec.in_rank != 4 || p->spec.in_dims[3] != 1 || p->spec.is_u8_in) return fail(MF_ERR_UNSUPPORTED_SHAPE, "the model does not take one-channel int8 images");
    const size_t ie = p->spec.in_elems;
    std::vector<uint8_t> staged(n * ie);
    for (size_t k = 0; k < n; ++k) {
        std::string err;
        int h = 0, w = 0;
        int rc = features_from_bmp_gray8((const uint8_t *)bmps[k], lens[k], staged.data() + k * ie, ie, &h, &w, err);
        if (rc) return fail(rc, "image " + std::to_string(k) + ": " + err);
        if (h != p->spec.in_dims[1] || w != p->spec.in_dims[2]) return fail(MF_ERR_UNSUPPORTED_SHAPE, "image " + std::to_string(k) + " is not " +
                                                                                std::to_string(p->spec.in_dims[1]) + "x" + std::to_string(p->spec.in_dims[2]));
    }
    if (p->layout == MF_LAYOUT_NALGEBRA) return fail(MF_ERR_INVALID_ARG, "mf_predict_many_bmp stages NHWC inputs: create the model with MF_LAYOUT_NHWC");
    return predict_host_any(m, staged.data(), nullptr, n, out, nullptr, nullptr);
}

int mf_host_alloc(void **p, size_t bytes) {
    if (!p) return fail(MF_ERR_INVALID_ARG, "null pointer");
    int rc = check_device(nullptr);
    if (rc) return rc;
    MF_CUDA(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocPortable));   // usable by every device of a multi-device model
    return MF_OK;
}
int mf_host_free(void *p) {
    if (p) MF_CUDA(cudaFreeHost(p));
    return MF_OK;
}

// -------------------------------------------------------------------------------------------------
// per-operator hooks: build a one-layer plan, run it on `batch` samples, copy back
// -------------------------------------------------------------------------------------------------
static int run_single_layer(LayerSpec &L, int impl, const void *in, void *out, size_t batch) {
    int rc = check_device(nullptr);
    if (rc) return rc;
    if (!in || !out) return fail(MF_ERR_INVALID_ARG, "null buffer");
    cudaDeviceProp prop{};
    int dev = 0;
    MF_CUDA(cudaGetDevice(&dev));
    MF_CUDA(cudaGetDeviceProperties(&prop, dev));
    activation_clamp(L.act, L.out_scale, L.out_zp, L.is_u8, L.act_lo, L.act_hi);
    LayerExec E;
    E.spec = L;
    BlobBuilder bb;
    E.plan(bb, impl == 1 ? 1 : 0, true);
    uint8_t *d_blob = nullptr, *d_in = nullptr, *d_out = nullptr;
    auto cleanup = [&] { cudaFree(d_blob); cudaFree(d_in); cudaFree(d_out); };
    cudaError_t e = cudaSuccess;
    std::string err;
    do {
        if ((e = cudaMalloc(&d_blob, bb.bytes().size() + 256)) != cudaSuccess) break;
        if ((e = cudaMemcpy(d_blob, bb.bytes().data(), bb.bytes().size(), cudaMemcpyHostToDevice)) != cudaSuccess) break;
        E.resolve(d_blob, &err);
        if (impl == 2 && (E.kernel == Kernel::ConvGeneric || E.kernel == Kernel::FcGeneric)) {
            cleanup();
            return fail(MF_ERR_UNSUPPORTED_SHAPE, "no fast kernel takes this operator: " + E.why_not_fast);
        }
        if ((e = cudaMalloc(&d_in, batch * L.in_elems + 256)) != cudaSuccess) break;
        if ((e = cudaMalloc(&d_out, batch * L.out_elems + 256)) != cudaSuccess) break;
        if ((e = cudaMemcpy(d_in, in, batch * L.in_elems, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = E.run(d_in, d_out, (long long)batch, prop.multiProcessorCount, nullptr, &err)) != cudaSuccess) break;
        if ((e = cudaDeviceSynchronize()) != cudaSuccess) break;
        e = cudaMemcpy(out, d_out, batch * L.out_elems, cudaMemcpyDeviceToHost);
    } while (false);
    const uint8_t *d_in_tag = d_in;   // only their alignment is inspected after the buffers are freed
    uint8_t *d_out_tag = d_out;
    cleanup();
    if (e != cudaSuccess) return fail(MF_ERR_CUDA, std::string(kernel_name(E.kernel)) + ": " + cudaGetErrorString(e) + " " + err);
    g_err = E.launched_name(d_in_tag, d_out_tag, (long long)batch);   // lets tests see which kernel ran
    return MF_OK;
}

static int conv_spec_from_desc(const mf_conv_desc *d, LayerSpec &L) {
    if (!d || !d->filters || !d->filter_zero_points || !d->c0 || !d->c1 || d->n_filter_zero_points < 1 || d->n_c1 < 1)
        return fail(MF_ERR_INVALID_ARG, "incomplete mf_conv_desc");
    if (d->dtype != MF_DTYPE_I8 && d->dtype != MF_DTYPE_U8) return fail(MF_ERR_UNSUPPORTED_TYPE, "dtype must be INT8 or UINT8");
    L.op = d->depthwise ? MF_OP_DEPTHWISE_CONV_2D : MF_OP_CONV_2D;
    L.is_u8 = d->dtype == MF_DTYPE_U8;
    L.H = d->in_h; L.W = d->in_w; L.Cin = d->in_c; L.OH = d->out_h; L.OW = d->out_w; L.Cout = d->out_c;
    L.KH = d->kh; L.KW = d->kw; L.sh = d->stride_h; L.sw = d->stride_w; L.pad = d->padding; L.act = d->activation;
    L.in_zp = d->in_zero_point; L.out_scale = d->out_scale; L.out_zp = d->out_zero_point;
    if (L.H <= 0 || L.W <= 0 || L.Cin <= 0 || L.OH <= 0 || L.OW <= 0 || L.Cout <= 0 || L.KH <= 0 || L.KW <= 0 || L.sh <= 0 || L.sw <= 0)
        return fail(MF_ERR_UNSUPPORTED_SHAPE, "non-positive dimension");
    if (L.act != MF_ACT_NONE && L.act != MF_ACT_RELU && L.act != MF_ACT_RELU6) return fail(MF_ERR_UNSUPPORTED_ACTIVATION, "unsupported fused activation");
    if (L.pad != MF_PAD_SAME && L.pad != MF_PAD_VALID) return fail(MF_ERR_INVALID_ARG, "unknown padding");
    if (L.pad == MF_PAD_VALID && (L.sh * (L.OH - 1) + L.KH > L.H || L.sw * (L.OW - 1) + L.KW > L.W))
        return fail(MF_ERR_VIEW_OUT_OF_BOUNDS, "VALID view indexes outside the input (src/tensor.rs:222 would panic)");
    const size_t wn = d->depthwise ? (size_t)L.KH * L.KW * L.Cout : (size_t)L.Cout * L.KH * L.KW * L.Cin;
    L.w.assign((const uint8_t *)d->filters, (const uint8_t *)d->filters + wn);
    L.w_zp.assign(d->filter_zero_points, d->filter_zero_points + d->n_filter_zero_points);
    L.c0.assign(d->c0, d->c0 + L.Cout);
    L.c1.assign(d->c1, d->c1 + d->n_c1);
    L.in_elems = (size_t)L.H * L.W * L.Cin;
    L.out_elems = (size_t)L.OH * L.OW * L.Cout;
    L.macs = (uint64_t)L.OH * L.OW * L.Cout * L.KH * L.KW * (d->depthwise ? 1 : L.Cin);
    activation_clamp(L.act, L.out_scale, L.out_zp, L.is_u8, L.act_lo, L.act_hi);
    return MF_OK;
}

int mf_op_conv_2d(const mf_conv_desc *d, const void *in, void *out, size_t batch) {
    LayerSpec L;
    int rc = conv_spec_from_desc(d, L);
    if (rc) return rc;
    return run_single_layer(L, d->impl, in, out, batch);
}

// A sequence of conv_2d / depthwise_conv_2d operators, each consuming the previous one's output -- the straight-line chain the
// macro emits (microflow-macros/src/lib.rs:198-201) restricted to convolutions.  fuse = 0: one kernel per operator;
// fuse = 1: the whole sequence must be taken by ONE fused_chain_kernel launch (mf_fused.h), else MF_ERR_UNSUPPORTED_SHAPE.
int mf_op_conv_chain(const mf_conv_desc *descs, int n_ops, const void *in, void *out, size_t batch, int fuse) {
    int rc = check_device(nullptr);
    if (rc) return rc;
    if (!descs || n_ops < 1 || !in || !out) return fail(MF_ERR_INVALID_ARG, "null or empty operator list");
    cudaDeviceProp prop{};
    int dev = 0;
    MF_CUDA(cudaGetDevice(&dev));
    MF_CUDA(cudaGetDeviceProperties(&prop, dev));
    std::vector<LayerExec> layers((size_t)n_ops);
    BlobBuilder bb;
    size_t max_elems = 0;
    for (int i = 0; i < n_ops; ++i) {
        rc = conv_spec_from_desc(&descs[i], layers[(size_t)i].spec);
        if (rc) return rc;
        if (i && layers[(size_t)i].spec.in_elems != layers[(size_t)i - 1].spec.out_elems) return fail(MF_ERR_UNSUPPORTED_SHAPE, "operator input does not match the previous output");
        layers[(size_t)i].plan(bb, 0, true);
        max_elems = std::max(max_elems, std::max(layers[(size_t)i].spec.in_elems, layers[(size_t)i].spec.out_elems));
    }
    std::vector<Chain> chains;
    if (fuse) plan_chains(layers, bb, chains);
    uint8_t *d_blob = nullptr, *d_a = nullptr, *d_b = nullptr;
    auto cleanup = [&] { cudaFree(d_blob); cudaFree(d_a); cudaFree(d_b); };
    cudaError_t e = cudaSuccess;
    std::string err, names;
    do {
        if ((e = cudaMalloc(&d_blob, bb.bytes().size() + 256)) != cudaSuccess) break;
        if ((e = cudaMemcpy(d_blob, bb.bytes().data(), bb.bytes().size(), cudaMemcpyHostToDevice)) != cudaSuccess) break;
        for (auto &L : layers) L.resolve(d_blob, &err);
        resolve_chains(layers, chains, d_blob);
        if (fuse && (chains.size() != 1 || chains[0].first != 0 || chains[0].last != n_ops - 1)) {
            cleanup();
            return fail(MF_ERR_UNSUPPORTED_SHAPE, "the operator sequence is not one fusable chain of (depthwise 3x3 s1 SAME, 128 ch) -> (1x1 conv, 128 -> 128) pairs");
        }
        if ((e = cudaMalloc(&d_a, batch * max_elems + 256)) != cudaSuccess) break;
        if ((e = cudaMalloc(&d_b, batch * max_elems + 256)) != cudaSuccess) break;
        if ((e = cudaMemcpy(d_a, in, batch * layers[0].spec.in_elems, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        uint8_t *cur = d_a, *nxt = d_b;
        if (fuse) {
            e = fused_chain_launch(chains[0].plan, cur, nxt, (long long)batch, prop.multiProcessorCount, nullptr, 0);
            names = "fused_chain_kernel";
            std::swap(cur, nxt);
        } else {
            for (auto &L : layers) {
                if ((e = L.run(cur, nxt, (long long)batch, prop.multiProcessorCount, nullptr, &err)) != cudaSuccess) break;
                names += std::string(names.empty() ? "" : ",") + L.launched_name(cur, nxt, (long long)batch);
                std::swap(cur, nxt);
            }
        }
        if (e != cudaSuccess) break;
        if ((e = cudaDeviceSynchronize()) != cudaSuccess) break;
        e = cudaMemcpy(out, cur, batch * layers.back().spec.out_elems, cudaMemcpyDeviceToHost);
    } while (false);
    cleanup();
    if (e != cudaSuccess) return fail(MF_ERR_CUDA, std::string("conv chain: ") + cudaGetErrorString(e) + " " + err);
    g_err = names;   // lets tests see which kernels ran
    return MF_OK;
}

// ---- persistent operator object: plan once, run on device-resident buffers (benchmarks, pipelines) ----------------
struct mf_op {
    LayerExec exec;
    uint8_t *d_blob = nullptr;
    int num_sms = 148;
    const char *launched = nullptr;     // the kernel the most recent mf_op_run_device launched (static string)
};

int mf_op_conv_2d_create(const mf_conv_desc *d, mf_op **out) {
    if (!out) return fail(MF_ERR_INVALID_ARG, "null output pointer");
    *out = nullptr;
    int rc = check_device(nullptr);
    if (rc) return rc;
    LayerSpec L;
    rc = conv_spec_from_desc(d, L);
    if (rc) return rc;
    std::unique_ptr<mf_op, void (*)(mf_op *)> op(new mf_op(), mf_op_destroy);
    cudaDeviceProp prop{};
    int dev = 0;
    MF_CUDA(cudaGetDevice(&dev));
    MF_CUDA(cudaGetDeviceProperties(&prop, dev));
    op->num_sms = prop.multiProcessorCount;
    op->exec.spec = L;
    BlobBuilder bb;
    op->exec.plan(bb, d->impl == 1 ? 1 : 0, true);
    MF_CUDA(cudaMalloc(&op->d_blob, bb.bytes().size() + 256));
    MF_CUDA(cudaMemcpy(op->d_blob, bb.bytes().data(), bb.bytes().size(), cudaMemcpyHostToDevice));
    std::string err;
    op->exec.resolve(op->d_blob, &err);
    if (d->impl == 2 && op->exec.kernel == Kernel::ConvGeneric) return fail(MF_ERR_UNSUPPORTED_SHAPE, "no fast kernel takes this operator: " + op->exec.why_not_fast);
    *out = op.release();
    return MF_OK;
}
int mf_op_run_device(mf_op *op, const void *d_in, void *d_out, size_t batch, void *stream) {
    if (!op || !d_in || !d_out) return fail(MF_ERR_INVALID_ARG, "null argument");
    std::string err;
    cudaError_t e = op->exec.run((const uint8_t *)d_in, (uint8_t *)d_out, (long long)batch, op->num_sms, (cudaStream_t)stream, &err);
    if (e != cudaSuccess) return fail(MF_ERR_CUDA, std::string(kernel_name(op->exec.kernel)) + ": " + cudaGetErrorString(e) + " " + err);
    op->launched = op->exec.launched_name((const uint8_t *)d_in, (uint8_t *)d_out, (long long)batch);
    return MF_OK;
}
const char *mf_op_kernel_name(const mf_op *op) { return op ? (op->launched ? op->launched : kernel_name(op->exec.kernel)) : ""; }
void mf_op_destroy(mf_op *op) {
    if (!op) return;
    if (op->d_blob) cudaFree(op->d_blob);
    delete op;
}

int mf_op_fully_connected(const mf_fc_desc *d, const void *in, void *out, size_t batch) {
    if (!d || !d->weights_nk || !d->c0 || !d->c2) return fail(MF_ERR_INVALID_ARG, "incomplete mf_fc_desc");
    if (d->dtype != MF_DTYPE_I8 && d->dtype != MF_DTYPE_U8) return fail(MF_ERR_UNSUPPORTED_TYPE, "dtype must be INT8 or UINT8");
    if (d->in_features <= 0 || d->out_features <= 0) return fail(MF_ERR_UNSUPPORTED_SHAPE, "non-positive dimension");
    LayerSpec L;
    L.op = MF_OP_FULLY_CONNECTED;
    L.is_u8 = d->dtype == MF_DTYPE_U8;
    L.Cin = d->in_features; L.Cout = d->out_features;
    L.w.assign((const uint8_t *)d->weights_nk, (const uint8_t *)d->weights_nk + (size_t)L.Cin * L.Cout);
    L.w_zp.assign(1, d->weight_zero_point);
    L.out_scale = d->out_scale; L.out_zp = d->out_zero_point; L.act = d->activation;
    L.c0.assign(d->c0, d->c0 + L.Cout);
    L.c1.assign(1, d->c1);
    L.c2.assign(d->c2, d->c2 + L.Cout);
    L.c3 = d->c3;
    L.in_elems = (size_t)L.Cin; L.out_elems = (size_t)L.Cout;
    return run_single_layer(L, d->impl, in, out, batch);
}

int mf_op_average_pool_2d(const mf_pool_desc *d, const void *in, void *out, size_t batch) {
    if (!d) return fail(MF_ERR_INVALID_ARG, "null mf_pool_desc");
    if (d->dtype != MF_DTYPE_I8 && d->dtype != MF_DTYPE_U8) return fail(MF_ERR_UNSUPPORTED_TYPE, "dtype must be INT8 or UINT8");
    LayerSpec L;
    L.op = MF_OP_AVERAGE_POOL_2D;
    L.is_u8 = d->dtype == MF_DTYPE_U8;
    L.H = d->in_h; L.W = d->in_w; L.Cin = L.Cout = d->chans; L.OH = d->out_h; L.OW = d->out_w;
    L.KH = d->filter_h; L.KW = d->filter_w; L.sh = d->stride_h; L.sw = d->stride_w; L.pad = d->padding; L.act = d->activation;
    L.out_scale = d->out_scale; L.out_zp = d->out_zero_point;
    if (L.H <= 0 || L.W <= 0 || L.Cin <= 0 || L.OH <= 0 || L.OW <= 0 || L.KH <= 0 || L.KW <= 0 || L.sh <= 0 || L.sw <= 0)
        return fail(MF_ERR_UNSUPPORTED_SHAPE, "non-positive dimension");
    if (L.pad == MF_PAD_VALID && (L.sh * (L.OH - 1) + L.KH > L.H || L.sw * (L.OW - 1) + L.KW > L.W))
        return fail(MF_ERR_VIEW_OUT_OF_BOUNDS, "VALID view indexes outside the input (src/tensor.rs:222 would panic)");
    L.c0.assign(1, d->c0);
    L.c1.assign(1, d->c1);
    L.in_elems = (size_t)L.H * L.W * L.Cin;
    L.out_elems = (size_t)L.OH * L.OW * L.Cin;
    return run_single_layer(L, d->impl, in, out, batch);
}

int mf_op_softmax(int32_t dtype, int32_t rows, int32_t cols, float in_scale, float out_scale, int32_t out_zero_point, const void *in, void *out, size_t batch) {
    if (dtype != MF_DTYPE_I8 && dtype != MF_DTYPE_U8) return fail(MF_ERR_UNSUPPORTED_TYPE, "dtype must be INT8 or UINT8");
    if (rows <= 0 || cols <= 0) return fail(MF_ERR_UNSUPPORTED_SHAPE, "non-positive dimension");
    LayerSpec L;
    L.op = MF_OP_SOFTMAX;
    L.is_u8 = dtype == MF_DTYPE_U8;
    L.in_scale = in_scale; L.out_scale = out_scale; L.out_zp = out_zero_point;
    L.out_rank = 2; L.out_dims[0] = rows; L.out_dims[1] = cols;
    L.in_elems = L.out_elems = (size_t)rows * cols;
    L.exp_lut.resize(256);
    for (int b = 0; b < 256; ++b) L.exp_lut[(size_t)b] = libm_expf((float)(L.is_u8 ? b : (int)(int8_t)b) * in_scale);
    return run_single_layer(L, 0, in, out, batch);
}

int mf_op_layout_transpose(const void *in, void *out, size_t batch, int32_t rows, int32_t cols, int32_t elem_bytes, int32_t to_nalgebra) {
    int rc = check_device(nullptr);
    if (rc) return rc;
    if (!in || !out || rows < 1 || cols < 1 || elem_bytes < 1) return fail(MF_ERR_INVALID_ARG, "bad argument");
    const size_t bytes = batch * (size_t)rows * cols * elem_bytes;
    if (!bytes) return MF_OK;
    uint8_t *d_in = nullptr, *d_out = nullptr;
    cudaError_t e = cudaSuccess;
    do {
        if ((e = cudaMalloc(&d_in, bytes)) != cudaSuccess) break;
        if ((e = cudaMalloc(&d_out, bytes)) != cudaSuccess) break;
        if ((e = cudaMemcpy(d_in, in, bytes, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        // the kernel writes dst[r][c] = src[c][r] for its (R, C): towards column-major the destination's outer index is the column
        if ((e = launch_layout_transpose(d_in, d_out, (long long)batch, to_nalgebra ? cols : rows, to_nalgebra ? rows : cols, elem_bytes, nullptr)) != cudaSuccess) break;
        if ((e = cudaDeviceSynchronize()) != cudaSuccess) break;
        e = cudaMemcpy(out, d_out, bytes, cudaMemcpyDeviceToHost);
    } while (false);
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) return fail(MF_ERR_CUDA, std::string("layout_transpose_kernel: ") + cudaGetErrorString(e));
    return MF_OK;
}
int mf_op_quantize(int32_t dtype, float scale, int32_t zero_point, const float *in, void *out, size_t n) {
    int rc = check_device(nullptr);
    if (rc) return rc;
    if (!in || !out) return fail(MF_ERR_INVALID_ARG, "null buffer");
    if (dtype != MF_DTYPE_I8 && dtype != MF_DTYPE_U8) return fail(MF_ERR_UNSUPPORTED_TYPE, "dtype must be INT8 or UINT8");
    float *d_in = nullptr;
    uint8_t *d_out = nullptr;
    MF_CUDA(cudaMalloc(&d_in, n * 4 + 4));
    MF_CUDA(cudaMalloc(&d_out, n + 4));
    cudaError_t e = cudaMemcpy(d_in, in, n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_quantize(d_in, d_out, n, scale, (float)zero_point, dtype == MF_DTYPE_U8, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, n, cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) return fail(MF_ERR_CUDA, cudaGetErrorString(e));
    return MF_OK;
}
int mf_op_dequantize(int32_t dtype, float scale, int32_t zero_point, const void *in, float *out, size_t n) {
    int rc = check_device(nullptr);
    if (rc) return rc;
    if (!in || !out) return fail(MF_ERR_INVALID_ARG, "null buffer");
    if (dtype != MF_DTYPE_I8 && dtype != MF_DTYPE_U8) return fail(MF_ERR_UNSUPPORTED_TYPE, "dtype must be INT8 or UINT8");
    float *d_out = nullptr;
    uint8_t *d_in = nullptr;
    MF_CUDA(cudaMalloc(&d_in, n + 4));
    MF_CUDA(cudaMalloc(&d_out, n * 4 + 4));
    cudaError_t e = cudaMemcpy(d_in, in, n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_dequantize(d_in, d_out, n, scale, (float)zero_point, dtype == MF_DTYPE_U8, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, n * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) return fail(MF_ERR_CUDA, cudaGetErrorString(e));
    return MF_OK;
}

}  // extern "C"
