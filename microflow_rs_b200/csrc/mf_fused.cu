// mf_fused.cu -- fused low-resolution stage: a chain of (depthwise 3x3 s1 SAME, C = 128) -> (pointwise 1x1, 128 -> 128) layer
// pairs in ONE persistent kernel (see mf_fused.h for the why; reference chain:
// microflow-macros/src/lib.rs:198-201, per-op semantics src/ops/depthwise_conv_2d.rs:56-101, conv_2d.rs:56-104).
//
// One CTA per SM, 768 threads = 2 TEAMS of 12 warps (<= 80 registers per thread; with 8-warp teams the kernel sat at 62 % issue
// utilisation on fixed-latency dependency stalls, profiles/r02a_fused_chain_v1.txt).  A team owns a UNIT of U samples (U * H * W <= 256 pixel rows) and takes it
// through every layer of the chain without leaving the SM; the two teams run on different units, unsynchronised, so one team's
// tensor-core latency and barrier waits are filled with the other's CUDA-core work.
//
//   shared memory   sB      n_pairs x 16 KB   pointwise weights, SWIZZLE_128B K-major images, resident for the whole kernel
//                   sA[t]   32 KB per team    A operand of the pointwise GEMM: row = pixel of the unit, 128 channel bytes, SWIZZLE_128B
//                   X[t]    36 KB per team    the unit's activations, pixel-major; pitch 128 B as loaded from HBM (layer 0), 144 B
//                                             afterwards (the epilogue writes one pixel per lane: 144 = 128 + 16 keeps its 16-byte
//                                             stores conflict-free, the depthwise lanes read consecutive words at any pitch)
//                   cbuf[t] 3.7 KB per team   constants of the team's current layer pair (double use: the depthwise part is refilled
//                                             for the next pair while the tensor core works, the pointwise part during the depthwise)
//   tensor memory   512 columns = 2 teams x 2 tiles (M = 128 rows each) x 128 int32 accumulator columns
//
// Per layer pair and team:
//   depthwise    thread = (sample, output column pair, 4-channel word), walks down the rows: 4 LDS + 4x4 byte transpose (8 PRMT) per
//                input row, IDP.4A against (w0,w1,w2,0) / (0,w0,w1,w2) = three taps of two outputs, input-stationary over the three
//                kernel rows, pre-biased accumulators and the packed exact f32 epilogue of dwconv3x3_pair_kernel (mf_kernels.cu);
//                rows -1 and H and window columns outside the image are the input zero-point (so kcorr = in_zp * sum(w) is uniform)
//   pointwise    fence.proxy.async + team barrier; one thread issues 4 x tcgen05.mma (K = 32 each) per 128-row tile, commit -> mbarrier
//   epilogue     the three warps of a TMEM lane quarter share its (tile, 32-column chunk) jobs: tcgen05.ld 32 lanes x 32 columns, requant4_biased with the pair's tables (broadcast
//                LDS.128), 16-byte stores into X -- or, for the last pair, 32-byte stores straight to HBM
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "mf_device.cuh"
#include "mf_fused.h"
#include "mf_kernels.h"
#include "mf_tc_ptx.cuh"

namespace mf {

namespace {

using namespace tcptx;

constexpr int kTeams = 2;
constexpr uint32_t kXPitch = 144;                 // bytes per pixel row of X after the first epilogue
constexpr uint32_t kWImg = 128 * 128;             // one pointwise weight image
constexpr uint32_t kATeam = 256 * 128;            // A operand of a team: two 128-row tiles
constexpr uint32_t kSmemLimit = 232448;

struct FusedParams {
    const uint8_t *in;
    uint8_t *out;
    const uint8_t *wimg;
    const uint8_t *consts;
    long long batch;
    int n_pairs, H, W, HW, U, JJ;
    FastDiv fd_jj;
    uint32_t n_units;
    uint32_t off_A, off_X, x_stride, off_cbuf, off_bar;
    uint32_t idesc;
    int dw_zp[kFusedMaxPairs];
    float dw_lo[kFusedMaxPairs], dw_hi[kFusedMaxPairs], pw_lo[kFusedMaxPairs], pw_hi[kFusedMaxPairs];
};

__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void team_sync(int team, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(threads) : "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <bool FULL, int kTeamWarps>
__global__ void __launch_bounds__(kTeams * 32 * kTeamWarps, 1) fused_chain_kernel(const __grid_constant__ FusedParams p) {
    constexpr int kTeamThreads = 32 * kTeamWarps;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t s0 = smem_u32(smem);
    if ((s0 & 1023u) != 0) __trap();              // SWIZZLE_128B atoms need a 1024-byte aligned base
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);       // warp-uniform for the compiler
    const int team = warp >= kTeamWarps ? 1 : 0, tw = warp - team * kTeamWarps, tt = tid - team * kTeamThreads;
    // barriers: [0] weights landed, [1 + t] unit input landed, [3 + t] tensor core done; TMEM base address behind them
    const uint32_t bar0 = s0 + p.off_bar;
    const uint32_t wbar = bar0, xbar = bar0 + 8u * (1 + team), mbar = bar0 + 8u * (3 + team);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + p.off_bar + 64);
    const uint32_t sB = s0, sA = s0 + p.off_A + (uint32_t)team * kATeam, sX = s0 + p.off_X + (uint32_t)team * p.x_stride;
    uint32_t *cbuf = reinterpret_cast<uint32_t *>(smem + p.off_cbuf + (size_t)team * kFusedConstBytes);

    if (tid == 0) {
        for (int k = 0; k < 5; ++k) mbar_init(bar0 + 8u * k, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // depthwise constants of the first pair (static data: may be read before the previous kernel has finished)
    for (int k = tt; k < kFusedDwWords / 4; k += kTeamThreads)
        reinterpret_cast<uint4 *>(cbuf)[k] = __ldg(reinterpret_cast<const uint4 *>(p.consts) + k);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    if (tid == 0) {
        mbar_expect_tx(wbar, (uint32_t)p.n_pairs * kWImg);
        bulk_g2s(sB, p.wimg, (uint32_t)p.n_pairs * kWImg, wbar);
    }

    const uint32_t ustride = gridDim.x * kTeams;
    const uint32_t u0 = blockIdx.x * kTeams + (uint32_t)team;
    const uint32_t unit_bytes = (uint32_t)(p.U * p.HW) * 128u;
    auto request_unit = [&](uint32_t u) {           // one thread: the unit's samples are contiguous in HBM
        const long long first = (long long)u * p.U;
        const long long ns = p.batch - first < p.U ? p.batch - first : p.U;
        const uint32_t bytes = (uint32_t)ns * (uint32_t)p.HW * 128u;
        mbar_expect_tx(xbar, bytes);
        bulk_g2s(sX, p.in + (size_t)u * unit_bytes, bytes, xbar);
    };
    pdl_wait();                                     // activations: first access (and first store) after the previous kernel completed
    if (tt == 0 && u0 < p.n_units) request_unit(u0);

    const uint32_t lane4 = (uint32_t)lane * 4u;
    const uint32_t a_lq = (uint32_t)lane >> 2, a_lr = sA + ((uint32_t)lane & 3u) * 4u;
    uint32_t xph = 0, mph = 0;
    bool weights_ready = false;
    for (uint32_t u = u0; u < p.n_units; u += ustride) {
        const long long first = (long long)u * p.U;
        const int nsamp = (int)(p.batch - first < p.U ? p.batch - first : p.U);
        const int rows = nsamp * p.HW;
        const int ntiles = (rows + 127) >> 7;
        const int ntask = nsamp * p.JJ;
        mbar_wait(xbar, xph);
        xph ^= 1u;
        for (int l = 0; l < p.n_pairs; ++l) {
            const bool last = l + 1 == p.n_pairs;
            const uint32_t pitch = l == 0 ? 128u : kXPitch;
            // pointwise tables of this pair: fetched now, parked in shared memory after the depthwise (visible after its barrier)
            uint4 pwc = make_uint4(0, 0, 0, 0);
            if (tt < kFusedPwWords / 4) pwc = __ldg(reinterpret_cast<const uint4 *>(p.consts + (size_t)l * kFusedConstBytes) + kFusedDwWords / 4 + tt);

            // ================= depthwise 3x3, stride 1, SAME: X -> A =================
            {
                const int zp = p.dw_zp[l];
                const uint32_t zpw = (uint32_t)(zp & 0xff) * 0x01010101u;
                uint32_t wa[3][4], wb[3][4];
                int fresh[4];
                {
                    int wsum[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int T = 0; T < 3; ++T) {
                        uint32_t wq[4];
                        transpose_3x4(cbuf[(3 * T) * 32 + lane], cbuf[(3 * T + 1) * 32 + lane], cbuf[(3 * T + 2) * 32 + lane], wq);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const uint32_t w3 = wq[c] & 0x00ffffffu;
                            wa[T][c] = w3;
                            wb[T][c] = w3 << 8;
                            wsum[c] = __dp4a((int)w3, 0x01010101, wsum[c]);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) fresh[c] = kAccBias - zp * wsum[c];     // pre-biased accumulator (mf_device.cuh) minus in_zp * sum(w)
                }
                const float lo = p.dw_lo[l], hi = p.dw_hi[l];
                const int H = p.H, W = p.W;
                const uint32_t rstep = (uint32_t)W * pitch;
                for (int task = tw; task < ntask; task += kTeamWarps) {
                    uint32_t s, jj;
                    p.fd_jj.divmod((uint32_t)task, s, jj);
                    const int j0 = 2 * (int)jj, c0 = j0 - 1;
                    const bool ok0 = c0 >= 0, ok2 = c0 + 2 < W, ok3 = c0 + 3 < W;       // window column c0 + 1 = j0 is always inside
                    // shared address of (input row 0, window column c0); for c0 = -1 one pixel before the sample (never dereferenced)
                    uint32_t xr = sX + (uint32_t)((int)s * p.HW + c0) * pitch + lane4;
                    uint32_t orow = s * (uint32_t)p.HW + (uint32_t)j0;                  // A row of output (0, j0)
                    auto take = [&](uint32_t (&t)[4]) {                                 // the next input row, transposed; zero-point left / right of the image
                        const uint32_t v0 = ok0 ? lds_u32(xr) : zpw;
                        const uint32_t v1 = lds_u32(xr + pitch);
                        const uint32_t v2 = ok2 ? lds_u32(xr + 2 * pitch) : zpw;
                        const uint32_t v3 = ok3 ? lds_u32(xr + 3 * pitch) : zpw;
                        xr += rstep;
                        transpose_4x4(v0, v1, v2, v3, t);
                    };
                    struct Acc { int a[4], b[4]; };
                    auto mac = [&](Acc &A, const uint32_t (&t)[4], int T) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            A.a[c] = __dp4a((int)t[c], (int)wa[T][c], A.a[c]);
                            A.b[c] = __dp4a((int)t[c], (int)wb[T][c], A.b[c]);
                        }
                    };
                    auto open = [&](Acc &A, const uint32_t (&t)[4]) {                    // a new output row: accumulators start at `fresh`, kernel row 0
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            A.a[c] = __dp4a((int)t[c], (int)wa[0][c], fresh[c]);
                            A.b[c] = __dp4a((int)t[c], (int)wb[0][c], fresh[c]);
                        }
                    };
                    auto store = [&](const Acc &A) {                                    // two pixels of the A operand (SWIZZLE_128B: 16-byte chunk ^= row & 7)
                        const float4 z = *reinterpret_cast<const float4 *>(cbuf + 9 * 32 + 4 * lane);
                        const float4 sc = *reinterpret_cast<const float4 *>(cbuf + 13 * 32 + 4 * lane);
                        const uint32_t y0 = requant4_biased<FULL>(A.a[0], A.a[1], A.a[2], A.a[3], z, sc, lo, hi);
                        const uint32_t y1 = requant4_biased<FULL>(A.b[0], A.b[1], A.b[2], A.b[3], z, sc, lo, hi);
                        sts_u32(a_lr + orow * 128u + ((a_lq ^ (orow & 7u)) << 4), y0);
                        if (ok2) sts_u32(a_lr + (orow + 1u) * 128u + ((a_lq ^ ((orow + 1u) & 7u)) << 4), y1);
                        orow += (uint32_t)W;
                    };
                    // rows -1 and H are the zero-point: row -1 is kernel row 0 of output row 0, row H kernel row 2 of output row H - 1
                    const uint32_t zt[4] = {zpw, zpw, zpw, zpw};
                    uint32_t t[4];
                    Acc A, B, C;
                    open(A, zt);
                    take(t); mac(A, t, 1); open(B, t);
                    int left = H - 1;                                                    // output rows closed by a real input row
                    while (left > 0) {
                        take(t); mac(A, t, 2); mac(B, t, 1); open(C, t); store(A); if (--left == 0) { A = B; break; }
                        take(t); mac(B, t, 2); mac(C, t, 1); open(A, t); store(B); if (--left == 0) { A = C; break; }
                        take(t); mac(C, t, 2); mac(A, t, 1); open(B, t); store(C); if (--left == 0) break;
                    }
                    mac(A, zt, 2);                                                       // the last output row: its kernel row 2 lies below the image
                    store(A);
                }
            }
            if (tt < kFusedPwWords / 4) reinterpret_cast<uint4 *>(cbuf + kFusedDwWords)[tt] = pwc;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // A was written through the generic proxy, the tensor core reads it through the async one
            team_sync(team, kTeamThreads);

            // ================= pointwise 1x1: tcgen05.mma, A (smem) x W^T (smem) -> TMEM =================
            if (tw == 0 && elect_one()) {     // elect.sync, not `tt == 0`: the MMAs are then issued back to back (mf_tc_ptx.cuh)
                if (last && u + ustride < p.n_units) request_unit(u + ustride);   // X is free: the last depthwise of this unit has read it
                if (!weights_ready) { mbar_wait(wbar, 0); weights_ready = true; }
                tc_fence_after();
                const uint64_t desc_hi = make_desc(0);
                const uint32_t b16 = (sB + (uint32_t)l * kWImg) >> 4;
                for (int tile = 0; tile < ntiles; ++tile) {
                    const uint32_t a16 = (sA + (uint32_t)tile * 16384u) >> 4;
                    const uint32_t d_tmem = tmem_base + (uint32_t)team * 256u + (uint32_t)tile * 128u;
#pragma unroll
                    for (uint32_t ks = 0; ks < 4; ++ks) tc_mma_i8(d_tmem, desc_hi | (uint64_t)(a16 + 2 * ks), desc_hi | (uint64_t)(b16 + 2 * ks), p.idesc, ks);
                }
                tc_commit(mbar);
            }
            // depthwise constants of the next pair (or of pair 0 for the next unit) while the tensor core works
            {
                const int ln = last ? 0 : l + 1;
                if (tt < kFusedDwWords / 4) reinterpret_cast<uint4 *>(cbuf)[tt] = __ldg(reinterpret_cast<const uint4 *>(p.consts + (size_t)ln * kFusedConstBytes) + tt);
            }
            // one warp polls the mbarrier, the others wait at the team barrier (a hardware wait: no issue slots burnt on polling)
            if (tw == 0) mbar_wait(mbar, mph);
            mph ^= 1u;
            team_sync(team, kTeamThreads);
            tc_fence_after();

            // ================= epilogue: TMEM -> exact f32 requantize -> X (or HBM for the last pair) =================
            {
                const uint32_t q = (uint32_t)tw & 3u;                       // == warp % 4 (teams start at a multiple of 4 warps): the TMEM lane quarter
                const float *tz = reinterpret_cast<const float *>(cbuf + kFusedDwWords), *ts = tz + 128;
                const int *tk = reinterpret_cast<const int *>(ts + 128);
                const float lo = p.pw_lo[l], hi = p.pw_hi[l];
#pragma unroll 1
                for (uint32_t job = (uint32_t)tw >> 2; job < (uint32_t)ntiles * 4u; job += kTeamWarps / 4) {
                    const uint32_t tile = job >> 2, ch = job & 3u;
                    const uint32_t row = tile * 128u + q * 32u + (uint32_t)lane;
                    const bool valid = (int)row < rows;
                    uint32_t r[32];
                    tmem_ld32(tmem_base + (uint32_t)team * 256u + tile * 128u + ((q * 32u) << 16) + ch * 32u, r);
                    uint32_t w[8];
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const float4 zz = *reinterpret_cast<const float4 *>(tz + ch * 32 + 4 * g);
                        const float4 ss = *reinterpret_cast<const float4 *>(ts + ch * 32 + 4 * g);
                        const int4 kk = *reinterpret_cast<const int4 *>(tk + ch * 32 + 4 * g);
                        w[g] = requant4_biased<FULL>((int)r[4 * g] + kk.x, (int)r[4 * g + 1] + kk.y, (int)r[4 * g + 2] + kk.z, (int)r[4 * g + 3] + kk.w, zz, ss, lo, hi);
                    }
                    if (last) {
                        if (valid)
                            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p.out + ((size_t)first * p.HW + row) * 128u + ch * 32u),
                                         "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                                         : "memory");
                    } else if (valid) {
                        const uint32_t xrow = sX + row * kXPitch + ch * 32u;
                        sts_v4(xrow, w[0], w[1], w[2], w[3]);
                        sts_v4(xrow + 16u, w[4], w[5], w[6], w[7]);
                    }
                }
            }
            tc_fence_before();
            team_sync(team, kTeamThreads);        // X complete for the next depthwise; TMEM and A free for the next pointwise
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

struct Layout {
    uint32_t off_A, off_X, x_stride, off_cbuf, off_bar, total;
};
Layout make_layout(int n_pairs, int H, int W) {
    Layout L{};
    const int U = 256 / (H * W);
    const uint32_t rows = (uint32_t)(U * H * W);
    L.off_A = (uint32_t)n_pairs * kWImg;
    L.off_X = L.off_A + kTeams * kATeam + 256;                       // 256 bytes of slack: window column -1 of the first pixel
    L.x_stride = ((rows + 3) * kXPitch + 15u) & ~15u;                // + window columns W, W + 1 of the last pixel
    L.off_cbuf = L.off_X + kTeams * L.x_stride;
    L.off_bar = L.off_cbuf + kTeams * kFusedConstBytes;
    L.total = L.off_bar + 128;
    return L;
}

}  // namespace

void fused_pack_pointwise_image(const uint8_t *w, uint8_t *img) {
    for (int n = 0; n < 128; ++n)
        for (int k = 0; k < 128; ++k) img[n * 128 + (((k >> 4) ^ (n & 7)) << 4) + (k & 15)] = w[n * 128 + k];
}

void fused_pack_consts(const uint8_t *dw_w, const float *dw_c0z, const float *dw_c1, const float *pw_c0z, const float *pw_c1, const int32_t *pw_kcorr, uint8_t *out) {
    uint32_t *o = reinterpret_cast<uint32_t *>(out);
    std::memcpy(o, dw_w, 9 * 128);                                   // [tap][C] bytes = word (tap, g) at tap * 32 + g
    std::memcpy(o + 9 * 32, dw_c0z, 128 * 4);
    std::memcpy(o + 13 * 32, dw_c1, 128 * 4);
    std::memcpy(o + kFusedDwWords, pw_c0z, 128 * 4);
    std::memcpy(o + kFusedDwWords + 128, pw_c1, 128 * 4);
    for (int n = 0; n < 128; ++n) reinterpret_cast<int32_t *>(o + kFusedDwWords + 256)[n] = kAccBias - pw_kcorr[n];
}

size_t fused_chain_smem(int n_pairs, int H, int W) {
    if (n_pairs < 1 || n_pairs > kFusedMaxPairs || H < 1 || W < 1 || H * W > 128) return 0;
    const Layout L = make_layout(n_pairs, H, W);
    return L.total <= kSmemLimit ? L.total : 0;
}

bool fused_chain_finalize(FusedChainPlan &p, std::string *why) {
    auto no = [&](const char *m) { if (why) *why = m; return false; };
    if (p.n_pairs < 1 || p.n_pairs > kFusedMaxPairs) return no("chain length out of range");
    if (p.H < 1 || p.W < 1 || p.H * p.W > 128) return no("feature map too large for a 256-row unit");
    p.U = 256 / (p.H * p.W);
    p.smem_bytes = fused_chain_smem(p.n_pairs, p.H, p.W);
    if (!p.smem_bytes) return no("weights + activations do not fit in 227 KB of shared memory");
    p.full_clamp = true;
    for (int l = 0; l < p.n_pairs; ++l)
        p.full_clamp = p.full_clamp && p.dw_lo[l] == -128.f && p.dw_hi[l] == 127.f && p.pw_lo[l] == -128.f && p.pw_hi[l] == 127.f;
    return true;
}

cudaError_t fused_chain_launch(const FusedChainPlan &pl, const uint8_t *in, uint8_t *out, long long batch, int num_sms, cudaStream_t s, int pdl) {
    if (batch <= 0) return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(in) & 15u) != 0 || (reinterpret_cast<uintptr_t>(out) & 31u) != 0) return cudaErrorMisalignedAddress;
    const Layout L = make_layout(pl.n_pairs, pl.H, pl.W);
    FusedParams p{};
    p.in = in; p.out = out; p.wimg = pl.d_wimg; p.consts = pl.d_consts; p.batch = batch;
    p.n_pairs = pl.n_pairs; p.H = pl.H; p.W = pl.W; p.HW = pl.H * pl.W; p.U = pl.U; p.JJ = (pl.W + 1) / 2;
    p.fd_jj = FastDiv((uint32_t)p.JJ);
    const long long units = (batch + pl.U - 1) / pl.U;
    if (units >= (1ll << 31)) return cudaErrorInvalidValue;
    p.n_units = (uint32_t)units;
    p.off_A = L.off_A; p.off_X = L.off_X; p.x_stride = L.x_stride; p.off_cbuf = L.off_cbuf; p.off_bar = L.off_bar;
    // instruction descriptor (same fields as mf_conv_tc.cu): c S32, a/b signed int8, K-major, N = 128, M = 128
    p.idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    for (int l = 0; l < pl.n_pairs; ++l) {
        p.dw_zp[l] = pl.dw_zp[l];
        p.dw_lo[l] = pl.dw_lo[l]; p.dw_hi[l] = pl.dw_hi[l]; p.pw_lo[l] = pl.pw_lo[l]; p.pw_hi[l] = pl.pw_hi[l];
    }
    using Fn = void (*)(const FusedParams);
    // warps per team: 8 (119 registers per thread) or 12 (80); MF_FUSED_TW selects for experiments
    static const int env_tw = [] { const char *e = std::getenv("MF_FUSED_TW"); return e ? std::atoi(e) : 8; }();
    const int tw = env_tw == 12 ? 12 : 8;
    Fn fn = tw == 12 ? (pl.full_clamp ? fused_chain_kernel<true, 12> : fused_chain_kernel<false, 12>)
                     : (pl.full_clamp ? fused_chain_kernel<true, 8> : fused_chain_kernel<false, 8>);
    static std::mutex mu;
    static std::vector<std::pair<int, Fn>> done;
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        std::lock_guard<std::mutex> lock(mu);
        bool have = false;
        for (auto &d : done) have = have || (d.first == dev && d.second == fn);
        if (!have) {
            e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit);
            if (e != cudaSuccess) return e;
            done.emplace_back(dev, fn);
        }
    }
    const long long ctas = (units + kTeams - 1) / kTeams;
    const unsigned grid = (unsigned)(ctas < num_sms ? ctas : num_sms);
    return launch_pdl(fn, dim3(grid), dim3((unsigned)(kTeams * 32 * tw)), L.total, s, pdl, p);
}

}  // namespace mf
