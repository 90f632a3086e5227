// mf_conv_tc.cu -- hand-written tcgen05 int8 implicit-GEMM Conv2D for sm_100a (B200).
//
// Replaces src/ops/conv_2d.rs:28-108 for the shapes that are worth a tensor core:
//   * 3x3 stride-1 SAME convolutions with Cin a multiple of 128  (BASELINE config 5: 224x224x128 -> 128)
//   * every 1x1 convolution of person_detect but the last (256 -> 2), as a "packed pixel" GEMM (see mf_conv_tc.h)
//
// Pipeline inside one persistent CTA (640 threads, 1 CTA / SM):
//   warps 0..15        epilogue     : 16 warps = 4 TMEM lane quarters x 4 column groups; tcgen05.ld 32x32b -> registers
//                                     -> exact f32 requantize (mf_device.cuh) -> int8 pack -> 16-byte global stores.
//                                     (The epilogue is ~8 ALU instructions per output value; one warp per scheduler cannot
//                                     hide its own latencies, four can, and the MMA of the next tile runs underneath.)
//   warp 16            TMEM alloc / dealloc
//   warp 18 (one lane) TMA producer : cp.async.bulk.tensor.4d  global -> smem ring (SWIZZLE_128B), mbarrier tx
//   warp 19 (one lane) MMA issuer   : tcgen05.mma.cta_group::1.kind::i8, D in TMEM (2 accumulator buffers)
// The two single-lane roles are the HIGHEST warp ids on purpose: the warp scheduler prefers the highest eligible warp id
// (B300_MICROARCH.md), so the MMA issuer is never starved of issue slots by the four epilogue warps on its scheduler.
// The weights (B operand, <= 147 KB) are loaded once per CTA and stay in shared memory.
//
// Arithmetic: int32 accumulation is exact and order-independent, so any tiling is bit-identical to the
// reference; zero-filled out-of-bounds taps (TMA OOB fill) reproduce the reference's zero-filled view
// (src/tensor.rs:196-218) and the `in_zp * masked filter-sum` term (conv_2d.rs:83-89) becomes a per-border-class,
// per-channel int32 table subtracted before the f32 epilogue.  Weight zero-points must be 0 on this path.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <utility>
#include <vector>

#include "mf_conv_tc.h"
#include "mf_device.cuh"
#include "mf_kernels.h"   // FastDiv
#include "mf_tc_ptx.cuh"

namespace mf {

namespace {

using namespace tcptx;

constexpr int kMaxStages = 8;
constexpr int kEpiWarps = 16;
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr int kWarpAlloc = kEpiWarps, kWarpTma = kEpiWarps + 2, kWarpMma = kEpiWarps + 3;
constexpr uint32_t kSmemLimit = 232448;  // 227 KB usable per CTA on sm_100

// Per-channel epilogue tables, passed by value as a kernel parameter, i.e. they live in the constant bank.  How the
// epilogue reads them is the kernel's TAB mode:
//   kTabSmem    copied once per CTA into shared memory, read with broadcast LDS.128 (0.75 per output value; the MIO queue
//               this fills is the top stall of the big pointwise layers, profiles/r01e)
//   kTabPeriod  packed-pointwise layers with Cout | 32: every 32-column chunk sees the same 32 table entries, so ONE copy
//               of the epilogue code has compile-time table indices and the values arrive as uniform-register / constant
//               operands of the FMUL / FADD / IADD that use them -- no LDS at all
//   kTabGroup   N <= 128: one code copy per column group (4 x 32 columns), same idea; with N = 256 the four copies (two
//               chunks each) no longer fit the instruction cache of a sub-partition and this is slower than kTabSmem
// (A fourth mode -- kTabGroup for the 3x3 kernel with the eight non-interior border classes kept in shared memory as deltas -- removed
// the epilogue's 32 % share of the L1/shared data pipe and was still slower, 0.138 vs 0.103 ms on BASELINE config 5; it is not kept:
// profiles/r01j_conv3x3_experiments.txt.)
enum { kTabSmem = 0, kTabGroup = 1, kTabPeriod = 2, kTabGroupPB = 3, kTabPeriodPB = 4 };   // PB: + pre-biased accumulators in TMEM (below)
struct ConvTcTables {
    float c0z[256];
    float c1[256];
    int32_t corr[9 * 256];   // [ncls][N] with row pitch N
};

struct ConvTcParams {
    uint8_t *out;
    int N, CB, KH, KW, TW, TH, tw_log2, off_r, off_c, ncls, stages;
    int tiles_x, tiles_y;
    FastDiv fd_img, fd_tx;     // tile -> (image, ty, tx) without integer division
    long long num_tiles, OW, OH;
    float lo, hi;
    uint32_t idesc, stage_bytes, b_block_bytes, tmem_cols, nkb;
    uint32_t stage_tx;          // bytes one stage's TMA load delivers (== stage_bytes unless the stage is padded to 1024)
    uint32_t acc_mask, acc_shift;   // TMEM accumulator ring: nacc = acc_mask + 1 = 1 << acc_shift buffers (2, or 4 when 4 N <= 512 and MF_TC_NACC=4)
    int early;                  // release the accumulator right after tcgen05.ld (MF_TC_EARLY, default 1)
    int out_u8;                 // outputs are uint8 (F2IP.U8 in the general epilogue)
    uint32_t patch, patch_w;    // single-patch mode (ConvTcPlan::patch) and its patch width TW + KW - 1 in pixels
};

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
// KH_T/KW_T/CB_T != 0: compile-time loop bounds, so the single MMA-issuing thread spends ~3 scalar instructions per MMA
// (descriptor = base + constant); XU = F2I.S8 / I2F epilogue (needs the full int8 clamp range) instead of the XU-free one;
// TAB = how the epilogue reads its per-channel tables (above).
template <bool BIG, bool XU, int KH_T, int KW_T, int CB_T, int TABM>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ ConvTcTables tab,
               const ConvTcParams p) {
    constexpr int TAB = TABM == kTabGroupPB ? kTabGroup : (TABM == kTabPeriodPB ? kTabPeriod : TABM);
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();   // SWIZZLE_128B atoms need a 1024-byte aligned base
    uint8_t *sB = smem;
    uint8_t *sA = sB + (size_t)p.nkb * p.b_block_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sA + (size_t)p.stages * p.stage_bytes);   // 256 bytes reserved
    float *s_c0z = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(bars) + 256);          // kTabSmem only
    float *s_c1 = s_c0z + p.N;
    int32_t *s_corr = reinterpret_cast<int32_t *>(s_c1 + p.N);                                 // [ncls][N]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kMaxStages + 9);

    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (kMaxStages + s); };
    const uint32_t bfull_bar = bar0 + 8u * (2 * kMaxStages);
    auto tfull_bar = [&](uint32_t a) { return bar0 + 8u * (2 * kMaxStages + 1 + a); };
    auto tempty_bar = [&](uint32_t a) { return bar0 + 8u * (2 * kMaxStages + 5 + a); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KH = KH_T ? KH_T : p.KH, KW = KW_T ? KW_T : p.KW, CB = CB_T ? CB_T : p.CB;
    // the 1x1 instantiations are only launched on a single row of (packed) pixels (TW = 128, H = B = 1; conv_tc_launch checks):
    // a tile index is its x coordinate and there are no border classes
    constexpr bool LINEAR = KH_T == 1 && KW_T == 1;
    // Pre-biased accumulators in TMEM (constant-operand table modes of the packed pointwise GEMM): every epilogue warp owns fixed
    // accumulator columns whose correction kAccBias - kcorr[column] is the same for every tile, so it keeps those 32 words in
    // registers and writes them back into TMEM (one tcgen05.st) right after reading an accumulator; the MMAs of the next tile then
    // ACCUMULATE onto them and the epilogue needs neither the integer add per value nor the table operand for it.
    constexpr bool PREBIAS = (TABM == kTabPeriodPB || TABM == kTabGroupPB) && XU && !BIG;

    if (warp == kWarpTma && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
    }
    if (warp == kWarpMma && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(bfull_bar, 1);
        for (uint32_t a = 0; a < 4; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWarpAlloc) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (TAB == kTabSmem && warp < kEpiWarps) {
        for (int k = threadIdx.x; k < p.N; k += 32 * kEpiWarps) { s_c0z[k] = tab.c0z[k]; s_c1[k] = tab.c1[k]; }
        for (int k = threadIdx.x; k < p.ncls * p.N; k += 32 * kEpiWarps) s_corr[k] = tab.corr[k];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();

    if (warp == kWarpTma) {
        if (elect_one()) {
            // ===== TMA producer =====
            mbar_expect_tx(bfull_bar, p.nkb * p.b_block_bytes);
            for (uint32_t kb = 0; kb < p.nkb; ++kb) tma_load_2d(smem_u32(sB + (size_t)kb * p.b_block_bytes), &tmap_b, bfull_bar, (int)(kb * 128), 0);
            pdl_wait();       // the weights above are static; the activations below are the previous kernel's output
            uint32_t s = 0, ph = 0;
            for (long long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                uint32_t b = 0, rem, ty = 0, tx = (uint32_t)tile;
                if (!LINEAR) {
                    p.fd_img.divmod((uint32_t)tile, b, rem);
                    p.fd_tx.divmod(rem, ty, tx);
                }
                for (int n = 0; n < (p.patch ? 1 : KW); ++n)      // patch mode: one load per channel block brings every tap's data
                    for (int cb = 0; cb < CB; ++cb) {
                        mbar_wait(empty_bar(s), ph ^ 1);
                        mbar_expect_tx(full_bar(s), p.stage_tx);
                        tma_load_4d(smem_u32(sA + (size_t)s * p.stage_bytes), &tmap_a, full_bar(s), cb * 128, (int)tx * p.TW + n - p.off_c,
                                    (int)ty * p.TH - p.off_r, (int)b);
                        if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
                    }
            }
        }
    } else if (warp == kWarpMma) {
        if (elect_one()) {
            // ===== MMA issuer =====
            mbar_wait(bfull_bar, 0);
            tc_fence_after();
            uint32_t s = 0, ph = 0, it = 0;
            // descriptors differ only in their 14-bit start-address field: keep the constant part, add (bytes >> 4)
            const uint64_t desc_hi = make_desc(0);
            const uint32_t a0 = smem_u32(sA) >> 4, b0 = smem_u32(sB) >> 4;
            const uint32_t stage16 = p.stage_bytes >> 4, bblk16 = p.b_block_bytes >> 4, arow16 = (uint32_t)p.TW * 8u;   // TW rows * 128 B / 16
            if (a0 + (uint32_t)p.stages * stage16 >= (1u << 14) || b0 + p.nkb * bblk16 >= (1u << 14)) __trap();     // descriptor start field would overflow
            for (long long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const uint32_t acc = it & p.acc_mask, aph = (it >> p.acc_shift) & 1;
                mbar_wait(tempty_bar(acc), PREBIAS ? aph : (aph ^ 1));     // PREBIAS: phase 0 of each buffer is the epilogue warps' initial fill
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.N;
                uint32_t accumulate = PREBIAS ? 1 : 0;
                if (p.patch) {
                    // one stage per channel block holds the whole patch: tap (m, n) starts (m * patch_w + n) pixel rows into it and
                    // the 8-row groups of the tile (TW == 8: one tile row each) are patch_w rows apart (SBO)
                    const uint64_t desc_a = make_desc(0, p.patch_w * 128u);
                    for (int cb = 0; cb < CB; ++cb) {
                        mbar_wait(full_bar(s), ph);
                        tc_fence_after();
                        const uint64_t adesc = desc_a | (uint64_t)(a0 + s * stage16);
                        const uint64_t bdesc = desc_hi | (uint64_t)(b0 + (uint32_t)cb * bblk16);
#pragma unroll
                        for (int m = 0; m < KH; ++m)
#pragma unroll
                            for (int n = 0; n < KW; ++n)
#pragma unroll
                                for (uint32_t ks = 0; ks < 4; ++ks) {
                                    tc_mma_i8(d_tmem, adesc + (uint64_t)(((uint32_t)m * p.patch_w + (uint32_t)n) * 8u + 2 * ks),
                                              bdesc + (uint64_t)((uint32_t)((m * KW + n) * CB) * bblk16 + 2 * ks), p.idesc, accumulate);
                                    accumulate = 1;
                                }
                        tc_commit(empty_bar(s));
                        if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
                    }
                } else
#pragma unroll
                for (int n = 0; n < KW; ++n)
#pragma unroll
                    for (int cb = 0; cb < CB; ++cb) {
                        mbar_wait(full_bar(s), ph);
                        tc_fence_after();
                        const uint64_t adesc = desc_hi | (uint64_t)(a0 + s * stage16);                       // start field never carries:
                        const uint64_t bdesc = desc_hi | (uint64_t)(b0 + (uint32_t)(n * CB + cb) * bblk16);  // smem < 256 KB -> (addr >> 4) < 2^14
#pragma unroll
                        for (int m = 0; m < KH; ++m) {
#pragma unroll
                            for (uint32_t ks = 0; ks < 4; ++ks) {
                                tc_mma_i8(d_tmem, adesc + (uint64_t)((uint32_t)m * arow16 + 2 * ks), bdesc + (uint64_t)((uint32_t)(m * KW * CB) * bblk16 + 2 * ks),
                                          p.idesc, accumulate);
                                accumulate = 1;
                            }
                        }
                        tc_commit(empty_bar(s));  // smem slot is free once these MMAs retire
                        if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
                    }
                tc_commit(tfull_bar(acc));        // accumulator complete -> epilogue
            }
        }
    } else if (warp < kEpiWarps) {
        // ===== epilogue: warp e = (lane quarter q, column group cg); chunk c of 32 columns belongs to group c % 4 =====
        // cg_tag: the column group as a compile-time constant (kTabGroup), else -1 and the group is warp >> 2.
        auto epilogue = [&](auto cg_tag) {
            constexpr int CG = decltype(cg_tag)::value;
            const uint32_t q = (uint32_t)(warp & 3);               // == warp % 4: the TMEM lane quarter this warp may read
            const int cg = CG >= 0 ? CG : (warp >> 2);
            const int row = (int)(q * 32 + lane);
            const int rr = row >> p.tw_log2, rc = row & (p.TW - 1);
            const float lo = p.lo, hi = p.hi;
            uint32_t kkr[32];
            if (PREBIAS) {
#pragma unroll
                for (int j = 0; j < 32; ++j) kkr[j] = (uint32_t)tab.corr[(TAB == kTabPeriod ? 0 : 32 * CG) + j];
                for (uint32_t a = 0; a < 2; ++a) {                  // both accumulator buffers start out holding the correction
                    for (int c0 = 32 * cg; c0 < p.N; c0 += (TAB == kTabGroup ? 1 << 20 : 128)) tmem_st32(tmem_base + a * (uint32_t)p.N + ((q * 32u) << 16) + (uint32_t)c0, kkr);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(tempty_bar(0)); mbar_arrive(tempty_bar(1)); }
            }
            pdl_wait();       // stores must not overtake the previous kernel's reads of the ping-pong buffer
            uint32_t it = 0;
            // LINEAR: this thread's row and output pointer advance by a constant per tile (no multiplies in the tile loop)
            uint32_t lin_ox = blockIdx.x * 128u + (uint32_t)row;
            uint8_t *lin_row = p.out + (size_t)lin_ox * (size_t)p.N;
            const uint32_t lin_step = gridDim.x * 128u;
            const size_t lin_bytes = (size_t)lin_step * (size_t)p.N;
            const uint32_t ntiles = (uint32_t)p.num_tiles;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const uint32_t acc = it & p.acc_mask, aph = (it >> p.acc_shift) & 1;
                bool valid;
                const int32_t *corr = s_corr;
                uint8_t *orow;
                if (LINEAR) {                                       // one row of 128-byte (packed) pixels: tile t covers rows [128 t, 128 t + 128)
                    valid = lin_ox < (uint32_t)p.OW;                // conv_tc_launch keeps OW below 2^31 for the LINEAR instantiations
                    orow = lin_row;
                    lin_ox += lin_step;
                    lin_row += lin_bytes;
                } else {
                    uint32_t b, rem, ty, tx;
                    p.fd_img.divmod((uint32_t)tile, b, rem);
                    p.fd_tx.divmod(rem, ty, tx);
                    const long long oy = (long long)ty * p.TH + rr, ox = (long long)tx * p.TW + rc;
                    valid = oy < p.OH && ox < p.OW;
                    int cls = 0;
                    if (TAB == kTabSmem && p.ncls == 9) cls = 3 * (oy == 0 ? 0 : (oy == p.OH - 1 ? 2 : 1)) + (ox == 0 ? 0 : (ox == p.OW - 1 ? 2 : 1));
                    corr = s_corr + cls * p.N;
                    orow = p.out + (((long long)b * p.OH + oy) * p.OW + ox) * p.N;
                }

                mbar_wait(tfull_bar(acc), aph);
                tc_fence_after();
                const uint32_t t_base = tmem_base + acc * (uint32_t)p.N + ((q * 32u) << 16);
                constexpr bool ONE_CHUNK = TAB == kTabGroup;      // N <= 128: a warp's only chunk
                for (int c0 = 32 * cg; c0 < p.N; c0 += (ONE_CHUNK ? 1 << 20 : 128)) {
                    uint32_t r[32];
                    tmem_ld32(t_base + (uint32_t)c0, r);
                    if (PREBIAS) tmem_st32(t_base + (uint32_t)c0, kkr);     // re-arm the accumulator for the tile after next
                    if (!PREBIAS && p.early && (ONE_CHUNK || c0 + 128 >= p.N)) {
                        // the warp's last chunk of this accumulator is in registers: hand the buffer back BEFORE the math and the stores
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(tempty_bar(acc));
                    }
                    uint32_t w[8];
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        // PACKED (XU && !BIG): the correction table holds kAccBias - corr, so acc + table is the pre-biased
                        // accumulator of requant4_biased (packed FADD2 epilogue, no I2F); otherwise table = corr
                        constexpr bool PACKED = XU && !BIG;
                        float zz[4], ss[4];
                        int kk[4];
                        if (TAB == kTabSmem) {
                            const float4 z = *reinterpret_cast<const float4 *>(s_c0z + c0 + 4 * g);
                            const float4 sc = *reinterpret_cast<const float4 *>(s_c1 + c0 + 4 * g);
                            const int4 kc = *reinterpret_cast<const int4 *>(corr + c0 + 4 * g);
                            zz[0] = z.x; zz[1] = z.y; zz[2] = z.z; zz[3] = z.w;
                            ss[0] = sc.x; ss[1] = sc.y; ss[2] = sc.z; ss[3] = sc.w;
                            kk[0] = kc.x; kk[1] = kc.y; kk[2] = kc.z; kk[3] = kc.w;
                        } else {
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                // kTabPeriod: the tables repeat every 32 columns; kTabGroup: N <= 128, this warp's only chunk is 32 * CG
                                const int n = (TAB == kTabPeriod ? 0 : 32 * CG) + 4 * g + u;
                                zz[u] = tab.c0z[n]; ss[u] = tab.c1[n]; kk[u] = PREBIAS ? 0 : tab.corr[n];
                            }
                        }
                        if (PACKED) {
                            w[g] = requant4_biased<true>((int)r[4 * g] + kk[0], (int)r[4 * g + 1] + kk[1], (int)r[4 * g + 2] + kk[2], (int)r[4 * g + 3] + kk[3],
                                                         make_float4(zz[0], zz[1], zz[2], zz[3]), make_float4(ss[0], ss[1], ss[2], ss[3]), lo, hi);
                        } else if (XU) {        // any accumulator magnitude: I2F, then the same packed adds
                            w[g] = requant4_i2f((int)r[4 * g] - kk[0], (int)r[4 * g + 1] - kk[1], (int)r[4 * g + 2] - kk[2], (int)r[4 * g + 3] - kk[3],
                                                make_float4(zz[0], zz[1], zz[2], zz[3]), make_float4(ss[0], ss[1], ss[2], ss[3]));
                        } else {                // any clamp range, int8 or uint8 outputs
                            w[g] = requant4_clamp<BIG>((int)r[4 * g] - kk[0], (int)r[4 * g + 1] - kk[1], (int)r[4 * g + 2] - kk[2], (int)r[4 * g + 3] - kk[3],
                                                       make_float4(zz[0], zz[1], zz[2], zz[3]), make_float4(ss[0], ss[1], ss[2], ss[3]), lo, hi, p.out_u8 != 0);
                        }
                    }
                    if (valid)     // the thread's 32 output bytes = one aligned 32-byte sector: a single 256-bit store (STG.256), so L2 sees
                                   // full-sector writes (two 16-byte halves per sector kept L2 at 68 % of its request rate, profiles/r01l_tc_pw.txt)
                        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow + c0), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                                     "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                                     : "memory");
                }
                if (PREBIAS) tmem_st_wait();
                if (PREBIAS || !p.early) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(tempty_bar(acc));
                }
            }
        };
        if (TAB == kTabGroup) {
            switch (warp >> 2) {
                case 0: epilogue(std::integral_constant<int, 0>{}); break;
                case 1: epilogue(std::integral_constant<int, 1>{}); break;
                case 2: epilogue(std::integral_constant<int, 2>{}); break;
                default: epilogue(std::integral_constant<int, 3>{}); break;
            }
        } else {
            epilogue(std::integral_constant<int, -1>{});
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kWarpAlloc) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// 3x3 convolution on CTA PAIRS (tcgen05.mma.cta_group::2)
// ------------------------------------------------------------------------------------------------
// Why: the one-CTA M128 x N128 x K32 int8 instruction reads 4 KB of A and 4 KB of B from shared memory every 66 clocks -- 124 of the
// 128 B/clk the shared-memory port delivers (tools/ubench/mma_i8.cu: 4 423 TOP/s MMA-only) -- so every TMA write and epilogue table
// load steals tensor-core time; the kernel above stops at 60 % of that ceiling.  A pair of CTAs on one TPC shares the B operand:
// each keeps the weights of HALF of the output channels (72 KB instead of 147 KB: six pipeline stages instead of three), one thread of
// the even CTA issues M256 x N128 x K32 instructions that run on both SMs' tensor cores, and each SM reads 4 KB of A + 2 KB of B per
// instruction.  Everything else is the one-CTA kernel's: single-patch A staging by TMA (every tap a descriptor offset), zero-filled
// borders + per-class correction table, 16 epilogue warps per CTA with the exact f32 requantize.
//   barriers the LEADER waits on are fed by both CTAs: full[s] (TMA transaction bytes of both patches, cp.async.bulk.tensor
//   .cta_group::2 with the leader's mbarrier), bfull (both weight halves), tempty[a] (2 x 16 epilogue warps, the odd CTA's by
//   mapa + remote arrive); completions are multicast: tcgen05.commit.cta_group::2...multicast::cluster on empty[s] and tfull[a].
template <bool BIG, bool XU>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ ConvTcTables tab,
                    const ConvTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    const uint32_t NH = (uint32_t)p.N >> 1;                                   // output channels whose weights this CTA holds
    const uint32_t bblk = NH * 128u;                                          // one tap's weight block: [N/2][128 B]
    uint8_t *sB = smem;
    uint8_t *sA = sB + 9u * bblk;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sA + (size_t)p.stages * p.stage_bytes);
    float *s_c0z = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(bars) + 256);
    float *s_c1 = s_c0z + p.N;
    int32_t *s_corr = reinterpret_cast<int32_t *>(s_c1 + p.N);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kMaxStages + 9);

    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (kMaxStages + s); };
    const uint32_t bfull_bar = bar0 + 8u * (2 * kMaxStages);
    auto tfull_bar = [&](uint32_t a) { return bar0 + 8u * (2 * kMaxStages + 1 + a); };
    auto tempty_bar = [&](uint32_t a) { return bar0 + 8u * (2 * kMaxStages + 5 + a); };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();                                  // 0 = leader (issues the MMAs)
    const uint32_t npairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
    const uint32_t ntiles = (uint32_t)p.num_tiles;

    if (warp == kWarpTma && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
    }
    if (warp == kWarpMma && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }
        mbar_init(bfull_bar, 2);
        for (uint32_t a = 0; a < 4; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kWarpAlloc) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    if (warp < kEpiWarps) {
        for (int k = threadIdx.x; k < p.N; k += 32 * kEpiWarps) { s_c0z[k] = tab.c0z[k]; s_c1[k] = tab.c1[k]; }
        for (int k = threadIdx.x; k < p.ncls * p.N; k += 32 * kEpiWarps) s_corr[k] = tab.corr[k];
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                      // both CTAs' barriers exist before any remote arrive / transaction
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();

    // tile of this CTA in round t: 2 * t + rank; the odd CTA of the last pair may have none (it then recomputes the last tile and
    // stores nothing, so that the pair's barrier protocol stays symmetric)
    auto tile_of = [&](uint32_t t, bool &valid) {
        const uint32_t mine = 2u * t + rank;
        valid = mine < ntiles;
        return valid ? mine : ntiles - 1u;
    };

    if (warp == kWarpTma) {
        if (elect_one()) {
            // ===== TMA producer (both CTAs; transaction bytes are counted on the leader's barriers) =====
            const uint32_t l_bfull = mapa_rank(bfull_bar, 0);
            mbar_expect_tx_cluster(l_bfull, 9u * bblk);
            for (uint32_t kb = 0; kb < 9; ++kb) tma_load_2d_pair(smem_u32(sB + (size_t)kb * bblk), &tmap_b, l_bfull, (int)(kb * 128), (int)(rank * NH));
            pdl_wait();
            uint32_t s = 0, ph = 0;
            for (uint32_t t = pair; 2u * t < ntiles; t += npairs) {
                bool valid;
                const uint32_t tile = tile_of(t, valid);
                uint32_t b, rem, ty, tx;
                p.fd_img.divmod(tile, b, rem);
                p.fd_tx.divmod(rem, ty, tx);
                mbar_wait(empty_bar(s), ph ^ 1);
                const uint32_t l_full = mapa_rank(full_bar(s), 0);
                mbar_expect_tx_cluster(l_full, p.stage_tx);
                tma_load_4d_pair(smem_u32(sA + (size_t)s * p.stage_bytes), &tmap_a, l_full, 0, (int)tx * p.TW - p.off_c, (int)ty * p.TH - p.off_r, (int)b);
                if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == kWarpMma) {
        if (rank == 0 && elect_one()) {
            // ===== MMA issuer (leader CTA only) =====
            mbar_wait(bfull_bar, 0);
            tc_fence_after();
            uint32_t s = 0, ph = 0, it = 0;
            const uint64_t desc_hi = make_desc(0);
            const uint64_t desc_a = make_desc(0, p.patch_w * 128u);
            const uint32_t a0 = smem_u32(sA) >> 4, b0 = smem_u32(sB) >> 4;
            const uint32_t stage16 = p.stage_bytes >> 4, bblk16 = bblk >> 4;
            for (uint32_t t = pair; 2u * t < ntiles; t += npairs, ++it) {
                const uint32_t acc = it & p.acc_mask, aph = (it >> p.acc_shift) & 1;
                mbar_wait(tempty_bar(acc), aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (uint32_t)p.N;
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint64_t adesc = desc_a | (uint64_t)(a0 + s * stage16);
                const uint64_t bdesc = desc_hi | (uint64_t)b0;
                uint32_t accumulate = 0;
#pragma unroll
                for (int m = 0; m < 3; ++m)
#pragma unroll
                    for (int n = 0; n < 3; ++n)
#pragma unroll
                        for (uint32_t ks = 0; ks < 4; ++ks) {
                            tc_mma_i8_pair(d_tmem, adesc + (uint64_t)(((uint32_t)m * p.patch_w + (uint32_t)n) * 8u + 2 * ks),
                                           bdesc + (uint64_t)((uint32_t)(m * 3 + n) * bblk16 + 2 * ks), p.idesc, accumulate);
                            accumulate = 1;
                        }
                tc_commit_pair(empty_bar(s));            // both CTAs' slot s is free once these MMAs retire
                if (++s == (uint32_t)p.stages) { s = 0; ph ^= 1; }
                tc_commit_pair(tfull_bar(acc));          // both CTAs' accumulator halves are complete
            }
        }
    } else if (warp < kEpiWarps) {
        // ===== epilogue (both CTAs, each on its own 128 accumulator rows = its own tile) =====
        const uint32_t q = (uint32_t)(warp & 3);
        const int cg = warp >> 2;
        const int row = (int)(q * 32 + lane);
        const int rr = row >> p.tw_log2, rc = row & (p.TW - 1);
        pdl_wait();
        uint32_t it = 0;
        for (uint32_t t = pair; 2u * t < ntiles; t += npairs, ++it) {
            const uint32_t acc = it & p.acc_mask, aph = (it >> p.acc_shift) & 1;
            bool tvalid;
            const uint32_t tile = tile_of(t, tvalid);
            uint32_t b, rem, ty, tx;
            p.fd_img.divmod(tile, b, rem);
            p.fd_tx.divmod(rem, ty, tx);
            const long long oy = (long long)ty * p.TH + rr, ox = (long long)tx * p.TW + rc;
            const bool valid = tvalid && oy < p.OH && ox < p.OW;
            const int cls = 3 * (oy == 0 ? 0 : (oy == p.OH - 1 ? 2 : 1)) + (ox == 0 ? 0 : (ox == p.OW - 1 ? 2 : 1));
            const int32_t *corr = s_corr + cls * p.N;
            uint8_t *orow = p.out + (((long long)b * p.OH + oy) * p.OW + ox) * p.N;
            mbar_wait(tfull_bar(acc), aph);
            tc_fence_after();
            const uint32_t t_base = tmem_base + acc * (uint32_t)p.N + ((q * 32u) << 16);
            for (int c0 = 32 * cg; c0 < p.N; c0 += 128) {
                uint32_t r[32];
                tmem_ld32(t_base + (uint32_t)c0, r);
                if (p.early && c0 + 128 >= p.N) {          // the accumulator is in registers: release it before the math (as in conv_tc_kernel)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(mapa_rank(tempty_bar(acc), 0));
                }
                uint32_t w[8];
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const float4 z = *reinterpret_cast<const float4 *>(s_c0z + c0 + 4 * g);
                    const float4 sc = *reinterpret_cast<const float4 *>(s_c1 + c0 + 4 * g);
                    const int4 kc = *reinterpret_cast<const int4 *>(corr + c0 + 4 * g);
                    if (XU && !BIG) {
                        w[g] = requant4_biased<true>((int)r[4 * g] + kc.x, (int)r[4 * g + 1] + kc.y, (int)r[4 * g + 2] + kc.z, (int)r[4 * g + 3] + kc.w, z, sc, p.lo, p.hi);
                    } else if (XU) {
                        w[g] = requant4_i2f((int)r[4 * g] - kc.x, (int)r[4 * g + 1] - kc.y, (int)r[4 * g + 2] - kc.z, (int)r[4 * g + 3] - kc.w, z, sc);
                    } else {
                        w[g] = requant4_clamp<BIG>((int)r[4 * g] - kc.x, (int)r[4 * g + 1] - kc.y, (int)r[4 * g + 2] - kc.z, (int)r[4 * g + 3] - kc.w, z, sc, p.lo, p.hi, p.out_u8 != 0);
                    }
                }
                if (valid)
                    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(orow + c0), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                                 : "memory");
            }
            if (!p.early) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(mapa_rank(tempty_bar(acc), 0));      // the leader's MMA thread counts both CTAs' epilogue warps
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                      // no CTA of the pair leaves (or frees TMEM) while the other may still signal it
    if (warp == kWarpAlloc) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn(std::string *why) {
    static std::once_flag once;
    static EncodeTiledFn fn = nullptr;
    static std::string err;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
            err = std::string("cuTensorMapEncodeTiled unavailable: ") + cudaGetErrorString(e);
            (void)cudaGetLastError();
        } else {
            fn = reinterpret_cast<EncodeTiledFn>(p);
        }
    });
    if (!fn && why) *why = err;
    return fn;
}

bool encode_map(CUtensorMap *map, const void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides, const cuuint32_t *box, std::string *why) {
    EncodeTiledFn fn = get_encode_fn(why);
    if (!fn) return false;
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, (cuuint32_t)rank, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        if (why) *why = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
        return false;
    }
    return true;
}

size_t plan_smem(const ConvTcPlan &p, int stages) {
    const size_t b_bytes = (size_t)p.KH * p.KW * p.CB * p.N * 128;
    const size_t stage = p.patch ? ((((size_t)(p.TH + p.KH - 1) * (p.TW + p.KW - 1) * 128) + 1023) & ~(size_t)1023) : (size_t)(p.TH + p.KH - 1) * p.TW * 128;
    const size_t tables = (size_t)p.N * 8 + (size_t)p.ncls * p.N * 4;
    return b_bytes + stage * stages + 256 + tables;
}

}  // namespace

int conv_tc_pick_pack(int Cin, int Cout) {
    if (Cin <= 0 || Cout <= 0) return 0;
    if (Cin % 128 == 0) return (Cout % 32 == 0 && Cout <= 256) ? 1 : 0;
    if (128 % Cin != 0) return 0;
    const int P = 128 / Cin;
    if ((long long)P * Cout > 256 || (P * Cout) % 32 != 0) return 0;
    return P;
}

std::vector<uint8_t> conv_tc_pack_pointwise(const uint8_t *w, int Cout, int Cin, int P) {
    const size_t K = (size_t)P * Cin, N = (size_t)P * Cout;
    std::vector<uint8_t> m(N * K, 0);
    for (int pp = 0; pp < P; ++pp)
        for (int o = 0; o < Cout; ++o)
            std::memcpy(&m[((size_t)pp * Cout + o) * K + (size_t)pp * Cin], w + (size_t)o * Cin, (size_t)Cin);
    return m;
}

std::vector<int32_t> conv_tc_border_corr_3x3(const uint8_t *w, int Cout, int Cin, int in_zp, int H, int W, bool is_u8) {
    (void)H; (void)W;
    std::vector<int32_t> t((size_t)9 * Cout, 0);
    for (int rcls = 0; rcls < 3; ++rcls)
        for (int ccls = 0; ccls < 3; ++ccls)
            for (int o = 0; o < Cout; ++o) {
                int32_t s = 0;
                for (int m = 0; m < 3; ++m) {
                    if ((rcls == 0 && m == 0) || (rcls == 2 && m == 2)) continue;  // tap row outside the image
                    for (int n = 0; n < 3; ++n) {
                        if ((ccls == 0 && n == 0) || (ccls == 2 && n == 2)) continue;
                        const uint8_t *f = w + (((size_t)o * 3 + m) * 3 + n) * Cin;
                        for (int c = 0; c < Cin; ++c) s += is_u8 ? (int)f[c] : (int)(int8_t)f[c];
                    }
                }
                t[(size_t)(rcls * 3 + ccls) * Cout + o] = in_zp * s;
            }
    return t;
}

bool conv_tc_available(std::string *why) { return get_encode_fn(why) != nullptr; }

bool conv_tc_finalize_plan(ConvTcPlan &p, std::string *why) {
    auto no = [&](const char *m) { if (why) *why = m; return false; };
    if (p.N % 32 != 0 || p.N < 32 || p.N > 256) return no("N must be a multiple of 32 in [32,256]");
    if (p.ncls != 1 && p.ncls != 9) return no("border classes must be 1 or 9");
    if ((int)p.h_c0z.size() != p.N || (int)p.h_c1.size() != p.N || (int)p.h_corr.size() != p.ncls * p.N) return no("epilogue tables have the wrong size");
    if (p.TH * p.TW != 128 || (p.TW & (p.TW - 1)) != 0 || p.TW < 8) return no("tile must be TH x TW = 128 pixels with TW a power of two >= 8");
    if (p.C != 128 * p.CB) return no("row bytes must be 128 * CB");
    if (p.patch && (p.TW != 8 || (p.TW + p.KW - 1) * 128 >= (1 << 18))) return no("single-patch mode needs 8-pixel-wide tiles");
    if (p.TH + p.KH - 1 > 256) return no("patch too tall for one TMA box");
    int stages = 0;
    for (int s = 6; s >= 2; --s)
        if (plan_smem(p, s) <= kSmemLimit) { stages = s; break; }
    if (!stages) return no("weights + pipeline stages do not fit in 227 KB of shared memory");
    p.stages = stages;
    p.smem_bytes = plan_smem(p, stages);
    const cuuint64_t ktot = (cuuint64_t)p.KH * p.KW * p.C;
    cuuint64_t dims[2] = {ktot, (cuuint64_t)p.N};
    cuuint64_t strides[1] = {ktot};
    cuuint32_t box[2] = {128, (cuuint32_t)p.N};
    CUtensorMap m;
    if (!encode_map(&m, p.d_wmat, 2, dims, strides, box, why)) return false;
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    std::memcpy(p.tmap_b, &m, sizeof m);
    // CTA-pair variant of the 3x3 kernel: one 128-byte channel block, single-patch staging, the weights split by output channel
    p.pair_ok = false;
    if (p.KH == 3 && p.KW == 3 && p.CB == 1 && p.patch && p.ncls == 9 && p.N % 64 == 0) {
        const size_t b_half = (size_t)9 * (p.N / 2) * 128;
        const size_t stage = (((size_t)(p.TH + 2) * (p.TW + 2) * 128) + 1023) & ~(size_t)1023;
        const size_t fixed = b_half + 256 + (size_t)p.N * 8 + (size_t)9 * p.N * 4;
        int st = 0;
        for (int k = kMaxStages; k >= 2; --k)
            if (fixed + stage * k <= kSmemLimit) { st = k; break; }
        cuuint32_t box_h[2] = {128, (cuuint32_t)(p.N / 2)};
        CUtensorMap mh;
        if (st && encode_map(&mh, p.d_wmat, 2, dims, strides, box_h, nullptr)) {
            std::memcpy(p.tmap_b_half, &mh, sizeof mh);
            p.pair_ok = true;
            p.pair_stages = st;
            p.pair_smem_bytes = fixed + stage * st;
        }
    }
    return true;
}

namespace {
struct MapKey {
    const void *p; long long W, H, B; int C, TW, BH;
    bool operator==(const MapKey &o) const { return p == o.p && W == o.W && H == o.H && B == o.B && C == o.C && TW == o.TW && BH == o.BH; }
};
struct MapCache {   // activation tensor maps are re-used launch after launch (ping-pong buffers): encode each once
    std::mutex mu;
    std::vector<std::pair<MapKey, CUtensorMap>> v;
};
MapCache g_maps;
}  // namespace

cudaError_t conv_tc_launch(const ConvTcPlan &p, const ConvTcLaunch &l, int num_sms, cudaStream_t s, std::string *why) {
    if ((reinterpret_cast<uintptr_t>(l.out) & 31u) != 0 || (reinterpret_cast<uintptr_t>(l.in) & 15u) != 0) {   // STG.256 rows / TMA base
        if (why) *why = "tcgen05 conv: the input must be 16-byte and the output 32-byte aligned";
        return cudaErrorMisalignedAddress;
    }
    CUtensorMap ta, tb;
    std::memcpy(&tb, p.tmap_b, sizeof tb);
    const int box_w = p.patch ? p.TW + p.KW - 1 : p.TW;
    const MapKey key{l.in, l.W, l.H, l.B, p.C, box_w, p.TH + p.KH - 1};
    bool found = false;
    {
        std::lock_guard<std::mutex> lock(g_maps.mu);
        for (auto &e : g_maps.v)
            if (e.first == key) { ta = e.second; found = true; break; }
    }
    if (!found) {
        cuuint64_t dims[4] = {(cuuint64_t)p.C, (cuuint64_t)l.W, (cuuint64_t)l.H, (cuuint64_t)l.B};
        cuuint64_t strides[3] = {(cuuint64_t)p.C, (cuuint64_t)p.C * l.W, (cuuint64_t)p.C * l.W * l.H};
        cuuint32_t box[4] = {128, (cuuint32_t)box_w, (cuuint32_t)(p.TH + p.KH - 1), 1};
        if (!encode_map(&ta, l.in, 4, dims, strides, box, why)) return cudaErrorInvalidValue;
        std::lock_guard<std::mutex> lock(g_maps.mu);
        if (g_maps.v.size() >= 256) g_maps.v.clear();
        g_maps.v.emplace_back(key, ta);
    }

    ConvTcParams k{};
    k.out = l.out;
    ConvTcTables tab;
    std::memcpy(tab.c0z, p.h_c0z.data(), (size_t)p.N * 4);
    std::memcpy(tab.c1, p.h_c1.data(), (size_t)p.N * 4);
    const bool packed_epilogue = p.lo == -128.f && p.hi == 127.f && !p.big_acc && !p.is_u8;   // == the kernel's XU && !BIG (MF_TC_XUG=0 clears it below)
    std::memcpy(tab.corr, p.h_corr.data(), (size_t)p.ncls * p.N * 4);
    k.N = p.N; k.CB = p.CB; k.KH = p.KH; k.KW = p.KW; k.TW = p.TW; k.TH = p.TH;
    k.tw_log2 = 0;
    while ((1 << k.tw_log2) < p.TW) ++k.tw_log2;
    k.off_r = p.off_r; k.off_c = p.off_c; k.ncls = p.ncls; k.stages = p.stages;
    k.tiles_x = (int)((l.OW + p.TW - 1) / p.TW);
    k.tiles_y = (int)((l.OH + p.TH - 1) / p.TH);
    k.num_tiles = (long long)k.tiles_x * k.tiles_y * l.B;
    if (k.num_tiles >= (1ll << 31)) { if (why) *why = "too many tiles for one launch"; return cudaErrorInvalidValue; }
    k.fd_img = FastDiv((uint32_t)(k.tiles_x * k.tiles_y));
    k.fd_tx = FastDiv((uint32_t)k.tiles_x);
    k.OW = l.OW; k.OH = l.OH;
    k.lo = p.lo; k.hi = p.hi;
    // instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 = 2 @4, a/b format INT8 = 1 @7/@10,
    // K-major A and B (bits 15, 16 = 0), n_dim = N >> 3 @17, m_dim = 128 >> 4 @24
    // (uint8 tensors -- the reference's `T = u8` instantiation, microflow-macros/src/ops/conv_2d.rs:39-46 -- are a/b format 0)
    const uint32_t fmt = p.is_u8 ? 0u : 1u;
    k.idesc = (2u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
    k.stage_tx = (uint32_t)((p.TH + p.KH - 1) * box_w * 128);
    k.stage_bytes = p.patch ? ((k.stage_tx + 1023u) & ~1023u) : k.stage_tx;      // every stage base stays 1024-byte aligned (swizzle atom)
    k.patch = p.patch ? 1u : 0u;
    k.out_u8 = p.is_u8 ? 1 : 0;
    static const int env_early = [] { const char *e = std::getenv("MF_TC_EARLY"); return e ? std::atoi(e) : 1; }();
    k.early = env_early;
    k.patch_w = (uint32_t)(p.TW + p.KW - 1);
    k.b_block_bytes = (uint32_t)(p.N * 128);
    k.nkb = (uint32_t)(p.KH * p.KW * p.CB);
    // TMEM accumulator ring: four buffers where they fit the 512 columns (N <= 128), else two.  Round 1 measured four no faster (config 5:
    // 0.1010 vs 0.1000 ms, profiles/r01j_conv3x3_experiments.txt) -- the MMA issue loop was the limiter then.  With elect.sync issuers and the
    // early accumulator release: CTA-pair kernel 0.0794 -> 0.0777 ms at batch 16, 0.1509 -> 0.1492 at batch 32; one-CTA kernel and the
    // person_detect step unchanged (profiles/r02g_conv3x3_experiments.txt).  MF_TC_NACC=2 restores two.
    static const int env_nacc = [] { const char *e = std::getenv("MF_TC_NACC"); return e ? std::atoi(e) : 4; }();
    static const bool env_prebias = [] { const char *e = std::getenv("MF_TC_PREBIAS"); return e && std::atoi(e) != 0; }();      // its initial fill arms two buffers
    const uint32_t nacc = (env_nacc == 4 && !env_prebias && 4u * (uint32_t)p.N <= 512u) ? 4u : 2u;
    k.acc_mask = nacc - 1;
    k.acc_shift = nacc == 4 ? 2 : 1;
    const uint32_t need_cols = nacc * (uint32_t)p.N;
    k.tmem_cols = need_cols <= 32 ? 32 : (need_cols <= 64 ? 64 : (need_cols <= 128 ? 128 : (need_cols <= 256 ? 256 : 512)));
    if (k.num_tiles <= 0) return cudaSuccess;

    // the F2I.S8 / I2F epilogue needs the full int8 clamp range (F2I.S8 saturation is the clamp); MF_TC_XUG=0 forces the XU-free one
    static const int env_xug = [] { const char *e = std::getenv("MF_TC_XUG"); return e ? std::atoi(e) : -1; }();
    const bool xu = p.lo == -128.f && p.hi == 127.f && env_xug != 0 && !p.is_u8;   // uint8 outputs take the clamp-based (XU-free) epilogue
    if (xu && packed_epilogue)   // pre-biased accumulators: the table entry is added, not subtracted (conv_tc_kernel, PACKED)
        for (int k = 0; k < p.ncls * p.N; ++k) tab.corr[k] = kAccBias - tab.corr[k];
    const bool linear = p.KH == 1 && p.KW == 1 && p.TW == 128 && p.TH == 1 && l.H == 1 && l.B == 1 && l.OH == 1 && p.ncls == 1 && l.OW < (1ll << 31) - 128;
    const int shape = (p.KH == 3 && p.KW == 3 && p.CB == 1) ? 1 : ((linear && p.CB == 1) ? 2 : ((linear && p.CB == 2) ? 3 : 0));
    using KernelFn = void (*)(const CUtensorMap, const CUtensorMap, const ConvTcTables, const ConvTcParams);
    KernelFn fn = nullptr;
#define MF_TC_PICK(BIGV, XUV)                                                       \
    switch (shape) {                                                                \
        case 1: fn = conv_tc_kernel<BIGV, XUV, 3, 3, 1, kTabSmem>; break;           \
        case 2: fn = conv_tc_kernel<BIGV, XUV, 1, 1, 1, kTabSmem>; break;           \
        case 3: fn = conv_tc_kernel<BIGV, XUV, 1, 1, 2, kTabSmem>; break;           \
        default: fn = conv_tc_kernel<BIGV, XUV, 0, 0, 0, kTabSmem>; break;          \
    }
    if (p.big_acc) {
        if (xu) { MF_TC_PICK(true, true) } else { MF_TC_PICK(true, false) }
    } else {
        if (xu) { MF_TC_PICK(false, true) } else { MF_TC_PICK(false, false) }
    }
#undef MF_TC_PICK
    // constant-operand epilogues for the common pointwise case (one border class, one 128-byte channel block, XU epilogue)
    static const int env_tab = [] { const char *e = std::getenv("MF_TC_TAB"); return e ? std::atoi(e) : -1; }();
    if (shape == 2 && xu && !p.big_acc && p.ncls == 1 && env_tab != kTabSmem) {
        const bool periodic = p.Cout > 0 && 32 % p.Cout == 0 && p.N % p.Cout == 0 && p.P * p.Cout == p.N;
        // MF_TC_PREBIAS=1: accumulators pre-biased in TMEM by the epilogue warps (see PREBIAS in the kernel)
        static const bool env_pb = [] { const char *e = std::getenv("MF_TC_PREBIAS"); return e && std::atoi(e) != 0; }();
        if (periodic && env_tab != kTabGroup) fn = env_pb ? conv_tc_kernel<false, true, 1, 1, 1, kTabPeriodPB> : conv_tc_kernel<false, true, 1, 1, 1, kTabPeriod>;
        else if (p.N <= 128) fn = env_pb ? conv_tc_kernel<false, true, 1, 1, 1, kTabGroupPB> : conv_tc_kernel<false, true, 1, 1, 1, kTabGroup>;
    }
    // the opt-in to > 48 KB of dynamic shared memory is per device (context): remember it per (device, kernel)
    static std::mutex attr_mu;
    static std::vector<std::pair<int, KernelFn>> attr_done;
    {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        std::lock_guard<std::mutex> lock(attr_mu);
        bool have = false;
        for (auto &f : attr_done) have = have || (f.first == dev && f.second == fn);
        if (!have) {
            e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit);
            if (e != cudaSuccess) return e;
            attr_done.emplace_back(dev, fn);
        }
    }
    // 3x3 layers run on CTA pairs (conv3x3_pair_kernel) unless MF_TC_PAIR=0.  Until the MMA issue loop was fixed (elect_one) and the
    // accumulators were released right after tcgen05.ld, the pair kernel only equalled the one-CTA kernel (0.0900 vs 0.0894 ms on
    // BASELINE config 5); now its halved B-operand traffic and six pipeline stages show: 0.0819 vs 0.0879 ms at batch 16, 0.156 vs
    // 0.168 ms at batch 32 (profiles/r02g_conv3x3_experiments.txt).  The one-CTA kernel remains for odd shapes and single tiles.
    static const int env_pair = [] { const char *e = std::getenv("MF_TC_PAIR"); return e ? std::atoi(e) : 1; }();
    if (env_pair && shape == 1 && p.pair_ok && p.patch && k.num_tiles >= 2 && num_sms >= 2) {
        KernelFn pf = p.big_acc ? (xu ? conv3x3_pair_kernel<true, true> : conv3x3_pair_kernel<true, false>)
                                : (xu ? conv3x3_pair_kernel<false, true> : conv3x3_pair_kernel<false, false>);
        CUtensorMap tbh;
        std::memcpy(&tbh, p.tmap_b_half, sizeof tbh);
        ConvTcParams kp = k;
        kp.stages = p.pair_stages;
        kp.idesc = (2u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.N >> 3) << 17) | ((256u >> 4) << 24);     // M = 256 across the pair
        {
            int dev = 0;
            cudaError_t e = cudaGetDevice(&dev);
            if (e != cudaSuccess) return e;
            std::lock_guard<std::mutex> lock(attr_mu);
            bool have = false;
            for (auto &f : attr_done) have = have || (f.first == dev && f.second == pf);
            if (!have) {
                e = cudaFuncSetAttribute(pf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit);
                if (e != cudaSuccess) return e;
                attr_done.emplace_back(dev, pf);
            }
        }
        const long long pairs_needed = (k.num_tiles + 1) / 2;
        long long ctas = 2 * (pairs_needed < num_sms / 2 ? pairs_needed : num_sms / 2);
        return launch_pdl(pf, dim3((unsigned)ctas), dim3(kThreads), p.pair_smem_bytes, s, l.pdl, ta, tbh, tab, kp);
    }
    const unsigned grid = (unsigned)(k.num_tiles < num_sms ? k.num_tiles : num_sms);
    return launch_pdl(fn, dim3(grid), dim3(kThreads), p.smem_bytes, s, l.pdl, ta, tb, tab, k);
}

}  // namespace mf
