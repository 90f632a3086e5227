// mf_fused.h -- cross-layer fusion of the low-resolution stages.
//
// The reference chains its operators straight-line in the generated `predict_inner` (microflow-macros/src/lib.rs:198-201),
// every intermediate tensor moved by value between the calls.  On the GPU a layer-by-layer execution of the small feature maps
// (person_detect layers 13-22: ten layers on 6x6x128 = 4.6 KB per sample) is bound by launch ramps, per-sample fixed costs and
// HBM round trips, not by arithmetic.  A FusedChain runs a run of
//      [ depthwise_conv_2d 3x3 / stride 1 / SAME, C = 128 ]  ->  [ conv_2d 1x1, 128 -> 128 ]        (n_pairs times)
// in ONE persistent kernel: a unit of U = floor(256 / (H*W)) samples is loaded once into shared memory, every depthwise layer
// runs on CUDA cores (dp4a) from shared memory into the swizzled A-operand tile of the following pointwise layer, that layer is
// a tcgen05.mma (kind::i8, M128 N128 K128, accumulator in TMEM) against weights resident in shared memory, and its fused f32
// requantize epilogue writes the next layer's input back into shared memory.  Only the first input and the last output touch HBM.
// Arithmetic per layer is exactly the stand-alone kernels' (src/ops/depthwise_conv_2d.rs:56-101, conv_2d.rs:56-104).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace mf {

constexpr int kFusedMaxPairs = 6;
constexpr int kFusedC = 128;

struct FusedChainPlan {
    int n_pairs = 0;
    int H = 0, W = 0;
    int U = 0;                          // samples per unit
    size_t smem_bytes = 0;
    // per pair: depthwise (in_zp, clamp) and pointwise (clamp)
    int dw_zp[kFusedMaxPairs] = {};
    float dw_lo[kFusedMaxPairs] = {}, dw_hi[kFusedMaxPairs] = {}, pw_lo[kFusedMaxPairs] = {}, pw_hi[kFusedMaxPairs] = {};
    bool full_clamp = true;             // every clamp is the full int8 range: the saturating F2I.S8 is the clamp
    // static data in the model blob (device pointers after resolve)
    const uint8_t *d_wimg = nullptr;    // n_pairs x 16 KB: pointwise weights [Cout][Cin] as the SWIZZLE_128B shared-memory image
    const uint8_t *d_consts = nullptr;  // n_pairs x kFusedConstBytes: per-pair constant block (layout below)
};

// constant block of one pair, copied to shared memory as it is:
//   words [0, 9*32)        depthwise weights, word (tap k, channel group g) at k * 32 + g   (= the [KH][KW][C] bytes of the reference)
//   words [9*32, 13*32)    depthwise c0z   (f32(out_zp) + c0[ch]),  float4 of group g at 9*32 + 4*g
//   words [13*32, 17*32)   depthwise c1,   float4 of group g at 13*32 + 4*g
//   words [17*32, +128)    pointwise c0z[128]
//   words [.., +128)       pointwise c1[128]
//   words [.., +128)       pointwise kAccBias - kcorr[128]  (pre-biased accumulator correction, mf_device.cuh)
constexpr int kFusedDwWords = 17 * 32;
constexpr int kFusedPwWords = 3 * 128;
constexpr int kFusedConstBytes = (kFusedDwWords + kFusedPwWords) * 4;

// host-side packing (pure functions; CPU-tested)
// pointwise filters [128][128] (OHWI with 1x1 taps) -> 16 KB SWIZZLE_128B K-major image: byte (n, k) at n*128 + (((k>>4) ^ (n&7))<<4) + (k&15)
void fused_pack_pointwise_image(const uint8_t *w_ohwi, uint8_t *img16k);
// one pair's constant block
void fused_pack_consts(const uint8_t *dw_w, const float *dw_c0z, const float *dw_c1, const float *pw_c0z, const float *pw_c1, const int32_t *pw_kcorr,
                       uint8_t *out);

// shared-memory footprint of a chain; 0 if it cannot fit
size_t fused_chain_smem(int n_pairs, int H, int W);
bool fused_chain_finalize(FusedChainPlan &p, std::string *why);
cudaError_t fused_chain_launch(const FusedChainPlan &p, const uint8_t *in, uint8_t *out, long long batch, int num_sms, cudaStream_t s, int pdl);

}  // namespace mf
