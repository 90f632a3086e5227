// mf_conv_tc.h -- tcgen05 (5th-gen tensor core) int8 implicit-GEMM Conv2D / pointwise GEMM for sm_100a.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace mf {

// A launch-ready description of one Conv2D layer on the tensor cores.
//
// The kernel computes, per CTA tile, D[128 x N] (int32, TMEM) = A[128 x K] * B[N x K]^T with
//   A  = activations fetched by TMA through a 4-D tensor map (c, w, h, b), 128-byte rows, SWIZZLE_128B;
//        for a KHxKW stride-1 SAME conv one stage is the (TH+KH-1) x TW pixel patch shifted by one kernel
//        column, and the KH kernel rows are reached by offsetting the UMMA descriptor by whole patch rows
//   B  = the weight matrix [N][K_total], resident in shared memory for the whole (persistent) kernel
// followed by the reference's f32 requantize + zero-point + clamp epilogue and 16-byte int8 stores.
//
// Pointwise (1x1) layers are run as a plain GEMM over "packed pixels": P consecutive pixels form one
// 128-byte row (P * Cin == 128 * CB) against a block-diagonal weight matrix (N = P * Cout), so rows are always
// full swizzle atoms and the output row is P * Cout contiguous NHWC bytes.
struct ConvTcPlan {
    // logical GEMM
    int N = 0;          // UMMA N (= Cout, or P * Cout for packed pointwise); multiple of 32, <= 256
    int CB = 1;         // 128-byte channel blocks per pixel row
    int KH = 1, KW = 1; // kernel taps (stride 1)
    int TW = 128, TH = 1;  // output tile = TH x TW pixels (TH * TW == 128)
    int off_r = 0, off_c = 0;
    // tensor-map geometry of the activations: dims (C, W, H, B) in elements (bytes)
    int C = 128;        // bytes per (packed) pixel row = 128 * CB
    int P = 1;          // pixels packed per row (pointwise only)
    int Cout = 0;       // real output channels
    int ncls = 1;       // border classes (1, or 9 for 3x3)
    int stages = 2;
    size_t smem_bytes = 0;
    // KxK only: one stage = the whole (TH+KH-1) x (TW+KW-1) input patch of a tile, loaded ONCE; every kernel tap is a UMMA
    // descriptor whose start is offset by (m * (TW+KW-1) + n) pixels (128-byte rows) and whose 8-row group stride (SBO) is one
    // patch row -- which needs TW == 8 so that an 8-row core-matrix group is one row of the tile.  The input then crosses
    // L2 -> shared memory 1.4x (halo) instead of 3.75x (three column-shifted patches).
    bool patch = false;
    // device buffer owned by the model blob
    const uint8_t *d_wmat = nullptr;   // [N][K_total] bytes, K_total = KH*KW*C
    // epilogue tables (host copies; they travel to the kernel as __grid_constant__ parameters = constant bank)
    std::vector<float> h_c0z, h_c1;    // [N]
    std::vector<int32_t> h_corr;       // [ncls][N]  in_zp * (sum of weights over the taps valid for that border class)
    float lo = -128.f, hi = 127.f;
    bool big_acc = false;              // |acc - corr| may exceed 2^22: use the general exact int->float in the epilogue
    bool is_u8 = false;                // uint8 activations and weights: unsigned operand formats in the instruction descriptor
    alignas(64) unsigned char tmap_b[128];  // CUtensorMap of the weight matrix
    // 3x3 layers with Cin == 128 and Cout % 64 == 0 can run on CTA PAIRS (conv3x3_pair_kernel: tcgen05.mma.cta_group::2, M = 256):
    // each CTA of the pair keeps HALF of the output channels' weights (box {128, N/2}) and reads A + B/2 per instruction
    bool pair_ok = false;
    int pair_stages = 0;
    size_t pair_smem_bytes = 0;
    alignas(64) unsigned char tmap_b_half[128];
};

struct ConvTcLaunch {
    const uint8_t *in = nullptr;
    uint8_t *out = nullptr;
    long long W = 0, H = 1, B = 1;  // tensor-map extents of the activations (in packed rows for pointwise)
    long long OW = 0, OH = 1;       // output extents (== W, H for stride-1 SAME)
    int pdl = 0;                    // programmatic dependent launch (mf_kernels.h)
};

// Host-side packing helpers (pure functions, tested on CPU)
// pointwise: pick P so that P * Cin is a multiple of 128 and P * Cout <= 256; returns 0 if impossible
int conv_tc_pick_pack(int Cin, int Cout);
// builds the block-diagonal weight matrix [P*Cout][P*Cin] from OHWI 1x1 filters [Cout][Cin]
std::vector<uint8_t> conv_tc_pack_pointwise(const uint8_t *w, int Cout, int Cin, int P);
// 3x3 border-class table [9][Cout]: class = 3*row_cls + col_cls, cls 0 = first row/col, 1 = interior, 2 = last
std::vector<int32_t> conv_tc_border_corr_3x3(const uint8_t *w_ohwi, int Cout, int Cin, int in_zp, int H, int W, bool is_u8 = false);

bool conv_tc_available(std::string *why);      // driver entry point for cuTensorMapEncodeTiled present?
// fills plan.tmap_b / smem_bytes / stages; returns false (with reason) if the shape cannot run on this kernel
bool conv_tc_finalize_plan(ConvTcPlan &plan, std::string *why);
cudaError_t conv_tc_launch(const ConvTcPlan &plan, const ConvTcLaunch &l, int num_sms, cudaStream_t s, std::string *why);

}  // namespace mf
