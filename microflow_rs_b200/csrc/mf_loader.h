// mf_loader.h -- run-time TFLite loader: does at mf_model_create() what the reference's proc-macro
// (microflow-macros/src/lib.rs:46-208 and ops/<op>.rs `new` + `preprocess`) does at Rust compile time.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/microflow_cuda.h"

namespace mf {

struct LayerSpec {
    int op = 0;              // MF_OP_*
    bool is_u8 = false;
    int in_rank = 0, out_rank = 0;
    int in_dims[4] = {1, 1, 1, 1}, out_dims[4] = {1, 1, 1, 1};
    // geometry of the 4-D ops (conv / depthwise / pool); fc: Cin = K, Cout = N
    int H = 1, W = 1, Cin = 1, OH = 1, OW = 1, Cout = 1, KH = 1, KW = 1, sh = 1, sw = 1, pad = 0, act = 0;
    float in_scale = 0.f, out_scale = 0.f;
    int in_zp = 0, out_zp = 0;
    std::vector<uint8_t> w;       // conv: OHWI; depthwise: [KH][KW][Cout]; fc: [N][K] (TFLite bytes)
    std::vector<int32_t> w_zp;    // per-channel (len Cout) or single
    std::vector<float> c0, c1;    // constants.0 / constants.1 exactly as the macro computes them
    std::vector<int32_t> c2;      // fc only
    int32_t c3 = 0;               // fc only
    int act_lo = -128, act_hi = 127;  // saturation ∘ fused activation as one clamp (src/activation.rs:21-34)
    std::vector<float> exp_lut;   // softmax: expf(f32(q) * in_scale) for each of the 256 byte values
    size_t in_elems = 0, out_elems = 0;
    uint64_t macs = 0;
};

struct ModelSpec {
    std::vector<LayerSpec> layers;
    bool is_u8_in = false, is_u8_out = false;
    int in_rank = 0, out_rank = 0;
    int in_dims[4] = {1, 1, 1, 1}, out_dims[4] = {1, 1, 1, 1};
    float in_scale = 0.f, out_scale = 0.f;
    int in_zp = 0, out_zp = 0;
    size_t in_elems = 0, out_elems = 0, max_elems = 0;
};

// Returns MF_OK or the mf_status that corresponds to the reference's compile-time diagnostic.
int parse_tflite(const uint8_t *buf, size_t len, ModelSpec &out, std::string &err);

// Host staging of the reference's image sample format (samples/person.bmp -> the const PERSON of samples/features/person_detect.rs):
// an uncompressed 8-bit BMP with the identity gray palette; the feature tensor is the pixel bytes read as int8, top row first
// (BMP stores its rows bottom-up unless the height is negative).  Returns MF_OK or MF_ERR_INVALID_ARG / MF_ERR_UNSUPPORTED_TYPE.
int features_from_bmp_gray8(const uint8_t *bmp, size_t len, uint8_t *out, size_t cap, int *height, int *width, std::string &err);

// Scalar semantics shared by host-side preprocessing (bit-exact restatements).
float libm_expf(float x);                                   // Rust libm 0.2 expf (musl e_expf.c algorithm)
int quantize_scalar(float x, float scale, int zp, bool is_u8);  // src/quantize.rs:16-18
void activation_clamp(int act, float out_scale, int out_zp, bool is_u8, int &lo, int &hi);

}  // namespace mf
