// mf_engine.h -- per-layer execution plans and the device-side model.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "mf_conv_tc.h"
#include "mf_kernels.h"
#include "mf_loader.h"

namespace mf {

enum class Kernel {
    None,           // reshape: no data movement (src/ops/reshape.rs + tensor.rs:95-141 keep NHWC element order)
    ConvGeneric,
    ConvTcPointwise,  // tcgen05 packed-pixel GEMM
    ConvTc3x3,        // tcgen05 implicit GEMM
    PwConvDp4a,
    DwConvC4,
    DwConv3x3Rows,
    DwConvCin1,
    FcGeneric,
    FcWarp,
    FcTc,             // tcgen05 GEMM: [batch x K] x [N x K]^T (w_zp == 0, K % 128 == 0, N % 32 == 0)
    PoolGeneric,
    Softmax,
};
const char *kernel_name(Kernel k);

// Static (weights + constants) storage of all layers: one contiguous blob, built on the host, uploaded once.
class BlobBuilder {
  public:
    size_t add(const void *p, size_t bytes);   // returns the 256-byte aligned offset
    const std::vector<uint8_t> &bytes() const { return host_; }
  private:
    std::vector<uint8_t> host_;
};

struct LayerExec {
    LayerSpec spec;
    Kernel kernel = Kernel::None;
    std::string why_not_fast;  // reason the tensor-core / fast path was not taken (for mf_model_dump)
    // blob offsets (SIZE_MAX = absent)
    size_t o_w = SIZE_MAX, o_wzp = SIZE_MAX, o_c0z = SIZE_MAX, o_c1 = SIZE_MAX, o_kcorr = SIZE_MAX, o_c2 = SIZE_MAX, o_lut = SIZE_MAX;
    size_t o_tc_w = SIZE_MAX, o_tc_c0z = SIZE_MAX, o_tc_c1 = SIZE_MAX, o_tc_corr = SIZE_MAX;
    ConvTcPlan tc;
    int tc_P = 1;
    bool big_acc = false;   // |acc - kcorr| may exceed 2^22
    // resolved launch arguments
    ConvArgs conv;
    FcArgs fc;
    PoolArgs pool;
    SoftmaxArgs sm;
    uint64_t alg_bytes = 0, weight_bytes = 0;

    // impl: 0 auto, 1 generic only, 2 auto without tensor cores
    void plan(BlobBuilder &bb, int impl, bool have_device);
    bool resolve(const uint8_t *d_blob, std::string *err);
    cudaError_t run(const uint8_t *in, uint8_t *out, long long batch, int num_sms, cudaStream_t s, std::string *err, int pdl = 0) const;
    // name of the kernel run() launches for these buffers / this batch (two depthwise kernels pick a sample-resident variant at run time)
    const char *launched_name(const uint8_t *in, uint8_t *out, long long batch) const;
};

}  // namespace mf
