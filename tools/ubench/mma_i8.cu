// tcgen05.mma kind::i8 issue-rate microbenchmark (sm_100a): the MMA-only ceiling that BASELINE config 5 (3x3 Conv2D) is measured
// against.  One persistent CTA per SM; the A tile (128 rows x 128 bytes) and the B tile (N rows x 128 bytes) sit in shared memory as
// SWIZZLE_128B K-major images (contents irrelevant), two accumulators alternate in TMEM, ONE thread issues the instructions back
// to back (M128 x N x K32 each) and commits to an mbarrier every 64 instructions; there is no TMA, no epilogue, no global traffic.
//   int8 ops = 2 * 128 * N * 32 per instruction.   Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_i8 mma_i8.cu
// Usage: ./mma_i8 [iters]      prints one line per N in {64, 128, 256} (cta_group::1).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo = 1024) {     // K-major, SWIZZLE_128B, SBO = 1024 (see mf_tc_ptx.cuh)
    uint64_t d = (uint64_t)((addr >> 4) & 0x3FFFu);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// LD_WARPS > 0: that many extra warps (4 per TMEM lane quarter) read the accumulators with tcgen05.ld.32x32b.x32 in a loop while the
// MMAs run -- does epilogue-style TMEM traffic slow the tensor pipe (and how fast are the loads under MMA load)?
template <int N, int LD_WARPS>
__global__ void __launch_bounds__(128 + 32 * LD_WARPS, 1) mma_kernel(int iters, long long *cycles, unsigned *sink, unsigned long long *ld_count, uint32_t a_off, uint32_t sbo, int chain) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[8];
    __shared__ uint32_t tmem_slot;
    uint8_t *sA = smem, *sB = smem + 32768;      // A region 32 KB: room for shifted / strided views
    for (int i = threadIdx.x; i < (32768 + N * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01010101u * (uint32_t)(i & 3);
    const uint32_t bar_a = smem_u32(&bar[0]);
    if (threadIdx.x == 0) {
        reinterpret_cast<volatile unsigned *>(&bar[3])[4] = 0;
        for (int k = 0; k < 4; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a + 8 * k) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t da = make_desc(smem_u32(sA) + a_off, sbo), db = make_desc(smem_u32(sB));
        // batch `it` of 64 instructions commits to barrier it & 3 (phase (it >> 2) & 1); batch it - 2 is awaited before batch it + 1 is
        // issued, so three batches are in flight and a barrier never completes a phase that has not been waited for
        auto wait_batch = [&](int b) {
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(bar_a + 8u * (uint32_t)(b & 3)), "r"((uint32_t)(b >> 2) & 1u) : "memory");
        };
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 64; ++k)       // chain: all 64 MMAs of a batch accumulate into ONE accumulator (a long K loop), else 4 per accumulator
                mma_i8(tmem + (uint32_t)(chain ? (it & 1) : ((k >> 2) & 1)) * N, da + 2 * (k & 3), db + 2 * (k & 3), idesc, chain ? (uint32_t)(k != 0) : (uint32_t)(k & 3));
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a + 8u * (uint32_t)(it & 3)) : "memory");
            if (it >= 2) wait_batch(it - 2);
        }
        for (int b = iters >= 2 ? iters - 2 : 0; b < iters; ++b) wait_batch(b);
        cycles[blockIdx.x] = clock64() - t0;
        reinterpret_cast<volatile unsigned *>(&bar[3])[4] = 1;
    }
    if (LD_WARPS > 0 && threadIdx.x >= 128) {
        const unsigned q = (threadIdx.x >> 5) & 3;
        volatile unsigned *flag = reinterpret_cast<volatile unsigned *>(&bar[3]) + 4;   // set by thread 0 when it is done (spare shared word)
        unsigned acc = 0;
        unsigned long long n = 0;
        while (*flag == 0) {
            unsigned r[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                           "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                           "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                         : "r"(tmem + ((q * 32u) << 16) + (unsigned)((n & 3) * 32))
                         : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int k = 0; k < 32; ++k) acc ^= r[k];
            ++n;
        }
        if (acc == 0x12345u) sink[0] = acc;
        if ((threadIdx.x & 31) == 0) atomicAdd(ld_count, n);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int N, int LD_WARPS> static void run(int iters, int sms, uint32_t a_off = 0, uint32_t sbo = 1024, int chain = 0) {
    long long *cyc;
    unsigned *sink;
    unsigned long long *ldc;
    cudaMalloc(&cyc, sms * sizeof(long long));
    cudaMalloc(&sink, 4);
    cudaMalloc(&ldc, 8);
    const size_t smem = 32768 + (size_t)N * 128;
    cudaFuncSetAttribute(mma_kernel<N, LD_WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    mma_kernel<N, LD_WARPS><<<sms, 128 + 32 * LD_WARPS, smem>>>(iters / 8 + 1, cyc, sink, ldc, a_off, sbo, chain);
    cudaDeviceSynchronize();
    float best = 1e30f;
    unsigned long long loads = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaMemset(ldc, 0, 8);
        cudaEventRecord(e0);
        mma_kernel<N, LD_WARPS><<<sms, 128 + 32 * LD_WARPS, smem>>>(iters, cyc, sink, ldc, a_off, sbo, chain);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) { best = ms; cudaMemcpy(&loads, ldc, 8, cudaMemcpyDeviceToHost); }
    }
    cudaError_t e = cudaDeviceSynchronize();
    long long c0 = 0;
    cudaMemcpy(&c0, cyc, sizeof c0, cudaMemcpyDeviceToHost);
    const double ops = 2.0 * 128.0 * N * 32.0 * 64.0 * iters * sms;
    printf("tcgen05.mma.cta_group::1.kind::i8 M128 N%-3d K32, %2d warps of tcgen05.ld, A start +%u B, SBO %u%s: %8.3f ms -> %8.1f TOP/s  (%.1f clk per MMA on SM 0", N, LD_WARPS, a_off, sbo, chain ? ", one accumulator per 64" : "", best,
           ops / (best * 1e-3) / 1e12, (double)c0 / (64.0 * iters));
    if (LD_WARPS) printf("; %.1f clk per 32x32 tcgen05.ld per warp", (double)c0 * LD_WARPS / ((double)loads / sms));
    printf(", %s)\n", cudaGetErrorString(e));
    cudaFree(cyc); cudaFree(sink); cudaFree(ldc);
}

int main(int argc, char **argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 2000;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs, SM clock max %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    run<64, 0>(iters, p.multiProcessorCount);
    run<128, 0>(iters, p.multiProcessorCount);
    run<256, 0>(iters, p.multiProcessorCount);
    run<128, 4>(iters, p.multiProcessorCount);
    run<128, 16>(iters, p.multiProcessorCount);
    // the 3x3 kernel's single-patch A views: 8-row groups patch_w * 128 B apart, tap (m, n) starts (m * patch_w + n) * 128 B into the patch
    run<128, 0>(iters, p.multiProcessorCount, 0, 1024, 1);
    run<128, 0>(iters, p.multiProcessorCount, 128, 1024, 0);
    run<128, 0>(iters, p.multiProcessorCount, 0, 1280, 0);
    run<128, 0>(iters, p.multiProcessorCount, 1280 + 256, 1280, 1);
    return 0;
}
