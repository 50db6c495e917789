// mma_probe.cu — GPU microbenchmark (diagnostic, not product): cycles per tcgen05.mma for
// kind::tf32 / kind::f16, different N, and 1..4 independent accumulators issued round-robin by ONE
// thread from fixed shared-memory operands.  Answers: are dependent (same-accumulator) MMAs
// latency-serialised, and what is the per-instruction floor for small N?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I spleeterrt_b200/csrc tools/mma_probe.cu -o tools/mma_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "srt_ptx.cuh"
using namespace srt;

template <int KIND>   // 0 tf32, 1 f16
__device__ __forceinline__ void mma(uint32_t d, uint32_t alo, uint32_t blo, uint32_t idesc, uint32_t acc)
{
    if (KIND == 0) ptx::mma_tf32_ss_lo(d, alo, blo, idesc, acc);
    else
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(d), "r"(alo), "r"(blo), "r"(idesc), "r"(acc), "r"(ptx::kDescHiSw128) : "memory");
}

// ELECT = false: the issuing thread is selected with `threadIdx.x == 0` (every UTCHMMA gets wrapped in an
// ELECT / BRA.U.ANY waterfall loop by the compiler); ELECT = true: with elect.sync (straight-line UTCHMMA).
template <int KIND, bool ELECT>
__global__ void probe(int N, int nacc, int iters, int a_shift_rows, long long* out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const uint32_t base = (ptx::smem_u32(smem) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 40 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.0f;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) ptx::tmem_alloc<512>(&tbase);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (ELECT ? (threadIdx.x < 32 && ptx::elect_one()) : (threadIdx.x == 0)) {
        const uint32_t idesc = KIND == 0 ? ptx::umma_idesc_tf32(128, N)
                                         : ((1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24));
        const uint32_t alo = ptx::umma_desc_lo(base + a_shift_rows * 128), blo = ptx::umma_desc_lo(base + 20 * 1024);
        // warm-up
        for (int i = 0; i < 8; i++) mma<KIND>(tbase, alo, blo, idesc, 1);
        ptx::mma_commit(&bar);
        ptx::mbar_wait(&bar, 0);
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++)
            for (int a = 0; a < nacc; a++)
#pragma unroll
                for (int kk = 0; kk < 4; kk++) mma<KIND>(tbase + a * N, alo + kk * 2, blo + kk * 2, idesc, 1);
        const long long t1 = clock64();
        ptx::mma_commit(&bar);
        ptx::mbar_wait(&bar, 1);
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tbase);
}

// two issuing warps (warp 0 and warp 1), each on its own accumulators: is the ~60-100 cycle floor per issuing
// thread or per SM?
template <int KIND>
__global__ void probe2(int N, int nacc, int iters, long long* out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tbase;
    const uint32_t base = (ptx::smem_u32(smem) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 40 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.0f;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar[0], 1); ptx::mbar_init(&bar[1], 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) ptx::tmem_alloc<512>(&tbase);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const int w = threadIdx.x >> 5;
    if (w < 2 && (threadIdx.x & 31) == 0) {
        const uint32_t idesc = ptx::umma_idesc_tf32(128, N);
        const uint32_t alo = ptx::umma_desc_lo(base), blo = ptx::umma_desc_lo(base + 20 * 1024);
        const uint32_t t0a = tbase + w * nacc * N;
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++)
            for (int a = 0; a < nacc; a++)
#pragma unroll
                for (int kk = 0; kk < 4; kk++) mma<KIND>(t0a + a * N, alo + kk * 2, blo + kk * 2, idesc, 1);
        ptx::mma_commit(&bar[w]);
        ptx::mbar_wait(&bar[w], 0);
        out[w] = clock64() - t0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tbase);
}

int main()
{
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(probe<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(probe<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(probe<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(probe<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int elect = 0; elect < 2; elect++) {
    printf("issuer selected by %s\nkind N nacc shift | issue cyc/MMA | total cyc/MMA\n", elect ? "elect.sync" : "threadIdx.x == 0");
    for (int kind = 0; kind < 2; kind++)
        for (int N : {16, 32, 64, 128, 256})
            for (int nacc : {1, 2, 4})
                for (int shift : {0, 1}) {
                    if (nacc * N > 512) continue;
                    if (shift && !(N == 64)) continue;
                    const int iters = 200;
                    if (kind == 0 && !elect) probe<0, false><<<1, 128, 64 * 1024>>>(N, nacc, iters, shift, d);
                    else if (kind == 0) probe<0, true><<<1, 128, 64 * 1024>>>(N, nacc, iters, shift, d);
                    else if (!elect) probe<1, false><<<1, 128, 64 * 1024>>>(N, nacc, iters, shift, d);
                    else probe<1, true><<<1, 128, 64 * 1024>>>(N, nacc, iters, shift, d);
                    long long h[2];
                    cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    const double n = (double)iters * nacc * 4;
                    printf("%s %3d %d %d | %7.1f | %7.1f\n", kind ? "f16 " : "tf32", N, nacc, shift, h[0] / n, h[1] / n);
                }
    }
    cudaFuncSetAttribute(probe2<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    printf("two issuing warps: N nacc(per warp) | cycles per MMA per warp | aggregate cycles per MMA\n");
    for (int N : {32, 64, 128})
        for (int nacc : {1, 2}) {
            if (2 * nacc * N > 512) continue;
            const int iters = 200;
            probe2<0><<<1, 128, 64 * 1024>>>(N, nacc, iters, d);
            long long h[2];
            if (cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("error\n"); return 1; }
            const double n = (double)iters * nacc * 4;
            printf("tf32 %3d %d | %7.1f %7.1f | %7.1f\n", N, nacc, h[0] / n, h[1] / n, (h[0] > h[1] ? h[0] : h[1]) / (2 * n));
        }
    return 0;
}
