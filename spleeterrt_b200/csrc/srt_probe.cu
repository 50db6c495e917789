// srt_probe.cu — measured roofline denominators (srt_probe_tensor_peak, srt_probe_copy_bandwidth in include/srt_b200.h).
//
// The tensor-core layers are tcgen05.mma kind::tf32 (main term) and kind::f16 / bf16 (compensation term).  MEASURED_PEAKS.json
// only has a cuBLAS bf16 GEMM rate, so bench.py used "bf16 / 2" as the TF32 peak.  This probe measures the pipe itself: one
// persistent CTA per SM whose elected thread issues M = 128, N = 256 MMAs (the shape that runs at the documented 128 cycles
// per K = 8 TF32 step, tools/mma_probe.cu) back to back into two TMEM accumulators from shared-memory operands filled with
// random data (zeros would draw less power and clock higher), for a caller-chosen duration.  No global-memory traffic at all:
// the number is the issue-rate x clock ceiling of the MMA pipe on this board at this moment, i.e. the denominator a kernel
// that also has to feed the pipe can only approach.
#include <cuda_runtime.h>

#include <cstdio>

#include "../../include/srt_b200.h"
#include "srt_internal.h"
#include "srt_ptx.cuh"

using namespace srt;

namespace {

constexpr int kProbeN = 256;
constexpr int kProbeSmem = 16 * 1024 + kProbeN * 128 + 2048;   // A tile, B tile, alignment slack

template <int KIND>   // 0 = kind::tf32 (K = 8), 1 = kind::f16 with bf16 operands (K = 16)
__global__ void __launch_bounds__(128, 1) tensor_peak_kernel(int iters, unsigned seed)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint32_t* words = reinterpret_cast<uint32_t*>(smem_raw + (base - ptx::smem_u32(smem_raw)));
    // operands: normal-range values of either format (exponent field kept mid-range), different per SM
    unsigned x = seed * 2654435761u + blockIdx.x * 40503u + threadIdx.x;
    for (int i = threadIdx.x; i < (16 * 1024 + kProbeN * 128) / 4; i += blockDim.x) {
        x = x * 1664525u + 1013904223u;
        words[i] = KIND == 0 ? ((x & 0x807fe000u) | 0x3f000000u) : ((x & 0x807f807fu) | 0x3f003f00u);
    }
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
    if (threadIdx.x < 32) ptx::tmem_alloc<512>(&tbase);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    if (threadIdx.x < 32 && ptx::elect_one()) {
        const uint32_t idesc = KIND == 0 ? ptx::umma_idesc_tf32(128, kProbeN) : ptx::umma_idesc_bf16(128, kProbeN);
        const uint32_t alo = ptx::umma_desc_lo(base), blo = ptx::umma_desc_lo(base + 16 * 1024);
        for (int i = 0; i < iters; i++)
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
                    if (KIND == 0) ptx::mma_tf32_ss_lo(tbase + a * kProbeN, alo + kk * 2, blo + kk * 2, idesc, 1);
                    else ptx::mma_bf16_ss_lo(tbase + a * kProbeN, alo + kk * 2, blo + kk * 2, idesc, 1);
                }
        ptx::mma_commit(&bar);
        ptx::mbar_wait(&bar, 0);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tbase);
}

__global__ void copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = __ldg(src + i);
}

}  // namespace

extern "C" int srt_probe_tensor_peak(int device, int kind, double seconds, double* tflops_out)
{
    if (!tflops_out || (kind != 0 && kind != 1) || !(seconds > 0) || seconds > 30) return internal::set_error(SRT_ERR_ARG, "srt_probe_tensor_peak: bad argument");
    internal::DeviceGuard g(device);
    if (!g.ok) return internal::set_error(SRT_ERR_CUDA, "cudaSetDevice failed");
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return internal::set_error(SRT_ERR_CUDA, "no device");
    auto kern = kind == 0 ? tensor_peak_kernel<0> : tensor_peak_kernel<1>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kProbeSmem) != cudaSuccess)
        return internal::set_error(SRT_ERR_CUDA, "srt_probe_tensor_peak: not an sm_100 device");
    cudaStream_t st;
    cudaEvent_t e0, e1;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const double flop_per_iter = 2.0 * 128 * kProbeN * (kind == 0 ? 8 : 16) * 8 /* MMAs per iteration */ * sms;
    // calibrate on a short launch, then launches of ~20 ms until `seconds` have run; the reported rate covers all of them
    int iters = 2000;
    float ms = 0.f;
    double total_flop = 0, total_ms = 0;
    for (int round = 0; round < 2000; round++) {
        cudaEventRecord(e0, st);
        kern<<<sms, 128, kProbeSmem, st>>>(iters, 17u + round);
        cudaEventRecord(e1, st);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        cudaEventElapsedTime(&ms, e0, e1);
        if (round > 0) { total_flop += flop_per_iter * iters; total_ms += ms; }
        if (round == 0 && ms > 0) iters = (int)(iters * 20.0 / ms) + 1;
        if (total_ms >= seconds * 1e3) break;
    }
    const cudaError_t err = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    if (err != cudaSuccess || total_ms <= 0) return internal::set_error(SRT_ERR_CUDA, cudaGetErrorString(err));
    *tflops_out = total_flop / (total_ms * 1e-3) / 1e12;
    return 0;
}

extern "C" int srt_probe_copy_bandwidth(int device, size_t bytes, double* gbs_out)
{
    if (!gbs_out || bytes < (1u << 20)) return internal::set_error(SRT_ERR_ARG, "srt_probe_copy_bandwidth: bad argument");
    internal::DeviceGuard g(device);
    if (!g.ok) return internal::set_error(SRT_ERR_CUDA, "cudaSetDevice failed");
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    float4 *a = nullptr, *b = nullptr;
    const size_t n = bytes / 16;
    if (cudaMalloc((void**)&a, n * 16) != cudaSuccess || cudaMalloc((void**)&b, n * 16) != cudaSuccess) {
        if (a) cudaFree(a);
        cudaGetLastError();
        return internal::set_error(SRT_ERR_CUDA, "srt_probe_copy_bandwidth: out of memory");
    }
    cudaMemset(a, 1, n * 16);
    cudaStream_t st;
    cudaEvent_t e0, e1;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaDeviceSynchronize();
    float best = 0.f;
    for (int r = 0; r < 6; r++) {
        cudaEventRecord(e0, st);
        copy_kernel<<<sms * 8, 512, 0, st>>>(a, b, n);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const float gbs = (float)(2.0 * n * 16 / (ms * 1e-3) / 1e9);
        if (r > 0 && gbs > best) best = gbs;
    }
    const cudaError_t err = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    cudaFree(a);
    cudaFree(b);
    if (err != cudaSuccess) return internal::set_error(SRT_ERR_CUDA, cudaGetErrorString(err));
    *gbs_out = best;
    return 0;
}
