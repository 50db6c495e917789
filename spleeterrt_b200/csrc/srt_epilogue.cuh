// srt_epilogue.cuh — the fused layer epilogues, shared by the tcgen05 and the SIMT gather-GEMM.
//   encoder (Executable/spleeter.c:182-190): v = conv + bias -> skip tensor (NHWC);
//            act(scale*v + offset) -> next layer's input, written in space-to-depth form.
//   down6   (spleeter.c:232-238): bias only.
//   decoder (spleeter.c:240-247): scale*act(tconv + bias) + offset, scattered to the output
//            pixel (2Y+po, 2X+qo) of the phase.
// Values that feed a tensor-core layer are rounded to TF32 here (cvt.rna), so the MMA's
// implicit operand truncation is a no-op and the rounding is unbiased.
#pragma once
#include "srt_kernels.cuh"
#include "srt_ptx.cuh"

namespace srt {

__device__ __forceinline__ void load16(const float* __restrict__ src, float* v)
{
    const float4* q = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float4 t = __ldg(q + i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}

__device__ __forceinline__ void store16(float* dst, const float* v)
{
    float4* d = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int q = 0; q < 4; q++) d[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

// v[16]: accumulators for channels [c0, c0+16) of pixel (n, Y, X) in tile space.
__device__ __forceinline__ void epilogue16(const ConvParams& p, int s, int n, int Y, int X, int phase, int c0, float* v)
{
    float bias[16];
    load16(p.bias + s * p.cout + c0, bias);
    if (p.mode == 2) {
        float sc[16], of[16];
        load16(p.bn_scale + s * p.cout + c0, sc);
        load16(p.bn_offset + s * p.cout + c0, of);
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
            float t = sc[i] * apply_act(p.act[s], v[i] + bias[i]) + of[i];
            o[i] = p.round_act ? ptx::rna_tf32(t) : t;
        }
        const int oy = 2 * Y + (phase >> 1), ox = 2 * X + (phase & 1);
        store16(p.out_dec + (((size_t)n * (2 * p.Hs) + oy) * (2 * p.Ws) + ox) * p.cout + c0, o);
        return;
    }
    float raw[16];
#pragma unroll
    for (int i = 0; i < 16; i++) raw[i] = v[i] + bias[i];
    if (p.mode == 0) {
        float sc[16], of[16];
        load16(p.bn_scale + s * p.cout + c0, sc);
        load16(p.bn_offset + s * p.cout + c0, of);
        float a[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
            float t = apply_act(p.act[s], sc[i] * raw[i] + of[i]);
            a[i] = p.round_act ? ptx::rna_tf32(t) : t;
        }
        store16(p.out_act + ((((size_t)n * (p.Hs / 2) + Y / 2) * (p.Ws / 2) + X / 2) * 4 + (Y & 1) * 2 + (X & 1)) * p.cout + c0, a);
    }
    if (p.round_raw) {
#pragma unroll
        for (int i = 0; i < 16; i++) raw[i] = ptx::rna_tf32(raw[i]);
    }
    store16(p.out_raw + (((size_t)n * p.Hs + Y) * p.Ws + X) * p.cout + c0, raw);
}

}  // namespace srt
