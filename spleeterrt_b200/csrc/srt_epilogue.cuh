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

// Warp-cooperative store of one 16-channel chunk per lane (lane = pixel).  A lane's 64 bytes are contiguous in
// global memory but neighbouring lanes are a whole pixel (>= 64 B, usually >= 128 B) apart, so a plain STG.128
// touches 32 different lines per instruction: ncu (profiles/r1k) counts 32 sectors per store request and the
// stage-skip sweep (tools/stage_sweep.py rp) shows the epilogue alone costs as much as all the MMAs of the small-N
// layers - the LSU, not the math.  Here the chunk is transposed through a 2 KB per-warp shared-memory buffer so
// that four consecutive lanes write one pixel's 64 bytes: 8 lines per instruction instead of 32.
//   dst   : this lane's destination (16 floats), or nullptr if the pixel is outside the tensor
//   stage : 128 float4 of shared memory owned by this warp (XOR-swizzled, conflict-free both ways)
// All 32 lanes must call it.
__device__ __forceinline__ void store16_warp(float* dst, const float* v, float4* stage)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 4; k++) stage[lane * 4 + (k ^ ((lane >> 1) & 3))] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int i = j * 32 + lane, px = i >> 2, piece = i & 3;
        const float4 t = stage[px * 4 + (piece ^ ((px >> 1) & 3))];
        float* d = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(dst), px));
        if (d) *reinterpret_cast<float4*>(d + piece * 4) = t;
    }
    __syncwarp();
}

// 16 residuals v - tf32(v) of one pixel -> 32 bytes of bf16 (two 16-byte stores per lane; every lane writes one whole
// 32-byte sector).  dst == nullptr: nothing stored (pixel outside the tensor).
__device__ __forceinline__ void store16_lo(uint16_t* dst, const float* lo)
{
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = ptx::pack_bf16x2(lo[2 * i], lo[2 * i + 1]);
    if (dst) {
        uint4* d = reinterpret_cast<uint4*>(dst);
        d[0] = make_uint4(w[0], w[1], w[2], w[3]);
        d[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
}

// The same residuals in the 8-bit format: e5m2(4 lo), 16 bytes per pixel and chunk, one store.
__device__ __forceinline__ void store16_lo8(uint8_t* dst, const float* lo)
{
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; i++) w[i] = ptx::pack_e5m2x4(4.0f * lo[4 * i], 4.0f * lo[4 * i + 1], 4.0f * lo[4 * i + 2], 4.0f * lo[4 * i + 3]);
    if (dst) *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
}
// residual store in the destination's format; `elem` = element index of the pixel's first channel of this chunk
__device__ __forceinline__ void store16_residual(void* base, size_t elem, int fp8, bool valid, const float* lo)
{
    if (fp8) store16_lo8(valid ? reinterpret_cast<uint8_t*>(base) + elem : nullptr, lo);
    else store16_lo(valid ? reinterpret_cast<uint16_t*>(base) + elem : nullptr, lo);
}

// Branch-free activations for the tensor-core epilogues.  The epilogue warps are instruction-latency
// bound (ncu / tools/stage_sweep.py: the epilogue, not the MMAs, dominates down1 and costs as much as the
// MMAs of the small-N layers), so the activation kind is resolved once per 16-channel chunk and ELU is
// 5 instructions: FMUL, MUFU.EX2 (ex2.approx.ftz), FADD, FSETP, FSEL.  __expf() compiles to the non-ftz
// ex2.approx plus a denormal-range fix-up (FSETP/PLOP3/FMUL x0.5/square: ~13 instructions per element,
// cuobjdump).  With .ftz the result for x*log2(e) < -126 is 0, i.e. ELU = -1: the value the reference's
// clamp returns there (spleeter.c:51-56).  Absolute error ~1e-7, far below the TF32 rounding applied to
// the stored value right after.
__device__ __forceinline__ float exp_fast(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}
template <int ACT>
__device__ __forceinline__ float act_fast(float x)
{
    if (ACT == ACT_LEAKY) return x >= 0.0f ? x : 0.2f * x;
    if (ACT == ACT_RELU) return fmaxf(x, 0.0f);
    if (ACT == ACT_ELU_CLAMP) { const float e = x < -15.0f ? -1.0f : exp_fast(x) - 1.0f; return x >= 0.0f ? x : e; }
    if (ACT == ACT_ELU) { const float e = exp_fast(x) - 1.0f; return x >= 0.0f ? x : e; }
    return x;
}

// o = the stored value (TF32-rounded when `round`), l = what the rounding dropped (exact in fp32; stored as bf16 for the
// consumer's compensation blocks when the layer runs in compensated precision)
template <int ACT>
__device__ __forceinline__ void dec16(const float* v, const float* bias, const float* sc, const float* of, bool round, float* o, float* l)
{
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const float t = fmaf(sc[i], act_fast<ACT>(v[i] + bias[i]), of[i]);
        o[i] = round ? ptx::rna_tf32(t) : t;
        l[i] = t - o[i];
    }
}
template <int ACT>
__device__ __forceinline__ void enc16(const float* raw, const float* sc, const float* of, bool round, float* a, float* l)
{
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const float t = act_fast<ACT>(fmaf(sc[i], raw[i], of[i]));
        a[i] = round ? ptx::rna_tf32(t) : t;
        l[i] = t - a[i];
    }
}

// v[16]: accumulators for channels [c0, c0+16) of pixel (n, Y, X) in tile space.
// stage == nullptr: per-thread stores, callable from divergent code (`valid` must be true; SIMT check kernel).
// stage != nullptr: warp-cooperative stores through the warp's 2 KB staging buffer; all 32 lanes call, lanes
// whose pixel is outside the tensor pass valid = false.
__device__ __forceinline__ void epilogue16(const ConvParams& p, int s, int n, int Y, int X, int phase, int c0, float* v,
                                           float4* stage = nullptr, bool valid = true)
{
    const int act = p.act[s];
    float bias[16];
    load16(p.bias + s * p.cout + c0, bias);
    if (p.mode == 2) {
        float sc[16], of[16], o[16], l[16];
        load16(p.bn_scale + s * p.cout + c0, sc);
        load16(p.bn_offset + s * p.cout + c0, of);
        if (act == ACT_ELU_CLAMP) dec16<ACT_ELU_CLAMP>(v, bias, sc, of, p.round_act, o, l);
        else if (act == ACT_RELU) dec16<ACT_RELU>(v, bias, sc, of, p.round_act, o, l);
        else dec16<ACT_ELU>(v, bias, sc, of, p.round_act, o, l);
        const int oy = 2 * Y + (phase >> 1), ox = 2 * X + (phase & 1);
        const size_t opix = ((size_t)n * (2 * p.Hs) + oy) * (2 * p.Ws) + ox;
        float* dst = p.out_dec + opix * p.cout + c0;
        if (stage) store16_warp(valid ? dst : nullptr, o, stage);
        else store16(dst, o);
        if (p.lo_dec) store16_residual(p.lo_dec, opix * p.lo_dec_C + p.lo_dec_coff + c0, p.lo_dec_fp8, valid, l);
        return;
    }
    float raw[16];
#pragma unroll
    for (int i = 0; i < 16; i++) raw[i] = v[i] + bias[i];
    if (p.mode == 0) {
        float sc[16], of[16], a[16], l[16];
        load16(p.bn_scale + s * p.cout + c0, sc);
        load16(p.bn_offset + s * p.cout + c0, of);
        if (act == ACT_ELU_CLAMP) enc16<ACT_ELU_CLAMP>(raw, sc, of, p.round_act, a, l);
        else if (act == ACT_LEAKY) enc16<ACT_LEAKY>(raw, sc, of, p.round_act, a, l);
        else enc16<ACT_ELU>(raw, sc, of, p.round_act, a, l);
        const size_t apix = (((size_t)n * (p.Hs / 2) + Y / 2) * (p.Ws / 2) + X / 2) * 4 + (Y & 1) * 2 + (X & 1);
        float* dst = p.out_act + apix * p.cout + c0;
        if (stage) store16_warp(valid ? dst : nullptr, a, stage);
        else store16(dst, a);
        if (p.lo_act) store16_residual(p.lo_act, apix * p.cout + c0, p.lo_act_fp8, valid, l);
    }
    const size_t rpix = ((size_t)n * p.Hs + Y) * p.Ws + X;
    if (p.round_raw) {
        float l[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const float hi = ptx::rna_tf32(raw[i]);
            l[i] = raw[i] - hi;
            raw[i] = hi;
        }
        if (p.lo_raw) store16_residual(p.lo_raw, rpix * p.lo_raw_C + p.lo_raw_coff + c0, p.lo_raw_fp8, valid, l);
    }
    float* dst = p.out_raw + rpix * p.cout + c0;
    if (stage) store16_warp(valid ? dst : nullptr, raw, stage);
    else store16(dst, raw);
}

}  // namespace srt
