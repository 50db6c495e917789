// srt_fft.cuh — 4096-point complex FFT shared by the offline transforms (srt_stft.cu) and the
// streaming flavour (srt_stream.cu).  Three radix-16 Stockham passes: 256 threads x 16 points in
// registers, padded shared-memory exchanges, twiddles from one table load per pass + powers.
#pragma once
#include "srt_kernels.cuh"

namespace srt {

constexpr int kFftThreads = 256;
constexpr int kPadLen = kFFT + kFFT / 16;

__device__ __forceinline__ int pad_idx(int i) { return i + (i >> 4); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// ---- packed complex arithmetic ------------------------------------------------------------------------------------------
// A complex number travels as one 64-bit register pair (re, im) and is worked on with the packed fp32 instructions of sm_100
// (FADD2 / FMUL2 / FFMA2: two IEEE fp32 operations per issue slot).  The transforms are bound by fp32 instruction issue
// (~700 of ~1000 instructions per thread and frame were scalar FADD / FMUL / FFMA), so halving their count is what speeds
// them up.  What makes it free: ptxas folds the half swap and per-half negation of an operand into the instruction
// (R.F32x2.LO_HI, .NP) and a scalar used twice into a broadcast operand (R.F32), so
//     a + b, a - b          1 FADD2            -i a = (a.im, -a.re)     0 (operand modifier of its consumer)
//     a * w                 1 FMUL2 + 1 FFMA2  (re = a.re w.re - a.im w.im, im = a.re w.im + a.im w.re: the same roundings as the
//                                               scalar form fma(a.re, w.re, -(a.im w.im)), fma(a.re, w.im, a.im w.re))
typedef unsigned long long c64;
__device__ __forceinline__ c64 cpk(float re, float im)
{
    c64 r;
    asm("mov.b64 %0, {%1, %2};\n" : "=l"(r) : "f"(re), "f"(im));
    return r;
}
__device__ __forceinline__ float2 cup(c64 v)
{
    float2 r;
    asm("mov.b64 {%0, %1}, %2;\n" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ c64 cadd(c64 a, c64 b)
{
    c64 r;
    asm("add.rn.f32x2 %0, %1, %2;\n" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ c64 csub(c64 a, c64 b)
{
    c64 r;
    asm("sub.rn.f32x2 %0, %1, %2;\n" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ c64 cmul(c64 a, c64 w)
{
    const float2 av = cup(a), wv = cup(w);
    c64 t, r;
    asm("mul.rn.f32x2 %0, %1, %2;\n" : "=l"(t) : "l"(cpk(av.y, av.y)), "l"(cpk(wv.y, wv.x)));   // (a.im w.im, a.im w.re)
    const float2 tv = cup(t);
    asm("fma.rn.f32x2 %0, %1, %2, %3;\n" : "=l"(r) : "l"(cpk(av.x, av.x)), "l"(w), "l"(cpk(-tv.x, tv.y)));
    return r;
}
__device__ __forceinline__ c64 cmulc(c64 a, float wr, float wi) { return cmul(a, cpk(wr, wi)); }
__device__ __forceinline__ c64 cmul_mi(c64 a)   // -i a
{
    const float2 v = cup(a);
    return cpk(v.y, -v.x);
}

// forward DFT-4 in place, natural order out
__device__ __forceinline__ void fft4(c64& a0, c64& a1, c64& a2, c64& a3)
{
    const c64 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = cmul_mi(csub(a1, a3));
    a0 = cadd(t0, t2);
    a2 = csub(t0, t2);
    a1 = cadd(t1, t3);
    a3 = csub(t1, t3);
}

// forward DFT-16 of v[0..15] (natural order in, natural order out)
__device__ __forceinline__ void fft16(c64* v)
{
    // n = 4*n1 + n2, k = k1 + 4*k2
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) fft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);   // v[4*k1 + n2] = y[n2][k1]
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
    // W16^m = exp(-2 pi i m / 16)
    v[4 * 1 + 1] = cmulc(v[4 * 1 + 1], c1, -s1);    // m = 1
    v[4 * 1 + 2] = cmulc(v[4 * 1 + 2], h, -h);      // m = 2
    v[4 * 1 + 3] = cmulc(v[4 * 1 + 3], s1, -c1);    // m = 3
    v[4 * 2 + 1] = cmulc(v[4 * 2 + 1], h, -h);      // m = 2
    v[4 * 2 + 2] = cmul_mi(v[4 * 2 + 2]);           // m = 4: -i
    v[4 * 2 + 3] = cmulc(v[4 * 2 + 3], -h, -h);     // m = 6
    v[4 * 3 + 1] = cmulc(v[4 * 3 + 1], s1, -c1);    // m = 3
    v[4 * 3 + 2] = cmulc(v[4 * 3 + 2], -h, -h);     // m = 6
    v[4 * 3 + 3] = cmulc(v[4 * 3 + 3], -c1, s1);    // m = 9
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) fft4(v[4 * k1 + 0], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);   // v[4*k1 + k2] = X[k1 + 4*k2]
    c64 t[16];
#pragma unroll
    for (int r = 0; r < 16; r++) t[r] = v[(r & 3) * 4 + (r >> 2)];
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = t[r];
}

// v[r] *= w^r for r = 1..15, powers built by squaring / one multiply (depth <= 4, error ~4 ulp).
// One coalesced table load per thread instead of 15 gathers: the scattered twiddle loads were the
// main L1 traffic of the transform kernels (ncu r1c: l1tex 91%, 480 of ~700 wavefronts per FFT).
__device__ __forceinline__ void twiddle_powers(c64* v, c64 w)
{
    c64 p[16];
    p[1] = w;
#pragma unroll
    for (int r = 2; r < 16; r++) p[r] = (r & 1) ? cmul(p[r - 1], w) : cmul(p[r >> 1], p[r >> 1]);
#pragma unroll
    for (int r = 1; r < 16; r++) v[r] = cmul(v[r], p[r]);
}

// Shared-memory state of one transform: the padded exchange buffer (complex, so each exchange is one 64-bit
// access per point; every pattern below touches 16 consecutive float2 per half-warp -> conflict-free) and the
// pass-2 twiddles.  Pass 2 needs w^r with w = W4096^(16 (j & 15)): only 16 x 15 distinct values, kept as a
// [r][j & 15] table (one conflict-free LDS.64 per factor instead of the power chain).
struct FftSmem {
    float2 x[kPadLen];
    float2 tw2[16 * 16];
};

// Fills the pass-2 table.  Call once per CTA, before the first fft4096 (any barrier inside it publishes the table).
__device__ __forceinline__ void fft_smem_init(FftSmem& sm, const float2* __restrict__ tw, int j)
{
    const int r = j >> 4, jj = j & 15;
    sm.tw2[j] = __ldg(&tw[(jj * r * 16) & (kFFT - 1)]);
}

// 4096-point forward FFT.  In: v[r] = x[j + 256 r].  Out: v[r] = X[j + 256 r].
__device__ __forceinline__ void fft4096(float2* vf, FftSmem& sm, const float2* __restrict__ tw, int j)
{
    c64* sx = reinterpret_cast<c64*>(sm.x);
    const c64* tw2 = reinterpret_cast<const c64*>(sm.tw2);
    c64 v[16];
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = cpk(vf[r].x, vf[r].y);
    // pass 1 (Ns = 1)
    fft16(v);
#pragma unroll
    for (int r = 0; r < 16; r++) sx[17 * j + r] = v[r];   // pad_idx(16 j + r)
    __syncthreads();
    // pass 2 (Ns = 16)
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = sx[pad_idx(j + 256 * r)];
#pragma unroll
    for (int r = 1; r < 16; r++) v[r] = cmul(v[r], tw2[r * 16 + (j & 15)]);
    fft16(v);
    __syncthreads();
    {
        const int base = (j >> 4) * 256 + (j & 15);
#pragma unroll
        for (int r = 0; r < 16; r++) sx[pad_idx(base + 16 * r)] = v[r];
    }
    __syncthreads();
    // pass 3 (Ns = 256)
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = sx[pad_idx(j + 256 * r)];
    const float2 w = __ldg(&tw[j]);
    twiddle_powers(v, cpk(w.x, w.y));
    fft16(v);
#pragma unroll
    for (int r = 0; r < 16; r++) vf[r] = cup(v[r]);
}

}  // namespace srt
