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

// forward DFT-4 in place, natural order out
__device__ __forceinline__ void fft4(float2& a0, float2& a1, float2& a2, float2& a3)
{
    const float2 t0 = make_float2(a0.x + a2.x, a0.y + a2.y);
    const float2 t1 = make_float2(a0.x - a2.x, a0.y - a2.y);
    const float2 t2 = make_float2(a1.x + a3.x, a1.y + a3.y);
    const float2 t3 = make_float2(a1.y - a3.y, -(a1.x - a3.x));   // -i (a1 - a3)
    a0 = make_float2(t0.x + t2.x, t0.y + t2.y);
    a2 = make_float2(t0.x - t2.x, t0.y - t2.y);
    a1 = make_float2(t1.x + t3.x, t1.y + t3.y);
    a3 = make_float2(t1.x - t3.x, t1.y - t3.y);
}

// forward DFT-16 of v[0..15] (natural order in, natural order out)
__device__ __forceinline__ void fft16(float2* v)
{
    // n = 4*n1 + n2, k = k1 + 4*k2
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) fft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);   // v[4*k1 + n2] = y[n2][k1]
    const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
    // W16^m = exp(-2 pi i m / 16)
    v[4 * 1 + 1] = cmul(v[4 * 1 + 1], make_float2(c1, -s1));    // m = 1
    v[4 * 1 + 2] = cmul(v[4 * 1 + 2], make_float2(h, -h));      // m = 2
    v[4 * 1 + 3] = cmul(v[4 * 1 + 3], make_float2(s1, -c1));    // m = 3
    v[4 * 2 + 1] = cmul(v[4 * 2 + 1], make_float2(h, -h));      // m = 2
    v[4 * 2 + 2] = make_float2(v[4 * 2 + 2].y, -v[4 * 2 + 2].x);   // m = 4: -i
    v[4 * 2 + 3] = cmul(v[4 * 2 + 3], make_float2(-h, -h));     // m = 6
    v[4 * 3 + 1] = cmul(v[4 * 3 + 1], make_float2(s1, -c1));    // m = 3
    v[4 * 3 + 2] = cmul(v[4 * 3 + 2], make_float2(-h, -h));     // m = 6
    v[4 * 3 + 3] = cmul(v[4 * 3 + 3], make_float2(-c1, s1));    // m = 9
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) fft4(v[4 * k1 + 0], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);   // v[4*k1 + k2] = X[k1 + 4*k2]
    float2 t[16];
#pragma unroll
    for (int r = 0; r < 16; r++) t[r] = v[(r & 3) * 4 + (r >> 2)];
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = t[r];
}

// v[r] *= w^r for r = 1..15, powers built by squaring / one multiply (depth <= 4, error ~4 ulp).
// One coalesced table load per thread instead of 15 gathers: the scattered twiddle loads were the
// main L1 traffic of the transform kernels (ncu r1c: l1tex 91%, 480 of ~700 wavefronts per FFT).
__device__ __forceinline__ void twiddle_powers(float2* v, float2 w)
{
    float2 p[16];
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
__device__ __forceinline__ void fft4096(float2* v, FftSmem& sm, const float2* __restrict__ tw, int j)
{
    float2* sx = sm.x;
    // pass 1 (Ns = 1)
    fft16(v);
#pragma unroll
    for (int r = 0; r < 16; r++) sx[17 * j + r] = v[r];   // pad_idx(16 j + r)
    __syncthreads();
    // pass 2 (Ns = 16)
#pragma unroll
    for (int r = 0; r < 16; r++) v[r] = sx[pad_idx(j + 256 * r)];
#pragma unroll
    for (int r = 1; r < 16; r++) v[r] = cmul(v[r], sm.tw2[r * 16 + (j & 15)]);
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
    twiddle_powers(v, __ldg(&tw[j]));
    fft16(v);
}

}  // namespace srt
