// srt_stft.cu — STFT framer, mask·spectrum + inverse transform, overlap-add.
//
// Replaces stft()/istft()/DFT4096 of the reference (Executable/stftFix.c:363-579,
// codelet.c:2-271) and the magnitude / mask loops of the tile driver (main.c:459-494).
// The reference runs a real 4096-point Hartley transform per channel; here the two
// channels of a stereo frame ride one complex 4096-point FFT (z = L + iR), evaluated as
// three radix-16 Stockham passes: 256 threads x 16 points in registers, exchanges through
// padded shared memory, twiddles from a 4096-entry table.
//
// Conventions kept from the reference: re = Re FFT(x*hann)/4096,
// im = -Im FFT(x*hann)/4096 (conjugate), magnitude = hypot(re, im) * 4096.
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

// 4096-point forward FFT.  In: v[r] = x[j + 256 r].  Out: v[r] = X[j + 256 r].
__device__ __forceinline__ void fft4096(float2* v, float* sre, float* sim, const float2* __restrict__ tw, int j)
{
    // pass 1 (Ns = 1)
    fft16(v);
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const int i = 17 * j + r;   // pad_idx(16 j + r)
        sre[i] = v[r].x;
        sim[i] = v[r].y;
    }
    __syncthreads();
    // pass 2 (Ns = 16)
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const int i = pad_idx(j + 256 * r);
        v[r] = make_float2(sre[i], sim[i]);
    }
    twiddle_powers(v, __ldg(&tw[(j & 15) * 16]));
    fft16(v);
    __syncthreads();
    {
        const int base = (j >> 4) * 256 + (j & 15);
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const int i = pad_idx(base + 16 * r);
            sre[i] = v[r].x;
            sim[i] = v[r].y;
        }
    }
    __syncthreads();
    // pass 3 (Ns = 256)
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const int i = pad_idx(j + 256 * r);
        v[r] = make_float2(sre[i], sim[i]);
    }
    twiddle_powers(v, __ldg(&tw[j]));
    fft16(v);
}

// =========================================================================================
// STFT: one CTA per (tile image, frame).  Writes the spectrum row (both channels) and the
// magnitude row the U-Net consumes.
// =========================================================================================
__global__ void __launch_bounds__(kFftThreads) stft_kernel(const StftParams p)
{
    __shared__ float sre[kPadLen], sim[kPadLen];
    const int img = blockIdx.x / p.T, t = blockIdx.x % p.T;
    const ImgDesc d = p.imgs[img];
    const int f = d.f0 + t;
    const int j = threadIdx.x;
    float4* srow = p.spec + ((size_t)img * p.T + t) * kBins;
    float2* mrow = reinterpret_cast<float2*>(p.mag) + ((size_t)img * p.T + t) * p.F;
    const int nfr = p.n_frames[d.stream];
    if (f >= nfr) {
        // zero-padded tail of the last tile (main.c:507-514)
        for (int k = j; k < kBins; k += kFftThreads) srow[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = j; k < p.F; k += kFftThreads) mrow[k] = make_float2(0.f, 0.f);
        return;
    }
    const int n = p.n_samples[d.stream];
    const float* L = p.pcmL[d.stream];
    const float* R = p.pcmR[d.stream];
    const long base = (long)f * kHop - p.front_pad;
    float2 v[16];
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const int i = j + 256 * r;
        const long si = base + i;
        float2 x = make_float2(0.f, 0.f);
        if (si >= 0 && si < n) {
            const float w = __ldg(&p.window[i]);
            x = make_float2(L[si] * w, R[si] * w);
        }
        v[r] = x;
    }
    fft4096(v, sre, sim, p.twiddle, j);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const int i = pad_idx(j + 256 * r);
        sre[i] = v[r].x;
        sim[i] = v[r].y;
    }
    __syncthreads();
    // Z = FFT(l + i r):  XL[k] = (Z[k] + conj Z[N-k]) / 2,  XR[k] = (Z[k] - conj Z[N-k]) / (2i)
    for (int k = j; k < kBins; k += kFftThreads) {
        const int ia = pad_idx(k), ib = pad_idx((kFFT - k) & (kFFT - 1));
        const float ar = sre[ia], ai = sim[ia], br = sre[ib], bi = sim[ib];
        float4 o;
        o.x = 0.5f * (ar + br);          // Re XL
        o.y = -0.5f * (ai - bi);         // -Im XL   (reference stores the conjugate)
        o.z = 0.5f * (ai + bi);          // Re XR
        o.w = 0.5f * (ar - br);          // -Im XR = -(-(ar - br)/2)
        if (k == 0 || k == kFFT / 2) { o.y = 0.f; o.w = 0.f; }
        srow[k] = o;
        if (k < p.F) mrow[k] = make_float2(hypotf(o.x, o.y) * (float)kFFT, hypotf(o.z, o.w) * (float)kFFT);
    }
}

void launch_stft(const StftParams& p, cudaStream_t st)
{
    if (p.n_img == 0) return;
    stft_kernel<<<p.n_img * p.T, kFftThreads, 0, st>>>(p);
}

// =========================================================================================
// mask·spectrum -> inverse FFT -> synthesis window.  One CTA per (frame, stem); both channels
// ride one complex transform.  Windowed time frames go to a scratch buffer; ola_kernel sums
// the 4 overlapping frames per output sample in frame order (stftFix.c:570-575).
// =========================================================================================
__global__ void __launch_bounds__(kFftThreads) istft_kernel(const IstftParams p)
{
    __shared__ float sre[kPadLen], sim[kPadLen];
    const int li = blockIdx.x / p.T, t = blockIdx.x % p.T;
    const int img = p.img_first + li;
    const int s = blockIdx.y;
    const ImgDesc d = p.imgs[img];
    const int f = d.f0 + t;
    if (f >= p.n_frames[d.stream]) return;
    const int j = threadIdx.x;
    const float4* srow = p.spec + ((size_t)img * p.T + t) * kBins;
    const float2* mrow = p.mask ? reinterpret_cast<const float2*>(p.mask) + (((size_t)s * p.mask_stem_stride + img) * p.T + t) * p.F : nullptr;
    const float uw = p.unaffected[s];
    float2 v[16];
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const int k = j + 256 * r;
        const int kk = k <= kFFT / 2 ? k : kFFT - k;
        const float4 sp = srow[kk];
        float mL = uw, mR = uw;
        if (mrow) {
            if (kk < p.F) { const float2 m = mrow[kk]; mL = m.x; mR = m.y; }
        } else {
            mL = mR = 1.0f;
        }
        // true spectra: XL = reL - i imL, XR = reR - i imR (masked, main.c:476-493)
        const float xlr = sp.x * mL, xli = -(sp.y * mL), xrr = sp.z * mR, xri = -(sp.w * mR);
        float2 z;
        if (k <= kFFT / 2) z = make_float2(xlr - xri, xli + xrr);        // XL + i XR
        else z = make_float2(xlr + xri, -xli + xrr);                      // conj(XL) + i conj(XR)
        v[r] = make_float2(z.x, -z.y);                                    // conj(Z): inverse via forward FFT
    }
    fft4096(v, sre, sim, p.twiddle, j);
    float2* out = p.frames_out + (((size_t)s * p.frames_stem_stride + li) * p.T + t) * kFFT;
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const int i = j + 256 * r;
        const float w = __ldg(&p.postwin[i]);
        out[i] = make_float2(v[r].x * w, -v[r].y * w);                    // z = conj(Y): l = Re, r = -Im
    }
}

void launch_istft(const IstftParams& p, cudaStream_t st)
{
    if (p.n_img == 0) return;
    dim3 grid(p.n_img * p.T, p.S);
    istft_kernel<<<grid, kFftThreads, 0, st>>>(p);
}

// =========================================================================================
// overlap-add + un-framing (channel_joinFloat preshift, main.c:806): out[i] = ola[front_pad + i]
// =========================================================================================
__global__ void __launch_bounds__(256) ola_kernel(const OlaParams p)
{
    const int st = p.stream_first + blockIdx.z, s = blockIdx.y;
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int n = p.n_samples[st];
    if (i >= n) return;
    const int nfr = p.n_frames[st];
    const int img0 = p.stream_img0[st] - p.img_first;
    const int r = i + p.front_pad;
    const int fhi = min(r / kHop, nfr - 1);
    const int flo = max(r / kHop - 3, 0);
    float aL = 0.f, aR = 0.f;
    for (int f = flo; f <= fhi; f++) {
        const int img = img0 + f / p.T, t = f % p.T;
        const float2 v = p.frames[(((size_t)s * p.frames_stem_stride + img) * p.T + t) * kFFT + (r - f * kHop)];
        aL += v.x;
        aR += v.y;
    }
    float* const* o = p.out + (size_t)st * p.S * 2 + s * 2;
    o[0][i] = aL;
    o[1][i] = aR;
}

void launch_ola(const OlaParams& p, cudaStream_t st)
{
    if (p.n_streams == 0 || p.max_samples == 0) return;
    dim3 grid((p.max_samples + 255) / 256, p.S, p.n_streams);
    ola_kernel<<<grid, 256, 0, st>>>(p);
}

}  // namespace srt
