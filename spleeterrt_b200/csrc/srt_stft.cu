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
#include <algorithm>

#include "srt_fft.cuh"
#include "srt_kernels.cuh"
#include "srt_ptx.cuh"

namespace srt {

// =========================================================================================
// STFT: one CTA per (tile image, frame).  Writes the spectrum row (both channels) and the
// magnitude row the U-Net consumes.
// =========================================================================================
__global__ void __launch_bounds__(kFftThreads) stft_kernel(const StftParams p)
{
    __shared__ FftSmem sm;
    const int img = blockIdx.x / p.T, t = blockIdx.x % p.T;
    const ImgDesc d = p.imgs[img];
    const int f = d.f0 + t;
    const int j = threadIdx.x;
    float4* srow = p.spec + ((size_t)img * p.T + t) * kBins;
    float2* mimg = reinterpret_cast<float2*>(p.mag) + (size_t)img * p.T * p.F;   // space-to-depth image
    const int nfr = p.n_frames[d.stream];
    if (f >= nfr) {
        // zero-padded tail of the last tile (main.c:507-514)
        for (int k = j; k < kBins; k += kFftThreads) srow[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = j; k < p.F; k += kFftThreads) {
            const size_t mi = mag_s2d_index(p.T, p.F, t, k);
            mimg[mi] = make_float2(0.f, 0.f);
            mimg[mi + p.mag_lo_off / 2] = make_float2(0.f, 0.f);
        }
        return;
    }
    const int n = p.n_samples[d.stream];
    const float* L = p.pcmL[d.stream];
    const float* R = p.pcmR[d.stream];
    const int ps = p.pcm_stride ? p.pcm_stride[d.stream] : 1;
    const long base = (long)f * kHop - p.front_pad;
    float2 v[16];
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const int i = j + 256 * r;
        const long si = base + i;
        float2 x = make_float2(0.f, 0.f);
        if (si >= 0 && si < n) {
            const float w = __ldg(&p.window[i]);
            x = make_float2(L[si * ps] * w, R[si * ps] * w);
        }
        v[r] = x;
    }
    fft_smem_init(sm, p.twiddle, j);
    fft4096(v, sm, p.twiddle, j);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; r++) sm.x[pad_idx(j + 256 * r)] = v[r];
    __syncthreads();
    // Z = FFT(l + i r):  XL[k] = (Z[k] + conj Z[N-k]) / 2,  XR[k] = (Z[k] - conj Z[N-k]) / (2i)
    for (int k = j; k < kBins; k += kFftThreads) {
        const float2 za = sm.x[pad_idx(k)], zb = sm.x[pad_idx((kFFT - k) & (kFFT - 1))];
        const float ar = za.x, ai = za.y, br = zb.x, bi = zb.y;
        float4 o;
        o.x = 0.5f * (ar + br);          // Re XL
        o.y = -0.5f * (ai - bi);         // -Im XL   (reference stores the conjugate)
        o.z = 0.5f * (ai + bi);          // Re XR
        o.w = 0.5f * (ar - br);          // -Im XR = -(-(ar - br)/2)
        if (k == 0 || k == kFFT / 2) { o.y = 0.f; o.w = 0.f; }
        srow[k] = o;
        if (k < p.F) {
            const float mL = hypotf(o.x, o.y) * (float)kFFT, mR = hypotf(o.z, o.w) * (float)kFFT;
            const float hL = ptx::rna_tf32(mL), hR = ptx::rna_tf32(mR);
            const size_t mi = mag_s2d_index(p.T, p.F, t, k);
            mimg[mi] = make_float2(hL, hR);
            mimg[mi + p.mag_lo_off / 2] = make_float2(mL - hL, mR - hR);
        }
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
__global__ void __launch_bounds__(kFftThreads, 4) istft_kernel(const IstftParams p)
{
    __shared__ FftSmem sm;
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
    fft_smem_init(sm, p.twiddle, j);
    fft4096(v, sm, p.twiddle, j);
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
// Fused mask*spectrum -> inverse FFT -> synthesis window -> overlap-add -> un-framing.
// One CTA owns hops [h0, h0+G) of one (stream, stem): it transforms frames h0-3 .. h0+G-1 in order and
// overlap-adds them in REGISTERS: thread j holds samples j + 256 r of every frame, and a frame advances the
// output position by 1024 = 4 * 256 samples, so the 4096-sample overlap window is 16 thread-private
// accumulators that shift down by 4 per frame.  acc[0..3] is complete after frame f has been added (frames
// f-3 .. f, the summation order of stftFix.c:570-575) and is emitted as segment f.  (G+3)/G of the FFT work,
// no 32 KB-per-frame scratch round trip, no separate OLA kernel, and no shared-memory accumulator: the
// previous shared-memory ring and the per-frame window loads were 28 % of this kernel's L1 wavefronts, its
// bound (ncu r1k/r1m: l1tex 67 %, dram 15 %).
// =========================================================================================
__global__ void __launch_bounds__(kFftThreads, 2) istft_ola_kernel(const IstftOlaParams p)
{
    __shared__ FftSmem sm;
    const int st = p.stream_first + blockIdx.z, s = blockIdx.y;
    const int nfr = p.n_frames[st];
    const int h0 = blockIdx.x * p.hops_per_cta;
    if (h0 >= nfr) return;
    const int h1 = min(h0 + p.hops_per_cta, nfr);
    const int j = threadIdx.x;
    const int n = p.n_samples[st], img0 = p.stream_img0[st];
    float* outL = p.out[((size_t)st * p.out_pairs + p.pair_first + s) * 2];
    float* outR = p.out[((size_t)st * p.out_pairs + p.pair_first + s) * 2 + 1];
    const bool masked = s < p.S_masked;   // else: the plain inverse transform of the spectrum (main.c:881)
    const float uw = masked ? p.unaffected[s] : 1.0f;
    const int os = p.out_stride;
    float w[16];
    float2 acc[16];
#pragma unroll
    for (int r = 0; r < 16; r++) {
        w[r] = __ldg(&p.postwin[j + 256 * r]);
        acc[r] = make_float2(0.f, 0.f);
    }
    fft_smem_init(sm, p.twiddle, j);
    for (int f = max(h0 - 3, 0); f < h1; f++) {
        const int img = img0 + f / p.T, t = f % p.T;
        const float4* srow = p.spec + ((size_t)img * p.T + t) * kBins;
        const float2* mrow = reinterpret_cast<const float2*>(p.mask) + (((size_t)(masked ? s : 0) * p.mask_stem_stride + img) * p.T + t) * p.F;
        float2 v[16];
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const int k = j + 256 * r;
            const int kk = k <= kFFT / 2 ? k : kFFT - k;
            const float4 sp = srow[kk];
            float mL = uw, mR = uw;
            if (masked && kk < p.F) { const float2 m = mrow[kk]; mL = m.x; mR = m.y; }
            const float xlr = sp.x * mL, xli = -(sp.y * mL), xrr = sp.z * mR, xri = -(sp.w * mR);
            float2 z;
            if (k <= kFFT / 2) z = make_float2(xlr - xri, xli + xrr);
            else z = make_float2(xlr + xri, -xli + xrr);
            v[r] = make_float2(z.x, -z.y);
        }
        __syncthreads();   // the previous frame's pass-3 reads of the exchange buffer are done (also publishes tw2)
        fft4096(v, sm, p.twiddle, j);
#pragma unroll
        for (int r = 0; r < 16; r++) {
            acc[r].x += v[r].x * w[r];
            acc[r].y += -v[r].y * w[r];
        }
        if (f >= h0) {   // segment f is complete: frames f-3 .. f have been added
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const long io = (long)f * kHop + j + 256 * k - p.front_pad;
                if (io >= 0 && io < n) { outL[io * os] = acc[k].x; outR[io * os] = acc[k].y; }
            }
        }
#pragma unroll
        for (int r = 0; r < 12; r++) acc[r] = acc[r + 4];
#pragma unroll
        for (int r = 12; r < 16; r++) acc[r] = make_float2(0.f, 0.f);
    }
}

// Same walk, with the next frame's rows staged by the copy engine while the current frame is transformed.
// The register kernel above exposes one global-load round trip per frame (16 LDG.128 + 16 LDG.64 per thread in
// front of the first butterfly; two 8-warp CTAs per SM cannot hide it: ncu r1o issue slots 56 % busy, DRAM 21 %).
// Here one thread issues two 1-D bulk copies (spectrum row 32 784 B, mask row 8 F B) for frame f+1 right after the
// CTA has finished reading frame f's rows, an mbarrier publishes them, and the gather of the Hermitian extension
// reads shared memory (each bin is needed twice, for k and N-k).
constexpr uint32_t kSpecRowBytes = kBins * sizeof(float4);
static size_t istft_pf_smem(int F) { return sizeof(FftSmem) + kSpecRowBytes + (size_t)F * sizeof(float2) + 16; }

__global__ void __launch_bounds__(kFftThreads, 2) istft_ola_pf_kernel(const IstftOlaParams p)
{
    extern __shared__ __align__(128) uint8_t dsm[];
    FftSmem& sm = *reinterpret_cast<FftSmem*>(dsm);
    const float4* sspec = reinterpret_cast<const float4*>(dsm + sizeof(FftSmem));
    const float2* smask = reinterpret_cast<const float2*>(dsm + sizeof(FftSmem) + kSpecRowBytes);
    uint64_t* bar = reinterpret_cast<uint64_t*>(dsm + sizeof(FftSmem) + kSpecRowBytes + (size_t)p.F * sizeof(float2));
    const int st = p.stream_first + blockIdx.z, s = blockIdx.y;
    const int nfr = p.n_frames[st];
    const int h0 = blockIdx.x * p.hops_per_cta;
    if (h0 >= nfr) return;
    const int h1 = min(h0 + p.hops_per_cta, nfr);
    const int j = threadIdx.x;
    const int n = p.n_samples[st], img0 = p.stream_img0[st];
    float* outL = p.out[((size_t)st * p.out_pairs + p.pair_first + s) * 2];
    float* outR = p.out[((size_t)st * p.out_pairs + p.pair_first + s) * 2 + 1];
    const bool masked = s < p.S_masked;   // else: the plain inverse transform of the spectrum (main.c:881)
    const float uw = masked ? p.unaffected[s] : 1.0f;
    const int os = p.out_stride;
    const int Fm = masked ? p.F : 0;      // bins below Fm take the staged mask
    auto stage = [&](int f) {             // one thread: rows of frame f -> shared memory
        const int img = img0 + f / p.T, t = f % p.T;
        const float4* srow = p.spec + ((size_t)img * p.T + t) * kBins;
        ptx::mbar_arrive_expect_tx(bar, kSpecRowBytes + (uint32_t)Fm * (uint32_t)sizeof(float2));
        ptx::bulk_load_1d((void*)sspec, srow, kSpecRowBytes, bar);
        if (masked) {
            const float2* mrow = reinterpret_cast<const float2*>(p.mask) + (((size_t)s * p.mask_stem_stride + img) * p.T + t) * p.F;
            ptx::bulk_load_1d((void*)smask, mrow, (uint32_t)p.F * (uint32_t)sizeof(float2), bar);
        }
    };
    const int f_first = max(h0 - 3, 0);
    if (j == 0) {
        ptx::mbar_init(bar, 1);
        ptx::fence_barrier_init();
        stage(f_first);
    }
    float w[16];
    float2 acc[16];
#pragma unroll
    for (int r = 0; r < 16; r++) {
        w[r] = __ldg(&p.postwin[j + 256 * r]);
        acc[r] = make_float2(0.f, 0.f);
    }
    fft_smem_init(sm, p.twiddle, j);
    __syncthreads();                      // barrier initialised before anyone polls it
    uint32_t phase = 0;
    for (int f = f_first; f < h1; f++) {
        ptx::mbar_wait(bar, phase);
        phase ^= 1;
        float2 v[16];
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const int k = j + 256 * r;
            const int kk = k <= kFFT / 2 ? k : kFFT - k;
            const float4 sp = sspec[kk];
            float mL = uw, mR = uw;
            if (kk < Fm) { const float2 m = smask[kk]; mL = m.x; mR = m.y; }
            const float xlr = sp.x * mL, xli = -(sp.y * mL), xrr = sp.z * mR, xri = -(sp.w * mR);
            float2 z;
            if (k <= kFFT / 2) z = make_float2(xlr - xri, xli + xrr);
            else z = make_float2(xlr + xri, -xli + xrr);
            v[r] = make_float2(z.x, -z.y);
        }
        __syncthreads();   // staged rows consumed; the previous frame's pass-3 reads of the exchange buffer are done
        if (j == 0 && f + 1 < h1) stage(f + 1);
        fft4096(v, sm, p.twiddle, j);
#pragma unroll
        for (int r = 0; r < 16; r++) {
            acc[r].x += v[r].x * w[r];
            acc[r].y += -v[r].y * w[r];
        }
        if (f >= h0) {   // segment f is complete: frames f-3 .. f have been added
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const long io = (long)f * kHop + j + 256 * k - p.front_pad;
                if (io >= 0 && io < n) { outL[io * os] = acc[k].x; outR[io * os] = acc[k].y; }
            }
        }
#pragma unroll
        for (int r = 0; r < 12; r++) acc[r] = acc[r + 4];
#pragma unroll
        for (int r = 12; r < 16; r++) acc[r] = make_float2(0.f, 0.f);
    }
}

void launch_istft_ola(const IstftOlaParams& p, int n_streams, int max_frames, cudaStream_t st)
{
    if (n_streams == 0 || max_frames == 0) return;
    dim3 grid((max_frames + p.hops_per_cta - 1) / p.hops_per_cta, p.S, n_streams);
    static const int prefetch = [] { const char* e = getenv("SRT_ISTFT_PREFETCH"); return e ? atoi(e) : 1; }();
    if (prefetch) {
        static LaunchState state;
        const size_t smem = istft_pf_smem(p.F);
        state.prepare(istft_ola_pf_kernel, istft_pf_smem(2048));
        istft_ola_pf_kernel<<<grid, kFftThreads, smem, st>>>(p);
    } else {
        istft_ola_kernel<<<grid, kFftThreads, 0, st>>>(p);
    }
}

// =========================================================================================
// Cascade stage 2 (main.c:849-865, 911): residual = spec - mask * spec, and its magnitudes.  One CTA per
// (tile image, frame), like stft_kernel; the subtraction is done on the products the reference forms
// (orig - re*mask), not on (1 - mask).
// =========================================================================================
__global__ void __launch_bounds__(256) residual_kernel(const ResidualParams p)
{
    const int img = blockIdx.x / p.T, t = blockIdx.x % p.T;
    const ImgDesc d = p.imgs[img];
    const bool live = d.f0 + t < p.n_frames[d.stream];
    const float4* in = p.spec_in + ((size_t)img * p.T + t) * kBins;
    const float2* mrow = reinterpret_cast<const float2*>(p.mask_in) + ((size_t)img * p.T + t) * p.F;
    float4* out = p.spec_out + ((size_t)img * p.T + t) * kBins;
    float2* mimg = reinterpret_cast<float2*>(p.mag) + (size_t)img * p.T * p.F;
    for (int k = threadIdx.x; k < kBins; k += 256) {
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) {
            const float4 sp = in[k];
            float mL = p.unaffected, mR = p.unaffected;
            if (k < p.F) { const float2 m = mrow[k]; mL = m.x; mR = m.y; }
            // two roundings like the reference (product stored, then subtracted), not one FMA
            o = make_float4(__fsub_rn(sp.x, __fmul_rn(sp.x, mL)), __fsub_rn(sp.y, __fmul_rn(sp.y, mL)), __fsub_rn(sp.z, __fmul_rn(sp.z, mR)),
                            __fsub_rn(sp.w, __fmul_rn(sp.w, mR)));
        }
        out[k] = o;
        if (k < p.F) {
            const float mL = hypotf(o.x, o.y) * (float)kFFT, mR = hypotf(o.z, o.w) * (float)kFFT;
            const float hL = ptx::rna_tf32(mL), hR = ptx::rna_tf32(mR);
            const size_t mi = mag_s2d_index(p.T, p.F, t, k);
            mimg[mi] = make_float2(hL, hR);
            mimg[mi + p.mag_lo_off / 2] = make_float2(mL - hL, mR - hR);
        }
    }
}

void launch_residual(const ResidualParams& p, cudaStream_t st)
{
    if (p.n_img == 0) return;
    residual_kernel<<<p.n_img * p.T, 256, 0, st>>>(p);
}

__global__ void __launch_bounds__(256) diff_kernel(const DiffParams p)
{
    const int st = blockIdx.z, c = blockIdx.y;
    const int n = p.n_samples[st];
    float* dst = p.out[((size_t)st * p.out_pairs + p.dst_pair) * 2 + c];
    const float* b = p.out[((size_t)st * p.out_pairs + p.sub_pair) * 2 + c];
    const float* a = p.pcmL ? (c ? p.pcmR[st] : p.pcmL[st]) : dst;
    const long os = p.out_stride, as = p.pcmL ? (p.pcm_stride ? p.pcm_stride[st] : 1) : os;
    for (long i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) dst[i * os] = a[i * as] - b[i * os];
}

void launch_diff(const DiffParams& p, cudaStream_t st)
{
    if (p.n_streams == 0 || p.max_samples == 0) return;
    dim3 grid(std::min((p.max_samples + 255) / 256, 1024), 2, p.n_streams);
    diff_kernel<<<grid, 256, 0, st>>>(p);
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
