// srt_stream.cu — the real-time streaming flavour of the path (VST/Source/Spleeter4Stems.c) on the GPU.
//
// Per 1024-sample hop ONE kernel launch does what LLPAMSProcessNPR (Spleeter4Stems.c:257-379) and its
// helper thread task_type1 (:40-113) do with ten Hartley transforms on the CPU:
//   * CTA S        : analysis — asymmetric window (:383-416), packed stereo FFT, spectrum + magnitude of the
//                    NEW frame into the tile being collected (:322-349);
//   * CTAs 0..S-1  : synthesis — mask * spectrum of the frame recorded two tiles ago (:272-297, 64-89),
//                    inverse FFT, synthesis window on the last 2048 samples (:303-309), 50 % overlap-add
//                    with the previous hop (:313-320).
// Every T hops the S U-Nets run on the finished tile on the context's own stream (:351-371); the hop
// stream only waits for them (on the device, via an event) at the tile boundary where the reference
// joins its NN threads.  Buffers: spectra ring of 3 tiles (write k, read k-2), masks / magnitudes x2.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/Spleeter4Stems.h"
#undef FFTSIZE
#undef HALFWNDLEN
#include "../../include/srt_b200.h"
#include "srt_fft.cuh"
#include "srt_internal.h"
#include "srt_kernels.cuh"
#include "srt_ptx.cuh"

using namespace srt;

namespace {

struct HopParams {
    const float* ring;        // device input ring [2][4096]
    int pos;                  // index of the oldest sample of the current frame
    const float* awin;        // analysis window, 2 * reference analysisWnd (so the FFT yields the reference's re / -im directly)
    const float* swin;        // synthesis window, pre-shifted: 2048 entries
    const float2* twiddle;
    float4* spec_new;         // [2049] row being recorded
    float2* mag_new;          // space-to-depth magnitude image being recorded (row = cursor), hi part
    size_t mag_lo_off;        // float2 offset of the lo part
    const float4* spec_old;   // [2049] row recorded two tiles ago
    const float* mask_old;    // masks of that tile: [S][T][F][2]
    int cursor, T, F, S;
    float uw[8];
    const float* overlap_in;  // [2S][1024]
    float* overlap_out;       // [2S][1024]
    float* out;               // [2S][1024]
};

__global__ void __launch_bounds__(kFftThreads) stream_hop_kernel(const HopParams p)
{
    __shared__ FftSmem sm;
    const int j = threadIdx.x;
    const int role = blockIdx.x;
    float2 v[16];
    fft_smem_init(sm, p.twiddle, j);
    if (role == p.S) {
        // ---- analysis of the newest frame (Spleeter4Stems.c:261-267, 322-349) --------------------
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const int i = j + 256 * r;
            const int k = (i + p.pos) & (kFFT - 1);
            const float w = __ldg(&p.awin[i]);
            v[r] = make_float2(p.ring[k] * w, p.ring[kFFT + k] * w);
        }
        fft4096(v, sm, p.twiddle, j);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; r++) sm.x[pad_idx(j + 256 * r)] = v[r];
        __syncthreads();
        for (int k = j; k < kBins; k += kFftThreads) {
            const float2 za = sm.x[pad_idx(k)], zb = sm.x[pad_idx((kFFT - k) & (kFFT - 1))];
            const float ar = za.x, ai = za.y, br = zb.x, bi = zb.y;
            float4 o;
            o.x = 0.5f * (ar + br);
            o.y = -0.5f * (ai - bi);
            o.z = 0.5f * (ai + bi);
            o.w = 0.5f * (ar - br);
            if (k == 0 || k == kFFT / 2) { o.y = 0.f; o.w = 0.f; }
            p.spec_new[k] = o;
            if (k < p.F) {
                const float mL = hypotf(o.x, o.y) * (float)kFFT, mR = hypotf(o.z, o.w) * (float)kFFT;
                const float hL = ptx::rna_tf32(mL), hR = ptx::rna_tf32(mR);
                const size_t mi = mag_s2d_index(p.T, p.F, p.cursor, k);
                p.mag_new[mi] = make_float2(hL, hR);
                p.mag_new[mi + p.mag_lo_off] = make_float2(mL - hL, mR - hR);
            }
        }
        return;
    }
    // ---- synthesis of stem `role` from the frame recorded two tiles ago --------------------------
    const int s = role;
    const float2* mrow = reinterpret_cast<const float2*>(p.mask_old) + ((size_t)s * p.T + p.cursor) * p.F;
    const float uw = p.uw[s];
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const int k = j + 256 * r;
        const int kk = k <= kFFT / 2 ? k : kFFT - k;
        const float4 sp = p.spec_old[kk];
        float mL = uw, mR = uw;
        if (kk < p.F) { const float2 m = mrow[kk]; mL = m.x; mR = m.y; }
        const float xlr = sp.x * mL, xli = -(sp.y * mL), xrr = sp.z * mR, xri = -(sp.w * mR);
        float2 z;
        if (k <= kFFT / 2) z = make_float2(xlr - xri, xli + xrr);
        else z = make_float2(xlr + xri, -xli + xrr);
        v[r] = make_float2(z.x, -z.y);
    }
    fft4096(v, sm, p.twiddle, j);
    // keep samples 2048..4095 (SAMPLESHIFT), synthesis window, 50 % overlap-add (:303-320)
#pragma unroll
    for (int r = 8; r < 16; r++) {
        const int q = j + 256 * r - 2048;          // 0..2047
        const float w = __ldg(&p.swin[q]);
        const float l = v[r].x * w, rr = -v[r].y * w;
        if (q < 1024) {
            p.out[(2 * s) * 1024 + q] = p.overlap_in[(2 * s) * 1024 + q] + l;
            p.out[(2 * s + 1) * 1024 + q] = p.overlap_in[(2 * s + 1) * 1024 + q] + rr;
        } else {
            p.overlap_out[(2 * s) * 1024 + q - 1024] = l;
            p.overlap_out[(2 * s + 1) * 1024 + q - 1024] = rr;
        }
    }
}

__global__ void fill_kernel(float* p, float v, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace

struct srt_stream {
    srt_ctx* ctx = nullptr;
    int S = 0, T = 0, F = 0, device = 0;
    cudaStream_t hop_stream = nullptr;
    cudaEvent_t ev_tile[2]{}, ev_nn[2]{};
    float *d_ring = nullptr, *d_awin = nullptr, *d_swin = nullptr, *d_mask = nullptr, *d_overlap = nullptr, *d_out = nullptr;
    float4* d_spec = nullptr;
    float2* d_mag = nullptr;   // = the context's magnitude buffer (2 images), not owned
    float *h_in = nullptr, *h_out = nullptr;   // pinned staging
    float uw[8]{};
    // host state (mirrors mInputPos / mInputSamplesNeeded / nnMaskCursor / output buffers of the reference)
    int in_pos = 0, filled = 0, cursor = 0;
    long long tile = 0, hops = 0, launches = 0;
    // finished hops waiting to be handed out: a fixed ring of pinned blocks [kOutRing][2S][1024] the hop's D2H lands in
    // directly - nothing is allocated on the audio thread (srt_stream_process drains after every hop, so at most two
    // blocks are ever pending)
    static constexpr int kOutRing = 8;
    long long out_head = 0, out_tail = 0;   // blocks produced / blocks fully handed out
    int read_off = 0;
};

#define SCK(call)                                                                              \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) return internal::set_error(SRT_ERR_CUDA, cudaGetErrorString(e_)); \
    } while (0)

extern "C" void srt_stream_destroy(srt_stream* st)
{
    if (!st) return;
    internal::DeviceGuard dev_guard(st->device);
    if (st->hop_stream) cudaStreamSynchronize(st->hop_stream);
    if (st->ctx) srt_synchronize(st->ctx);
    for (void* p : {(void*)st->d_ring, (void*)st->d_awin, (void*)st->d_swin, (void*)st->d_mask, (void*)st->d_overlap, (void*)st->d_out,
                    (void*)st->d_spec})
        if (p) cudaFree(p);
    if (st->h_in) cudaFreeHost(st->h_in);
    if (st->h_out) cudaFreeHost(st->h_out);
    for (int i = 0; i < 2; i++) {
        if (st->ev_tile[i]) cudaEventDestroy(st->ev_tile[i]);
        if (st->ev_nn[i]) cudaEventDestroy(st->ev_nn[i]);
    }
    if (st->hop_stream) cudaStreamDestroy(st->hop_stream);
    if (st->ctx) srt_destroy(st->ctx);
    delete st;
}

extern "C" long long srt_stream_launch_count(const srt_stream* st) { return st ? st->launches + srt_launch_count(st->ctx) : 0; }

extern "C" int srt_stream_create(const srt_config* cfg, const float* const* coeffs, const float* unaffected, srt_stream** out)
{
    if (!cfg || !out || !coeffs) return internal::set_error(SRT_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->n_stems < 1 || cfg->n_stems > SRT_MAX_STEMS) return internal::set_error(SRT_ERR_ARG, "n_stems out of range");
    srt_config c2 = *cfg;
    c2.flavour = 1;            // VST flavour: exact sigmoid, unclamped ELU (VST/Source/spleeter.c:56-77)
    c2.max_images = 1;
    c2.max_batch_images = 2;   // the context's magnitude buffer doubles as the tile double buffer
    c2.cuda_stream = nullptr;  // the nets get their own stream
    int modes[SRT_MAX_STEMS];
    for (int s = 0; s < cfg->n_stems; s++) modes[s] = 1;   // all ELU (Spleeter4Stems.c:444-447)
    srt_stream* st = new srt_stream();
    st->device = cfg->device;
    int r = srt_create(&c2, coeffs, modes, &st->ctx);
    internal::DeviceGuard dev_guard(cfg->device);
    if (r) { delete st; return r; }
    st->S = cfg->n_stems; st->T = cfg->time_step; st->F = cfg->bin_limit;
    const int S = st->S, T = st->T, F = st->F;
    static const float kDefaultUw[4] = {0.25f, 0.0f, 0.25f, 0.25f};
    for (int s = 0; s < S; s++) st->uw[s] = unaffected ? unaffected[s] : (s < 4 ? kDefaultUw[s] : 0.25f);
    auto bail = [&](int code) { srt_stream_destroy(st); return code; };
    if (cudaStreamCreateWithFlags(&st->hop_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(internal::set_error(SRT_ERR_CUDA, "stream"));
    for (int i = 0; i < 2; i++) {
        cudaEventCreateWithFlags(&st->ev_tile[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&st->ev_nn[i], cudaEventDisableTiming);
    }
    const size_t nspec = (size_t)3 * T * kBins, nmag = (size_t)2 * T * F, nmask = (size_t)2 * S * T * F * 2;
    if (cudaMalloc(&st->d_ring, 2 * kFFT * 4) || cudaMalloc(&st->d_awin, kFFT * 4) || cudaMalloc(&st->d_swin, 2048 * 4) ||
        cudaMalloc(&st->d_mask, nmask * 4) || cudaMalloc(&st->d_overlap, (size_t)2 * 2 * S * 1024 * 4) || cudaMalloc(&st->d_out, (size_t)2 * S * 1024 * 4) ||
        cudaMalloc(&st->d_spec, nspec * sizeof(float4)) ||
        cudaMallocHost(&st->h_in, 2 * 1024 * 4) || cudaMallocHost(&st->h_out, (size_t)srt_stream::kOutRing * 2 * S * 1024 * 4))
        return bail(internal::set_error(SRT_ERR_CUDA, "stream buffers: out of memory"));
    cudaMemsetAsync(st->d_ring, 0, 2 * kFFT * 4, st->hop_stream);
    cudaMemsetAsync(st->d_spec, 0, nspec * sizeof(float4), st->hop_stream);
    st->d_mag = reinterpret_cast<float2*>(internal::ctx_mag(st->ctx));
    cudaMemsetAsync(st->d_mag, 0, 2 * nmag * sizeof(float2), st->hop_stream);
    cudaMemsetAsync(st->d_overlap, 0, (size_t)2 * 2 * S * 1024 * 4, st->hop_stream);
    fill_kernel<<<(unsigned)((nmask + 255) / 256), 256, 0, st->hop_stream>>>(st->d_mask, 1.0f, nmask);   // masks start at 1 (Spleeter4Stems.c:455-466)
    // windows: getAsymmetricWindow(analysis, synthesis, k = 4096, m = 1024, 1.0) (Spleeter4Stems.c:383-401, 414-416)
    {
        const int k = kFFT, m = 1024;
        std::vector<float> an(k), sy(k, 0.0f), aw(k), sw(2048);
        int n = ((k - m) << 1) + 2;
        for (int i = 0; i < k - m; ++i) an[i] = (float)pow(0.5 * (1.0 - cos(2.0 * M_PI * (i + 1.0) / (double)n)), 1.0);
        n = (m << 1) + 2;
        for (int i = k - m; i < k; ++i) an[i] = (float)pow(sqrt(0.5 * (1.0 - cos(2.0 * M_PI * ((m + i - (k - m)) + 1.0) / (double)n))), 1.0);
        n = m << 1;
        for (int i = k - (m << 1); i < k; ++i) sy[i] = (float)(0.5 * (1.0 - cos(2.0 * M_PI * (double)(i - (k - (m << 1))) / (double)n))) / an[i];
        for (int i = 0; i < 2048; i++) sw[i] = sy[i + 2048];                       // pre-shift by SAMPLESHIFT
        for (int i = 0; i < k; i++) {
            float a = an[i];
            a *= (1.0 / kFFT) * 0.5f;                                               // reference analysisWnd (:415-416)
            aw[i] = a * 2.0f;                                                       // x2: the packed complex FFT then yields re / -im directly
        }
        cudaMemcpyAsync(st->d_awin, aw.data(), k * 4, cudaMemcpyHostToDevice, st->hop_stream);
        cudaMemcpyAsync(st->d_swin, sw.data(), 2048 * 4, cudaMemcpyHostToDevice, st->hop_stream);
        cudaStreamSynchronize(st->hop_stream);      // the host vectors go out of scope here
    }
    if (cudaStreamSynchronize(st->hop_stream) != cudaSuccess) return bail(internal::set_error(SRT_ERR_CUDA, "stream init failed"));
    *out = st;
    return 0;
}

// one hop: the host has just written 1024 new samples at in_pos (ring position) into h_in
static int do_hop(srt_stream* st)
{
    const int S = st->S, T = st->T, F = st->F;
    const long long k = st->tile;
    SCK(cudaMemcpyAsync(st->d_ring + st->in_pos, st->h_in, 1024 * 4, cudaMemcpyHostToDevice, st->hop_stream));
    SCK(cudaMemcpyAsync(st->d_ring + kFFT + st->in_pos, st->h_in + 1024, 1024 * 4, cudaMemcpyHostToDevice, st->hop_stream));
    st->in_pos = (st->in_pos + 1024) & (kFFT - 1);
    HopParams p{};
    p.ring = st->d_ring; p.pos = st->in_pos;                                  // oldest sample follows the newest
    p.awin = st->d_awin; p.swin = st->d_swin; p.twiddle = internal::ctx_twiddle(st->ctx);
    p.spec_new = st->d_spec + ((size_t)(k % 3) * T + st->cursor) * kBins;
    p.mag_new = st->d_mag + (size_t)(k % 2) * T * F;
    p.mag_lo_off = (size_t)2 * T * F;   // lo images follow the two hi images
    p.spec_old = st->d_spec + ((size_t)((k + 1) % 3) * T + st->cursor) * kBins;   // tile k-2
    p.mask_old = st->d_mask + (size_t)(k % 2) * S * T * F * 2;                    // masks of tile k-2
    p.cursor = st->cursor; p.T = T; p.F = F; p.S = S;
    for (int s = 0; s < S; s++) p.uw[s] = st->uw[s];
    p.overlap_in = st->d_overlap + (size_t)(st->hops & 1) * 2 * S * 1024;
    p.overlap_out = st->d_overlap + (size_t)((st->hops + 1) & 1) * 2 * S * 1024;
    p.out = st->d_out;
    stream_hop_kernel<<<S + 1, kFftThreads, 0, st->hop_stream>>>(p);
    st->launches++;
    if (st->out_head - st->out_tail >= srt_stream::kOutRing) return internal::set_error(SRT_ERR_STATE, "output ring overflow");
    SCK(cudaMemcpyAsync(st->h_out + (size_t)(st->out_head % srt_stream::kOutRing) * 2 * S * 1024, st->d_out, (size_t)2 * S * 1024 * 4,
                        cudaMemcpyDeviceToHost, st->hop_stream));
    st->hops++;
    if (++st->cursor >= T) {
        // tile k is complete (Spleeter4Stems.c:351-371): launch the nets on it, and make the hop stream wait
        // for the masks of tile k-1, which the next tile's synthesis reads.
        st->cursor = 0;
        cudaStream_t nn = internal::ctx_stream(st->ctx);
        SCK(cudaEventRecord(st->ev_tile[k % 2], st->hop_stream));
        SCK(cudaStreamWaitEvent(nn, st->ev_tile[k % 2], 0));
        int r = internal::ctx_run_unet(st->ctx, (int)(k % 2), 1, st->d_mask + (size_t)(k % 2) * S * T * F * 2, 1, 0);
        if (r) return r;
        SCK(cudaEventRecord(st->ev_nn[k % 2], nn));
        if (k >= 1) SCK(cudaStreamWaitEvent(st->hop_stream, st->ev_nn[(k - 1) % 2], 0));
        st->tile++;
    }
    SCK(cudaStreamSynchronize(st->hop_stream));
    st->out_head++;
    return 0;
}

extern "C" int srt_stream_process(srt_stream* st, const float* inL, const float* inR, int n, float* const* components)
{
    if (!st || n < 0) return internal::set_error(SRT_ERR_ARG, "bad argument");
    internal::DeviceGuard dev_guard(st->device);
    if (!dev_guard.ok) return internal::set_error(SRT_ERR_CUDA, "cudaSetDevice failed");
    const int want = n;
    int done = 0;
    // hand out what is ready, at most `want` samples in total; `components` stays untouched beyond that (:538-581)
    auto drain = [&]() {
        while (st->out_tail < st->out_head && done < want) {
            const float* blk = st->h_out + (size_t)(st->out_tail % srt_stream::kOutRing) * 2 * st->S * 1024;
            const int c = std::min(1024 - st->read_off, want - done);
            for (int q = 0; q < 2 * st->S; q++) std::memcpy(components[q] + done, blk + (size_t)q * 1024 + st->read_off, (size_t)c * 4);
            done += c;
            st->read_off += c;
            if (st->read_off == 1024) { st->read_off = 0; st->out_tail++; }
        }
    };
    while (n > 0) {
        const int c = std::min(1024 - st->filled, n);
        std::memcpy(st->h_in + st->filled, inL, (size_t)c * 4);
        std::memcpy(st->h_in + 1024 + st->filled, inR, (size_t)c * 4);
        inL += c; inR += c; n -= c;
        st->filled += c;
        if (st->filled == 1024) {
            st->filled = 0;
            drain();               // keeps the pending blocks <= 2 whatever the call size: no allocation on the audio thread
            int r = do_hop(st);
            if (r) return r;
        }
    }
    drain();
    return 0;
}

// ---- tier A (VST/Source/Spleeter4Stems.h:67-69) ---------------------------------------------------
static void die_stream(const char* where)
{
    fprintf(stderr, "[spleeterrt_b200] %s failed: %s\n", where, srt_last_error());
    abort();
}

extern "C" void Spleeter4StemsInit(Spleeter4Stems* msr, int initSpectralBinLimit, int initTimeStep, void* coeffProvider[4])
{
    memset(msr, 0, sizeof(*msr));
    srt_config cfg;
    memset(&cfg, 0, sizeof cfg);
    const char* dev = getenv("SRT_DEVICE");
    cfg.device = dev ? atoi(dev) : 0;
    cfg.n_stems = 4;
    cfg.time_step = initTimeStep;
    cfg.bin_limit = initSpectralBinLimit;
    const float* cp[4] = {(const float*)coeffProvider[0], (const float*)coeffProvider[1], (const float*)coeffProvider[2], (const float*)coeffProvider[3]};
    srt_stream* st = nullptr;
    if (srt_stream_create(&cfg, cp, nullptr, &st)) die_stream("Spleeter4StemsInit");
    msr->impl = st;
    msr->analyseBinLimit = initSpectralBinLimit;
    msr->timeStep = initTimeStep;
}

extern "C" void Spleeter4StemsFree(Spleeter4Stems* msr)
{
    srt_stream_destroy((srt_stream*)msr->impl);
    msr->impl = nullptr;
}

extern "C" void Spleeter4StemsProcessSamples(Spleeter4Stems* msr, const float* inLeft, const float* inRight, int inSampleCount, float** components)
{
    if (srt_stream_process((srt_stream*)msr->impl, inLeft, inRight, inSampleCount, components)) die_stream("Spleeter4StemsProcessSamples");
}
