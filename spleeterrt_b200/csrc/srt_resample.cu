// srt_resample.cu — sample-rate conversion in front of the path.
//
// The reference CLI brings every decoded file to 44.1 kHz with JamesDSPOfflineResampling (Executable/main.c:209-224,
// 264-270) = libsamplerate's one-shot src_simple() on the windowed-sinc interpolator of libsamplerate/src_sinc.c, fed
// with a coefficient table the HOST owns (`decompressedCoefficients`, main.c:277, 693-694; geometry src_sinc.c:141-143:
// 22438 floats, half length 22436, 491 table steps per input sample).  Like the network weights the table is caller
// data here: it is passed in, not embedded.
//
// The converter is a sequential state machine only in its bookkeeping: output frame k reads input frames around
// b_k with sub-sample phase frac_k, where (b, frac) follow  frac += 1/ratio; b += round-down carry  in double
// (src_sinc.c:347-352 / 495-500), and the stream ends where the reference's buffer logic says so (prepare_data,
// src_sinc.c:1102-1172; termination tests :323-326 mono `>`, :479-482 stereo `>=`).  That scalar recurrence runs on the
// host (a few ns per frame) and yields, per output frame, the input frame index and the fixed-point start offset into
// the table.  The arithmetic - ~92 interpolated taps per frame and channel, accumulated in double in the reference's
// order (calc_output_single :218-271, calc_output_stereo :366-420) - is one GPU thread per output sample, written
// with explicit round-to-nearest operations (no FMA contraction) so that the result is bit-identical to the reference.
#include <cmath>
#include <cstdio>
#include <vector>

#include "srt_internal.h"
#include "srt_kernels.cuh"

namespace srt {

struct ResamplePlanEntry {
    int32_t frame;      // input frame under the centre tap (b_k)
    int32_t start;      // start_filter_index: lrint(frac * index_inc * min(ratio, 1) * 4096), 12-bit fixed point
};

struct ResampleParams {
    const float* in;                    // n_in frames of ch interleaved floats
    const float* coeffs;                // half_len + 2 floats
    const ResamplePlanEntry* plan;      // n_gen entries
    float* out;                         // n_gen frames of ch interleaved floats
    long n_in_floats;
    int n_gen, ch;
    int32_t inc, max_idx;               // table step per input sample, half_len << 12
    double scale;                       // min(ratio, 1), as the reference forms it (float_increment / index_inc)
};

__device__ __forceinline__ double rs_tap(const float* __restrict__ coeffs, int32_t fidx)
{
    const double fraction = (double)(fidx & 4095) * (1.0 / 4096.0);
    const int k = fidx >> 12;
    const float c0 = __ldg(coeffs + k), c1 = __ldg(coeffs + k + 1);
    return __dadd_rn((double)c0, __dmul_rn(fraction, (double)__fsub_rn(c1, c0)));
}

__global__ void __launch_bounds__(256) resample_kernel(const ResampleParams p)
{
    const long t = (long)blockIdx.x * 256 + threadIdx.x;
    if (t >= (long)p.n_gen * p.ch) return;
    const int k = (int)(t / p.ch), c = (int)(t % p.ch);
    const ResamplePlanEntry e = p.plan[k];
    const long centre = (long)e.frame * p.ch + c;
    // left wing, far tap first (table offsets start + j * inc, samples b - j)
    int32_t fidx = e.start;
    int32_t n = (p.max_idx - fidx) / p.inc;
    fidx += n * p.inc;
    long di = centre - (long)p.ch * n;
    double left = 0.0;
    do {
        const double x = (di >= 0 && di < p.n_in_floats) ? (double)__ldg(p.in + di) : 0.0;
        left = __dadd_rn(left, __dmul_rn(rs_tap(p.coeffs, fidx), x));
        fidx -= p.inc;
        di += p.ch;
    } while (fidx >= 0);
    // right wing (table offsets inc - start + j * inc, samples b + 1 + j)
    fidx = p.inc - e.start;
    n = (p.max_idx - fidx) / p.inc;
    fidx += n * p.inc;
    di = centre + (long)p.ch * (1 + n);
    double right = 0.0;
    do {
        const double x = (di >= 0 && di < p.n_in_floats) ? (double)__ldg(p.in + di) : 0.0;
        right = __dadd_rn(right, __dmul_rn(rs_tap(p.coeffs, fidx), x));
        fidx -= p.inc;
        di -= p.ch;
    } while (fidx > 0);
    p.out[t] = __double2float_rn(__dmul_rn(p.scale, __dadd_rn(left, right)));
}

static double fmod_one(double v)   // common.h:137-145
{
    const double r = v - (double)lrint(v);
    return r < 0.0 ? r + 1.0 : r;
}

// The converter's bookkeeping for a one-shot conversion of n_in frames into at most n_out frames.
static void build_plan(long n_in, int ch, double ratio, int half_len, int index_inc, long n_out, std::vector<ResamplePlanEntry>& plan,
                       int32_t* inc_out, double* scale_out)
{
    long b_len = 3 * lrint((half_len + 2.0) / index_inc * 256.0 + 1);     // sinc_set_converter, SRC_MAX_RATIO = 256
    if (b_len < 4096) b_len = 4096;
    b_len = b_len * ch + 1;
    double count = (half_len + 2.0) / index_inc;
    if (ratio < 1.0) count /= ratio;
    const long half = ch * (lrint(count) + 1);
    const long in_count = n_in * ch;
    long b_cur = 0, b_end = 0, b_real_end = -1, in_used = 0, frame = 0;
    double frac = 0.0;
    const double terminate = 1.0 / ratio + 1e-20;
    const double step_in = 1.0 / ratio;
    const double float_inc = index_inc * (ratio < 1.0 ? ratio : 1.0);
    *inc_out = (int32_t)lrint(float_inc * 4096.0);
    *scale_out = float_inc / index_inc;
    plan.clear();
    plan.reserve((size_t)n_out);
    while ((long)plan.size() < n_out) {
        long in_hand = (b_end - b_cur + b_len) % b_len;
        if (in_hand <= half) {
            if (b_real_end < 0) {   // refill: only the indices of prepare_data matter here
                long len;
                if (b_cur == 0) {
                    len = b_len - 2 * half;
                    b_cur = b_end = half;
                } else if (b_end + half + ch < b_len) {
                    len = std::max(b_len - b_cur - half, 0L);
                } else {
                    len = b_end - b_cur;
                    b_cur = half;
                    b_end = b_cur + len;
                    len = std::max(b_len - b_cur - half, 0L);
                }
                len = std::min(in_count - in_used, len);
                len -= len % ch;
                b_end += len;
                in_used += len;
                if (in_used == in_count && b_end - b_cur < 2 * half) {   // src_simple sets end_of_input
                    if (b_len - b_end < half + 5) {
                        len = b_end - b_cur;
                        b_cur = half;
                        b_end = b_cur + len;
                    }
                    b_real_end = b_end;
                    len = half + 5;
                    if (b_end + len > b_len) len = b_len - b_end;
                    b_end += len;
                }
            }
            in_hand = (b_end - b_cur + b_len) % b_len;
            if (in_hand <= half) break;
        }
        if (b_real_end >= 0) {
            const double pos = (double)b_cur + frac + terminate;
            if (ch == 1 ? pos > (double)b_real_end : pos >= (double)b_real_end) break;
        }
        plan.push_back(ResamplePlanEntry{(int32_t)frame, (int32_t)lrint(frac * float_inc * 4096.0)});
        frac += step_in;
        const double rem = fmod_one(frac);
        const long carry = lrint(frac - rem);
        b_cur = (b_cur + ch * carry) % b_len;
        frame += carry;
        frac = rem;
    }
}

}  // namespace srt

using namespace srt;

static int rs_fail(int code, const char* msg) { return internal::set_error(code, msg); }

extern "C" size_t srt_resample_frames(size_t n_in, double ratio)
{
    return (size_t)(long)ceil((double)n_in * ratio);   // main.c:265
}

// Host-only view of the bookkeeping: which input frame and which table offset every output frame uses, and how many
// frames the converter produces.  frames / starts may be NULL.
extern "C" long long srt_resample_plan(size_t n_in, int channels, double ratio, int coeff_count, int index_inc, size_t n_out,
                                       int32_t* frames, int32_t* starts)
{
    if (channels != 1 && channels != 2) return rs_fail(SRT_ERR_ARG, "srt_resample: 1 or 2 channels (main.c:764-769)");
    if (!(ratio >= 1.0 / 256.0 && ratio <= 256.0)) return rs_fail(SRT_ERR_ARG, "srt_resample: ratio outside [1/256, 256] (samplerate.c:144)");
    if (coeff_count < 3 || index_inc < 1 || coeff_count > (1 << 18)) return rs_fail(SRT_ERR_ARG, "srt_resample: bad coefficient table geometry");
    if (n_in == 0 || n_in > ((size_t)1 << 30) || n_out > ((size_t)1 << 30)) return rs_fail(SRT_ERR_ARG, "srt_resample: bad length");
    std::vector<ResamplePlanEntry> plan;
    int32_t inc = 0;
    double scale = 0.0;
    build_plan((long)n_in, channels, ratio, coeff_count - 2, index_inc, (long)n_out, plan, &inc, &scale);
    for (size_t i = 0; i < plan.size(); i++) {
        if (frames) frames[i] = plan[i].frame;
        if (starts) starts[i] = plan[i].start;
    }
    return (long long)plan.size();
}

// d_in / d_coeffs / d_out are device pointers on ctx's device; the call returns when the result is complete.
extern "C" int srt_resample_device(srt_ctx* ctx, const float* d_in, size_t n_in, int channels, double ratio, const float* d_coeffs,
                                   int coeff_count, int index_inc, float* d_out, size_t n_out, size_t* n_generated)
{
    if (!ctx) return rs_fail(SRT_ERR_STATE, "srt_resample: null context");
    if (!d_in || !d_coeffs || !d_out) return rs_fail(SRT_ERR_ARG, "srt_resample: null buffer");
    if (channels != 1 && channels != 2) return rs_fail(SRT_ERR_ARG, "srt_resample: 1 or 2 channels (main.c:764-769)");
    if (!(ratio >= 1.0 / 256.0 && ratio <= 256.0)) return rs_fail(SRT_ERR_ARG, "srt_resample: ratio outside [1/256, 256] (samplerate.c:144)");
    if (coeff_count < 3 || index_inc < 1 || coeff_count > (1 << 18)) return rs_fail(SRT_ERR_ARG, "srt_resample: bad coefficient table geometry");
    if (n_in == 0 || n_in > ((size_t)1 << 30) || n_out > ((size_t)1 << 30)) return rs_fail(SRT_ERR_ARG, "srt_resample: bad length");
    internal::DeviceGuard dev_guard(internal::ctx_device(ctx));
    if (!dev_guard.ok) return rs_fail(SRT_ERR_CUDA, "srt_resample: cudaSetDevice failed");
    const int half_len = coeff_count - 2;               // src_sinc.c:142
    std::vector<ResamplePlanEntry> plan;
    int32_t inc = 0;
    double scale = 0.0;
    build_plan((long)n_in, channels, ratio, half_len, index_inc, (long)n_out, plan, &inc, &scale);
    if (n_generated) *n_generated = plan.size();
    if (plan.empty()) return 0;
    if (inc < 1) return rs_fail(SRT_ERR_ARG, "srt_resample: table step rounds to zero");
    cudaStream_t st = internal::ctx_stream(ctx);
    ResamplePlanEntry* d_plan = nullptr;
    if (cudaMalloc((void**)&d_plan, plan.size() * sizeof(ResamplePlanEntry)) != cudaSuccess) return rs_fail(SRT_ERR_CUDA, "srt_resample: cudaMalloc failed");
    cudaError_t e = cudaMemcpyAsync(d_plan, plan.data(), plan.size() * sizeof(ResamplePlanEntry), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        ResampleParams p{};
        p.in = d_in; p.coeffs = d_coeffs; p.plan = d_plan; p.out = d_out;
        p.n_in_floats = (long)n_in * channels;
        p.n_gen = (int)plan.size(); p.ch = channels;
        p.inc = inc; p.max_idx = (int32_t)half_len << 12;
        p.scale = scale;
        const long threads = (long)p.n_gen * channels;
        resample_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(p);
        internal::ctx_count_launch(ctx, 1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_plan);
    if (e != cudaSuccess) return rs_fail(SRT_ERR_CUDA, cudaGetErrorString(e));
    return 0;
}

// Host buffers: `out` holds n_out frames and is zero-filled beyond the generated ones, like the reference's tmpBuf
// (main.c:266-268).
extern "C" int srt_resample_host(srt_ctx* ctx, const float* in, size_t n_in, int channels, double ratio, const float* coeffs,
                                 int coeff_count, int index_inc, float* out, size_t n_out, size_t* n_generated)
{
    if (!ctx) return rs_fail(SRT_ERR_STATE, "srt_resample: null context");
    if (!in || !coeffs || !out) return rs_fail(SRT_ERR_ARG, "srt_resample: null buffer");
    if (channels != 1 && channels != 2) return rs_fail(SRT_ERR_ARG, "srt_resample: 1 or 2 channels (main.c:764-769)");
    if (coeff_count < 3 || coeff_count > (1 << 18)) return rs_fail(SRT_ERR_ARG, "srt_resample: bad coefficient table geometry");
    if (n_in == 0 || n_in > ((size_t)1 << 30) || n_out > ((size_t)1 << 30)) return rs_fail(SRT_ERR_ARG, "srt_resample: bad length");
    internal::DeviceGuard dev_guard(internal::ctx_device(ctx));
    if (!dev_guard.ok) return rs_fail(SRT_ERR_CUDA, "srt_resample: cudaSetDevice failed");
    float *d_in = nullptr, *d_c = nullptr, *d_out = nullptr;
    const size_t in_b = n_in * channels * sizeof(float), out_b = (n_out ? n_out : 1) * channels * sizeof(float);
    int rc = 0;
    size_t gen = 0;
    cudaStream_t st = internal::ctx_stream(ctx);
    if (cudaMalloc((void**)&d_in, in_b) != cudaSuccess || cudaMalloc((void**)&d_c, (size_t)coeff_count * 4) != cudaSuccess ||
        cudaMalloc((void**)&d_out, out_b) != cudaSuccess) {
        rc = rs_fail(SRT_ERR_CUDA, "srt_resample: cudaMalloc failed");
    } else if (cudaMemcpyAsync(d_in, in, in_b, cudaMemcpyHostToDevice, st) != cudaSuccess ||
               cudaMemcpyAsync(d_c, coeffs, (size_t)coeff_count * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
               cudaMemsetAsync(d_out, 0, out_b, st) != cudaSuccess) {
        rc = rs_fail(SRT_ERR_CUDA, "srt_resample: upload failed");
    } else {
        rc = srt_resample_device(ctx, d_in, n_in, channels, ratio, d_c, coeff_count, index_inc, d_out, n_out, &gen);
        if (rc == 0 && n_out && cudaMemcpy(out, d_out, n_out * channels * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess)
            rc = rs_fail(SRT_ERR_CUDA, "srt_resample: download failed");
    }
    cudaFree(d_in); cudaFree(d_c); cudaFree(d_out);
    if (n_generated) *n_generated = gen;
    return rc;
}
