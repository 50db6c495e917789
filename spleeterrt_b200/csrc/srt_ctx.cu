// srt_ctx.cu — context, memory plan and orchestration behind the tier-B C ABI (include/srt_b200.h).
// One context per GPU and per (n_stems, T, F) configuration; all work is enqueued on one CUDA
// stream.  No CPU compute path exists: if the device or the kernels are unavailable the calls
// return SRT_ERR_CUDA.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <map>
#include <memory>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/srt_b200.h"
#include "srt_internal.h"
#include "srt_kernels.cuh"
#include "srt_plan.h"
#include "srt_sigmoid_table.h"

using namespace srt;

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return fail(SRT_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)

extern "C" const char* srt_last_error(void) { return g_err.c_str(); }

// ------------------------------------------------------------------------------------------
// driver entry point for TMA descriptors (no link-time dependency on libcuda)
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// activation tensor [n][H][W][C] fp32 -> 4-D map, box {32, tw, th, nb}, 128B swizzle, zero OOB fill
static int make_tmap(CUtensorMap* m, const float* base, int C, int W, int H, int N, int tw, int th, int nb, int kbw = kKB)
{
    // box = {kbw channels (32 -> SWIZZLE_128B, 16 -> SWIZZLE_64B, 8 -> SWIZZLE_32B), tw pixels, th rows, nb images}; may overhang the tensor (zero fill)
    PFN_encodeTiled enc = get_encode();
    if (!enc) return fail(SRT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)kbw, (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)nb};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     kbw == kKB ? CU_TENSOR_MAP_SWIZZLE_128B : kbw == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SRT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) C=%d W=%d H=%d N=%d box=%d,%d,%d", (int)r, C, W, H, N, tw, th, nb);
    return 0;
}

// bf16 residual tensor [n][H][W][C] of a compensated layer -> 4-D map, box {64, tw, th, nb} = the same 128-byte swizzled rows
static int make_tmap_lo(CUtensorMap* m, const void* base, int C, int W, int H, int N, int tw, int th, int nb, int lo_fmt)
{
    const bool fp8 = lo_fmt == LO_FP8 || lo_fmt == LO_FP8N, narrow = lo_fmt == LO_FP8N;
    const int esize = fp8 ? 1 : 2;            // e5m2 bytes, 128 (narrow: 64, in 64-byte rows) channels per box row - or bf16, 64
    PFN_encodeTiled enc = get_encode();
    if (!enc) return fail(SRT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * esize, (cuuint64_t)W * C * esize, (cuuint64_t)H * W * C * esize};
    cuuint32_t box[4] = {(cuuint32_t)(narrow ? 64 : fp8 ? kKBlo8 : kKBlo), (cuuint32_t)tw, (cuuint32_t)th, (cuuint32_t)nb};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, fp8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     narrow ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SRT_ERR_CUDA, "cuTensorMapEncodeTiled (bf16 residual) failed (%d) C=%d W=%d H=%d N=%d box=%d,%d,%d", (int)r, C, W, H, N, tw, th, nb);
    return 0;
}

// ------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------
static const int kEnc[7] = {2, 16, 32, 64, 128, 256, 512};
static const int kDecOut[6] = {256, 128, 64, 32, 16, 1};

struct Span {
    int cat;
    cudaEvent_t a, b;
};

constexpr int kMetaSlots = 16;    // per-call metadata ring
constexpr int kTicketMod = 3 << 28;   // tickets wrap here (multiple of kBatchSlots)
constexpr int kBatchSlots = 3;    // host-pointer batches in flight (upload of k+1 | kernels of k | download of k-1)

// srt_config.share_weights: the constant device data of a context (packed weights, k-block tables, epilogue vectors,
// transform tables - everything build() uploads) keyed by what it is a function of.  Contexts with an equal key take the
// uploads of the first one, in build()'s call order.
struct SharedUploads {
    std::vector<void*> ptrs;
    int device = 0;
    ~SharedUploads()
    {
        int prev = -1;
        cudaGetDevice(&prev);
        cudaSetDevice(device);
        for (void* p : ptrs) cudaFree(p);
        if (prev >= 0) cudaSetDevice(prev);
    }
};
struct SharedKey {
    std::vector<const void*> coeffs;
    std::vector<int> ints;          // device, S, T, F, B, NB, flavour, conv_impl, precision, modes...
    std::vector<float> fingerprint; // sampled weights: a pointer that was freed and reused for other weights must not hit
    bool operator<(const SharedKey& o) const
    {
        if (coeffs != o.coeffs) return coeffs < o.coeffs;
        if (ints != o.ints) return ints < o.ints;
        return fingerprint < o.fingerprint;
    }
};
static std::mutex g_shared_mu;
static std::map<SharedKey, std::weak_ptr<SharedUploads>> g_shared;

struct srt_ctx {
    srt_config cfg{};
    std::shared_ptr<SharedUploads> shared;   // non-null: uploads go through / come from the shared set
    bool shared_hit = false;                 // the set already existed: skip packing, take pointers in order
    size_t shared_next = 0;
    int S = 0, T = 0, F = 0, B = 0, NB = 0;   // B = U-Net batch capacity, NB = batch images capacity
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int act_enc[8]{}, act_dec[8]{};
    std::vector<void*> allocs;
    // tables
    float *d_window = nullptr, *d_postwin = nullptr, *d_lut = nullptr;
    float2* d_twiddle = nullptr;
    // weights
    float *d_w1 = nullptr, *d_b1 = nullptr, *d_s1 = nullptr, *d_o1 = nullptr;          // down1
    float *d_w6 = nullptr, *d_b6 = nullptr, *d_s6 = nullptr, *d_o6 = nullptr;          // up6
    float *d_w7 = nullptr, *d_b7 = nullptr;                                            // up7
    std::vector<float> h_w1, h_b1;  // down1 weights [stem][tap][cin][cout] and {bias, scale, offset}
    std::vector<float> h_w6, h_w7;  // up6 / up7 weights, host copies (ride in the kernel parameter bank)
    Up6TcParams up6tc{};            // tensor-core up6 (srt_up6_tc.cu)
    bool use_up6tc = false;
    int sm_count = 148;
    std::vector<LayerPlan> plans;
    std::vector<ConvParams> conv;   // 10 tensor-core layers
    RowConvParams rp[10];           // row-patch form of the small-N layers (down2, down3, up4, up5)
    bool use_rp[10]{};
    RowConvParams d1[4];            // down1 on the tensor cores: groups of up to 4 stems fused into N
    int n_d1 = 0;
    bool split_weights = false;     // two-term weights (not TF32-exact)
    // activations
    float* E[7]{};    // E[1..6] raw skips (NHWC)
    float* A[6]{};    // A[1..5] activated, space-to-depth
    float* U[7]{};    // U[1..5] decoder outputs (NHWC), U[6] = up6 output [n][T][F]
    // compensated precision (srt_config.precision 0): bf16 residuals v - tf32(v) of the tensors above, per consumer layer
    unsigned comp_mask = 0;        // bit i: tensor-core layer i (0..4 = down2..down6, 5..9 = up1..up5) contracts the residuals too
    int lo_want = LO_BF16;         // residual format asked for; layer i uses layer_lo_format(i, lo_want)
    uint8_t* Alo[6]{};             // Alo[i]: residual of A[i] (same S2D layout), read by down{i+1}: bf16 or e5m2 bytes (the consumer's format)
    uint8_t* Clo[5]{};             // Clo[d], d = 1..4: [skip E{6-d} | up U{d}] residuals per pixel, read by up{d+1}; Clo[0]: E6's, read by up1
    bool will_rp[10]{};            // the layer runs in the row-patch kernel (decided before anything is allocated: it selects the format)
    bool lo_narrow_off = false;    // SRT_LO_NARROW=0: down2 / up5 keep bf16 residuals (A/B timing)
    int lo_fmt_of(int layer) const { return ((comp_mask >> layer) & 1u) ? layer_lo_format(layer, lo_want, will_rp[layer] && !lo_narrow_off) : LO_NONE; }
    bool lo_is_fp8(int layer) const { const int f = lo_fmt_of(layer); return f == LO_FP8 || f == LO_FP8N; }
    // batch buffers
    float* d_mag = nullptr;       // [NB][T/2][F/2][(py,px)][c] space-to-depth, TF32-rounded
    float4* d_spec = nullptr;     // [NB][T][2049]
    float* d_mask = nullptr;      // [S][NB][T][F][2]
    float2* d_frames = nullptr;   // [max(S,1)][B][T][4096]
    // per-call metadata (device + pinned host mirror)
    // a ring of kMetaSlots slots so back-to-back (and in-flight asynchronous) calls never wait on the host
    uint8_t *d_meta_base = nullptr, *h_meta_base = nullptr, *h_meta_dev = nullptr;   // h_meta_dev: device view of the pinned mirror
    uint8_t *d_meta = nullptr, *h_meta = nullptr;   // current slot
    size_t meta_cap = 0;                            // bytes per slot
    cudaEvent_t meta_ev[kMetaSlots]{};
    int meta_slot = 0;
    // staging for the host-pointer transform helpers (srt_stft_host / srt_istft_host)
    float *d_pcm = nullptr, *d_out = nullptr;
    size_t pcm_cap = 0, out_cap = 0;
    // staging for srt_separate_batch[_async]: three slots, so batch k+1 uploads while batch k computes and
    // batch k-1 is still being copied back
    float *d_bpcm[kBatchSlots]{}, *d_bout[kBatchSlots]{};
    size_t bpcm_cap[kBatchSlots]{}, bout_cap[kBatchSlots]{};
    cudaEvent_t ev_cdone[kBatchSlots]{}, ev_d2h[kBatchSlots]{};   // last compute / last D2H that used the slot
    bool slot_busy[kBatchSlots]{};
    int slot_ticket[kBatchSlots]{};                               // ticket of the batch that owns the slot
    long long batch_seq = 0;
    cudaStream_t s_in = nullptr, s_out = nullptr;   // copy streams of the host-pointer API (H2D / D2H overlap compute)
    cudaEvent_t ev_in[8]{}, ev_c[8]{};
    // CUDA graphs of the U-Net pass, one per distinct (first image, images, mask destination): the 15-19 launches of a pass
    // replay as one graph launch (their parameter blocks are baked into the nodes, and a pass with the same key has the same
    // parameters).  Matters for small batches: one 10 s stream is ~0.6 ms of kernels and the launch gaps were visible.
    struct UnetGraph {
        int mag_img0, Bv, mask_stride, mask_img0;
        float* mask_base;
        cudaGraphExec_t exec;
        int kernels;
    };
    std::vector<UnetGraph> graphs;
    bool use_graphs = true;
    // bookkeeping
    long long launches = 0;
    bool timing = false;
    std::vector<Span> spans;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    int last_Bv = 0;
    // CLI output modes (srt_create_cli): 0 = one output pair per net; 2 = [vocal, input - vocal] (main.c:776-844);
    // 3 = [drum, vocal, accompaniment] cascade (main.c:845-970) whose second net lives in `next`
    int cli_mode = 0;
    srt_ctx* next = nullptr;      // owned; enqueues on this context's stream
    int pairs() const { return cli_mode ? cli_mode : S; }
};

template <class Tp>
static int dalloc(srt_ctx* c, Tp** p, size_t count)
{
    void* q = nullptr;
    if (count == 0) count = 1;
    CK(cudaMalloc(&q, count * sizeof(Tp)));
    c->allocs.push_back(q);
    *p = (Tp*)q;
    return 0;
}
template <class Tp>
static int upload(srt_ctx* c, Tp** p, const std::vector<Tp>& h)
{
    if (c->shared && c->shared_hit) {          // same key, same call order: the data is already on the device
        if (c->shared_next >= c->shared->ptrs.size()) return fail(SRT_ERR_STATE, "shared weight set is shorter than this context's uploads");
        *p = (Tp*)c->shared->ptrs[c->shared_next++];
        return 0;
    }
    if (c->shared) {
        void* q = nullptr;
        CK(cudaMalloc(&q, std::max<size_t>(h.size(), 1) * sizeof(Tp)));
        c->shared->ptrs.push_back(q);          // owned by the set, freed with its last context
        *p = (Tp*)q;
    } else {
        int r = dalloc(c, p, h.size());
        if (r) return r;
    }
    CK(cudaMemcpy(*p, h.data(), h.size() * sizeof(Tp), cudaMemcpyHostToDevice));
    return 0;
}

static cudaEvent_t get_event(srt_ctx* c)
{
    if (c->ev_used == c->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        c->ev_pool.push_back(e);
    }
    return c->ev_pool[c->ev_used++];
}
struct Timed {
    srt_ctx* c;
    Span sp;
    bool on;
    Timed(srt_ctx* c_, int cat) : c(c_), on(c_->timing)
    {
        if (on) {
            sp.cat = cat;
            sp.a = get_event(c);
            sp.b = get_event(c);
            cudaEventRecord(sp.a, c->stream);
        }
    }
    ~Timed()
    {
        if (on) {
            cudaEventRecord(sp.b, c->stream);
            c->spans.push_back(sp);
        }
    }
};

extern "C" void srt_half_to_float(const uint16_t* in, float* out, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        const uint32_t h = in[i];
        uint32_t bits = ((h & 0x7c00u) == 0) ? 0u : (((h & 0x7fffu) << 13) + 0x38000000u);
        bits |= (h & 0x8000u) << 16;
        std::memcpy(&out[i], &bits, 4);
    }
}

static size_t act_floats(const srt_ctx* c, int level, int ch)   // tensor at resolution T>>level
{
    return (size_t)c->S * c->B * (c->T >> level) * (c->F >> level) * ch;
}

extern "C" void srt_destroy(srt_ctx* c)
{
    if (!c) return;
    internal::DeviceGuard dev_guard(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->next) { srt_destroy(c->next); c->next = nullptr; }
    for (auto& g : c->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    for (void* p : c->allocs) cudaFree(p);
    if (c->d_meta_base) cudaFree(c->d_meta_base);
    if (c->h_meta_base) cudaFreeHost(c->h_meta_base);
    for (int i = 0; i < kMetaSlots; i++)
        if (c->meta_ev[i]) cudaEventDestroy(c->meta_ev[i]);
    if (c->s_in) cudaStreamSynchronize(c->s_in);
    if (c->s_out) cudaStreamSynchronize(c->s_out);
    if (c->d_pcm) cudaFree(c->d_pcm);
    if (c->d_out) cudaFree(c->d_out);
    for (int i = 0; i < kBatchSlots; i++) {
        if (c->d_bpcm[i]) cudaFree(c->d_bpcm[i]);
        if (c->d_bout[i]) cudaFree(c->d_bout[i]);
        if (c->ev_cdone[i]) cudaEventDestroy(c->ev_cdone[i]);
        if (c->ev_d2h[i]) cudaEventDestroy(c->ev_d2h[i]);
    }
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    if (c->s_in) cudaStreamDestroy(c->s_in);
    if (c->s_out) cudaStreamDestroy(c->s_out);
    for (int i = 0; i < 8; i++) {
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_c[i]) cudaEventDestroy(c->ev_c[i]);
    }
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

static int build(srt_ctx* c, const float* const* coeffs, const int* modes)
{
    const int S = c->S, T = c->T, F = c->F;
    // ---- transform tables (InitSTFT, stftFix.c:302-312), computed in double like the reference
    {
        std::vector<float> win(kFFT), post(kFFT);
        std::vector<float2> tw(kFFT);
        const double w = 6.283185307179586476925286766559 / kFFT;
        for (int i = 0; i < kFFT; i++) {
            const float rc = (float)((1.0 / kFFT) * (0.5 * (1.0 - cos(w * (i + 0.5)))));   // LLraisedCosTblFloat
            win[i] = rc;                                                                     // = 2 * mPreWindow
            post[i] = rc * ((float)kFFT * ((1.0f / 2.0f) / (3.0f / 8.0f))) * 0.5f;           // mPostWindow
            tw[i] = make_float2((float)cos(w * i), (float)-sin(w * i));
        }
        int r;
        if ((r = upload(c, &c->d_window, win))) return r;
        if ((r = upload(c, &c->d_postwin, post))) return r;
        if ((r = upload(c, &c->d_twiddle, tw))) return r;
        if (c->cfg.flavour == 0) {
            // fastSigmoid's table (reference data, spleeter.c:29, carried as bit patterns in srt_sigmoid_table.h) and its
            // per-interval slope / origin, evaluated with the reference's float expression (spleeter.c:38-40)
            std::vector<float> t(1026);
            std::memcpy(t.data(), kSigmoidTableBits, 1026 * sizeof(float));
            std::vector<float> lut(1025 * 4);
            const volatile float step = 0.01367188f;
            for (int i = 0; i < 1025; i++) {
                volatile float m1 = step * (float)i, m2 = step * (float)(i + 1);
                volatile float x1 = -7.0f + m1, x2 = -7.0f + m2;
                volatile float den = x2 - x1, num = t[i + 1] - t[i];
                volatile float slope = num / den;
                lut[4 * i + 0] = t[i]; lut[4 * i + 1] = slope; lut[4 * i + 2] = x1; lut[4 * i + 3] = 0.0f;
            }
            if ((r = upload(c, &c->d_lut, lut))) return r;
        }
    }
    int r;
    // ---- batch buffers
    if ((r = dalloc(c, &c->d_spec, (size_t)c->NB * T * kBins))) return r;
    if ((r = dalloc(c, &c->d_frames, (size_t)(S ? S : 1) * c->B * T * kFFT))) return r;
    if (S == 0) return 0;
    if ((r = dalloc(c, &c->d_mag, (size_t)2 * c->NB * T * F * 2))) return r;   // hi images, then lo images
    if ((r = dalloc(c, &c->d_mask, (size_t)S * c->NB * T * F * 2))) return r;
    // ---- activations
    for (int i = 1; i <= 6; i++)
        if ((r = dalloc(c, &c->E[i], act_floats(c, i, kEnc[i])))) return r;
    for (int i = 1; i <= 5; i++)
        if ((r = dalloc(c, &c->A[i], act_floats(c, i, kEnc[i])))) return r;
    for (int d = 1; d <= 5; d++)
        if ((r = dalloc(c, &c->U[d], act_floats(c, 6 - d, kDecOut[d - 1])))) return r;
    if ((r = dalloc(c, &c->U[6], act_floats(c, 0, 1)))) return r;
    // ---- precision: which tensor-core layers also contract the bf16 residuals of their inputs
    int prec = c->cfg.precision;
    if (const char* pe = getenv("SRT_PRECISION")) prec = atoi(pe);
    if (prec != SRT_PRECISION_COMPENSATED && prec != SRT_PRECISION_TF32 && prec != SRT_PRECISION_COMPENSATED_BF16)
        return fail(SRT_ERR_ARG, "precision %d: 0 (compensated), 1 (TF32) or 2 (compensated, bf16 residuals)", prec);
    c->comp_mask = prec == SRT_PRECISION_TF32 ? 0u : 0x3ffu;
    c->lo_want = prec == SRT_PRECISION_COMPENSATED ? LO_FP8 : LO_BF16;
    if (const char* me = getenv("SRT_COMP_MASK")) c->comp_mask = (unsigned)strtoul(me, nullptr, 0) & 0x3ffu;   // experiments: per-layer selection
    // which layers run in the row-patch kernel (SRT_CONV_RP: "0" = never, "1" = whenever supported, unset = when the tile row is wide enough to pay)
    {
        const char* rpe = getenv("SRT_CONV_RP");
        c->lo_narrow_off = getenv("SRT_LO_NARROW") && atoi(getenv("SRT_LO_NARROW")) == 0;
        for (int li = 0; li < 10; li++) {
            if (!row_plan_supported(li) || c->cfg.conv_impl == 1) continue;
            const RowPlan rpl = build_row_plan(NetGeom{T, F}, li, false, false);
            c->will_rp[li] = (rpe ? atoi(rpe) != 0 : rpl.Ws >= 96);
        }
    }
    for (int i = 1; i <= 5; i++)
        if (c->lo_fmt_of(i - 1) != LO_NONE) {
            const size_t bytes = act_floats(c, i, kEnc[i]) * (c->lo_is_fp8(i - 1) ? 1 : 2);
            if ((r = dalloc(c, &c->Alo[i], bytes))) return r;
            CK(cudaMemset(c->Alo[i], 0, bytes));
        }
    for (int d = 0; d <= 4; d++)
        if (c->lo_fmt_of(5 + d) != LO_NONE) {
            const size_t bytes = act_floats(c, 6 - d, d == 0 ? 512 : 2 * kEnc[6 - d]) * (c->lo_is_fp8(5 + d) ? 1 : 2);
            if ((r = dalloc(c, &c->Clo[d], bytes))) return r;
            CK(cudaMemset(c->Clo[d], 0, bytes));
        }
    // ---- weights
    const CoeffLayout cl = coeff_layout();
    for (int s = 0; s < S; s++) {
        const bool elu = modes[s] != 0;
        c->act_enc[s] = elu ? (c->cfg.flavour ? ACT_ELU : ACT_ELU_CLAMP) : ACT_LEAKY;
        c->act_dec[s] = elu ? (c->cfg.flavour ? ACT_ELU : ACT_ELU_CLAMP) : ACT_RELU;
    }
    {
        std::vector<float> w1((size_t)S * 800), b1(S * 16), s1(S * 16), o1(S * 16);
        std::vector<float> w6((size_t)S * 800), b6(S), s6(S), o6(S), w7(S * 32), b7(S * 2);
        for (int s = 0; s < S; s++) {
            const float* k = coeffs[s];
            std::memcpy(&w1[(size_t)s * 800], k + cl.down_w[0], 800 * 4);
            std::memcpy(&b1[s * 16], k + cl.down_b[0], 16 * 4);
            std::memcpy(&o1[s * 16], k + cl.down_bn[0], 16 * 4);        // [0,C) offset
            std::memcpy(&s1[s * 16], k + cl.down_bn[0] + 16, 16 * 4);   // [C,2C) scale (spleeter.c:188)
            std::memcpy(&w6[(size_t)s * 800], k + cl.up_w[5], 800 * 4);
            b6[s] = k[cl.up_b[5]];
            o6[s] = k[cl.up_bn[5]];
            s6[s] = k[cl.up_bn[5] + 1];
            std::memcpy(&w7[s * 32], k + cl.w7, 32 * 4);
            std::memcpy(&b7[s * 2], k + cl.b7, 2 * 4);
        }
        if ((r = upload(c, &c->d_w1, w1)) || (r = upload(c, &c->d_b1, b1)) || (r = upload(c, &c->d_s1, s1)) || (r = upload(c, &c->d_o1, o1))) return r;
        c->h_w6 = w6;
        c->h_w1.assign((size_t)S * 800, 0.0f);
        c->h_b1.assign((size_t)S * 48, 0.0f);
        for (int s = 0; s < S; s++) {
            for (int o = 0; o < 16; o++)
                for (int ci = 0; ci < 2; ci++)
                    for (int tap = 0; tap < 25; tap++) c->h_w1[(size_t)s * 800 + (tap * 2 + ci) * 16 + o] = w1[((size_t)s * 16 + o) * 50 + ci * 25 + tap];
            std::memcpy(&c->h_b1[(size_t)s * 48], &b1[s * 16], 64);
            std::memcpy(&c->h_b1[(size_t)s * 48 + 16], &s1[s * 16], 64);
            std::memcpy(&c->h_b1[(size_t)s * 48 + 32], &o1[s * 16], 64);
        }
        c->h_w7.assign((size_t)S * 36, 0.0f);
        for (int s = 0; s < S; s++) {
            std::memcpy(&c->h_w7[(size_t)s * 36], &w7[s * 32], 32 * 4);
            std::memcpy(&c->h_w7[(size_t)s * 36 + 32], &b7[s * 2], 2 * 4);
        }
        if ((r = upload(c, &c->d_w6, w6)) || (r = upload(c, &c->d_b6, b6)) || (r = upload(c, &c->d_s6, s6)) || (r = upload(c, &c->d_o6, o6))) return r;
        if ((r = upload(c, &c->d_w7, w7)) || (r = upload(c, &c->d_b7, b7))) return r;
        // ---- tensor-core up6: GEMM + col2im (SRT_UP6_TC=0 keeps the SIMT kernel; conv_impl 1 is the all-SIMT check path)
        const char* u6e = getenv("SRT_UP6_TC");
        if (c->cfg.conv_impl != 1 && !(u6e && atoi(u6e) == 0) && S <= 8 && up6_tc_fits(S)) {
            Up6TcParams& q = c->up6tc;
            std::memset(&q, 0, sizeof q);
            bool exact = true;
            for (float v : w6) exact = exact && round_tf32(v) == v;
            q.w_terms = exact ? 1 : 2;
            static_assert(kUp6TcWFloatsPerStem == kUp6PackFloats, "up6 weight block");
            std::vector<float> wpk((size_t)S * kUp6TcWFloatsPerStem, 0.0f);
            for (int s = 0; s < S; s++) pack_up6_weights(&w6[(size_t)s * 800], &wpk[(size_t)s * kUp6TcWFloatsPerStem]);
            float* dw;
            if ((r = upload(c, &dw, wpk))) return r;
            q.w = dw;
            q.out = c->U[6];
            q.T = T; q.F = F; q.B = c->B; q.S = S;
            const char* pfe = getenv("SRT_UP6_PREFETCH");
            const char* dbe = getenv("SRT_UP6_DBG");
            q.prefetch_rows = pfe ? atoi(pfe) : 12;
            q.dbg = dbe ? atoi(dbe) : 0;
            const char* ste = getenv("SRT_UP6_STAGES");
            const char* ace = getenv("SRT_UP6_ACC");
            const char* l8e = getenv("SRT_UP6_LO8");
            q.lo8 = l8e ? atoi(l8e) != 0 : 1;   // every precision mode ("0": the fp32 residual tile, four rows in flight)
            const char* pre = getenv("SRT_UP6_PAIR");
            q.pair = pre ? atoi(pre) != 0 : 1;   // the two forms time the same (profiles/r2_up6_sweep.txt)
            const int max_stages = q.pair ? (q.lo8 ? 6 : 4) : (q.lo8 ? 8 : 5);
            q.stages = ste ? std::max(2, std::min(max_stages, atoi(ste))) : max_stages;
            q.acc_slots = ace ? std::max(2, std::min(8, atoi(ace) & ~1)) : 4;   // even: the rows go through in pairs
            const int W = F / 2;
            q.blocks_x = (W + 125) / 126;
            q.bw = (W + q.blocks_x - 1) / q.blocks_x;
            for (int s = 0; s < S; s++) {
                q.bias[s] = b6[s]; q.bn_scale[s] = s6[s]; q.bn_offset[s] = o6[s];
                q.act[s] = c->act_dec[s];
            }
            if ((r = make_tmap(&q.tmap[0], c->E[1], 16, W, T / 2, S * c->B, 128, 1, 1, 16))) return r;
            if ((r = make_tmap(&q.tmap[1], c->U[5], 16, W, T / 2, S * c->B, 128, 1, 1, 16))) return r;
            c->use_up6tc = true;
        }
    }
    // ---- tensor-core layers
    // Weights that are exactly TF32-representable (the reference's fp16 model) need one MMA term; anything else
    // (fp32 `.dat` dumps) is contracted as tf32(w) + tf32(w - tf32(w)) so rounding the weights never costs parity.
    bool split = false;
    for (int s = 0; s < S; s++) split = split || !weights_tf32_exact(coeffs[s]);     // ~10 ms per net: cheap next to packing
    if (const char* we = getenv("SRT_WEIGHT_SPLIT")) split = atoi(we) != 0;
    c->split_weights = split;
    // small batches: narrower N tiles so that the deep layers fill the SMs (SRT_TC_NARROW=0 keeps the wide tiles)
    const char* nwe = getenv("SRT_TC_NARROW");
    const char* fze = getenv("SRT_TC_FUSE");      // "0": decoder layers stay phase-separated (A/B timing)
    c->plans = build_plans(NetGeom{T, F}, c->B, split, S, (nwe && atoi(nwe) == 0) ? 0 : c->sm_count, c->comp_mask, !(fze && atoi(fze) == 0), c->lo_want);
    c->conv.resize(c->plans.size());
    for (size_t li = 0; li < c->plans.size(); li++) {
        const LayerPlan& L = c->plans[li];
        ConvParams& p = c->conv[li];
        std::memset(&p, 0, sizeof p);
        if (!conv_tc_supported(L.n_tile)) return fail(SRT_ERR_STATE, "layer %zu: no tensor-core kernel for an N tile of %d columns", li, L.n_tile);
        for (int ph = 0; ph < L.phases; ph++)
            if (L.kb[ph].size() > 512) return fail(SRT_ERR_STATE, "layer %zu: %zu k-blocks exceed the kernel's table", li, L.kb[ph].size());
        // weights + epilogue vectors
        std::vector<float> wpk((size_t)S * L.w_floats_per_stem), bias((size_t)S * L.cout), sc((size_t)S * L.cout, 1.0f), of((size_t)S * L.cout, 0.0f);
        for (int s = 0; s < S; s++) {
            if (!c->shared_hit) pack_layer(L, coeffs[s], &wpk[(size_t)s * L.w_floats_per_stem]);
            const float* k = coeffs[s];
            const size_t bo = L.transposed ? cl.up_b[L.index - 5] : cl.down_b[L.index + 1];
            const size_t bn = L.transposed ? cl.up_bn[L.index - 5] : cl.down_bn[L.index + 1];
            std::memcpy(&bias[(size_t)s * L.cout], k + bo, L.cout * 4);
            if (L.transposed || L.index < 4) {
                std::memcpy(&of[(size_t)s * L.cout], k + bn, L.cout * 4);
                std::memcpy(&sc[(size_t)s * L.cout], k + bn + L.cout, L.cout * 4);
            }
        }
        float *dw, *db, *ds, *dofs;
        if ((r = upload(c, &dw, wpk)) || (r = upload(c, &db, bias)) || (r = upload(c, &ds, sc)) || (r = upload(c, &dofs, of))) return r;
        // k-block tables
        std::vector<KBlock> kb;
        for (int ph = 0; ph < L.phases; ph++) {
            p.kb_off[ph] = (int)kb.size();
            p.nkb[ph] = (int)L.kb[ph].size();
            p.w_phase_off[ph] = L.w_phase_off[ph];
            kb.insert(kb.end(), L.kb[ph].begin(), L.kb[ph].end());
        }
        KBlock* dkb;
        if ((r = upload(c, &dkb, kb))) return r;
        p.kb = dkb;
        p.w = dw;
        p.w_stem_stride = L.w_floats_per_stem;
        p.bias = db;
        p.bn_scale = ds;
        p.bn_offset = dofs;
        p.n_tile = L.n_tile;
        p.n_tiles = L.n_tiles;
        p.phases = L.phases;
        p.fused = L.fused ? 1 : 0;
        p.Hs = L.Hs;
        p.Ws = L.Ws;
        p.B = c->B;
        p.S = S;
        p.tw = L.tw;
        p.th = L.th;
        p.nb = L.nb;
        p.tiles_x = (L.Ws + L.tw - 1) / L.tw;
        p.tiles_y = (L.Hs + L.th - 1) / L.th;
        p.cout = L.cout;
        for (int s = 0; s < S; s++) p.act[s] = L.transposed ? c->act_dec[s] : c->act_enc[s];
        // sources and destinations
        const float* src[2] = {nullptr, nullptr};
        if (!L.transposed) {
            const int i = L.index + 1;          // input = activated output of down{i}
            src[0] = c->A[i];
            p.mode = (L.index == 4) ? 1 : 0;
            p.out_raw = c->E[i + 1];
            p.out_act = (L.index == 4) ? nullptr : c->A[i + 1];
            p.round_raw = 1;                    // E2..E6 feed decoder tensor-core layers
            p.round_act = 1;
            // residuals for compensated consumers: the raw skip E{i+1} is read by up{6-i} (its [skip | up] residual tensor;
            // E6 by up1), the activation A{i+1} by down{i+2}
            const int dcons = 5 - i;            // Clo index of the decoder layer that reads E{i+1}
            if (c->Clo[dcons]) {
                p.lo_raw = c->Clo[dcons]; p.lo_raw_C = dcons == 0 ? 512 : 2 * L.cout; p.lo_raw_coff = 0;
                p.lo_raw_fp8 = c->lo_is_fp8(5 + dcons);
            }
            if (L.index < 4 && c->Alo[i + 1]) { p.lo_act = c->Alo[i + 1]; p.lo_act_fp8 = c->lo_is_fp8(L.index + 1); }
            if (L.comp) { p.lo_ptr = c->Alo[i]; p.lo_C = 4 * L.cin; p.lo_fp8 = L.lo_fmt == LO_FP8; }
        } else {
            const int d = L.index - 5;          // up{d+1}
            if (d == 0) src[0] = c->E[6];
            else { src[0] = c->E[6 - d]; src[1] = c->U[d]; }
            p.mode = 2;
            p.out_dec = c->U[d + 1];
            p.round_act = (d < 4) ? 1 : 0;      // up5 feeds the SIMT up6 kernel: keep fp32
            if (d < 4 && c->Clo[d + 1]) {       // U{d+1}: second half of up{d+2}'s residuals
                p.lo_dec = c->Clo[d + 1]; p.lo_dec_C = 2 * L.cout; p.lo_dec_coff = L.cout;
                p.lo_dec_fp8 = c->lo_is_fp8(5 + d + 1);
            }
            if (L.comp) { p.lo_ptr = c->Clo[d]; p.lo_C = L.cin; p.lo_fp8 = L.lo_fmt == LO_FP8; }
        }
        for (int q = 0; q < L.nsrc; q++) {
            p.src_ptr[q] = src[q];
            p.src_C[q] = L.src[q].C;
            if ((r = make_tmap(&p.tmap[q], src[q], L.src[q].C, L.src[q].W, L.src[q].H, S * c->B, L.tw, L.th, L.nb))) return r;
        }
        if (L.nsrc == 1) p.tmap[1] = p.tmap[0];
        p.tmap[2] = p.tmap[0];                  // always a valid descriptor (the kernels prefetch all three)
        if (L.comp && (r = make_tmap_lo(&p.tmap[2], p.lo_ptr, L.lo_src.C, L.lo_src.W, L.lo_src.H, S * c->B, L.tw, L.th, L.nb, L.lo_fmt))) return r;
    }
    // ---- row-patch variants of the small-N layers ------------------------------------------
    // SRT_CONV_RP: "0" = never, "1" = whenever supported, unset = when the tile row is wide enough to pay
    const char* rpe = getenv("SRT_CONV_RP");
    const char* boe = getenv("SRT_RP_BO");
    for (size_t li = 0; li < c->plans.size(); li++) {
        if (!row_plan_supported((int)li) || c->cfg.conv_impl == 1) continue;
        if (!c->will_rp[li]) continue;
        // the residual format was fixed when the tensors were allocated: the row-patch plan has to agree with it
        const int lf = c->lo_fmt_of((int)li);
        const RowPlan rpl = build_row_plan(NetGeom{T, F}, (int)li, c->split_weights, c->plans[li].comp, lf == LO_BF16 ? LO_BF16 : LO_FP8);
        const bool narrow = rpl.comp && rpl.lo_fmt == LO_FP8N;
        if (rpl.comp && rpl.lo_fmt != lf) return fail(SRT_ERR_STATE, "layer %zu: residual format mismatch (%d vs %d)", li, rpl.lo_fmt, lf);
        if (!conv_rp_fits((int)rpl.chunks.size(), (int)rpl.kb.size())) {
            if (narrow) return fail(SRT_ERR_STATE, "layer %zu: row-patch tables do not fit and its residuals are in the row-patch-only format", li);
            continue;
        }
        RowConvParams& q = c->rp[li];
        std::memset(&q, 0, sizeof q);
        std::vector<float> wpk((size_t)S * rpl.w_floats_per_stem);
        for (int s = 0; s < S && !c->shared_hit; s++) pack_row_layer(rpl, coeffs[s], &wpk[(size_t)s * rpl.w_floats_per_stem]);
        float* dw;
        RowChunk* dch;
        KBlock* dkb;
        if ((r = upload(c, &dw, wpk)) || (r = upload(c, &dch, rpl.chunks)) || (r = upload(c, &dkb, rpl.kb))) return r;
        q.chunks = dch; q.n_chunks = (int)rpl.chunks.size();
        q.kb = dkb; q.nkb = (int)rpl.kb.size();
        q.w = dw; q.w_stem_stride = rpl.w_floats_per_stem;
        q.N = rpl.N; q.R = rpl.R;
        q.tiles_x = (rpl.Ws + kTileM - 1) / kTileM;
        q.tiles_y = (rpl.Hs + rpl.R - 1) / rpl.R;
        q.bo_mode = boe ? atoi(boe) : 0;   // measured on B200: the swizzle is a function of the absolute smem address, base offset stays 0
        q.ep = c->conv[li];
        q.dbg = getenv("SRT_RP_DBG") ? atoi(getenv("SRT_RP_DBG")) : 0;
        bool ok = true;
        for (int k = 0; k < rpl.nsrc; k++)
            if (make_tmap(&q.tmap[k], c->conv[li].src_ptr[k], rpl.src[k].C, rpl.src[k].W, rpl.src[k].H, S * c->B, kPatchW, rpl.R + 2, 1)) ok = false;
        if (rpl.nsrc == 1) q.tmap[1] = q.tmap[0];
        q.tmap[2] = q.tmap[0];
        if (ok && rpl.comp && make_tmap_lo(&q.tmap[2], c->conv[li].lo_ptr, rpl.lo_src.C, rpl.lo_src.W, rpl.lo_src.H, S * c->B, kPatchW, rpl.R + 2, 1, rpl.lo_fmt)) ok = false;
        if (!ok && narrow) return fail(SRT_ERR_CUDA, "layer %zu: row-patch tensor map rejected (%s)", li, g_err.c_str());
        if (!ok) { fprintf(stderr, "[spleeterrt_b200] row-patch tensor map rejected for layer %zu (%s); using the generic kernel\n", li, g_err.c_str()); continue; }
        c->use_rp[li] = true;
    }
    // ---- down1 on the tensor cores: the magnitude image is shared by all stems, so up to 4 stems ride one MMA
    // (SRT_DOWN1_TC=0 keeps the SIMT kernel)
    const char* d1e = getenv("SRT_DOWN1_TC");
    if (c->cfg.conv_impl != 1 && !(d1e && atoi(d1e) == 0)) {
        const Down1Plan dp = build_down1_plan(NetGeom{T, F}, c->split_weights);
        KBlock* dkb;
        const int ntap = (int)dp.kb.size() / 2;
        std::vector<RowChunk> chunk = {RowChunk{0, 0, 0, ntap}, RowChunk{1, 0, ntap, ntap}};
        RowChunk* dch;
        if ((r = upload(c, &dkb, dp.kb)) || (r = upload(c, &dch, chunk))) return r;
        std::vector<float> one(S * 16, 1.0f), zero(S * 16, 0.0f);
        for (int s0 = 0; s0 < S;) {
            const int g = (S - s0 >= 4) ? 4 : (S - s0 >= 2) ? 2 : 1;
            RowConvParams& q = c->d1[c->n_d1];
            std::memset(&q, 0, sizeof q);
            std::vector<float> wpk((size_t)dp.kb.size() * 16 * g * kKB1);
            if (!c->shared_hit) pack_down1(dp, coeffs + s0, g, wpk.data());
            float* dw;
            if ((r = upload(c, &dw, wpk))) return r;
            q.chunks = dch; q.n_chunks = 2; q.kb = dkb; q.nkb = (int)dp.kb.size();
            q.w = dw; q.w_stem_stride = 0;
            q.N = 16 * g; q.R = 4; q.kb_width = kKB1;
            q.tiles_x = (dp.Ws + kTileM - 1) / kTileM;
            q.tiles_y = (dp.Hs + q.R - 1) / q.R;
            q.stems_per_tile = g; q.stem0 = s0;
            q.dbg = getenv("SRT_RP_DBG") ? atoi(getenv("SRT_RP_DBG")) : 0;
            ConvParams& e = q.ep;
            e.Hs = dp.Hs; e.Ws = dp.Ws; e.B = c->B; e.S = 1; e.cout = 16;
            e.bias = c->d_b1; e.bn_scale = c->d_s1; e.bn_offset = c->d_o1;
            for (int s = 0; s < S; s++) e.act[s] = c->act_enc[s];
            e.mode = 0; e.out_raw = c->E[1]; e.out_act = c->A[1];
            e.round_raw = 0;            // the skip feeds the fp32 SIMT up6 kernel
            e.round_act = 1;
            e.lo_act = c->Alo[1];       // residual of A1 for a compensated down2 (nullptr otherwise)
            e.lo_act_fp8 = c->lo_is_fp8(0);
            if (make_tmap(&q.tmap[0], c->d_mag, 8, F / 2, T / 2, c->NB, kPatchW, q.R + 2, 1, kKB1) ||
                make_tmap(&q.tmap[1], c->d_mag + (size_t)c->NB * T * F * 2, 8, F / 2, T / 2, c->NB, kPatchW, q.R + 2, 1, kKB1)) {
                fprintf(stderr, "[spleeterrt_b200] down1 tensor map rejected (%s); using the SIMT kernel\n", g_err.c_str());
                c->n_d1 = 0;
                break;
            }
            q.tmap[2] = q.tmap[0];
            c->n_d1++;
            s0 += g;
        }
    }
    return 0;
}

extern "C" int srt_create(const srt_config* cfg, const float* const* coeffs, const int* stem_modes, srt_ctx** out)
{
    if (!cfg || !out) return fail(SRT_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->n_stems < 0 || cfg->n_stems > SRT_MAX_STEMS) return fail(SRT_ERR_ARG, "n_stems %d out of range", cfg->n_stems);
    if (cfg->n_stems > 0 && (!coeffs || !stem_modes)) return fail(SRT_ERR_ARG, "weights missing");
    if (cfg->time_step < 1 || cfg->max_images < 1) return fail(SRT_ERR_ARG, "time_step/max_images must be positive");
    if (cfg->n_stems > 0) {
        // initSpleeter needs six halvings (spleeter.c:113-119); the CLI clamps F to [512, 2048] (main.c:733-748)
        if (cfg->time_step % 64 || cfg->bin_limit % 64 || cfg->bin_limit < 64 || cfg->bin_limit > 2048)
            return fail(SRT_ERR_ARG, "time_step (%d) and bin_limit (%d) must be multiples of 64, bin_limit <= 2048", cfg->time_step, cfg->bin_limit);
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(SRT_ERR_CUDA, "no CUDA device: this library has no CPU path");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(SRT_ERR_ARG, "device %d out of range (%d devices)", cfg->device, ndev);
    internal::DeviceGuard dev_guard(cfg->device);
    if (!dev_guard.ok) return fail(SRT_ERR_CUDA, "cudaSetDevice(%d) failed", cfg->device);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return fail(SRT_ERR_CUDA, "device %d is sm_%d%d; this build contains sm_100a code only", cfg->device, prop.major, prop.minor);
    srt_ctx* c = new srt_ctx();
    c->cfg = *cfg;
    c->sm_count = prop.multiProcessorCount;
    c->S = cfg->n_stems;
    c->T = cfg->time_step;
    c->F = cfg->n_stems ? cfg->bin_limit : 0;
    c->B = cfg->max_images;
    c->NB = cfg->max_batch_images > cfg->max_images ? cfg->max_batch_images : cfg->max_images;
    if (cfg->cuda_stream) c->stream = (cudaStream_t)cfg->cuda_stream;
    else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return fail(SRT_ERR_CUDA, "stream creation failed"); }
        c->own_stream = true;
    }
    int r = build(c, coeffs, stem_modes);
    // build() uploads with blocking copies / memsets on the legacy stream: wait for that stream and the context's own, not for the
    // whole device (it may be busy with the caller's other streams)
    if (r == 0 && (cudaStreamSynchronize(cudaStreamLegacy) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess)) r = fail(SRT_ERR_CUDA, "context build: %s", cudaGetErrorString(cudaGetLastError()));
    if (r) { std::string keep = g_err; srt_destroy(c); g_err = keep; return r; }
    *out = c;
    return 0;
}

// The CLI's output modes (Executable/main.c:776-970) as device-resident contexts: one single-net context per cascade
// stage, the second stage enqueuing on the first one's stream.  Net roles and activations as main.c: the drum net is
// coeffProvPtr2 with stemMode 1 (:858), the vocal net coeffProvPtr1 with stemMode 0 (:782, :911).
extern "C" int srt_create_cli(const srt_config* cfg, int n_outputs, const float* const* coeffs, srt_ctx** out)
{
    if (!cfg || !out || !coeffs) return fail(SRT_ERR_ARG, "null argument");
    *out = nullptr;
    if (n_outputs != 2 && n_outputs != 3) return fail(SRT_ERR_ARG, "n_outputs must be 2 or 3 (main.c:776, 845), got %d", n_outputs);
    srt_config c1 = *cfg;
    c1.n_stems = 1;
    const int mode_first = n_outputs == 3 ? 1 : 0;
    srt_ctx* a = nullptr;
    int r = srt_create(&c1, coeffs, &mode_first, &a);
    if (r) return r;
    a->cli_mode = n_outputs;
    if (n_outputs == 3) {
        srt_config c2 = c1;
        c2.cuda_stream = (void*)a->stream;
        const int mode_second = 0;
        if ((r = srt_create(&c2, coeffs + 1, &mode_second, &a->next))) {
            std::string keep = g_err; srt_destroy(a); g_err = keep;
            return r;
        }
    }
    *out = a;
    return 0;
}

// ------------------------------------------------------------------------------------------
// U-Net on Bv images whose magnitudes sit at d_mag (layout [Bv][T][F][2]); masks go to
// mask_base[s][mask_img0 + b] with `mask_stride` images between stems.
// ------------------------------------------------------------------------------------------
static int run_unet_launches(srt_ctx* c, int mag_img0, int Bv, float* mask_base, int mask_stride, int mask_img0)
{
    const int S = c->S;
    c->last_Bv = Bv;
    {
        Timed t(c, 10);
        if (c->n_d1 > 0) {
            for (int g = 0; g < c->n_d1; g++) {
                RowConvParams& q = c->d1[g];
                q.ep.Bv = Bv;
                q.src_img0 = mag_img0;
                launch_conv_rp(q, c->stream);
                c->launches++;
            }
        } else {
            Down1Params p{};
            p.mag = c->d_mag + (size_t)mag_img0 * c->T * c->F * 2;
            p.mag_lo = p.mag + (size_t)c->NB * c->T * c->F * 2;
            p.w = c->d_w1; p.bias = c->d_b1; p.bn_scale = c->d_s1; p.bn_offset = c->d_o1;
            p.out_raw = c->E[1]; p.out_act = c->A[1]; p.lo_act = reinterpret_cast<uint16_t*>(c->Alo[1]);   // down2's residuals are always bf16 (64 channels)
            p.T = c->T; p.F = c->F; p.B = c->B; p.Bv = Bv; p.S = S;
            for (int s = 0; s < S; s++) {
                p.stem = s;
                p.act[0] = c->act_enc[s];
                std::memcpy(p.wk, &c->h_w1[(size_t)s * 800], 800 * sizeof(float));
                std::memcpy(p.bk, &c->h_b1[(size_t)s * 48], 48 * sizeof(float));
                launch_down1(p, c->stream);
                c->launches++;
            }
        }
    }
    for (size_t li = 0; li < c->conv.size(); li++) {
        Timed t(c, (int)li);
        ConvParams& p = c->conv[li];
        p.Bv = Bv;
        p.tiles_n = (Bv + p.nb - 1) / p.nb;
        if (c->cfg.conv_impl == 1) launch_conv_simt(p, c->stream);
        else if (c->use_rp[li]) { c->rp[li].ep.Bv = Bv; launch_conv_rp(c->rp[li], c->stream); }
        else launch_conv_tc(p, c->stream, c->sm_count);
        c->launches++;
    }
    if (c->use_up6tc) {
        Timed t(c, 11);
        Up6TcParams& q = c->up6tc;
        q.Bv = Bv;
        // row chunks: enough units for ~8 waves of persistent CTAs, chunks no shorter than 32 rows
        const int H = c->T / 2;
        int chunks = 1;
        while (q.blocks_x * chunks * Bv * S < 8 * c->sm_count && H / (chunks * 2) >= 32) chunks *= 2;
        q.rows_per_unit = (((H + chunks - 1) / chunks) + 1) & ~1;
        q.chunks = (H + q.rows_per_unit - 1) / q.rows_per_unit;
        launch_up6_tc(q, c->stream);
        c->launches++;
    } else {
        Timed t(c, 11);
        Up6Params p{};
        p.skip = c->E[1]; p.up = c->U[5]; p.w = c->d_w6; p.bias = c->d_b6; p.bn_scale = c->d_s6; p.bn_offset = c->d_o6;
        p.out = c->U[6];
        p.T = c->T; p.F = c->F; p.B = c->B; p.Bv = Bv; p.S = S;
        for (int s = 0; s < S; s++) {
            p.stem = s;
            p.act[0] = c->act_dec[s];
            std::memcpy(p.wk, &c->h_w6[(size_t)s * 800], 800 * sizeof(float));
            launch_up6(p, c->stream);
            c->launches++;
        }
    }
    {
        Timed t(c, 12);
        Up7Params p{};
        p.in = c->U[6]; p.w = c->d_w7; p.bias = c->d_b7; p.lut = c->d_lut; p.mask = mask_base;
        p.T = c->T; p.F = c->F; p.B = c->B; p.Bv = Bv; p.S = S;
        p.mask_stem_stride = mask_stride; p.mask_img0 = mask_img0;
        // per-stem launches keep the weights as immediates of the constant bank; when one stem's grid would not even fill
        // the SMs (small batches) all stems share one launch instead of queueing behind each other
        const long ctas_per_stem = (long)((c->F + 127) / 128) * ((c->T + 63) / 64) * Bv;
        const char* m7 = getenv("SRT_UP7_MERGED");
        if (S > 1 && S <= 8 && (m7 ? atoi(m7) != 0 : ctas_per_stem < 2L * c->sm_count)) {
            p.merged = 1;
            for (int s = 0; s < S; s++) std::memcpy(p.wk_all[s], &c->h_w7[(size_t)s * 36], 36 * sizeof(float));
            launch_up7(p, c->stream);
            c->launches++;
        } else {
            for (int s = 0; s < S; s++) {
                p.stem = s;
                std::memcpy(p.wk, &c->h_w7[(size_t)s * 36], 36 * sizeof(float));
                launch_up7(p, c->stream);
                c->launches++;
            }
        }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SRT_ERR_CUDA, "U-Net launch: %s", cudaGetErrorString(e));
    return 0;
}

// The pass as a CUDA graph (captured on first use per key, then replayed).  Per-kernel timing (srt_set_timing) and the SIMT
// verification path launch directly.  SRT_GRAPHS=0 disables.
static int run_unet(srt_ctx* c, int mag_img0, int Bv, float* mask_base, int mask_stride, int mask_img0)
{
    if (!c->use_graphs || c->timing || c->cfg.conv_impl == 1) return run_unet_launches(c, mag_img0, Bv, mask_base, mask_stride, mask_img0);
    srt_ctx::UnetGraph* slot = nullptr;
    for (auto& g : c->graphs)
        if (g.mag_img0 == mag_img0 && g.Bv == Bv && g.mask_base == mask_base && g.mask_stride == mask_stride && g.mask_img0 == mask_img0) { slot = &g; break; }
    if (slot && slot->exec) {
        c->last_Bv = Bv;
        CK(cudaGraphLaunch(slot->exec, c->stream));
        c->launches += slot->kernels;
        return 0;
    }
    if (!slot) {
        // first pass with this key: plain launches (they also do the kernels' one-time setup - opt-in shared-memory attributes -
        // outside any capture); a key that comes back is captured on its second pass
        if (c->graphs.size() < 64) c->graphs.push_back(srt_ctx::UnetGraph{mag_img0, Bv, mask_stride, mask_img0, mask_base, nullptr, 0});
        return run_unet_launches(c, mag_img0, Bv, mask_base, mask_stride, mask_img0);
    }
    const long long before = c->launches;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        (void)cudaGetLastError();
        c->use_graphs = false;                                   // e.g. the caller is already capturing this stream
        return run_unet_launches(c, mag_img0, Bv, mask_base, mask_stride, mask_img0);
    }
    const int r = run_unet_launches(c, mag_img0, Bv, mask_base, mask_stride, mask_img0);
    const cudaError_t e = cudaStreamEndCapture(c->stream, &graph);
    const int kernels = (int)(c->launches - before);
    c->launches = before;
    if (r || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        (void)cudaGetLastError();
        c->use_graphs = false;
        return r ? r : run_unet_launches(c, mag_img0, Bv, mask_base, mask_stride, mask_img0);
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess || !exec) {
        (void)cudaGetLastError();
        c->use_graphs = false;
        return run_unet_launches(c, mag_img0, Bv, mask_base, mask_stride, mask_img0);
    }
    slot->exec = exec;
    slot->kernels = kernels;
    CK(cudaGraphLaunch(exec, c->stream));
    c->launches += kernels;
    return 0;
}

// Spans accumulate over calls while timing is on (so a caller can time K back-to-back steps
// without a host sync in between) and are cleared by srt_set_timing().
namespace srt {
namespace internal {
int ctx_run_unet(srt_ctx* c, int mag_img0, int Bv, float* mask_base, int mask_stride, int mask_img0)
{
    return run_unet(c, mag_img0, Bv, mask_base, mask_stride, mask_img0);
}
float* ctx_mag(srt_ctx* c) { return c->d_mag; }
cudaStream_t ctx_stream(srt_ctx* c) { return c->stream; }
int ctx_device(srt_ctx* c) { return c->cfg.device; }
const float2* ctx_twiddle(srt_ctx* c) { return c->d_twiddle; }
void ctx_count_launch(srt_ctx* c, int n) { c->launches += n; }
int set_error(int code, const char* msg) { return fail(code, "%s", msg); }
}  // namespace internal
}  // namespace srt

static void reset_spans(srt_ctx* c)
{
    if (c->timing) return;
    c->spans.clear();
    c->ev_used = 0;
}

extern "C" int srt_unet_device(srt_ctx* c, const float* d_mag, int n_img, float* d_mask)
{
    if (!c || c->S == 0) return fail(SRT_ERR_STATE, "context has no nets");
    if (n_img < 1 || n_img > c->B) return fail(SRT_ERR_CAPACITY, "n_img %d exceeds max_images %d", n_img, c->B);
    internal::DeviceGuard dev_guard(c->cfg.device);
    if (!dev_guard.ok) return fail(SRT_ERR_CUDA, "cudaSetDevice(%d) failed", c->cfg.device);
    reset_spans(c);
    launch_mag_to_s2d(d_mag, c->d_mag, c->d_mag + (size_t)c->NB * c->T * c->F * 2, c->T, c->F, n_img, c->stream);   // API layout -> space-to-depth hi/lo
    c->launches++;
    return run_unet(c, 0, n_img, d_mask, n_img, 0);
}

extern "C" int srt_unet_host(srt_ctx* c, const float* x, int n_img, float* y)
{
    if (!c || c->S == 0) return fail(SRT_ERR_STATE, "context has no nets");
    if (n_img < 1 || n_img > c->B) return fail(SRT_ERR_CAPACITY, "n_img %d exceeds max_images %d", n_img, c->B);
    internal::DeviceGuard dev_guard(c->cfg.device);
    if (!dev_guard.ok) return fail(SRT_ERR_CUDA, "cudaSetDevice(%d) failed", c->cfg.device);
    reset_spans(c);
    const size_t P = (size_t)c->T * c->F;
    std::vector<float> xi((size_t)n_img * P * 2), xlo((size_t)n_img * P * 2);   // space-to-depth hi / lo parts (what the device path stores)
    for (int b = 0; b < n_img; b++)
        for (int t = 0; t < c->T; t++)
            for (int f = 0; f < c->F; f++) {
                const size_t o = ((size_t)b * P + mag_s2d_index(c->T, c->F, t, f)) * 2, i = (size_t)t * c->F + f;
                for (int ch = 0; ch < 2; ch++) {
                    const float m = x[((size_t)b * 2 + ch) * P + i], hi = round_tf32(m);
                    xi[o + ch] = hi;
                    xlo[o + ch] = m - hi;
                }
            }
    CK(cudaMemcpyAsync(c->d_mag, xi.data(), xi.size() * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_mag + (size_t)c->NB * P * 2, xlo.data(), xlo.size() * 4, cudaMemcpyHostToDevice, c->stream));
    int r = run_unet(c, 0, n_img, c->d_mask, c->NB, 0);
    if (r) return r;
    std::vector<float> mi((size_t)n_img * P * 2);
    for (int s = 0; s < c->S; s++) {
        CK(cudaMemcpyAsync(mi.data(), c->d_mask + (size_t)s * c->NB * P * 2, mi.size() * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int b = 0; b < n_img; b++)
            for (size_t i = 0; i < P; i++) {
                y[(((size_t)s * n_img + b) * 2 + 0) * P + i] = mi[((size_t)b * P + i) * 2 + 0];
                y[(((size_t)s * n_img + b) * 2 + 1) * P + i] = mi[((size_t)b * P + i) * 2 + 1];
            }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// full path
// ------------------------------------------------------------------------------------------
struct BatchMeta {
    std::vector<int> n, nfr, img0;
    std::vector<ImgDesc> imgs;
    int total = 0;
    size_t max_n = 0;
};

static size_t padded_len(size_t n) { return (size_t)kFFT * ((n + kFFT - 1) / kFFT) + 2 * kFFT; }   // main.c:762-763

// Picks the next metadata slot (device block + pinned host mirror).  Waits only if that slot's
// previous upload, kMetaSlots calls ago, has not been consumed yet.
static int ensure_meta(srt_ctx* c, size_t bytes)
{
    if (bytes > c->meta_cap) {
        CK(cudaStreamSynchronize(c->stream));
        if (c->d_meta_base) cudaFree(c->d_meta_base);
        if (c->h_meta_base) cudaFreeHost(c->h_meta_base);
        c->meta_cap = (bytes * 2 + 4095) & ~(size_t)4095;
        CK(cudaMalloc((void**)&c->d_meta_base, c->meta_cap * kMetaSlots));
        CK(cudaHostAlloc((void**)&c->h_meta_base, c->meta_cap * kMetaSlots, cudaHostAllocMapped));
        CK(cudaHostGetDevicePointer((void**)&c->h_meta_dev, c->h_meta_base, 0));
        for (int i = 0; i < kMetaSlots; i++)
            if (!c->meta_ev[i]) CK(cudaEventCreateWithFlags(&c->meta_ev[i], cudaEventDisableTiming));
    }
    c->meta_slot = (c->meta_slot + 1) % kMetaSlots;
    CK(cudaEventSynchronize(c->meta_ev[c->meta_slot]));
    c->d_meta = c->d_meta_base + (size_t)c->meta_slot * c->meta_cap;
    c->h_meta = c->h_meta_base + (size_t)c->meta_slot * c->meta_cap;
    return 0;
}
// The metadata block goes up with a one-CTA kernel that reads the mapped pinned mirror, not with a memcpy: a
// copy-engine transfer on the compute stream queues behind the bulk PCM uploads of the next batches (one merged
// DMA of ~100 MB takes 2-3 ms) and stalled the first kernel of every batch (tools/e2e_probe.py, r1).
__global__ void meta_upload_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n16)
{
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
}
static int commit_meta(srt_ctx* c, size_t bytes)
{
    const size_t slot_off = (size_t)c->meta_slot * c->meta_cap;
    meta_upload_kernel<<<1, 256, 0, c->stream>>>((const uint4*)(c->h_meta_dev + slot_off), (uint4*)c->d_meta, (int)((bytes + 15) / 16));
    CK(cudaGetLastError());
    c->launches++;
    CK(cudaEventRecord(c->meta_ev[c->meta_slot], c->stream));
    return 0;
}

// pcm_stride (nullable = planar): per-stream distance in floats between consecutive samples of a channel; out_stride:
// 1 = planar outputs, 2 = interleaved stereo frames (d_out still lists an L and an R pointer per pair, R = L + 1).
static int separate_core(srt_ctx* c, const float* const* d_pcmL, const float* const* d_pcmR, const size_t* n_samples,
                         int n_streams, const float* unaffected, float* const* d_out, int front_pad,
                         const int* pcm_stride = nullptr, int out_stride = 1)
{
    const int T = c->T, S = c->S, P = c->pairs();
    const char* fo = getenv("SRT_FUSED_OLA");
    const bool fused = !(fo && atoi(fo) == 0);
    BatchMeta m;
    for (int i = 0; i < n_streams; i++) {
        if (n_samples[i] == 0 || n_samples[i] > (size_t)1 << 30) return fail(SRT_ERR_ARG, "stream %d: bad length", i);
        const int nfr = (int)(padded_len(n_samples[i]) / kHop);
        const int tiles = (nfr + T - 1) / T;
        // the two-kernel synthesis (SRT_FUSED_OLA=0) keeps a stream's frames in a max_images-sized scratch; the fused kernel
        // only needs the whole batch to fit max_batch_images
        if (tiles > c->B && !fused) return fail(SRT_ERR_CAPACITY, "stream %d needs %d tiles > max_images %d", i, tiles, c->B);
        m.n.push_back((int)n_samples[i]);
        m.nfr.push_back(nfr);
        m.img0.push_back(m.total);
        for (int j = 0; j < tiles; j++) m.imgs.push_back(ImgDesc{i, j * T});
        m.total += tiles;
        if (n_samples[i] > m.max_n) m.max_n = n_samples[i];
    }
    if (m.total > c->NB) return fail(SRT_ERR_CAPACITY, "batch needs %d tiles > max_batch_images %d", m.total, c->NB);
    // ---- metadata block: [pcmL ptrs][pcmR ptrs][out ptrs][n][nfr][img0][imgs]
    const size_t o_pl = 0, o_pr = o_pl + 8 * (size_t)n_streams, o_out = o_pr + 8 * (size_t)n_streams;
    const size_t o_n = o_out + 8 * (size_t)n_streams * P * 2, o_nfr = o_n + 4 * (size_t)n_streams, o_i0 = o_nfr + 4 * (size_t)n_streams;
    const size_t o_ps = o_i0 + 4 * (size_t)n_streams;
    const size_t o_img = (o_ps + 4 * (size_t)n_streams + 7) & ~(size_t)7, total_b = o_img + sizeof(ImgDesc) * m.imgs.size();
    int r = ensure_meta(c, total_b);
    if (r) return r;
    std::memcpy(c->h_meta + o_pl, d_pcmL, 8 * (size_t)n_streams);
    std::memcpy(c->h_meta + o_pr, d_pcmR, 8 * (size_t)n_streams);
    std::memcpy(c->h_meta + o_out, d_out, 8 * (size_t)n_streams * P * 2);
    std::memcpy(c->h_meta + o_n, m.n.data(), 4 * (size_t)n_streams);
    std::memcpy(c->h_meta + o_nfr, m.nfr.data(), 4 * (size_t)n_streams);
    std::memcpy(c->h_meta + o_i0, m.img0.data(), 4 * (size_t)n_streams);
    for (int i = 0; i < n_streams; i++) ((int*)(c->h_meta + o_ps))[i] = pcm_stride ? pcm_stride[i] : 1;
    std::memcpy(c->h_meta + o_img, m.imgs.data(), sizeof(ImgDesc) * m.imgs.size());
    if ((r = commit_meta(c, total_b))) return r;
    const ImgDesc* d_imgs = (const ImgDesc*)(c->d_meta + o_img);
    const int* d_n = (const int*)(c->d_meta + o_n);
    const int* d_nfr = (const int*)(c->d_meta + o_nfr);
    // ---- STFT + magnitude for every tile of the batch
    {
        Timed t(c, 13);
        StftParams p{};
        p.pcmL = (const float* const*)(c->d_meta + o_pl);
        p.pcmR = (const float* const*)(c->d_meta + o_pr);
        p.n_samples = d_n; p.n_frames = d_nfr; p.imgs = d_imgs;
        p.pcm_stride = pcm_stride ? (const int*)(c->d_meta + o_ps) : nullptr;
        p.window = c->d_window; p.twiddle = c->d_twiddle;
        p.spec = c->d_spec; p.mag = c->d_mag; p.mag_lo_off = (size_t)c->NB * T * c->F * 2;
        p.T = T; p.F = c->F; p.n_img = m.total; p.front_pad = front_pad;
        launch_stft(p, c->stream);
        c->launches++;
    }
    // ---- U-Net, max_images tiles per pass
    for (int i0 = 0; i0 < m.total; i0 += c->B) {
        const int Bv = std::min(c->B, m.total - i0);
        if ((r = run_unet(c, i0, Bv, c->d_mask, c->NB, i0))) return r;
    }
    // ---- mask * spectrum -> inverse FFT -> window -> overlap-add -> un-framing, one fused kernel
    // (SRT_FUSED_OLA=0 selects the two-kernel path through the scratch frames, kept for the tier-A istft())
    if ((c->cli_mode || out_stride != 1) && !fused)
        return fail(SRT_ERR_STATE, "the CLI output modes and interleaved outputs need the fused iSTFT+OLA kernel (unset SRT_FUSED_OLA)");
    int max_fr = 0;
    for (int i = 0; i < n_streams; i++) max_fr = std::max(max_fr, m.nfr[i]);
    // Hops per CTA: a CTA spends h + 3 transforms on h hops (3 warm-up frames of overlap), and the grid runs
    // in waves of 2 CTAs per SM (128 registers x 256 threads).  Pick the h that minimises waves x (h + 3):
    // small h for a single stream (fill the SMs), large h for a full batch (amortise the warm-up) while
    // keeping the last wave full.  SRT_ISTFT_HOPS overrides.
    auto hops_per_cta = [&](int transforms) {
        const long long slots = 2LL * c->sm_count;
        long long best_cost = -1;
        int best_h = 16;
        for (int h = 4; h <= 64; h++) {
            long long ctas = 0;
            for (int i = 0; i < n_streams; i++) ctas += (m.nfr[i] + h - 1) / h;
            ctas *= transforms;
            const long long cost = ((ctas + slots - 1) / slots) * (h + 3);
            if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best_h = h; }
        }
        const char* he = getenv("SRT_ISTFT_HOPS");
        return he ? std::max(1, atoi(he)) : best_h;
    };
    // one fused synthesis launch: `transforms` inverse transforms per frame, the first `masked` of them through
    // the masks of `src`'s nets, writing output pairs pair_first...
    auto synthesis = [&](srt_ctx* src, int transforms, int masked, int pair_first) {
        Timed t(c, 14);
        IstftOlaParams p{};
        p.spec = src->d_spec; p.mask = src->d_mask;
        p.stream_img0 = (const int*)(c->d_meta + o_i0); p.n_frames = d_nfr; p.n_samples = d_n;
        p.postwin = c->d_postwin; p.twiddle = c->d_twiddle; p.out = (float* const*)(c->d_meta + o_out);
        for (int s = 0; s < masked; s++) p.unaffected[s] = c->cli_mode ? (unaffected ? unaffected[0] : 0.1f) : (unaffected ? unaffected[s] : 0.1f);
        p.T = T; p.F = c->F; p.S = transforms; p.S_masked = masked; p.out_pairs = P; p.pair_first = pair_first; p.out_stride = out_stride;
        p.mask_stem_stride = c->NB; p.stream_first = 0; p.front_pad = front_pad;
        p.hops_per_cta = hops_per_cta(transforms);
        launch_istft_ola(p, n_streams, max_fr, c->stream);
        c->launches++;
    };
    auto difference = [&](bool from_pcm, int dst_pair, int sub_pair) {
        Timed t(c, 15);
        DiffParams p{};
        p.pcmL = from_pcm ? (const float* const*)(c->d_meta + o_pl) : nullptr;
        p.pcmR = from_pcm ? (const float* const*)(c->d_meta + o_pr) : nullptr;
        p.out = (float* const*)(c->d_meta + o_out);
        p.n_samples = d_n;
        p.pcm_stride = pcm_stride ? (const int*)(c->d_meta + o_ps) : nullptr;
        p.out_stride = out_stride;
        p.out_pairs = P; p.dst_pair = dst_pair; p.sub_pair = sub_pair;
        p.n_streams = n_streams; p.max_samples = (int)m.max_n;
        launch_diff(p, c->stream);
        c->launches++;
    };
    if (c->cli_mode == 2) {
        // vocal, then accompaniment = input - vocal in the time domain (main.c:782-794)
        synthesis(c, 1, 1, 0);
        difference(true, 1, 0);
    } else if (c->cli_mode == 3) {
        // drum net -> residual spectrum -> vocal net on the residual -> accompaniment = residual - vocal (main.c:845-927)
        srt_ctx* n = c->next;
        synthesis(c, 1, 1, 0);
        {
            Timed t(c, 13);
            ResidualParams p{};
            p.spec_in = c->d_spec; p.mask_in = c->d_mask; p.imgs = d_imgs; p.n_frames = d_nfr;
            p.spec_out = n->d_spec; p.mag = n->d_mag; p.mag_lo_off = (size_t)n->NB * T * c->F * 2;
            p.unaffected = unaffected ? unaffected[0] : 0.1f;
            p.T = T; p.F = c->F; p.n_img = m.total;
            launch_residual(p, c->stream);
            c->launches++;
        }
        for (int i0 = 0; i0 < m.total; i0 += n->B) {
            const int Bv = std::min(n->B, m.total - i0);
            if ((r = run_unet(n, i0, Bv, n->d_mask, n->NB, i0))) return r;
        }
        synthesis(n, 2, 1, 1);      // pair 1 = vocal (masked), pair 2 = the residual itself (main.c:881)
        difference(false, 2, 1);
    } else if (fused) {
        synthesis(c, S, S, 0);
    } else {
    int s0 = 0;
    while (s0 < n_streams) {
        int s1 = s0, imgs = 0;
        size_t gmax = 0;
        while (s1 < n_streams) {
            const int tiles = (m.nfr[s1] + T - 1) / T;
            if (imgs + tiles > c->B) break;
            imgs += tiles;
            gmax = std::max(gmax, (size_t)m.n[s1]);
            s1++;
        }
        {
            Timed t(c, 14);
            IstftParams p{};
            p.spec = c->d_spec; p.mask = c->d_mask; p.imgs = d_imgs; p.n_frames = d_nfr;
            p.postwin = c->d_postwin; p.twiddle = c->d_twiddle; p.frames_out = c->d_frames;
            for (int s = 0; s < S; s++) p.unaffected[s] = unaffected ? unaffected[s] : 0.1f;
            p.T = T; p.F = c->F; p.S = S;
            p.img_first = m.img0[s0]; p.n_img = imgs;
            p.mask_stem_stride = c->NB; p.frames_stem_stride = c->B;
            launch_istft(p, c->stream);
            c->launches++;
        }
        {
            Timed t(c, 15);
            OlaParams p{};
            p.frames = c->d_frames; p.stream_img0 = (const int*)(c->d_meta + o_i0); p.n_frames = d_nfr; p.n_samples = d_n;
            p.out = (float* const*)(c->d_meta + o_out);
            p.T = T; p.S = S; p.stream_first = s0; p.n_streams = s1 - s0; p.img_first = m.img0[s0];
            p.frames_stem_stride = c->B; p.max_samples = (int)gmax; p.front_pad = front_pad;
            launch_ola(p, c->stream);
            c->launches++;
        }
        s0 = s1;
    }
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SRT_ERR_CUDA, "separate launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int srt_separate_device(srt_ctx* c, const float* const* d_pcmL, const float* const* d_pcmR, const size_t* n_samples,
                                   int n_streams, const float* unaffected, float* const* d_stems_out)
{
    if (!c || c->S == 0) return fail(SRT_ERR_STATE, "context has no nets");
    if (n_streams < 1) return fail(SRT_ERR_ARG, "n_streams < 1");
    internal::DeviceGuard dev_guard(c->cfg.device);
    if (!dev_guard.ok) return fail(SRT_ERR_CUDA, "cudaSetDevice(%d) failed", c->cfg.device);
    reset_spans(c);
    return separate_core(c, d_pcmL, d_pcmR, n_samples, n_streams, unaffected, d_stems_out, kFFT);
}

// Interleaved frames either side of the path (the decoder's buffer before channel_splitFloat, main.c:767; the WAV
// writer's buffer after channel_joinFloat, main.c:806): the split and the join are strided accesses of the STFT loads
// and the overlap-add stores instead of host loops.
static int check_channels(const int* channels, int n_streams)
{
    if (!channels) return fail(SRT_ERR_ARG, "channels is NULL");
    for (int i = 0; i < n_streams; i++)
        if (channels[i] != 1 && channels[i] != 2) return fail(SRT_ERR_ARG, "stream %d: %d channels (1 or 2 supported, main.c:764-769)", i, channels[i]);
    return 0;
}

extern "C" int srt_separate_device_interleaved(srt_ctx* c, const float* const* d_pcm, const int* channels, const size_t* n_samples,
                                               int n_streams, const float* unaffected, float* const* d_out)
{
    if (!c || c->S == 0) return fail(SRT_ERR_STATE, "context has no nets");
    if (n_streams < 1) return fail(SRT_ERR_ARG, "n_streams < 1");
    int r = check_channels(channels, n_streams);
    if (r) return r;
    internal::DeviceGuard dev_guard(c->cfg.device);
    if (!dev_guard.ok) return fail(SRT_ERR_CUDA, "cudaSetDevice(%d) failed", c->cfg.device);
    reset_spans(c);
    const int P = c->pairs();
    std::vector<const float*> L(n_streams), R(n_streams);
    std::vector<float*> out((size_t)n_streams * P * 2);
    for (int i = 0; i < n_streams; i++) {
        L[i] = d_pcm[i];
        R[i] = d_pcm[i] + (channels[i] == 2 ? 1 : 0);       // mono: both channels read the same samples (main.c:768-769)
        for (int q = 0; q < P; q++) {
            out[((size_t)i * P + q) * 2] = d_out[(size_t)i * P + q];
            out[((size_t)i * P + q) * 2 + 1] = d_out[(size_t)i * P + q] + 1;
        }
    }
    return separate_core(c, L.data(), R.data(), n_samples, n_streams, unaffected, out.data(), kFFT, channels, 2);
}

static int batch_wait_slot(srt_ctx* c, int slot);

// Enqueues one host-pointer batch into staging slot `slot`: H2D on s_in, compute on the context's stream,
// D2H on s_out, software-pipelined over groups of streams (PCIe is full duplex; the copies would
// otherwise serialise with the kernels).  Returns without waiting; ev_d2h[slot] marks completion.
// Hazards across in-flight batches (kBatchSlots = 3):
//   H2D into d_bpcm[slot]  waits for the compute that last read it        (ev_cdone[slot])
//   compute into d_bout[slot] waits for the D2H that last read it         (ev_d2h[slot])
// everything else (spectra, activations, masks) lives on the context's stream and is ordered by it.
// channels != nullptr selects the interleaved formats: pcmL[i] = n_samples[i] frames of channels[i] interleaved floats
// (pcmR unused), stems_out[i * pairs + q] = n_samples[i] interleaved stereo frames.  The staging layout is the same
// (2 * np floats of PCM and 2 * np floats per output pair and stream), so each stream is one DMA per direction and pair.
static int batch_enqueue(srt_ctx* c, int slot, const float* const* pcmL, const float* const* pcmR, const size_t* n_samples,
                         int n_streams, const float* unaffected, float* const* stems_out, const int* channels = nullptr)
{
    const int P = c->pairs();
    if (!c->s_in) {
        CK(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 8; i++) {
            CK(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c->ev_c[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < kBatchSlots; i++) {
            CK(cudaEventCreateWithFlags(&c->ev_cdone[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c->ev_d2h[i], cudaEventDisableTiming));
        }
    }
    size_t tot = 0;
    for (int i = 0; i < n_streams; i++) tot += (n_samples[i] + 3) & ~(size_t)3;
    if (tot * 2 > c->bpcm_cap[slot] || tot * 2 * P > c->bout_cap[slot]) {
        // grow every slot at once, so only the first call of a new size pays for allocation
        for (int i = 0; i < kBatchSlots; i++) {
            int r = batch_wait_slot(c, i);
            if (r) return r;
        }
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaStreamSynchronize(c->s_in));
        CK(cudaStreamSynchronize(c->s_out));
        for (int i = 0; i < kBatchSlots; i++) {
            if (tot * 2 > c->bpcm_cap[i]) {
                if (c->d_bpcm[i]) cudaFree(c->d_bpcm[i]);
                c->d_bpcm[i] = nullptr; c->bpcm_cap[i] = 0;
                CK(cudaMalloc((void**)&c->d_bpcm[i], tot * 2 * 4));
                c->bpcm_cap[i] = tot * 2;
            }
            if (tot * 2 * P > c->bout_cap[i]) {
                if (c->d_bout[i]) cudaFree(c->d_bout[i]);
                c->d_bout[i] = nullptr; c->bout_cap[i] = 0;
                CK(cudaMalloc((void**)&c->d_bout[i], tot * 2 * P * 4));
                c->bout_cap[i] = tot * 2 * P;
            }
        }
    }
    float* const d_pcm = c->d_bpcm[slot];
    float* const d_out = c->d_bout[slot];
    std::vector<const float*> dl(n_streams), dr(n_streams);
    std::vector<float*> dout((size_t)n_streams * P * 2);
    size_t off = 0;
    for (int i = 0; i < n_streams; i++) {
        const size_t np = (n_samples[i] + 3) & ~(size_t)3;
        dl[i] = d_pcm + off * 2;
        dr[i] = channels ? dl[i] + (channels[i] == 2 ? 1 : 0) : d_pcm + off * 2 + np;
        for (int q = 0; q < P * 2; q++)
            dout[(size_t)i * P * 2 + q] = channels ? d_out + off * 2 * P + (size_t)(q >> 1) * 2 * np + (q & 1) : d_out + off * 2 * P + (size_t)q * np;
        off += np;
    }
    // Groups shorten the latency of a lone call (the first download starts after 1/groups of the kernels) but
    // cost batch efficiency.  Once other batches are in flight most of the overlap comes from them; two groups still
    // pay there when a step's kernels take about as long as its download (compensated precision: 8.3 ms of kernels
    // against 7.9-8.7 ms of D2H): the copy engine then always has half a batch queued and never waits for the end of
    // a whole step.  Measured per 32-stream step (r2, ceiling 57 GB/s = 7.9 ms): blocking 17.9 / 13.4 / 11.6 ms and
    // pipelined 9.3-9.6 / 9.0 / 9.4 ms for 1 / 2 / 4 groups.
    bool others_in_flight = false;
    for (int i = 0; i < kBatchSlots; i++)
        if (i != slot && c->slot_busy[i] && cudaEventQuery(c->ev_d2h[i]) == cudaErrorNotReady) others_in_flight = true;
    (void)cudaGetLastError();   // cudaErrorNotReady is not sticky, but keep the error state clean
    const char* ge = getenv("SRT_E2E_GROUPS");
    int groups = ge ? atoi(ge) : (others_in_flight ? 2 : 4);
    groups = std::max(1, std::min(std::min(groups, 8), n_streams));
    const int per = (n_streams + groups - 1) / groups;
    CK(cudaStreamWaitEvent(c->s_in, c->ev_cdone[slot], 0));     // no-op until the slot has been used once
    CK(cudaStreamWaitEvent(c->stream, c->ev_d2h[slot], 0));
    // Copies whose source and destination both continue the previous one are merged: a caller that keeps its
    // streams (and stems) back to back in one pinned allocation gets one DMA per group instead of one per
    // channel (55 vs 49.5 GB/s D2H on the bench box, profiles/r1_pcie_probe.json).
    struct Run {
        char* dst; const char* src; size_t bytes;
    };
    auto flush = [](std::vector<Run>& runs, cudaMemcpyKind kind, cudaStream_t st) -> cudaError_t {
        for (const Run& r : runs) {
            cudaError_t e = cudaMemcpyAsync(r.dst, r.src, r.bytes, kind, st);
            if (e != cudaSuccess) return e;
        }
        runs.clear();
        return cudaSuccess;
    };
    auto add = [](std::vector<Run>& runs, void* dst, const void* src, size_t bytes) {
        if (!runs.empty() && runs.back().dst + runs.back().bytes == (char*)dst && runs.back().src + runs.back().bytes == (const char*)src)
            runs.back().bytes += bytes;
        else
            runs.push_back(Run{(char*)dst, (const char*)src, bytes});
    };
    std::vector<Run> runs;
    for (int g = 0; g < groups; g++) {
        const int i0 = g * per, i1 = std::min(n_streams, i0 + per);
        for (int i = i0; i < i1; i++) {
            if (channels) {
                add(runs, (void*)dl[i], pcmL[i], n_samples[i] * 4 * (size_t)channels[i]);
            } else {
                add(runs, (void*)dl[i], pcmL[i], n_samples[i] * 4);
                add(runs, (void*)dr[i], pcmR[i], n_samples[i] * 4);
            }
        }
        CK(flush(runs, cudaMemcpyHostToDevice, c->s_in));
        CK(cudaEventRecord(c->ev_in[g], c->s_in));
    }
    for (int g = 0; g < groups; g++) {
        const int i0 = g * per, i1 = std::min(n_streams, i0 + per);
        if (i0 >= i1) break;
        CK(cudaStreamWaitEvent(c->stream, c->ev_in[g], 0));
        int r = separate_core(c, dl.data() + i0, dr.data() + i0, n_samples + i0, i1 - i0, unaffected, dout.data() + (size_t)i0 * P * 2, kFFT,
                              channels ? channels + i0 : nullptr, channels ? 2 : 1);
        if (r) return r;
        CK(cudaEventRecord(c->ev_c[g], c->stream));
        CK(cudaStreamWaitEvent(c->s_out, c->ev_c[g], 0));
        for (int i = i0; i < i1; i++) {
            if (channels) {
                for (int q = 0; q < P; q++) add(runs, stems_out[(size_t)i * P + q], dout[((size_t)i * P + q) * 2], n_samples[i] * 8);
            } else {
                for (int q = 0; q < P * 2; q++) add(runs, stems_out[(size_t)i * P * 2 + q], dout[(size_t)i * P * 2 + q], n_samples[i] * 4);
            }
        }
        CK(flush(runs, cudaMemcpyDeviceToHost, c->s_out));
    }
    CK(cudaEventRecord(c->ev_cdone[slot], c->stream));
    CK(cudaEventRecord(c->ev_d2h[slot], c->s_out));
    c->slot_busy[slot] = true;
    return 0;
}

static int batch_wait_slot(srt_ctx* c, int slot)
{
    if (!c->slot_busy[slot]) return 0;
    CK(cudaEventSynchronize(c->ev_d2h[slot]));
    c->slot_busy[slot] = false;
    return 0;
}

static int batch_submit(srt_ctx* c, const float* const* pcmL, const float* const* pcmR, const size_t* n_samples, int n_streams,
                        const float* unaffected, float* const* stems_out, int* ticket_out, const int* channels)
{
    if (!c || c->S == 0) return fail(SRT_ERR_STATE, "context has no nets");
    if (n_streams < 1) return fail(SRT_ERR_ARG, "n_streams < 1");
    if (!ticket_out) return fail(SRT_ERR_ARG, "ticket_out is NULL");
    internal::DeviceGuard dev_guard(c->cfg.device);
    if (!dev_guard.ok) return fail(SRT_ERR_CUDA, "cudaSetDevice(%d) failed", c->cfg.device);
    reset_spans(c);
    const int slot = (int)(c->batch_seq % kBatchSlots);
    int r = batch_wait_slot(c, slot);   // one batch too many in flight: the slot's previous owner must have drained
    if (r) return r;
    if ((r = batch_enqueue(c, slot, pcmL, pcmR, n_samples, n_streams, unaffected, stems_out, channels))) {
        // leave nothing half-enqueued behind an error return
        cudaStreamSynchronize(c->s_in); cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->s_out);
        return r;
    }
    const int ticket = (int)(c->batch_seq % kTicketMod);   // kTicketMod is a multiple of kBatchSlots: ticket % slots == slot
    c->slot_ticket[slot] = ticket;
    c->batch_seq++;
    *ticket_out = ticket;
    return 0;
}

extern "C" int srt_separate_batch_async(srt_ctx* c, const float* const* pcmL, const float* const* pcmR, const size_t* n_samples,
                                        int n_streams, const float* unaffected, float* const* stems_out, int* ticket_out)
{
    return batch_submit(c, pcmL, pcmR, n_samples, n_streams, unaffected, stems_out, ticket_out, nullptr);
}

extern "C" int srt_separate_batch_interleaved_async(srt_ctx* c, const float* const* pcm, const int* channels, const size_t* n_samples,
                                                    int n_streams, const float* unaffected, float* const* out, int* ticket_out)
{
    int r = check_channels(channels, n_streams < 0 ? 0 : n_streams);
    if (r) return r;
    return batch_submit(c, pcm, nullptr, n_samples, n_streams, unaffected, out, ticket_out, channels);
}

extern "C" int srt_batch_wait(srt_ctx* c, int ticket)
{
    if (!c) return fail(SRT_ERR_STATE, "null context");
    const long long age = ((c->batch_seq % kTicketMod) - 1 - ticket + kTicketMod) % kTicketMod;   // 0 = newest batch
    if (ticket < 0 || ticket >= kTicketMod || c->batch_seq == 0 || age >= c->batch_seq) return fail(SRT_ERR_ARG, "bad ticket %d", ticket);
    internal::DeviceGuard dev_guard(c->cfg.device);
    if (!dev_guard.ok) return fail(SRT_ERR_CUDA, "cudaSetDevice(%d) failed", c->cfg.device);
    const int slot = ticket % kBatchSlots;
    if (c->slot_ticket[slot] != ticket) return 0;   // an older batch: drained when its slot was handed on
    return batch_wait_slot(c, slot);
}

extern "C" int srt_separate_batch(srt_ctx* c, const float* const* pcmL, const float* const* pcmR, const size_t* n_samples,
                                  int n_streams, const float* unaffected, float* const* stems_out)
{
    int ticket = -1;
    int r = srt_separate_batch_async(c, pcmL, pcmR, n_samples, n_streams, unaffected, stems_out, &ticket);
    if (r) return r;
    if ((r = srt_batch_wait(c, ticket))) return r;
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int srt_separate_batch_interleaved(srt_ctx* c, const float* const* pcm, const int* channels, const size_t* n_samples,
                                              int n_streams, const float* unaffected, float* const* out)
{
    int ticket = -1;
    int r = srt_separate_batch_interleaved_async(c, pcm, channels, n_samples, n_streams, unaffected, out, &ticket);
    if (r) return r;
    if ((r = srt_batch_wait(c, ticket))) return r;
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------
// transforms for the tier-A shims (stft()/istft(), stftFix.c:363-579)
// ------------------------------------------------------------------------------------------
extern "C" size_t srt_stft_rows(size_t n) { return (n + kHop - 1) / kHop; }

extern "C" int srt_stft_host(srt_ctx* c, const float* L, const float* R, size_t n, float* reL, float* imL, float* reR, float* imR)
{
    if (!c) return fail(SRT_ERR_STATE, "null context");
    if (n < (size_t)kFFT) return fail(SRT_ERR_ARG, "stft needs at least 4096 samples");
    internal::DeviceGuard dev_guard(c->cfg.device);
    if (!dev_guard.ok) return fail(SRT_ERR_CUDA, "cudaSetDevice(%d) failed", c->cfg.device);
    reset_spans(c);
    const size_t rows = srt_stft_rows(n);
    const int computed = (int)((n - kFFT + kHop / 4) / kHop) + 1;   // stftFix.c:377 + the final frame (:460-493)
    const int cap = c->NB * c->T;                                    // rows the spectrum buffer holds
    std::vector<float4> hs((size_t)std::min<size_t>(cap, rows) * kBins);
    // device staging for the PCM
    const size_t np = (n + 3) & ~(size_t)3;
    if (np * 2 > c->pcm_cap) {
        if (c->d_pcm) cudaFree(c->d_pcm);
        c->pcm_cap = np * 2;
        CK(cudaMalloc((void**)&c->d_pcm, c->pcm_cap * 4));
    }
    CK(cudaMemcpyAsync(c->d_pcm, L, n * 4, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->d_pcm + np, R, n * 4, cudaMemcpyHostToDevice, c->stream));
    for (size_t q = 0; q < rows * kFFT; q++) reL[q] = imL[q] = reR[q] = imR[q] = 0.0f;
    // process in slabs of `cap` rows, each row its own single-frame "image"
    for (size_t r0 = 0; r0 < (size_t)computed; r0 += cap) {
        const int nr = (int)std::min<size_t>(cap, computed - r0);
        const size_t o_pl = 0, o_pr = 8, o_n = 16, o_nfr = 20, o_img = 24, total_b = o_img + sizeof(ImgDesc) * nr;
        int rr = ensure_meta(c, total_b);
        if (rr) return rr;
        const float* pl = c->d_pcm;
        const float* pr = c->d_pcm + np;
        const int ni = (int)n, nf = computed;
        std::memcpy(c->h_meta + o_pl, &pl, 8);
        std::memcpy(c->h_meta + o_pr, &pr, 8);
        std::memcpy(c->h_meta + o_n, &ni, 4);
        std::memcpy(c->h_meta + o_nfr, &nf, 4);
        ImgDesc* im = (ImgDesc*)(c->h_meta + o_img);
        for (int i = 0; i < nr; i++) im[i] = ImgDesc{0, (int)r0 + i};
        if ((rr = commit_meta(c, total_b))) return rr;
        StftParams p{};
        p.pcmL = (const float* const*)(c->d_meta + o_pl);
        p.pcmR = (const float* const*)(c->d_meta + o_pr);
        p.n_samples = (const int*)(c->d_meta + o_n);
        p.n_frames = (const int*)(c->d_meta + o_nfr);
        p.imgs = (const ImgDesc*)(c->d_meta + o_img);
        p.window = c->d_window; p.twiddle = c->d_twiddle; p.spec = c->d_spec; p.mag = nullptr;
        p.T = 1; p.F = 0; p.n_img = nr; p.front_pad = 0;
        launch_stft(p, c->stream);
        c->launches++;
        CK(cudaMemcpyAsync(hs.data(), c->d_spec, (size_t)nr * kBins * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < nr; i++)
            for (int k = 0; k < kBins; k++) {
                const float4 v = hs[(size_t)i * kBins + k];
                const size_t q = (r0 + i) * kFFT + k;
                reL[q] = v.x; imL[q] = v.y; reR[q] = v.z; imR[q] = v.w;
            }
    }
    return 0;
}

extern "C" int srt_istft_host(srt_ctx* c, const float* reL, const float* imL, const float* reR, const float* imR, size_t frames,
                              float* outL, float* outR)
{
    if (!c) return fail(SRT_ERR_STATE, "null context");
    internal::DeviceGuard dev_guard(c->cfg.device);
    if (!dev_guard.ok) return fail(SRT_ERR_CUDA, "cudaSetDevice(%d) failed", c->cfg.device);
    reset_spans(c);
    const size_t out_n = frames * kHop + (kFFT - kHop);
    const size_t cap = (size_t)c->B * c->T;   // frames the scratch holds
    if (frames > cap || frames > (size_t)c->NB * c->T) return fail(SRT_ERR_CAPACITY, "istft of %zu frames exceeds context capacity %zu", frames, cap);
    std::vector<float4> hs(frames * kBins);
    for (size_t f = 0; f < frames; f++)
        for (int k = 0; k < kBins; k++) hs[f * kBins + k] = make_float4(reL[f * kFFT + k], imL[f * kFFT + k], reR[f * kFFT + k], imR[f * kFFT + k]);
    if (out_n * 2 > c->out_cap) {
        if (c->d_out) cudaFree(c->d_out);
        c->out_cap = out_n * 2;
        CK(cudaMalloc((void**)&c->d_out, c->out_cap * 4));
    }
    const size_t o_out = 0, o_n = 16, o_nfr = 20, o_i0 = 24, o_img = 32, total_b = o_img + sizeof(ImgDesc) * frames;
    int r = ensure_meta(c, total_b);
    if (r) return r;
    float* po[2] = {c->d_out, c->d_out + out_n};
    const int ni = (int)out_n, nf = (int)frames, i0 = 0;
    std::memcpy(c->h_meta + o_out, po, 16);
    std::memcpy(c->h_meta + o_n, &ni, 4);
    std::memcpy(c->h_meta + o_nfr, &nf, 4);
    std::memcpy(c->h_meta + o_i0, &i0, 4);
    ImgDesc* im = (ImgDesc*)(c->h_meta + o_img);
    for (size_t i = 0; i < frames; i++) im[i] = ImgDesc{0, (int)i};
    if ((r = commit_meta(c, total_b))) return r;
    CK(cudaMemcpyAsync(c->d_spec, hs.data(), hs.size() * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    IstftParams p{};
    p.spec = c->d_spec; p.mask = nullptr; p.imgs = (const ImgDesc*)(c->d_meta + o_img); p.n_frames = (const int*)(c->d_meta + o_nfr);
    p.postwin = c->d_postwin; p.twiddle = c->d_twiddle; p.frames_out = c->d_frames;
    p.T = 1; p.F = 0; p.S = 1; p.img_first = 0; p.n_img = (int)frames; p.mask_stem_stride = 0; p.frames_stem_stride = (int)cap;
    launch_istft(p, c->stream);
    OlaParams q{};
    q.frames = c->d_frames; q.stream_img0 = (const int*)(c->d_meta + o_i0); q.n_frames = (const int*)(c->d_meta + o_nfr);
    q.n_samples = (const int*)(c->d_meta + o_n); q.out = (float* const*)(c->d_meta + o_out);
    q.T = 1; q.S = 1; q.stream_first = 0; q.n_streams = 1; q.img_first = 0; q.frames_stem_stride = (int)cap; q.max_samples = (int)out_n; q.front_pad = 0;
    launch_ola(q, c->stream);
    c->launches += 2;
    CK(cudaMemcpyAsync(outL, c->d_out, out_n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(outR, c->d_out + out_n, out_n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------
// introspection
// ------------------------------------------------------------------------------------------
extern "C" long long srt_launch_count(const srt_ctx* c) { return c ? c->launches + (c->next ? c->next->launches : 0) : 0; }
extern "C" int srt_output_pairs(const srt_ctx* c) { return c ? c->pairs() : 0; }
extern "C" int srt_set_timing(srt_ctx* c, int enable)
{
    if (!c) return SRT_ERR_STATE;
    cudaStreamSynchronize(c->stream);
    c->timing = false;
    reset_spans(c);
    c->timing = enable != 0;
    if (c->next) srt_set_timing(c->next, enable);
    return 0;
}
extern "C" int srt_last_timing(const srt_ctx* c, int which, float* ms_out)
{
    if (!c || !ms_out) return SRT_ERR_ARG;
    cudaStreamSynchronize(c->stream);
    float tot = 0.f;
    for (const Span& s : c->spans)
        if (s.cat == which) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) tot += ms;
        }
    if (c->next) {   // second stage of the cascade
        float more = 0.f;
        srt_last_timing(c->next, which, &more);
        tot += more;
    }
    *ms_out = tot;
    return 0;
}
extern "C" int srt_synchronize(srt_ctx* c)
{
    if (!c) return SRT_ERR_STATE;
    CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < kBatchSlots; i++) {
        int r = batch_wait_slot(c, i);
        if (r) return r;
    }
    return 0;
}
extern "C" void* srt_cuda_stream(srt_ctx* c) { return c ? (void*)c->stream : nullptr; }
extern "C" void* srt_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void srt_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" long long srt_debug_tensor(srt_ctx* c, const char* name, float* dst, size_t max_floats)
{
    if (!c || !name || c->S == 0) return fail(SRT_ERR_STATE, "no context");
    const float* src = nullptr;
    int level = 0, ch = 0;
    int idx = name[strlen(name) - 1] - '0';
    if (!strncmp(name, "skip", 4) && idx >= 1 && idx <= 6) { src = c->E[idx]; level = idx; ch = kEnc[idx]; }
    else if (!strncmp(name, "up", 2) && idx >= 1 && idx <= 5) { src = c->U[idx]; level = 6 - idx; ch = kDecOut[idx - 1]; }
    else if (!strcmp(name, "up6")) { src = c->U[6]; level = 0; ch = 1; }
    else return fail(SRT_ERR_ARG, "unknown tensor %s", name);
    const int H = c->T >> level, W = c->F >> level, Bv = c->last_Bv;
    const size_t per = (size_t)H * W * ch, need = per * c->S * Bv;
    if (need > max_floats) return fail(SRT_ERR_CAPACITY, "buffer too small");
    std::vector<float> h(per);
    CK(cudaStreamSynchronize(c->stream));
    for (int s = 0; s < c->S; s++)
        for (int b = 0; b < Bv; b++) {
            CK(cudaMemcpy(h.data(), src + ((size_t)s * c->B + b) * per, per * 4, cudaMemcpyDeviceToHost));
            float* d = dst + ((size_t)s * Bv + b) * per;
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++)
                    for (int k = 0; k < ch; k++) d[((size_t)k * H + y) * W + x] = h[((size_t)y * W + x) * ch + k];
        }
    return (long long)need;
}
