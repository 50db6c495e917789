// srt_kernels.cuh — parameter blocks shared by the host orchestration and the CUDA kernels.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "srt_plan.h"

#include <atomic>

namespace srt {

// Per-device launch state of one kernel: the opt-in dynamic shared-memory size (an attribute of the function ON A
// DEVICE, so a process that holds contexts on several GPUs must set it on each) and the SM count for persistent grids.
// One static instance per launcher; safe for concurrent host threads (tier A runs several instances from several
// threads, main.c:330, 592): the worst case is two threads setting the same attribute.
struct LaunchState {
    static constexpr int kMaxDevices = 64;
    std::atomic<int> smem[kMaxDevices];
    std::atomic<int> sms[kMaxDevices];
    // makes `bytes` of dynamic shared memory launchable for `kernel` on the current device; returns its SM count
    template <class K>
    int prepare(K kernel, size_t bytes)
    {
        int dev = 0;
        cudaGetDevice(&dev);
        const int d = (dev >= 0 && dev < kMaxDevices) ? dev : 0;
        if (smem[d].load(std::memory_order_acquire) < (int)bytes || d != dev) {
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            smem[d].store((int)bytes, std::memory_order_release);
        }
        int n = sms[d].load(std::memory_order_acquire);
        if (n == 0 || d != dev) {
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
            sms[d].store(n, std::memory_order_release);
        }
        return n;
    }
};

// ---------------------------------------------------------------------------------------
// Gather-GEMM layer launch parameters (tcgen05 kernel and SIMT verification kernel).
// Image index n runs over S stems x B images, stem-major: n = s*B + b.
// ---------------------------------------------------------------------------------------
struct alignas(64) ConvParams {
    CUtensorMap tmap[3];          // source activation tensors, dims {C, W, H, S*B}, box {32, tw, th, nb}, SW128; [2] = the bf16 residual
                                  // tensor of compensated layers (box {64, tw, th, nb}: the same 128-byte rows)
    const void* lo_ptr;           // the residual tensor as a raw pointer (SIMT verification path), lo_C channels per pixel,
    int lo_C;                     // bf16 or (lo_fp8) e5m2 bytes
    int lo_fp8;
    const float* src_ptr[2];      // same tensors as raw pointers (SIMT verification path)
    int src_C[2];
    const KBlock* kb;             // all phases back to back
    int kb_off[4], nkb[4];
    const float* w;               // packed weights, stem 0
    size_t w_stem_stride;         // floats between stems
    size_t w_phase_off[4];
    int n_tile, n_tiles, phases;
    int fused;                    // decoder with the four output parities fused into N (LayerPlan::fused): column = parity * cout + channel
    int Hs, Ws;                   // tile-space extent
    int B;                        // images per stem in the buffer layout (capacity)
    int Bv;                       // images per stem actually present in this launch
    int S;                        // stems
    int tw, th, nb;
    int tiles_x, tiles_y, tiles_n;
    int cout;
    // epilogue (Executable/spleeter.c:182-190 encoder, :240-247 decoder)
    const float* bias;            // [S][cout]
    const float* bn_scale;        // [S][cout]
    const float* bn_offset;       // [S][cout]
    int act[8];                   // srt::Act per stem (stemMode, spleeter.c:130-139)
    int mode;                     // 0 encoder (raw + activated), 1 encoder last (raw only), 2 decoder
    float* out_raw;               // encoder: NHWC [n][Hs][Ws][cout]   (conv + bias, the skip)
    float* out_act;               // encoder: S2D  [n][Hs/2][Ws/2][4*cout]  act(scale*v+offset)
    float* out_dec;               // decoder: NHWC [n][2Hs][2Ws][cout]  scale*act(v)+offset
    int round_raw, round_act;     // round stored values to TF32 (consumer is a tensor-core layer)
    // compensated precision: next to every TF32-rounded value hi = tf32(v) the epilogue stores bf16(v - hi) for the
    // consumer's compensation k-blocks (srt_plan.h build_plans).  nullptr = the consumer runs single-pass TF32.
    // Each residual tensor is in its CONSUMER's format: bf16, or e5m2 bytes holding 4 (v - hi) (srt_plan.h kPartLo8).
    void* lo_raw;                 // residual of out_raw:  [n][Hs][Ws][lo_raw_C], this layer's channels at lo_raw_coff
    void* lo_act;                 // residual of out_act:  same space-to-depth layout as out_act
    void* lo_dec;                 // residual of out_dec:  [n][2Hs][2Ws][lo_dec_C], this layer's channels at lo_dec_coff
    int lo_raw_C, lo_raw_coff, lo_dec_C, lo_dec_coff;
    int lo_raw_fp8, lo_act_fp8, lo_dec_fp8;
};

// Row-patch tensor-core kernel (srt_conv_rp.cu): small-N layers, see srt_plan.h RowPlan.
struct alignas(64) RowConvParams {
    CUtensorMap tmap[3];          // dims {C, W, H, S*B}, box {32, kPatchW, R+2, 1}, SW128; [2] = bf16 residual tensor, box {64, kPatchW, R+2, 1}
    const RowChunk* chunks;
    int n_chunks;
    const KBlock* kb;
    int nkb;
    const float* w;               // packed weights [stem][k-block][N][32]
    size_t w_stem_stride;
    int N, R;
    int tiles_x, tiles_y;
    int bo_mode;                  // (diagnostic) descriptor base-offset convention, 0 = product setting
    int dbg;                      // (diagnostic, SRT_RP_DBG) 1: skip epilogue math+stores, 2: skip MMAs, 4: skip patch loads, 8: skip weight loads
    int kb_width;                 // channels per k-block: 32 (SWIZZLE_128B rows) or 8 (down1: SWIZZLE_32B rows)
    int src_img0;                 // image coordinate offset of the source tensor (down1 reads the shared magnitude buffer)
    int stems_per_tile;           // 1, or (down1) stems fused into N: column c -> stem stem0 + c / cout
    int stem0;
    ConvParams ep;                // geometry + epilogue (tmap / k-block fields unused)
};
void launch_conv_rp(const RowConvParams& p, cudaStream_t st);
bool conv_rp_fits(int n_chunks, int nkb);   // table sizes the kernel's shared-memory header holds

// ---------------------------------------------------------------------------------------
// SIMT edge layers
// ---------------------------------------------------------------------------------------
struct Down1Params {              // 5x5 s2 conv 2->16 on the magnitude image, spleeter.c:181-190
    const float* mag;             // space-to-depth magnitudes of the launch's first image: [Bv][T/2][F/2][(py,px)][c], TF32 "hi" part
    const float* mag_lo;          // the matching "lo" part (mag = hi + lo exactly)
    const float* w;               // [S][16][2][5][5] (reference order)
    const float* bias;            // [S][16]
    const float* bn_scale;        // [S][16]
    const float* bn_offset;       // [S][16]
    float* out_raw;               // [S*B][T/2][F/2][16]
    float* out_act;               // S2D [S*B][T/4][F/4][64], TF32-rounded
    uint16_t* lo_act;             // bf16 residual of out_act (compensated down2: its 64 residual channels are always bf16) or nullptr
    int T, F, B, Bv, S;
    int act[8];
    int stem;                     // one launch per stem: weights / bias / BN ride in the constant bank
    float wk[800];                // [tap][cin][cout] of that stem
    float bk[48];                 // bias[16], bn scale[16], bn offset[16]
};

struct Up6Params {                // 5x5 s2 transposed conv 32->1 + act + BN, spleeter.c:289-294
    const float* skip;            // conv1 raw  [S*B][T/2][F/2][16]
    const float* up;              // up5 output [S*B][T/2][F/2][16]
    const float* w;               // [S][32][25]  (cin, kh*5+kw)
    const float* bias;            // [S]
    const float* bn_scale;        // [S]
    const float* bn_offset;       // [S]
    float* out;                   // [S*B][T][F]
    int T, F, B, Bv, S;
    int act[8];
    int stem;                     // the launch covers one stem: its weights ride in the parameter (constant) bank
    float wk[800];                // [32][25] of that stem: FFMA reads them as immediates, no shared-memory traffic
};

// Tensor-core up6 (srt_up6_tc.cu).  Weights: per stem [box = src*2 + channel half][term: 0 = tf32(w), 1 = tf32(w - tf32(w))]
// [32 rows = taps (25 used)][8 channels], SWIZZLE_32B pre-swizzled (swz32_index): 8 blocks of 256 floats.
constexpr int kUp6TcWFloatsPerStem = 9 * 256;   // 8 TF32 blocks + one [32 taps][32 channels] e5m2(w / 4) block (SWIZZLE_32B) for the 8-bit residual term
struct Up6TcParams {
    CUtensorMap tmap[2];          // skip1, up5: {16, F/2, T/2, S*B}, box {8, 128, 1, 1}, SWIZZLE_32B
    const float* w;               // [S][kUp6TcWFloatsPerStem]
    float* out;                   // [S*B][T][F]
    int T, F, B, Bv, S;
    int blocks_x, bw;             // column blocks per row and output pixels (input resolution) per block (<= 126)
    int chunks, rows_per_unit;    // row chunks per image, input rows per chunk (even)
    int w_terms;                  // 1: weights TF32-exact; 2: also the residual term
    int prefetch_rows;            // L2 prefetch distance in rows (0 = off)
    int stages, acc_slots;        // ring depths in use (input rows: see `pair`; <= 8 TMEM accumulators, an even number)
    int pair;                     // epilogue loop over row pairs (6 G rows, <= 6 / 4 input rows in flight) instead of single rows (4 G rows, <= 8 / 5)
    int lo8;                      // residual term in 8 bits: the split warps write e5m2(4 (a - trunc(a))) as ONE [128 px][32 ch] tile per row
                                  // (4 KB instead of 16 KB) that a single K = 32 MMA contracts; 0 = fp32 residual tile, four TF32 MMAs
    int dbg;                      // SRT_UP6_DBG stage-skip bits (1 gather, 2 MMA, 4 TMA, 8 split): timing experiments only
    float bias[8], bn_scale[8], bn_offset[8];
    int act[8];
};
void launch_up6_tc(const Up6TcParams& p, cudaStream_t st);
bool up6_tc_fits(int S);

struct Up7Params {                // 4x4 dilation-2 conv 1->2 + bias + sigmoid, spleeter.c:295-300
    const float* in;              // [S*B][T][F]
    const float* w;               // [S][2][16]
    const float* bias;            // [S][2]
    const float* lut;             // sigmoid table, 1025 x {t0, slope, x1, 0} (Executable flavour) or nullptr (exact)
    float* mask;                  // [S][mask_stem_stride images][T][F][2], first image of this launch = mask_img0
    int T, F, B, Bv, S;
    int mask_stem_stride, mask_img0;
    int stem;                     // one launch per stem
    float wk[36];                 // that stem's 32 weights + 2 biases (constant bank)
    int merged;                   // 1: one launch covers all S stems (blockIdx.z = stem * Bv + image), weights from wk_all[stem]:
    float wk_all[8][36];          // small batches, where four per-stem grids of a few dozen CTAs would run one after the other
};

// ---------------------------------------------------------------------------------------
// Transforms
// ---------------------------------------------------------------------------------------
struct ImgDesc {                  // one T-frame tile of one stream
    int stream;
    int f0;                       // first frame of the tile within the stream
};

struct StftParams {
    const float* const* pcmL;     // per stream device pointers (unpadded samples)
    const float* const* pcmR;
    const int* pcm_stride;        // per stream: floats between consecutive samples of a channel (1 planar, 2 interleaved stereo,
                                  // 1 with pcmR == pcmL for mono, main.c:767-769), or nullptr = 1
    const int* n_samples;         // per stream
    const int* n_frames;          // per stream: padded_len / 1024 (stftFix.c:367)
    const ImgDesc* imgs;          // [n_img]
    const float* window;          // hann(i+1/2)/4096
    const float2* twiddle;        // exp(-2 pi i m / 4096)
    float4* spec;                 // [n_img][T][2049] (reL, imL, reR, imR) in the reference's convention
    float* mag;                   // [n_img][T/2][F/2][(py,px)][c] space-to-depth, split in two tensors: hi = tf32(mag) here,
    size_t mag_lo_off;            // lo = mag - hi at mag + mag_lo_off floats (down1 contracts both: fp32-accurate first layer)
    int T, F, n_img;
    int front_pad;                // 4096 zeros in front (main.c:767) or 0 (raw stft())
};

struct IstftParams {
    const float4* spec;           // [all images][T][2049]
    const float* mask;            // [S][mask_stem_stride images][T][F][2] or nullptr (no masking)
    const ImgDesc* imgs;          // [all images]
    const int* n_frames;          // per stream
    const float* postwin;         // (2/3) hann(i+1/2)
    const float2* twiddle;
    float2* frames_out;           // scratch [S][frames_stem_stride images][T][4096] (l, r) windowed time frames
    float unaffected[8];          // per stem weight for bins >= F (main.c:486-493, Spleeter4Stems.c:73,281)
    int T, F, S;
    int img_first, n_img;         // images [img_first, img_first + n_img) are processed; scratch index is relative
    int mask_stem_stride, frames_stem_stride;
};

struct OlaParams {
    const float2* frames;         // scratch, see IstftParams
    const int* stream_img0;       // per stream: first image index (global)
    const int* n_frames;          // per stream
    const int* n_samples;         // per stream
    float* const* out;            // [stream][S*2] device pointers to planar outputs of n_samples
    int T, S;
    int stream_first, n_streams;  // streams [stream_first, stream_first + n_streams)
    int img_first;                // first image of the scratch
    int frames_stem_stride;
    int max_samples;
    int front_pad;
};

// mask*spectrum -> inverse FFT -> window -> overlap-add fused: one CTA walks G consecutive hops of one
// (stream, stem) and keeps the 4-frame overlap in shared memory (no scratch frames, no second kernel)
struct IstftOlaParams {
    const float4* spec;           // [all images][T][2049]
    const float* mask;            // [S][mask_stem_stride images][T][F][2]
    const int* stream_img0;       // per stream: first image index
    const int* n_frames;          // per stream
    const int* n_samples;         // per stream
    const float* postwin;
    const float2* twiddle;
    float* const* out;            // [stream][S*2] planar outputs
    float unaffected[8];
    int T, F, S;                  // S = inverse transforms per frame (grid.y)
    int S_masked;                 // transforms s < S_masked apply stem s's mask; the others are the plain inverse transform (main.c:881)
    int out_pairs;                // (L, R) pointer pairs per stream in `out` (>= S; the CLI modes keep extra pairs)
    int pair_first;               // transform s writes pair pair_first + s
    int out_stride;               // 1 = planar outputs, 2 = interleaved stereo frames (R pointer = L pointer + 1), main.c:806
    int mask_stem_stride;
    int stream_first;
    int front_pad;
    int hops_per_cta;
};
void launch_istft_ola(const IstftOlaParams& p, int n_streams, int max_frames, cudaStream_t st);

// Second stage of the CLI's 3-output cascade (main.c:849-865, 911): residual spectrum = spec - mask*spec of the first
// net (bins >= F: spec - unaffected*spec), and the magnitudes of that residual for the second net.
struct ResidualParams {
    const float4* spec_in;        // first stage: [n_img][T][2049]
    const float* mask_in;         // first stage, stem 0: [n_img][T][F][2]
    const ImgDesc* imgs;
    const int* n_frames;          // per stream
    float4* spec_out;             // [n_img][T][2049]
    float* mag;                   // space-to-depth hi part, lo at + mag_lo_off (as StftParams)
    size_t mag_lo_off;
    float unaffected;
    int T, F, n_img;
};
void launch_residual(const ResidualParams& p, cudaStream_t st);

// Time-domain differences of the CLI's output modes: pair `dst_pair` of every stream becomes a - (pair `sub_pair`),
// with a = the stream's input PCM (2 outputs: accompaniment = input - vocal, main.c:790-794) or pair `dst_pair`
// itself (3 outputs: accompaniment = (accompaniment + vocal) - vocal, main.c:923-927).
struct DiffParams {
    const float* const* pcmL;     // per stream, or nullptr: a = the destination pair
    const float* const* pcmR;
    float* const* out;            // [stream][out_pairs * 2]
    const int* pcm_stride;        // as StftParams
    const int* n_samples;
    int out_stride;               // as IstftOlaParams
    int out_pairs, dst_pair, sub_pair;
    int n_streams, max_samples;
};
void launch_diff(const DiffParams& p, cudaStream_t st);

// launchers (defined in the .cu files)
void launch_conv_tc(const ConvParams& p, cudaStream_t st, int sm_count);
bool conv_tc_supported(int n_tile);     // N tiles the generic kernel is instantiated for
void launch_conv_simt(const ConvParams& p, cudaStream_t st);
void launch_down1(const Down1Params& p, cudaStream_t st);
void launch_up6(const Up6Params& p, cudaStream_t st);
void launch_up7(const Up7Params& p, cudaStream_t st);
void launch_stft(const StftParams& p, cudaStream_t st);
void launch_istft(const IstftParams& p, cudaStream_t st);
void launch_ola(const OlaParams& p, cudaStream_t st);
// [n][T][F][2] (API layout) -> space-to-depth, TF32-rounded
void launch_mag_to_s2d(const float* in, float* out_hi, float* out_lo, int T, int F, int n_img, cudaStream_t st);
size_t conv_tc_smem_bytes(int n_tile, int mt, int* stages_out);

__device__ __forceinline__ float apply_act(int act, float x)
{
    switch (act) {
    case ACT_LEAKY: return x >= 0.0f ? x : 0.2f * x;                          // spleeter.c:43-46
    case ACT_RELU: return x >= 0.0f ? x : 0.0f;                               // spleeter.c:47-50
    case ACT_ELU_CLAMP: return x >= 0.0f ? x : (x < -15.0f ? -1.0f : expf(x) - 1.0f);   // spleeter.c:51-56
    case ACT_ELU: return x >= 0.0f ? x : expf(x) - 1.0f;                      // VST/Source/spleeter.c:74-77
    default: return x;
    }
}

}  // namespace srt
