// srt_plan.h — host-side description of the U-Net as a sequence of "gather-GEMM" layers.
//
// Every 5x5 stride-2 convolution (Executable/spleeter.c:96-100, im2col_dilated.c:10-33) and
// every 5x5 stride-2 transposed convolution (spleeter.c:73-78, im2col_dilated.c:42-65) of
// the reference is re-expressed as   D[pixel, cout] = sum over k-blocks  A_kb[pixel, 32] * W_kb[cout, 32]^T
// where each k-block is one 32-channel slab of the input read at a whole-pixel offset
// (dy, dx) of a *stride-1* window:
//   * encoder layers read their input in space-to-depth form (a 2x2 block of input pixels is
//     one "S2D pixel" with 4*Cin channels), which turns the stride-2 taps into stride-1 taps;
//   * decoder layers are split in the 4 output-pixel parities ("phases"), each of which is a
//     stride-1 convolution with 2 or 3 taps per axis over [skip | up] (never concatenated).
// The same tables drive the tcgen05 kernel, the SIMT verification kernel and the CPU model
// used by the host-logic tests.  Plain C++ (no CUDA) so it can be unit-tested on the CPU.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#if defined(__CUDACC__)
#define SRT_HD __host__ __device__
#else
#define SRT_HD
#endif

namespace srt {

constexpr int kFFT = 4096;
constexpr int kHop = 1024;
constexpr int kBins = 2049;
constexpr int kCoeffFloats = 9822725;   // sizeof(spleeterCoeff)/4, Executable/spleeter.h:5-31
constexpr int kKB = 32;                 // channels per k-block (= one 128-byte swizzle row of fp32)
constexpr int kKBlo = 64;               // channels per compensation k-block (= one 128-byte swizzle row of bf16)
constexpr int kTileM = 128;             // pixels per CTA tile (UMMA M)

enum Act : int { ACT_NONE = 0, ACT_LEAKY = 1, ACT_RELU = 2, ACT_ELU_CLAMP = 3, ACT_ELU = 4 };

// One k-block: which source tensor, which whole-pixel offset, which channel offset.
struct KBlock {
    int8_t src;      // 0 or 1 (decoder: 0 = skip tensor, 1 = up tensor)
    int8_t dy, dx;   // offset in tile-space pixels
    int8_t part;     // bit 0: weights tf32(w) / tf32(w - tf32(w)) (second term for weights that are not TF32-exact)
                     // bit 1 (kPartLo): compensation block - A = the bf16 residual tensor a - tf32(a) of the source (src = 2),
                     //        W = bf16(w), 64 channels wide, contracted with kind::f16 MMAs into the same accumulator
    int32_t c_off;   // channel coordinate in the source tensor (multiple of 32, or of 16 for paired taps; of 64 for compensation blocks)
};
constexpr int kPartLo = 2;
// bits 4-7 of KBlock::part (row-patch plans): K steps of the k-block (quarters of its 128-byte rows) whose weights are all zero -
// an encoder slab that spans several pixel parities carries a tap only in the parities that match the tap's offset - and that
// the MMA issuer therefore skips (down2: 75 of 96 K steps per row are left)
constexpr int kPartSkipShift = 4;
SRT_HD inline int kb_skip_mask(const KBlock& kb) { return ((unsigned char)kb.part >> kPartSkipShift) & 0xf; }
// bit 2 (with bit 1): the compensation block is in the 8-bit format: A = e5m2(4 (a - tf32(a))), W = e5m2(w / 4), 128 channels per
// 128-byte row, kind::f8f6f4 MMAs (K = 32).  Two mantissa bits on either side leave ~6 % of the rounding error (bf16: 0.4 %),
// about what the tensor core's own accumulation adds; half the MMAs and half the bytes of the bf16 form.  The factors 4 and
// 1/4 keep both operands inside e5m2's normal range (residuals of |a| >= 2^-4, weights >= 2^-12); smaller ones fade out.
constexpr int kPartLo8 = 4;
constexpr int kKBlo8 = 128;
// bit 3 (with bits 1, 2; row-patch plans only): the 8-bit block is 64 channels wide - a residual tensor with 64 channels per pixel
// (down2's, up5's) - in 64-byte rows (SWIZZLE_64B), two K = 32 steps.
constexpr int kPartLo8n = 8;
constexpr int kSrcLo = 2;               // KBlock::src / RowChunk::src of the residual tensor
SRT_HD inline int kb_channels(const KBlock& kb) { return (kb.part & kPartLo) ? ((kb.part & kPartLo8n) ? 64 : (kb.part & kPartLo8) ? kKBlo8 : kKBlo) : kKB; }
SRT_HD inline int kb_ksteps(const KBlock& kb) { return (kb.part & kPartLo8n) ? 2 : 4; }     // MMAs (K steps) per k-block and accumulator row
enum LoFormat : int { LO_NONE = 0, LO_BF16 = 1, LO_FP8 = 2, LO_FP8N = 3 };

// For the weight packer: where k-element j of a k-block comes from.
struct KElem {
    int32_t cin;     // input channel in the reference's (concatenated) numbering, -1 = zero
    int8_t kh, kw;
};

struct SrcDesc {     // a source activation tensor as the TMA sees it: [n][H][W][C] fp32
    int C, W, H;
};

struct LayerPlan {
    // identity
    int index;            // 0..4 = down2..down6, 5..9 = up1..up5
    bool transposed;
    int cin, cout;        // reference channel counts (cin = concatenated for the decoder)
    // tile space: encoder = output pixels, decoder = input-resolution pixels (per phase)
    int Hs, Ws;
    int tw, th, nb;       // tile = nb images x th rows x tw cols, product 128
    int n_tile;           // UMMA N per CTA
    int n_tiles;          // cout / n_tile
    int phases;           // 1 or 4
    int nsrc;
    SrcDesc src[2];
    std::vector<KBlock> kb[4];          // per phase
    std::vector<KElem> kelem[4];        // per phase, kb_channels() per k-block, k-block k starts at ke_off[k]
    std::vector<int32_t> ke_off[4];
    // compensated precision (see build_plans): the residual source, bf16, same pixel grid as src[]; encoder layers: the
    // space-to-depth tensor's twin; decoder layers: ONE tensor holding [skip residual | up residual] per pixel
    bool comp;
    int lo_fmt;           // LoFormat of this layer's residual tensor (LO_FP8 needs a channel count that is a multiple of 128)
    SrcDesc lo_src;
    // decoder layers with 4 * cout <= 256: the four output parities fused into N = 4 * cout (column = phase * cout + channel), ONE
    // k-block list over all 3 x 3 input offsets (a parity that does not use an offset gets zero weights).  An activation tile is then
    // fetched once per offset instead of once per (parity, tap): 9 instead of 25 fetches, each feeding a 4x wider MMA - the
    // phase-separated form of up3 (N = 64) took in 24 KB per 4 small MMAs and was bound by the SM's L2 intake at 40 % tensor use.
    bool fused;
    // weights blob: [phase][n_tile_idx][kb][n_tile*32 floats], pre-swizzled (see pack_layer)
    size_t w_floats_per_stem;
    size_t w_phase_off[4];              // float offset of each phase inside the per-stem blob
};

struct NetGeom {
    int T, F;             // image height (time frames) and width (frequency bins), multiples of 64
};

// Offsets (in floats) into one spleeterCoeff blob.
struct CoeffLayout {
    size_t down_w[6], down_b[6], down_bn[6];   // down_bn[5] unused (no BN on down6)
    size_t up_w[6], up_b[6], up_bn[6];
    size_t w7, b7;
};
CoeffLayout coeff_layout();

// Build the 10 tensor-core layers for a T x F image and a batch of n_img images per launch.
// split_weights: every k-block appears twice (part 0 / part 1) so that fp32 weights that are not exactly
// representable in TF32 (the VST's fp32 `.dat` dumps) contribute w = tf32(w) + tf32(w - tf32(w)).
// min_ctas > 0 (with the number of stems sharing a launch): a layer whose grid would hold fewer CTAs than that - the deep
// layers of a one-tile batch are 8-16 CTAs of 80 us each on a 148-SM part - gets narrower N tiles (down to 64 columns:
// below that an MMA does not get cheaper to issue), i.e. more and shorter CTAs.  0 keeps N = min(cout, 256).
// comp_mask bit i: layer i runs in compensated precision.  Its activation operands are rounded to TF32 by the producing
// epilogue (as always), which also stores the residual a - tf32(a) as bf16; the layer then contracts tf32(a) with w (kind::tf32)
// AND the residual with bf16(w) (kind::f16, 64 channels per 128-byte row, i.e. half the MMAs and half the bytes of the main
// term).  Operand error drops from 2^-12 to ~2^-19 relative: fp32-grade results for 1.5x the tensor work.
std::vector<LayerPlan> build_plans(NetGeom g, int n_img, bool split_weights = false, int n_stems = 1, int min_ctas = 0, unsigned comp_mask = 0,
                                   bool fuse_phases = true, int lo_fmt = LO_BF16);
// the residual format layer `index` (0..9) uses when `want` is asked for: LO_FP8 where the residual tensor has a multiple of 128
// channels per pixel; down2's and up5's have 64: LO_FP8N (64-byte rows) in the row-patch kernel, bf16 in the generic one
int layer_lo_format(int index, int want, bool row_patch = false);
// transposed conv: the kernel row that serves output parity `par` at input offset d (o = 2h + kh - 1), or -1
int dec_kh(int par, int d);

// Pack one stem's weights for a layer into the k-block-major, 128B-swizzled layout the MMA
// B operand is read from.  `coeff` is one spleeterCoeff blob.  Values are rounded to TF32
// (round-to-nearest, ties away) so fp32 `.dat` weights behave like cvt.rna; fp16-origin
// weights are exactly representable and pass through unchanged.
void pack_layer(const LayerPlan& L, const float* coeff, float* out);

float round_tf32(float x);
uint16_t bf16_rn(float x);              // round to nearest even (cvt.rn.bf16.f32)
float bf16_to_float(uint16_t h);
uint8_t e5m2_rn(float x);               // round to nearest even, saturating (cvt.rn.satfinite.e5m2x2.f32)
float e5m2_to_float(uint8_t b);
// value stored for a weight in a k-block of the given part
inline float weight_part(float w, int part) { const float hi = round_tf32(w); return (part & 1) == 0 ? hi : round_tf32(w - hi); }   // bit 0 of KBlock::part
bool weights_tf32_exact(const float* coeff);   // all tensor-core conv weights of one net representable in TF32?

// ---------------------------------------------------------------------------------------
// "Row-patch" form of the small-N layers (down2, down3, up4, up5).  A CTA owns R output rows x
// 128 columns of one image.  For every 32-channel slab ("chunk") of a source it loads ONE
// (R+2)-row x 136-pixel patch and issues the MMAs of all taps from shifted windows of that patch
// (A-operand descriptor start = patch + (r+dy+1)*row_pitch + (dx+1)*128 B), so every input pixel
// crosses L2->SM (R+2)/R times instead of once per tap.  Decoder layers fuse the 4 output
// parities into N = 4*cout (a tap that a parity does not use gets zero weights).
// ---------------------------------------------------------------------------------------
constexpr int kPatchW = 136;            // pixels per patch row (130 used; 136 keeps rows 1024B-aligned)

struct KElemP {
    int32_t cin;                        // -1 = zero slab element
    int8_t kh[4], kw[4];                // per fused phase; kh < 0: this phase does not use the tap
};
struct RowChunk {
    int8_t src;
    int32_t c_off;
    int32_t kb0, nkb;                   // its taps = k-blocks [kb0, kb0+nkb)
};
struct RowPlan {
    int index;                          // same numbering as LayerPlan::index
    bool transposed;
    int cin, cout, Hs, Ws;
    int N;                              // MMA N: cout (encoder) or 4*cout (decoder, phases fused)
    int R;                              // output rows per CTA
    int phases;                         // 1 or 4 (fused into N)
    int nsrc;
    SrcDesc src[2];
    std::vector<RowChunk> chunks;
    std::vector<KBlock> kb;             // dy, dx per k-block (src / c_off copied from the chunk)
    std::vector<KElemP> kelem;          // kb_channels() per k-block, k-block k starts at ke_off[k]
    std::vector<int32_t> ke_off;
    bool comp;                          // compensated precision: extra chunks with src = kSrcLo (see build_plans)
    int lo_fmt;
    SrcDesc lo_src;
    size_t w_floats_per_stem;           // kb.size() * N * 32
};
// rows per CTA for a given N (fixed by the shared-memory / TMEM budget, see srt_conv_rp.cu)
inline int row_plan_R(int N) { return N <= 64 ? 3 : 2; }
// ---- down1 on the tensor cores -------------------------------------------------------------
// The magnitude image is stored space-to-depth: [img][T/2][F/2][(py,px)][c] = 8 floats (32 B) per S2D
// pixel.  down1 (2 -> 16 ch, 5x5 stride 2) is then 9 stride-1 taps with K = 8 each.  All stems read the
// same magnitude, so up to 4 stems are fused into one MMA (N = 16 * stems): the k-block is 8 channels wide
// (one 32-byte SWIZZLE_32B row, a single K=8 MMA per tap).
constexpr int kKB1 = 8;
struct Down1Plan {
    int Hs, Ws;                         // output (= S2D input) extent: T/2 x F/2
    std::vector<KBlock> kb;             // 2 x 9 taps: src 0 = hi part of the magnitude, src 1 = lo part (same weights)
    std::vector<KElemP> kelem;          // 8 per tap (kh/kw in slot 0)
};
Down1Plan build_down1_plan(NetGeom g, bool split_weights = false);
// weights of `nstems` consecutive stems -> [tap][N = 16*nstems][8] fp32, SWIZZLE_32B pre-applied
void pack_down1(const Down1Plan& L, const float* const* coeffs, int nstems, float* out);
SRT_HD inline int swz32_index(int row, int j) { return row * 8 + ((((j >> 2) ^ ((row >> 2) & 1)) << 2) | (j & 3)); }
// float2 index of magnitude (t, f) inside one space-to-depth image
SRT_HD inline size_t mag_s2d_index(int T, int F, int t, int f) { return (((size_t)(t >> 1) * (F >> 1) + (f >> 1)) << 2) + ((t & 1) << 1) + (f & 1); }

bool row_plan_supported(int layer_index);                 // down2, down3, up4, up5
RowPlan build_row_plan(NetGeom g, int layer_index, bool split_weights = false, bool comp = false, int lo_fmt = LO_BF16);
void pack_row_layer(const RowPlan& L, const float* coeff, float* out);

// index of (row n, k-element j) inside a swizzled [rows][32] fp32 block
SRT_HD inline int swz128_index(int row, int j) { return row * 32 + ((((j >> 2) ^ (row & 7)) << 2) | (j & 3)); }
// same for a [rows][64] block of 2-byte elements (16-byte chunks of 8 elements)
SRT_HD inline int swz128_index16(int row, int j) { return row * 64 + ((((j >> 3) ^ (row & 7)) << 3) | (j & 7)); }
// and for a [rows][128] block of bytes (16-byte chunks of 16 elements)
SRT_HD inline int swz128_index8(int row, int j) { return row * 128 + ((((j >> 4) ^ (row & 7)) << 4) | (j & 15)); }
// a [rows][64] block of bytes in SWIZZLE_64B (chunk index ^= address bits 7-8 = (row >> 1) & 3)
SRT_HD inline int swz64_index8(int row, int j) { return row * 64 + ((((j >> 4) ^ ((row >> 1) & 3)) << 4) | (j & 15)); }
// byte j of row `row` of a [rows][32] byte tile in the SWIZZLE_32B layout (16-byte chunk ^= bit 2 of the row: address bit 4 ^= bit 7)
SRT_HD inline int swz32_index8(int row, int j) { return row * 32 + ((((j >> 4) ^ ((row >> 2) & 1)) << 4) | (j & 15)); }

// ---- up6 (transposed 5x5, 32 -> 1) for up6_tc_kernel: per stem kUp6PackFloats floats =
//   8 blocks [box b = source * 2 + channel half][term: tf32(w), tf32(w - tf32(w))][32 taps][8 channels] fp32, rows in the SWIZZLE_32B layout
//   (B operand of the K = 8 TF32 MMAs; taps 25..31 zero), then ONE block [32 taps][32 channels = skip1 16 | up5 16] of e5m2(w / 4) bytes,
//   SWIZZLE_32B (B operand of the K = 32 MMA that contracts the 8-bit residual tile).  w6 = the layer's [32 cin][25 taps] weights
//   (spleeter.c:289: channels [skip1 | up5]).
constexpr int kUp6PackFloats = 9 * 256;
void pack_up6_weights(const float* w6, float* out);

}  // namespace srt
