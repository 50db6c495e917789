// srt_unet_simt.cu — the HBM-bound edge layers of the U-Net as SIMT kernels (down1, up6, up7)
// plus a slow SIMT evaluation of the gather-GEMM layers that consumes exactly the same
// tables / packed weights / activation layouts as the tcgen05 kernel.  The latter is a
// verification aid (SRT_CONV_IMPL=simt), not a CPU fallback: everything here runs on the GPU.
#include <cuda_fp16.h>

#include "srt_epilogue.cuh"
#include "srt_kernels.cuh"
#include "srt_ptx.cuh"

namespace srt {

// =========================================================================================
// SIMT gather-GEMM (verification path)
// =========================================================================================
__global__ void __launch_bounds__(128) conv_simt_kernel(const __grid_constant__ ConvParams p)
{
    const int m = threadIdx.x;
    const int tx = blockIdx.x % p.tiles_x, ty = blockIdx.x / p.tiles_x;
    const int nt = blockIdx.y % p.n_tiles, phase = blockIdx.y / p.n_tiles;
    const int s = blockIdx.z / p.tiles_n, tz = blockIdx.z % p.tiles_n;
    const int x = m % p.tw, y = (m / p.tw) % p.th, nn = m / (p.tw * p.th);
    const int X = tx * p.tw + x, Y = ty * p.th + y, b = tz * p.nb + nn;
    const bool valid = X < p.Ws && Y < p.Hs && b < p.Bv;
    const int n = s * p.B + b;
    const KBlock* kbt = p.kb + p.kb_off[phase];
    const int nkb = p.nkb[phase];
    const float* wbase = p.w + (size_t)s * p.w_stem_stride + p.w_phase_off[phase] + (size_t)nt * nkb * p.n_tile * kKB;
    for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
        float acc[16];
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = 0.0f;
        for (int k = 0; k < nkb; k++) {
            const KBlock kb = kbt[k];
            const int yy = Y + kb.dy, xx = X + kb.dx;
            if (!valid || yy < 0 || yy >= p.Hs || xx < 0 || xx >= p.Ws) continue;   // zero padding
            const float* wb = wbase + (size_t)k * p.n_tile * kKB;
            if (kb.part & kPartLo8) {   // compensation block, 8-bit form: e5m2(4 lo) x e5m2(w / 4), 128 channels
                const uint8_t* a = reinterpret_cast<const uint8_t*>(p.lo_ptr) + (((size_t)n * p.Hs + yy) * p.Ws + xx) * p.lo_C + kb.c_off;
                const uint8_t* wh = reinterpret_cast<const uint8_t*>(wb);
                for (int j = 0; j < kKBlo8; j++) {
                    const float av = __half2float(__ushort_as_half((unsigned short)((unsigned)a[j] << 8)));   // e5m2 = the top byte of a half
#pragma unroll
                    for (int i = 0; i < 16; i++)
                        acc[i] = fmaf(av, __half2float(__ushort_as_half((unsigned short)((unsigned)wh[swz128_index8(c0 + i, j)] << 8))), acc[i]);
                }
                continue;
            }
            if (kb.part & kPartLo) {   // compensation block: bf16 residuals x bf16 weights, 64 channels
                const uint16_t* a = reinterpret_cast<const uint16_t*>(p.lo_ptr) + (((size_t)n * p.Hs + yy) * p.Ws + xx) * p.lo_C + kb.c_off;
                const uint16_t* wh = reinterpret_cast<const uint16_t*>(wb);
                for (int j = 0; j < kKBlo; j++) {
                    const float av = __uint_as_float((uint32_t)a[j] << 16);
#pragma unroll
                    for (int i = 0; i < 16; i++) acc[i] = fmaf(av, __uint_as_float((uint32_t)wh[swz128_index16(c0 + i, j)] << 16), acc[i]);
                }
                continue;
            }
            const int C = p.src_C[kb.src];
            const float* a = p.src_ptr[kb.src] + (((size_t)n * p.Hs + yy) * p.Ws + xx) * C + kb.c_off;
            for (int j = 0; j < kKB; j++) {
                const float av = a[j];
#pragma unroll
                for (int i = 0; i < 16; i++) acc[i] = fmaf(av, wb[swz128_index(c0 + i, j)], acc[i]);
            }
        }
        const int col = nt * p.n_tile + c0;
        if (valid) epilogue16(p, s, n, Y, X, p.fused ? col / p.cout : phase, p.fused ? col % p.cout : col, acc);
    }
}

void launch_conv_simt(const ConvParams& p, cudaStream_t st)
{
    dim3 grid(p.tiles_x * p.tiles_y, p.n_tiles * p.phases, p.S * p.tiles_n);
    conv_simt_kernel<<<grid, 128, 0, st>>>(p);
}

// =========================================================================================
// down1: 5x5 stride-2 conv, 2 -> 16 channels, on the magnitude image (spleeter.c:181-190).
// Input row 2*oh + kh - 1, col 2*ow + kw - 1 (im2col_dilated.c:19-27).  Writes the raw skip
// (conv + bias, fp32) and the activated feature in space-to-depth form, TF32-rounded, which is
// what down2's tensor-core kernel reads.
// Each thread produces a 2x2 block of output pixels x 16 channels (64 accumulators), so every
// broadcast weight load feeds 4 pixels and the kernel is FMA-bound rather than LDS-bound.
// =========================================================================================
constexpr int D1_BX = 16, D1_BY = 8;                    // threads
constexpr int D1_TW = 2 * D1_BX, D1_TH = 2 * D1_BY;     // output pixels per block: 32 x 16
constexpr int D1_PW = 2 * D1_TW + 3, D1_PH = 2 * D1_TH + 3;

__global__ void __launch_bounds__(D1_BX* D1_BY) down1_kernel(const __grid_constant__ Down1Params p)
{
    __shared__ float2 patch[D1_PH][D1_PW + 1];
    const int s = p.stem, b = blockIdx.z, n = s * p.B + b;
    const int Ho = p.T / 2, Wo = p.F / 2;
    const int ow0 = blockIdx.x * D1_TW, oh0 = blockIdx.y * D1_TH;
    const int tid = threadIdx.x;
    const float2* img = reinterpret_cast<const float2*>(p.mag) + (size_t)b * p.T * p.F;
    const float2* img_lo = reinterpret_cast<const float2*>(p.mag_lo) + (size_t)b * p.T * p.F;
    for (int i = tid; i < D1_PH * D1_PW; i += blockDim.x) {
        const int r = i / D1_PW, c = i % D1_PW;
        const int ih = 2 * oh0 - 1 + r, iw = 2 * ow0 - 1 + c;
        float2 v = make_float2(0.f, 0.f);
        if (ih >= 0 && ih < p.T && iw >= 0 && iw < p.F) {
            const size_t mi = mag_s2d_index(p.T, p.F, ih, iw);
            const float2 hi = img[mi], lo = img_lo[mi];
            v = make_float2(hi.x + lo.x, hi.y + lo.y);
        }
        patch[r][c] = v;
    }
    __syncthreads();
    const int tx = tid % D1_BX, ty = tid / D1_BX;
    // this thread's 2x2 outputs are (ty + 8a, tx + 16c): interleaved so that neighbouring lanes read
    // neighbouring patch columns (no shared-memory bank conflicts).  Weights are compile-time offsets
    // into the kernel-parameter (constant) bank: the FFMAs take them as uniform operands.
    float acc[2][2][16];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int i = 0; i < 16; i++) acc[a][c][i] = 0.0f;
#pragma unroll
    for (int kh = 0; kh < 5; kh++) {
#pragma unroll
        for (int kw = 0; kw < 5; kw++) {
            float2 v[2][2];
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int c = 0; c < 2; c++) v[a][c] = patch[2 * (ty + D1_BY * a) + kh][2 * (tx + D1_BX * c) + kw];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const float wl = p.wk[((kh * 5 + kw) * 2 + 0) * 16 + i], wr = p.wk[((kh * 5 + kw) * 2 + 1) * 16 + i];
#pragma unroll
                for (int a = 0; a < 2; a++)
#pragma unroll
                    for (int c = 0; c < 2; c++) acc[a][c][i] = fmaf(wl, v[a][c].x, fmaf(wr, v[a][c].y, acc[a][c][i]));
            }
        }
    }
    const int act = p.act[0];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int oh = oh0 + ty + D1_BY * a, ow = ow0 + tx + D1_BX * c;
            if (ow >= Wo || oh >= Ho) continue;
            float raw[16], av[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const float t = acc[a][c][i] + p.bk[i];
                raw[i] = t;
                const float full = apply_act(act, p.bk[16 + i] * t + p.bk[32 + i]);
                av[i] = ptx::rna_tf32(full);
                lo[i] = full - av[i];
            }
            float4* d0 = reinterpret_cast<float4*>(p.out_raw + (((size_t)n * Ho + oh) * Wo + ow) * 16);
            float4* d1 = reinterpret_cast<float4*>(
                p.out_act + ((((size_t)n * (Ho / 2) + oh / 2) * (Wo / 2) + ow / 2) * 4 + (oh & 1) * 2 + (ow & 1)) * 16);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                d0[q] = make_float4(raw[4 * q], raw[4 * q + 1], raw[4 * q + 2], raw[4 * q + 3]);
                d1[q] = make_float4(av[4 * q], av[4 * q + 1], av[4 * q + 2], av[4 * q + 3]);
            }
            if (p.lo_act)
                store16_lo(p.lo_act + ((((size_t)n * (Ho / 2) + oh / 2) * (Wo / 2) + ow / 2) * 4 + (oh & 1) * 2 + (ow & 1)) * 16, lo);
        }
}

void launch_down1(const Down1Params& p, cudaStream_t st)
{
    dim3 grid((p.F / 2 + D1_TW - 1) / D1_TW, (p.T / 2 + D1_TH - 1) / D1_TH, p.Bv);   // one launch per stem
    down1_kernel<<<grid, D1_BX * D1_BY, 0, st>>>(p);
}

// =========================================================================================
// up6: 5x5 stride-2 transposed conv, [skip1 | up5] (32 ch) -> 1 ch, then act, then BN
// (spleeter.c:289-294).  Output (2h+kh-1, 2w+kw-1) (im2col_dilated.c:57-58, 37-40).
// One thread handles two horizontally adjacent input-resolution pixels (a 2x4 output block), 16
// channels at a time (skip half, then up5 half), weights read as broadcast float4.
// =========================================================================================
constexpr int U6_BX = 32, U6_BY = 8;                  // threads
constexpr int U6_TW = 2 * U6_BX, U6_TH = U6_BY;       // input pixels per block: 64 x 8
constexpr int U6_PW = U6_TW + 2, U6_PH = U6_TH + 2, U6_PS = U6_PW + 2;   // row stride 68 floats (even)

__global__ void __launch_bounds__(U6_BX* U6_BY) up6_kernel(const __grid_constant__ Up6Params p)
{
    __shared__ __align__(16) float patch[16 * U6_PH * U6_PS];   // [16 ch][10][68]
    const int s = p.stem, n = s * p.B + blockIdx.z;
    const int H = p.T / 2, W = p.F / 2;
    const int x0 = blockIdx.x * U6_TW, y0 = blockIdx.y * U6_TH;
    const int tid = threadIdx.x;
    const int tx = tid % U6_BX, ty = tid / U6_BX;
    float o[2][4];   // [output row parity][output col 0..3] for input pixels (2tx, 2tx+1)
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) o[a][b] = 0.0f;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        __syncthreads();   // previous half fully consumed
        const float* src = half ? p.up : p.skip;
        for (int i = tid; i < U6_PH * U6_PW * 4; i += blockDim.x) {
            const int q = i & 3, pix = i >> 2;
            const int r = pix / U6_PW, c = pix % U6_PW;
            const int yy = y0 - 1 + r, xx = x0 - 1 + c;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = *reinterpret_cast<const float4*>(src + (((size_t)n * H + yy) * W + xx) * 16 + q * 4);
            float* d = patch + ((q * 4) * U6_PH + r) * U6_PS + c;
            d[0] = v.x;
            d[U6_PH * U6_PS] = v.y;
            d[2 * U6_PH * U6_PS] = v.z;
            d[3 * U6_PH * U6_PS] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 16; c++) {
            // neighbourhood: rows ty..ty+2 (dy = -1..1), cols 2tx..2tx+3 (input x-1 .. x+2)
            float nb[3][4];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const float2* rp = reinterpret_cast<const float2*>(patch + (c * U6_PH + ty + a) * U6_PS + 2 * tx);
                const float2 u = rp[0], v = rp[1];
                nb[a][0] = u.x; nb[a][1] = u.y; nb[a][2] = v.x; nb[a][3] = v.y;
            }
            // output row parity 0: kh=1 (dy=0), kh=3 (dy=-1); parity 1: kh=0 (dy=+1), kh=2 (dy=0), kh=4 (dy=-1)
            // output col parity likewise.  j = 0/1 selects the input pixel; its neighbourhood column for dx is j+1+dx.
            // The weight index is a compile-time constant: the operand comes straight from the constant bank.
#define TAP(kh, kw, dy, dx) (p.wk[(half * 16 + c) * 25 + (kh) * 5 + (kw)] * nb[(dy) + 1][j + 1 + (dx)])
#pragma unroll
            for (int j = 0; j < 2; j++) {
                o[0][2 * j + 0] += TAP(1, 1, 0, 0) + TAP(1, 3, 0, -1) + TAP(3, 1, -1, 0) + TAP(3, 3, -1, -1);
                o[0][2 * j + 1] += TAP(1, 0, 0, 1) + TAP(1, 2, 0, 0) + TAP(1, 4, 0, -1) + TAP(3, 0, -1, 1) + TAP(3, 2, -1, 0) + TAP(3, 4, -1, -1);
                o[1][2 * j + 0] += TAP(0, 1, 1, 0) + TAP(0, 3, 1, -1) + TAP(2, 1, 0, 0) + TAP(2, 3, 0, -1) + TAP(4, 1, -1, 0) + TAP(4, 3, -1, -1);
                o[1][2 * j + 1] += TAP(0, 0, 1, 1) + TAP(0, 2, 1, 0) + TAP(0, 4, 1, -1) + TAP(2, 0, 0, 1) + TAP(2, 2, 0, 0) + TAP(2, 4, 0, -1) +
                                   TAP(4, 0, -1, 1) + TAP(4, 2, -1, 0) + TAP(4, 4, -1, -1);
            }
#undef TAP
        }
    }
    const int X = x0 + 2 * tx, Y = y0 + ty;
    if (X >= W || Y >= H) return;
    const float bias = p.bias[s], sc = p.bn_scale[s], of = p.bn_offset[s];
    const int act = p.act[0];   // launch_up6 puts this stem's activation in slot 0
    float4 r0, r1;
    r0.x = sc * apply_act(act, o[0][0] + bias) + of;
    r0.y = sc * apply_act(act, o[0][1] + bias) + of;
    r0.z = sc * apply_act(act, o[0][2] + bias) + of;
    r0.w = sc * apply_act(act, o[0][3] + bias) + of;
    r1.x = sc * apply_act(act, o[1][0] + bias) + of;
    r1.y = sc * apply_act(act, o[1][1] + bias) + of;
    r1.z = sc * apply_act(act, o[1][2] + bias) + of;
    r1.w = sc * apply_act(act, o[1][3] + bias) + of;
    float* dst = p.out + ((size_t)n * p.T + 2 * Y) * p.F + 2 * X;
    *reinterpret_cast<float4*>(dst) = r0;            // W is even and X is even: 16-byte aligned
    *reinterpret_cast<float4*>(dst + p.F) = r1;
}

void launch_up6(const Up6Params& p, cudaStream_t st)
{
    // one launch per stem; p.wk / p.stem / p.act[0] are filled by the caller (srt_ctx.cu)
    dim3 grid((p.F / 2 + U6_TW - 1) / U6_TW, (p.T / 2 + U6_TH - 1) / U6_TH, p.Bv);
    up6_kernel<<<grid, U6_BX * U6_BY, 0, st>>>(p);
}

// =========================================================================================
// up7: 4x4, dilation 2, stride 1 conv 1 -> 2 (rows t + 2kh - 3, cols f + 2kw - 3; spleeter.c:156,295)
// + bias + sigmoid (LUT flavour spleeter.c:29-42 or exact VST/Source/spleeter.c:56-65).
// =========================================================================================
constexpr int U7_TW = 128, U7_TH = 16, U7_ROWS = 64;   // CTA: 64 rows x 128 columns, walked as 4 tiles of 16 rows
constexpr int U7_PITCH = U7_TW + 8;                      // tile columns f0-4 .. f0+131 (float4-aligned halo)

// fastSigmoid (spleeter.c:30-42) with the per-interval constants precomputed on the host with the
// same float operations: entry = {tbl[i], (tbl[i+1]-tbl[i]) / ((-7+step*(i+1)) - (-7+step*i))}; the interval
// origin -7+step*i is re-evaluated here with the reference's two roundings (multiply, then add).
// Only the index division is replaced (reciprocal multiply): on the rare inputs where the rounded product
// lands on the other side of an integer the neighbouring interval is used, which changes the (continuous,
// piecewise-linear) result by ~1e-8.
// The table lives in shared memory: gathered from global memory it cost ~28 L1 wavefronts per warp and
// lookup (every lane another line), which made the L1 the bound of this kernel (ncu r1k: 0.155 ms per
// stem launch against 0.03 ms of HBM time).
__device__ __forceinline__ float sigmoid_lut(const float2* __restrict__ tbl, float x)
{
    const float step = 0.01367188f;
    if (x > 7.0f) return 1.0f;
    if (x < -7.0f) return 0.0f;
    const int idx = min((int)(__fadd_rn(x, 7.0f) * (1.0f / step)), 1023);
    const float2 e = tbl[idx];
    const float x1 = __fadd_rn(-7.0f, __fmul_rn(step, (float)idx));
    return __fadd_rn(e.x, __fmul_rn(e.y, __fsub_rn(x, x1)));
}

__device__ __forceinline__ float sigmoid_exact(float a)
{
    return a >= 0.f ? 1.0f / (1.0f + expf(-a)) : expf(a) / (1.0f + expf(a));
}

// Each thread produces 4 consecutive columns of output rows r0 and r0+2: the two rows share 3 of their 4
// input rows (dilation 2) and the 4 columns share their 10-float input span, read as 3 LDS.128 per row:
// 15 vector loads feed 8 pixels x 32 FMAs (the one-pixel-per-thread version issued 16 scalar LDS per pixel).
// Accumulation order per output (kh-major, kw-minor, bias last) is the reference's (gemm row order).
template <bool MERGED>
__global__ void __launch_bounds__(256) up7_kernel(const __grid_constant__ Up7Params p)
{
    __shared__ __align__(16) float tile[(U7_TH + 6) * U7_PITCH];
    __shared__ float2 lut_s[1025];
    const int s = MERGED ? (int)blockIdx.z / p.Bv : p.stem, b = MERGED ? (int)blockIdx.z % p.Bv : (int)blockIdx.z, n = s * p.B + b;
    const float* wk = MERGED ? p.wk_all[s] : p.wk;     // block-uniform: constant-bank operands either way
    const int f0 = blockIdx.x * U7_TW;
    const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
    const float* img = p.in + (size_t)n * p.T * p.F;
    if (p.lut) {
        const float4* lut4 = reinterpret_cast<const float4*>(p.lut);
        for (int i = tid; i < 1025; i += 256) {
            const float4 e = __ldg(lut4 + i);
            lut_s[i] = make_float2(e.x, e.y);
        }
    }
    const int fx = 4 * lane;
    const int r0 = (g >> 1) * 4 + (g & 1);   // tile-relative output rows r0, r0 + 2
    float2* mimg = reinterpret_cast<float2*>(p.mask) + ((size_t)s * p.mask_stem_stride + p.mask_img0 + b) * p.T * p.F;
    for (int t0 = blockIdx.y * U7_ROWS; t0 < min((int)(blockIdx.y + 1) * U7_ROWS, p.T); t0 += U7_TH) {
        __syncthreads();   // previous tile fully consumed (and, first time, nothing to wait for)
        for (int i = tid; i < (U7_TH + 6) * (U7_PITCH / 4); i += 256) {
            const int r = i / (U7_PITCH / 4), c4 = i % (U7_PITCH / 4);
            const int t = t0 - 3 + r, f = f0 - 4 + 4 * c4;   // F is a multiple of 4: a float4 is entirely in or out
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t >= 0 && t < p.T && f >= 0 && f < p.F) v = *reinterpret_cast<const float4*>(img + (size_t)t * p.F + f);
            *reinterpret_cast<float4*>(tile + r * U7_PITCH + 4 * c4) = v;
        }
        __syncthreads();
        // (mask L, mask R) of a pixel accumulate as one packed pair: acc += (w_L, w_R) * (x, x) is ONE FFMA2 (the weight pairs sit in
        // uniform registers, x is a broadcast operand), 16 instead of 32 FMA issue slots per pixel; each half rounds like the scalar
        // FFMA, so the result is bit-identical to the scalar loop
        unsigned long long acc[2][4];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[a][j] = 0ull;
#pragma unroll
        for (int q = 0; q < 5; q++) {
            // tile row r0 + 2q = input row of output row r0 for kh = q and of output row r0 + 2 for kh = q - 1
            const float4* rp = reinterpret_cast<const float4*>(tile + (r0 + 2 * q) * U7_PITCH + fx);
            const float4 va = rp[0], vb = rp[1], vc = rp[2];
            const float v[12] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w, vc.x, vc.y, vc.z, vc.w};
#pragma unroll
            for (int kw = 0; kw < 4; kw++)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float x = v[1 + j + 2 * kw];   // input column f + 2 kw - 3
                    const unsigned long long xx = ptx::pack_f32x2(x, x);
                    if (q < 4) acc[0][j] = ptx::fma_f32x2(ptx::pack_f32x2(wk[q * 4 + kw], wk[16 + q * 4 + kw]), xx, acc[0][j]);
                    if (q >= 1) acc[1][j] = ptx::fma_f32x2(ptx::pack_f32x2(wk[(q - 1) * 4 + kw], wk[16 + (q - 1) * 4 + kw]), xx, acc[1][j]);
                }
        }
        const int f = f0 + fx;
        if (f < p.F) {
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const int t = t0 + r0 + 2 * a;
                float m[8];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float2 av = ptx::unpack_f32x2(acc[a][j]);
                    const float a0 = av.x + wk[32], a1 = av.y + wk[33];
                    if (p.lut) { m[2 * j] = sigmoid_lut(lut_s, a0); m[2 * j + 1] = sigmoid_lut(lut_s, a1); }
                    else { m[2 * j] = sigmoid_exact(a0); m[2 * j + 1] = sigmoid_exact(a1); }
                }
                float4* d = reinterpret_cast<float4*>(mimg + (size_t)t * p.F + f);
                d[0] = make_float4(m[0], m[1], m[2], m[3]);
                d[1] = make_float4(m[4], m[5], m[6], m[7]);
            }
        }
    }
}

void launch_up7(const Up7Params& p, cudaStream_t st)
{
    dim3 grid((p.F + U7_TW - 1) / U7_TW, (p.T + U7_ROWS - 1) / U7_ROWS, p.Bv);   // one launch per stem (p.stem, p.wk) ...
    if (p.merged) {                                                               // ... or one for all stems (p.wk_all)
        grid.z = p.Bv * p.S;
        up7_kernel<true><<<grid, 256, 0, st>>>(p);
    } else {
        up7_kernel<false><<<grid, 256, 0, st>>>(p);
    }
}

// API layout [n][T][F][2] -> internal space-to-depth layout, TF32-rounded (srt_unet_device)
__global__ void mag_to_s2d_kernel(const float2* __restrict__ in, float2* __restrict__ out, float2* __restrict__ out_lo, int T, int F, int n_img)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, per = (size_t)T * F;
    if (i >= per * n_img) return;
    const int img = (int)(i / per), t = (int)((i % per) / F), f = (int)(i % F);
    const float2 v = in[i];
    const float2 hi = make_float2(ptx::rna_tf32(v.x), ptx::rna_tf32(v.y));
    out[(size_t)img * per + mag_s2d_index(T, F, t, f)] = hi;
    out_lo[(size_t)img * per + mag_s2d_index(T, F, t, f)] = make_float2(v.x - hi.x, v.y - hi.y);
}
void launch_mag_to_s2d(const float* in, float* out, float* out_lo, int T, int F, int n_img, cudaStream_t st)
{
    const size_t n = (size_t)T * F * n_img;
    mag_to_s2d_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float2*>(in), reinterpret_cast<float2*>(out), reinterpret_cast<float2*>(out_lo), T, F, n_img);
}

}  // namespace srt
