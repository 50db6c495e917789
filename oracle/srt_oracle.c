/*
 * srt_oracle.c — CPU restatement of SpleeterRT's spectrogram-to-mask hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path; it is never
 * linked into, imported by, or executed from the product (spleeterrt_b200/, include/,
 * libspleeterrt_b200.so).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may touch anything under oracle/.
 *
 * Parity pin: the reference publishes no golden vectors, so this
 * restatement is pinned against the reference's own C code compiled from
 * /root/reference into oracle/_ref/ (oracle/build_ref.py) — see tests/test_oracle.py —
 * and against the fixtures under tests/golden/ that were generated from that build
 * (tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it follows.  The structure is deliberately
 * NOT the reference's (no im2col buffer, no sgemm): convolutions are evaluated directly, but
 * in the same accumulation order as the reference's naive gemm so results agree to the bit
 * (modulo the compiler), which makes the pin sharp.
 *
 * All tensors are planar [channel][row][col] float32 exactly like the reference
 * ("row" = time frame t, "col" = frequency bin f).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SRT_FFT 4096
#define SRT_HOP 1024
#define SRT_BINS 2049

/* ------------------------------------------------------------------------------------
 * Weight blob layout: one net = 9 822 725 floats in the order of `spleeterCoeff`
 * (Executable/spleeter.h:5-31).  Offsets are derived, not copied.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    const float *w, *b, *bn; /* bn: [0,C) offset, [C,2C) scale (Executable/spleeter.c:188) */
    int cin, cout;
} layer_t;

typedef struct {
    layer_t down[6];
    layer_t up[6];
    const float *w7, *b7;
} net_t;

static const int kEncCh[7] = {2, 16, 32, 64, 128, 256, 512};
/* decoder: input channels (after concat) and output channels, Executable/spleeter.c:150-155 */
static const int kDecIn[6] = {512, 512, 256, 128, 64, 32};
static const int kDecOut[6] = {256, 128, 64, 32, 16, 1};

size_t srt_oracle_coeff_floats(void)
{
    size_t n = 0;
    for (int i = 0; i < 6; i++) {
        n += (size_t)25 * kEncCh[i] * kEncCh[i + 1] + kEncCh[i + 1];
        if (i < 5) n += 2 * kEncCh[i + 1]; /* down6 has no batch norm (spleeter.h:16) */
    }
    for (int i = 0; i < 6; i++) n += (size_t)25 * kDecIn[i] * kDecOut[i] + kDecOut[i] + 2 * kDecOut[i];
    n += 4 * 4 * 1 * 2 + 2;
    return n;
}

static void net_bind(net_t *n, const float *p)
{
    for (int i = 0; i < 6; i++) {
        layer_t *l = &n->down[i];
        l->cin = kEncCh[i];
        l->cout = kEncCh[i + 1];
        l->w = p; p += (size_t)25 * l->cin * l->cout;
        l->b = p; p += l->cout;
        if (i < 5) { l->bn = p; p += 2 * l->cout; } else l->bn = NULL;
    }
    for (int i = 0; i < 6; i++) {
        layer_t *l = &n->up[i];
        l->cin = kDecIn[i];
        l->cout = kDecOut[i];
        l->w = p; p += (size_t)25 * l->cin * l->cout;
        l->b = p; p += l->cout;
        l->bn = p; p += 2 * l->cout;
    }
    n->w7 = p; p += 32;
    n->b7 = p;
}

/* ------------------------------------------------------------------------------------
 * Activations.  flavour 0 = Executable (LUT sigmoid spleeter.c:29-42, ELU clamp :51-56),
 * flavour 1 = VST (exact logistic VST/Source/spleeter.c:56-65, unclamped ELU :74-77).
 * ---------------------------------------------------------------------------------- */
static float g_sig_tbl[1026];
static int g_sig_ready = 0;

/* The reference's table (spleeter.c:29) is DATA: ~sigma(-7 + i*14/1024), i = 0..1024, plus a final 1.0, but printed from
 * some float evaluation - regenerating it in double and rounding to 8 decimals misses 307 entries by one ulp.  It is
 * carried as bit patterns (oracle/sigmoid_table.h, written by tools/gen_sigmoid_table.py from the reference); the pin test
 * demands fastSigmoid() of oracle/_ref == srt_oracle_sigmoid_lut() bit for bit on a dense grid. */
#include "sigmoid_table.h"
static void sig_init(void)
{
    if (g_sig_ready) return;
    memcpy(g_sig_tbl, kSigmoidTableBits, sizeof g_sig_tbl);
    g_sig_ready = 1;
}

const float *srt_oracle_sigmoid_table(void) { sig_init(); return g_sig_tbl; }

float srt_oracle_sigmoid_lut(float x)
{
    const float step = 0.01367188f; /* spleeter.c:38 */
    sig_init();
    if (x > 7.0f) return 1.0f;
    if (x < -7.0f) return 0.0f;
    short idx = (short)((x + 7.0f) / step);
    float x1 = -7.0f + step * idx;
    return g_sig_tbl[idx] + (g_sig_tbl[idx + 1] - g_sig_tbl[idx]) / (-7.0f + step * (idx + 1) - x1) * (x - x1);
}

static float sigmoid_exact(float x)
{
    /* VST/Source/spleeter.c:56-65: numerically split logistic */
    if (x >= 0.0f) return 1.0f / (1.0f + expf(-x));
    float z = expf(x);
    return z / (1.0f + z);
}

static inline float act_apply(int kind, float x)
{
    switch (kind) {
    case 0: return x >= 0.0f ? x : 0.2f * x;                       /* leakyReLU spleeter.c:43-46 */
    case 1: return x >= 0.0f ? x : 0.0f;                           /* ReLU      spleeter.c:47-50 */
    case 2: if (x < -15.0f) return -1.0f;                          /* ELU exec  spleeter.c:51-56 */
            return x >= 0.0f ? x : expf(x) - 1.0f;
    default: return x >= 0.0f ? x : expf(x) - 1.0f;                /* ELU vst */
    }
}

/* ------------------------------------------------------------------------------------
 * 5x5 stride-2 "SAME" convolution (processConv2dLayer spleeter.c:96-100 =
 * im2col_dilated_cpu im2col_dilated.c:10-33 + gemm_nn gemm.c:6-18): input row 2h+kh-1,
 * col 2w+kw-1, zero outside; fp32 accumulation in k = (c, kh, kw) ascending order.
 * ---------------------------------------------------------------------------------- */
static void conv5x5_s2(const float *x, int cin, int H, int W, const float *wgt, int cout, float *y)
{
    const int Ho = H / 2, Wo = W / 2;
#pragma omp parallel for schedule(dynamic, 1)
    for (int o = 0; o < cout; o++) {
        float *yo = y + (size_t)o * Ho * Wo;
        memset(yo, 0, sizeof(float) * (size_t)Ho * Wo);
        for (int c = 0; c < cin; c++) {
            const float *xc = x + (size_t)c * H * W;
            for (int kh = 0; kh < 5; kh++)
                for (int kw = 0; kw < 5; kw++) {
                    const float wv = wgt[(((size_t)o * cin + c) * 5 + kh) * 5 + kw];
                    /* valid output range so that 0 <= 2h+kh-1 < H */
                    int h0 = (kh == 0) ? 1 : 0, h1 = Ho;
                    while (h1 > h0 && 2 * (h1 - 1) + kh - 1 >= H) h1--;
                    int w0 = (kw == 0) ? 1 : 0, w1 = Wo;
                    while (w1 > w0 && 2 * (w1 - 1) + kw - 1 >= W) w1--;
                    for (int h = h0; h < h1; h++) {
                        const float *xr = xc + (size_t)(2 * h + kh - 1) * W + (kw - 1);
                        float *yr = yo + (size_t)h * Wo;
                        for (int w = w0; w < w1; w++) yr[w] += wv * xr[2 * w];
                    }
                }
        }
    }
}

/* ------------------------------------------------------------------------------------
 * 5x5 stride-2 transposed convolution (processTransposeConv2dLayer spleeter.c:73-78 =
 * gemm_tn gemm.c:33-45 + col2im_dilated_cpu im2col_dilated.c:42-65): output (2h+kh-1,
 * 2w+kw-1) += sum_cin W[cin][co][kh][kw] x[cin][h][w]; per-tap partial sums over cin are
 * formed first (the gemm) and then added in (kh, kw) ascending order (the col2im).
 * ---------------------------------------------------------------------------------- */
static void tconv5x5_s2(const float *x, int cin, int H, int W, const float *wgt, int cout, float *y)
{
    const int Ho = 2 * H, Wo = 2 * W;
#pragma omp parallel
    {
        float *part = (float *)malloc(sizeof(float) * (size_t)H * W);
#pragma omp for schedule(dynamic, 1)
        for (int o = 0; o < cout; o++) {
            float *yo = y + (size_t)o * Ho * Wo;
            memset(yo, 0, sizeof(float) * (size_t)Ho * Wo);
            for (int kh = 0; kh < 5; kh++)
                for (int kw = 0; kw < 5; kw++) {
                    memset(part, 0, sizeof(float) * (size_t)H * W);
                    for (int c = 0; c < cin; c++) {
                        const float wv = wgt[(((size_t)c * cout + o) * 5 + kh) * 5 + kw];
                        const float *xc = x + (size_t)c * H * W;
                        for (int i = 0; i < H * W; i++) part[i] += wv * xc[i];
                    }
                    for (int h = 0; h < H; h++) {
                        int oh = 2 * h + kh - 1;
                        if (oh < 0 || oh >= Ho) continue;
                        for (int w = 0; w < W; w++) {
                            int ow = 2 * w + kw - 1;
                            if (ow < 0 || ow >= Wo) continue;
                            yo[(size_t)oh * Wo + ow] += part[(size_t)h * W + w];
                        }
                    }
                }
        }
        free(part);
    }
}

/* 4x4, dilation 2, stride 1 head (spleeter.c:156, 295): rows h+2kh-3, cols w+2kw-3. */
static void conv4x4_d2(const float *x, int H, int W, const float *wgt, float *y)
{
#pragma omp parallel for
    for (int o = 0; o < 2; o++) {
        float *yo = y + (size_t)o * H * W;
        memset(yo, 0, sizeof(float) * (size_t)H * W);
        for (int kh = 0; kh < 4; kh++)
            for (int kw = 0; kw < 4; kw++) {
                const float wv = wgt[(o * 4 + kh) * 4 + kw];
                for (int h = 0; h < H; h++) {
                    int ih = h + 2 * kh - 3;
                    if (ih < 0 || ih >= H) continue;
                    for (int w = 0; w < W; w++) {
                        int iw = w + 2 * kw - 3;
                        if (iw < 0 || iw >= W) continue;
                        yo[(size_t)h * W + w] += wv * x[(size_t)ih * W + iw];
                    }
                }
            }
    }
}

/* ------------------------------------------------------------------------------------
 * processSpleeter (Executable/spleeter.c:177-301).  x, y: [2][T][F].
 *   stemMode 0: encoder leakyReLU(0.2), decoder ReLU; !=0: ELU/ELU (spleeter.c:130-139)
 *   flavour  0: Executable (LUT sigmoid, ELU clamp); 1: VST (exact sigmoid, no clamp)
 *   taps (optional): receives, back to back, the 6 raw encoder outputs (conv+bias, the
 *   skips), then the 6 decoder outputs after act+BN (up1..up6), planar NCHW.
 * ---------------------------------------------------------------------------------- */
size_t srt_oracle_taps_floats(int F, int T)
{
    size_t P = (size_t)F * T, n = 0;
    for (int i = 0; i < 6; i++) n += (P >> (2 * (i + 1))) * kEncCh[i + 1];
    for (int i = 0; i < 6; i++) n += (P >> (2 * (5 - i))) * kDecOut[i];
    return n;
}

void srt_oracle_unet(const float *coeff, int F, int T, int stemMode, int flavour,
                     const float *x, float *y, float *taps)
{
    net_t net;
    net_bind(&net, coeff);
    const int actEnc = stemMode ? (flavour ? 3 : 2) : 0;
    const int actDec = stemMode ? (flavour ? 3 : 2) : 1;
    const size_t P = (size_t)F * T;
    float *skip[6];
    float *cur = (float *)malloc(sizeof(float) * P * 8);   /* activated encoder feature */
    float *cat = (float *)malloc(sizeof(float) * P * 32);  /* [skip | up] concat buffer */
    int H = T, W = F;
    const float *in = x;
    for (int i = 0; i < 6; i++) {
        const layer_t *l = &net.down[i];
        const int Ho = H / 2, Wo = W / 2;
        const size_t hw = (size_t)Ho * Wo;
        skip[i] = (float *)malloc(sizeof(float) * hw * l->cout);
        conv5x5_s2(in, l->cin, H, W, l->w, l->cout, skip[i]);
        for (int c = 0; c < l->cout; c++)
            for (size_t p = 0; p < hw; p++) {
                float v = skip[i][c * hw + p] + l->b[c];          /* spleeter.c:186-187 */
                skip[i][c * hw + p] = v;
                if (l->bn) cur[c * hw + p] = act_apply(actEnc, l->bn[l->cout + c] * v + l->bn[c]);
            }
        if (taps) { memcpy(taps, skip[i], sizeof(float) * hw * l->cout); taps += hw * l->cout; }
        in = cur;
        H = Ho; W = Wo;
    }
    /* decoder: input of up1 is conv6; afterwards [skip_k | up] (spleeter.c:239-289) */
    const float *din = skip[5];
    for (int i = 0; i < 6; i++) {
        const layer_t *l = &net.up[i];
        const int Ho = 2 * H, Wo = 2 * W;
        const size_t hw = (size_t)Ho * Wo;
        const int nskip = (i < 5) ? l->cout : 0;   /* channels of the skip placed in front */
        float *up = cat + hw * nskip;
        float *tmp = (i == 5) ? (float *)malloc(sizeof(float) * hw) : up;
        tconv5x5_s2(din, l->cin, H, W, l->w, l->cout, tmp);
        for (int c = 0; c < l->cout; c++)
            for (size_t p = 0; p < hw; p++) {
                float v = act_apply(actDec, tmp[c * hw + p] + l->b[c]);   /* spleeter.c:244 */
                tmp[c * hw + p] = l->bn[l->cout + c] * v + l->bn[c];      /* spleeter.c:245 */
            }
        if (taps) { memcpy(taps, tmp, sizeof(float) * hw * l->cout); taps += hw * l->cout; }
        if (i < 5) {
            memcpy(cat, skip[4 - i], sizeof(float) * hw * nskip);        /* spleeter.c:248 */
            din = cat;
        } else {
            memcpy(cat, tmp, sizeof(float) * hw);
            free(tmp);
        }
        H = Ho; W = Wo;
    }
    float *head = (float *)malloc(sizeof(float) * P * 2);
    conv4x4_d2(cat, T, F, net.w7, head);
    for (int c = 0; c < 2; c++)
        for (size_t p = 0; p < P; p++) {
            float v = head[c * P + p] + net.b7[c];
            y[c * P + p] = flavour ? sigmoid_exact(v) : srt_oracle_sigmoid_lut(v);   /* spleeter.c:299 */
        }
    free(head);
    for (int i = 0; i < 6; i++) free(skip[i]);
    free(cur);
    free(cat);
}

/* ------------------------------------------------------------------------------------
 * Transforms.  Tables follow InitSTFT (stftFix.c:302-312): pre-window 0.5*hann(i+1/2)/4096,
 * post-window (2/3)*hann(i+1/2), sine table sin(2*pi*i/4096), 12-bit bit reversal.
 * ---------------------------------------------------------------------------------- */
static float g_pre[SRT_FFT], g_post[SRT_FFT], g_sin[SRT_FFT];
static unsigned g_rev[SRT_FFT];
static int g_tab_ready = 0;

static void tables_init(void)
{
    if (g_tab_ready) return;
    const double w = 6.283185307179586476925286766559 / SRT_FFT;
    for (int i = 0; i < SRT_FFT; i++) {
        /* LLraisedCosTblFloat(n, LAP=4): (1/n) * hann(i + 0.5)   stftFix.c:48-57 */
        float rc = (float)((1.0 / SRT_FFT) * (0.5 * (1.0 - cos(w * (i + 0.5)))));
        g_pre[i] = rc * (2.0f / 4.0f);                            /* stftFix.c:307-308 */
        /* LLCreatePostWindowFloat: scalefac = n * (1/2)/(3/8)   stftFix.c:64-75, then *0.5 :311-312 */
        const float scalefac = (float)SRT_FFT * ((1.0f / 2.0f) / (3.0f / 8.0f));
        g_post[i] = rc * scalefac * 0.5f;
        g_sin[i] = (float)sin(w * i);                             /* stftFix.c:58-63 */
        unsigned r = 0, v = (unsigned)i;
        for (int b = 0; b < 12; b++) { r = (r << 1) | (v & 1); v >>= 1; }
        g_rev[i] = r;
    }
    g_tab_ready = 1;
}

const float *srt_oracle_prewindow(void) { tables_init(); return g_pre; }
const float *srt_oracle_postwindow(void) { tables_init(); return g_post; }

/* In-place 4096-point discrete Hartley transform of bit-reversed input: the radix-2
 * decimation-in-time scheme of DFT4096 (codelet.c:2-271) written as generic loops.
 * H[k] = sum_n a[n] cas(2 pi n k / N).  Self-inverse up to a factor N. */
static void dht4096(float *a)
{
    for (int len = 2; len <= SRT_FFT; len <<= 1) {
        const int half = len >> 1, quarter = len >> 2, tstep = SRT_FFT / len;
        for (int i = 0; i < SRT_FFT; i += len) {
            float p = a[i], q = a[i + half];
            a[i] = p + q; a[i + half] = p - q;
            if (quarter) {
                p = a[i + quarter]; q = a[i + half + quarter];
                a[i + quarter] = p + q; a[i + half + quarter] = p - q;
            }
            for (int j = 1; j < quarter; j++) {
                const float s = g_sin[j * tstep], c = g_sin[j * tstep + 1024];
                const float e = a[i + half + j], f = a[i + len - j];
                const float b1 = e * c + f * s, b2 = e * s - f * c;
                p = a[i + j]; q = a[i + half - j];
                a[i + j] = p + b1; a[i + half + j] = p - b1;
                a[i + half - j] = q + b2; a[i + len - j] = q - b2;
            }
        }
    }
}

/* stft (stftFix.c:363-495), single-thread branch :429-494.  Rows = ceil(n/1024); rows are
 * 4096 floats wide, bins 0..2048 written, everything else (incl. trailing rows) zero.
 * Output buffers are caller-allocated and must be zero-filled. */
size_t srt_oracle_stft(const float *L, const float *R, size_t n,
                       float *reL, float *imL, float *reR, float *imR)
{
    tables_init();
    const size_t rows = (n + SRT_HOP - 1) / SRT_HOP;
    const size_t rangeM = ((n - SRT_FFT + SRT_HOP / 4) / SRT_HOP) * SRT_HOP;   /* stftFix.c:377 */
    const size_t nfull = rangeM / SRT_HOP;
#pragma omp parallel
    {
        float *buf = (float *)malloc(sizeof(float) * 2 * SRT_FFT);
#pragma omp for schedule(static)
        for (size_t f = 0; f <= nfull; f++) {
            const size_t pos = f * SRT_HOP;
            float *b[2] = {buf, buf + SRT_FFT};
            for (int i = 0; i < SRT_FFT; i++) {
                const int ok = pos + i < n;       /* only the last frame can run past n (:460-473) */
                b[0][g_rev[i]] = ok ? L[pos + i] * g_pre[i] : 0.0f;
                b[1][g_rev[i]] = ok ? R[pos + i] * g_pre[i] : 0.0f;
            }
            dht4096(b[0]);
            dht4096(b[1]);
            float *re[2] = {reL + f * SRT_FFT, reR + f * SRT_FFT};
            float *im[2] = {imL + f * SRT_FFT, imR + f * SRT_FFT};
            for (int c = 0; c < 2; c++) {
                re[c][0] = b[c][0] * 2.0f;        /* stftFix.c:441-444 */
                im[c][0] = 0.0f;
                for (int k = 1; k < SRT_BINS; k++) {
                    re[c][k] = b[c][k] + b[c][SRT_FFT - k];   /* :448-455 */
                    im[c][k] = b[c][k] - b[c][SRT_FFT - k];
                }
            }
        }
        free(buf);
    }
    return rows;
}

/* istft (stftFix.c:496-579), single-thread branch :552-577.  Outputs caller-allocated,
 * zero-filled, length frames*1024 + 3072.  Inputs are not modified (the reference's
 * multi-thread branch clobbers them, :537-538; callers must not rely on either). */
size_t srt_oracle_istft(const float *reL, const float *imL, const float *reR, const float *imR,
                        size_t frames, float *outL, float *outR)
{
    tables_init();
    float *tmp = (float *)malloc(sizeof(float) * 2 * SRT_FFT * (frames ? frames : 1));
#pragma omp parallel for schedule(static)
    for (size_t f = 0; f < frames; f++) {
        const float *re[2] = {reL + f * SRT_FFT, reR + f * SRT_FFT};
        const float *im[2] = {imL + f * SRT_FFT, imR + f * SRT_FFT};
        for (int c = 0; c < 2; c++) {
            float *h = tmp + (2 * f + c) * SRT_FFT;
            h[0] = re[c][0];
            for (int j = 1; j < SRT_BINS; j++) {
                h[g_rev[j]] = re[c][j] + im[c][j];               /* :563-566 */
                h[g_rev[SRT_FFT - j]] = re[c][j] - im[c][j];
            }
            dht4096(h);
        }
    }
    for (size_t f = 0; f < frames; f++)                           /* serial OLA, frame order (:570-575) */
        for (int p = 0; p < SRT_FFT; p++) {
            outL[f * SRT_HOP + p] += tmp[(2 * f) * SRT_FFT + p] * g_post[p];
            outR[f * SRT_HOP + p] += tmp[(2 * f + 1) * SRT_FFT + p] * g_post[p];
        }
    free(tmp);
    return frames * SRT_HOP + (SRT_FFT - SRT_HOP);
}

/* ------------------------------------------------------------------------------------
 * Tile driver = processMT single-thread branch (main.c:447-541) generalised to nStems nets
 * that each mask the same mixture spectrum (as the VST's 4 nets do, Spleeter4Stems.c:135),
 * with the CLI's host framing (main.c:762-767: 4096 zeros in front, padded length
 * 4096*ceil(n/4096)+8192) and un-framing (channel_joinFloat preshift 4096, main.c:806).
 * stems: nStems*2 planar outputs of n samples, order [stem][L,R].
 * masks_out (optional): nStems * nTiles * 2*T*F floats.
 * ---------------------------------------------------------------------------------- */
size_t srt_oracle_padded_len(size_t n) { return (size_t)SRT_FFT * ((n + SRT_FFT - 1) / SRT_FFT) + 2 * SRT_FFT; }

int srt_oracle_separate(const float *const *coeffs, const int *stemModes, int nStems, int flavour,
                        const float *pcmL, const float *pcmR, size_t n, int T, int F,
                        float unaffectedWeight, float *const *stems, float *masks_out)
{
    const size_t padded = srt_oracle_padded_len(n);
    const size_t frames = padded / SRT_HOP;
    float *pl = (float *)calloc(padded, sizeof(float)), *pr = (float *)calloc(padded, sizeof(float));
    memcpy(pl + SRT_FFT, pcmL, n * sizeof(float));
    memcpy(pr + SRT_FFT, pcmR, n * sizeof(float));
    float *spec[4];
    for (int i = 0; i < 4; i++) spec[i] = (float *)calloc(frames * SRT_FFT, sizeof(float));
    srt_oracle_stft(pl, pr, padded, spec[0], spec[1], spec[2], spec[3]);
    const size_t P = (size_t)T * F;
    const size_t tiles = (frames + T - 1) / T;
    float *mag = (float *)malloc(sizeof(float) * 2 * P);
    float *mask = (float *)malloc(sizeof(float) * 2 * P);
    float *ms[4];
    for (int i = 0; i < 4; i++) ms[i] = (float *)malloc(sizeof(float) * frames * SRT_FFT);
    const size_t outLen = frames * SRT_HOP + (SRT_FFT - SRT_HOP);
    float *oL = (float *)malloc(sizeof(float) * outLen), *oR = (float *)malloc(sizeof(float) * outLen);
    for (int s = 0; s < nStems; s++) {
        for (int i = 0; i < 4; i++) memcpy(ms[i], spec[i], sizeof(float) * frames * SRT_FFT);
        for (size_t j = 0; j < tiles; j++) {
            const size_t f0 = j * T;
            for (int t = 0; t < T; t++)
                for (int i = 0; i < F; i++) {
                    const size_t idx = (f0 + t) * SRT_FFT + i;
                    const int live = f0 + t < frames;            /* tail tile rows zeroed, main.c:507-514 */
                    mag[0 * P + (size_t)t * F + i] = live ? hypotf(spec[0][idx], spec[1][idx]) * (float)SRT_FFT : 0.0f;
                    mag[1 * P + (size_t)t * F + i] = live ? hypotf(spec[2][idx], spec[3][idx]) * (float)SRT_FFT : 0.0f;
                }
            srt_oracle_unet(coeffs[s], F, T, stemModes[s], flavour, mag, mask, NULL);
            if (masks_out) memcpy(masks_out + ((size_t)s * tiles + j) * 2 * P, mask, sizeof(float) * 2 * P);
            for (int t = 0; t < T && f0 + t < frames; t++) {
                const size_t off = (f0 + t) * SRT_FFT;
                for (int i = 0; i < F; i++) {                    /* main.c:476-485 */
                    const float mL = mask[0 * P + (size_t)t * F + i], mR = mask[1 * P + (size_t)t * F + i];
                    ms[0][off + i] *= mL; ms[1][off + i] *= mL;
                    ms[2][off + i] *= mR; ms[3][off + i] *= mR;
                }
                for (int i = F; i < SRT_BINS; i++)               /* main.c:486-493 */
                    for (int q = 0; q < 4; q++) ms[q][off + i] *= unaffectedWeight;
            }
        }
        memset(oL, 0, sizeof(float) * outLen);
        memset(oR, 0, sizeof(float) * outLen);
        srt_oracle_istft(ms[0], ms[1], ms[2], ms[3], frames, oL, oR);
        memcpy(stems[2 * s + 0], oL + SRT_FFT, n * sizeof(float));
        memcpy(stems[2 * s + 1], oR + SRT_FFT, n * sizeof(float));
    }
    for (int i = 0; i < 4; i++) { free(spec[i]); free(ms[i]); }
    free(mag); free(mask); free(oL); free(oR); free(pl); free(pr);
    return 0;
}

/* ------------------------------------------------------------------------------------
 * One net over a whole spectrum, in place = processMT's single-thread branch (main.c:447-541):
 * per T-frame tile magnitude (:459-471), U-Net, mask multiply below binLimit (:476-485),
 * unaffectedWeight above (:486-493); tail tile zero-padded (:507-514).
 * ---------------------------------------------------------------------------------- */
static void net_over_spectrum(const float *coeff, int stemMode, int flavour, float *const sp[4], size_t frames,
                              int T, int F, float unaffectedWeight)
{
    const size_t P = (size_t)T * F;
    const size_t tiles = (frames + T - 1) / T;
    float *mag = (float *)malloc(sizeof(float) * 2 * P);
    float *mask = (float *)malloc(sizeof(float) * 2 * P);
    for (size_t j = 0; j < tiles; j++) {
        const size_t f0 = j * T;
        for (int t = 0; t < T; t++)
            for (int i = 0; i < F; i++) {
                const size_t idx = (f0 + t) * SRT_FFT + i;
                const int live = f0 + t < frames;
                mag[0 * P + (size_t)t * F + i] = live ? hypotf(sp[0][idx], sp[1][idx]) * (float)SRT_FFT : 0.0f;
                mag[1 * P + (size_t)t * F + i] = live ? hypotf(sp[2][idx], sp[3][idx]) * (float)SRT_FFT : 0.0f;
            }
        srt_oracle_unet(coeff, F, T, stemMode, flavour, mag, mask, NULL);
        for (int t = 0; t < T && f0 + t < frames; t++) {
            const size_t off = (f0 + t) * SRT_FFT;
            for (int i = 0; i < F; i++) {
                const float mL = mask[0 * P + (size_t)t * F + i], mR = mask[1 * P + (size_t)t * F + i];
                sp[0][off + i] *= mL; sp[1][off + i] *= mL;
                sp[2][off + i] *= mR; sp[3][off + i] *= mR;
            }
            for (int i = F; i < SRT_BINS; i++)
                for (int q = 0; q < 4; q++) sp[q][off + i] *= unaffectedWeight;
        }
    }
    free(mag); free(mask);
}

/* ------------------------------------------------------------------------------------
 * The CLI's two output modes around the tile driver (main.c:776-970), with its host framing.
 *   nOut == 2 (main.c:777-796): vocal = istft(mask * spec) with coeffs[0]; accompaniment = input - vocal
 *             in the time domain.  stems = {vocal L, R, accompaniment L, R}.
 *   nOut == 3 (main.c:845-936): drum net (coeffs[0]) masks the mixture; the residual spectrum
 *             orig - drum (:859-865) is inverse-transformed (accompaniment + vocal, :879) and also
 *             fed to the vocal net (coeffs[1], :911); accompaniment = that sum - vocal in the time
 *             domain (:923-927).  stems = {drum L, R, vocal L, R, accompaniment L, R}.
 * ---------------------------------------------------------------------------------- */
int srt_oracle_separate_cli(const float *const *coeffs, const int *stemModes, int nOut,
                            const float *pcmL, const float *pcmR, size_t n, int T, int F,
                            float unaffectedWeight, float *const *stems)
{
    if (nOut != 2 && nOut != 3) return -1;
    const size_t padded = srt_oracle_padded_len(n);
    const size_t frames = padded / SRT_HOP;
    const size_t plane = frames * SRT_FFT;
    float *pl = (float *)calloc(padded, sizeof(float)), *pr = (float *)calloc(padded, sizeof(float));
    memcpy(pl + SRT_FFT, pcmL, n * sizeof(float));
    memcpy(pr + SRT_FFT, pcmR, n * sizeof(float));
    float *a[4], *b[4];
    for (int i = 0; i < 4; i++) { a[i] = (float *)calloc(plane, sizeof(float)); b[i] = (float *)calloc(plane, sizeof(float)); }
    srt_oracle_stft(pl, pr, padded, a[0], a[1], a[2], a[3]);
    const size_t outLen = frames * SRT_HOP + (SRT_FFT - SRT_HOP);
    float *o[3][2];
    for (int k = 0; k < 3; k++)
        for (int c = 0; c < 2; c++) o[k][c] = (float *)calloc(outLen, sizeof(float));
    if (nOut == 2) {
        net_over_spectrum(coeffs[0], stemModes[0], 0, a, frames, T, F, unaffectedWeight);
        srt_oracle_istft(a[0], a[1], a[2], a[3], frames, o[0][0], o[0][1]);
        for (size_t i = 0; i < n; i++) {                         /* main.c:790-794 then channel_joinFloat's preshift */
            stems[0][i] = o[0][0][SRT_FFT + i];
            stems[1][i] = o[0][1][SRT_FFT + i];
            stems[2][i] = pl[SRT_FFT + i] - o[0][0][SRT_FFT + i];
            stems[3][i] = pr[SRT_FFT + i] - o[0][1][SRT_FFT + i];
        }
    } else {
        for (int q = 0; q < 4; q++) memcpy(b[q], a[q], sizeof(float) * plane);   /* orig_* (:849-856) */
        net_over_spectrum(coeffs[0], stemModes[0], 0, a, frames, T, F, unaffectedWeight);   /* drum (:858) */
        for (int q = 0; q < 4; q++)
            for (size_t i = 0; i < plane; i++) b[q][i] = b[q][i] - a[q][i];       /* residual (:860-865) */
        srt_oracle_istft(a[0], a[1], a[2], a[3], frames, o[0][0], o[0][1]);       /* drum (:867) */
        srt_oracle_istft(b[0], b[1], b[2], b[3], frames, o[2][0], o[2][1]);       /* accompaniment + vocal (:881) */
        net_over_spectrum(coeffs[1], stemModes[1], 0, b, frames, T, F, unaffectedWeight);   /* vocal (:911) */
        srt_oracle_istft(b[0], b[1], b[2], b[3], frames, o[1][0], o[1][1]);       /* (:914) */
        for (size_t i = 0; i < n; i++)
            for (int c = 0; c < 2; c++) {
                stems[0 + c][i] = o[0][c][SRT_FFT + i];
                stems[2 + c][i] = o[1][c][SRT_FFT + i];
                stems[4 + c][i] = o[2][c][SRT_FFT + i] - o[1][c][SRT_FFT + i];    /* (:923-927) */
            }
    }
    for (int i = 0; i < 4; i++) { free(a[i]); free(b[i]); }
    for (int k = 0; k < 3; k++)
        for (int c = 0; c < 2; c++) free(o[k][c]);
    free(pl); free(pr);
    return 0;
}

/* half -> float with denormals flushed to zero (f32Decompress, main.c:423-434). */
void srt_oracle_half_to_float(const uint16_t *in, float *out, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        uint32_t h = in[i], mag = (h & 0x7fffu) << 13, sign = (h & 0x8000u) << 16;
        uint32_t bits = ((h & 0x7c00u) == 0) ? 0u : mag + 0x38000000u;
        bits |= sign;
        memcpy(&out[i], &bits, 4);
    }
}

/* ====================================================================================
 * Real-time streaming flavour: restatement of VST/Source/Spleeter4Stems.c.
 * Same state machine, written single-threaded: the reference's NN threads are joined before
 * their masks are used (Spleeter4Stems.c:357-360), so running the nets synchronously at the tile
 * boundary gives identical output.  VST flavour of the net (exact sigmoid, unclamped ELU).
 * ================================================================================== */
typedef struct {
    int S, T, F;
    const float **coeff;
    float uw[8];
    float awin[SRT_FFT], swin[SRT_FFT];
    float in[2][SRT_FFT];
    int in_pos, need, cursor, ofp;
    float *spec[2][4];              /* [buffer][reL, imL, reR, imR][T][2049]   (:423-438) */
    float *mag;                     /* [2][T][F]                                (:419-420) */
    float *mask[2][8];              /* [buffer][stem][2][T][F], start at 1.0    (:447-466) */
    float *overlap;                 /* [2S][1024]  mOverlapStage2dash */
    float *ready;                   /* finished hop blocks, interleaved like mOutputBuffer: [1024][2S] */
    int n_ready, read_off;          /* number of queued blocks (each 1024 * 2S floats), read offset in the first */
    int cap_ready;
} orc_vst;

/* getAsymmetricWindow(k = 4096, m = 1024, freq_temporal = 1) + analysis scaling (Spleeter4Stems.c:383-416) */
static void vst_windows(float *an, float *sy)
{
    const int k = SRT_FFT, m = 1024;
    memset(sy, 0, sizeof(float) * k);
    int n = ((k - m) << 1) + 2;
    for (int i = 0; i < k - m; ++i) an[i] = (float)pow(0.5 * (1.0 - cos(2.0 * M_PI * (i + 1.0) / (double)n)), 1.0);
    n = (m << 1) + 2;
    for (int i = k - m; i < k; ++i) an[i] = (float)pow(sqrt(0.5 * (1.0 - cos(2.0 * M_PI * ((m + i - (k - m)) + 1.0) / (double)n))), 1.0);
    n = m << 1;
    for (int i = k - (m << 1); i < k; ++i) sy[i] = (float)(0.5 * (1.0 - cos(2.0 * M_PI * (double)(i - (k - (m << 1))) / (double)n))) / an[i];
    for (int i = 0; i < k - 2048; i++) sy[i] = sy[i + 2048];          /* pre-shift by SAMPLESHIFT (:399-400) */
    for (int i = 0; i < k; i++) an[i] *= (1.0 / SRT_FFT) * 0.5f;      /* :415-416 */
}

void *srt_oracle_vst_create(const float *const *coeffs, int nStems, int T, int F, const float *unaffected)
{
    tables_init();
    orc_vst *v = (orc_vst *)calloc(1, sizeof(orc_vst));
    v->S = nStems; v->T = T; v->F = F;
    v->coeff = (const float **)malloc(sizeof(float *) * nStems);
    static const float kUw[4] = {0.25f, 0.0f, 0.25f, 0.25f};          /* :73, :281 */
    for (int s = 0; s < nStems; s++) { v->coeff[s] = coeffs[s]; v->uw[s] = unaffected ? unaffected[s] : (s < 4 ? kUw[s] : 0.25f); }
    vst_windows(v->awin, v->swin);
    v->need = SRT_HOP;
    for (int b = 0; b < 2; b++) {
        for (int q = 0; q < 4; q++) v->spec[b][q] = (float *)calloc((size_t)T * SRT_BINS, sizeof(float));
        for (int s = 0; s < nStems; s++) {
            v->mask[b][s] = (float *)malloc(sizeof(float) * 2 * T * F);
            for (size_t i = 0; i < (size_t)2 * T * F; i++) v->mask[b][s][i] = 1.0f;
        }
    }
    v->mag = (float *)calloc((size_t)2 * T * F, sizeof(float));
    v->overlap = (float *)calloc((size_t)2 * nStems * 1024, sizeof(float));
    v->cap_ready = 4;
    v->ready = (float *)malloc(sizeof(float) * v->cap_ready * 1024 * 2 * nStems);
    return v;
}

void srt_oracle_vst_destroy(void *h)
{
    orc_vst *v = (orc_vst *)h;
    for (int b = 0; b < 2; b++) {
        for (int q = 0; q < 4; q++) free(v->spec[b][q]);
        for (int s = 0; s < v->S; s++) free(v->mask[b][s]);
    }
    free(v->mag); free(v->overlap); free(v->ready); free((void *)v->coeff); free(v);
}

/* LLPAMSProcessNPR (Spleeter4Stems.c:257-379) */
static void vst_hop(orc_vst *v)
{
    const int S = v->S, T = v->T, F = v->F, C = 2 * S;
    float fl[SRT_FFT], fr[SRT_FFT];
    float *td = (float *)malloc(sizeof(float) * C * SRT_FFT);
    for (int i = 0; i < SRT_FFT; i++) {                                /* :261-267 */
        const int k = (i + v->in_pos) & (SRT_FFT - 1);
        fl[g_rev[i]] = v->in[0][k] * v->awin[i];
        fr[g_rev[i]] = v->in[1][k] * v->awin[i];
    }
    float **sp = v->spec[v->ofp];
    const size_t row = (size_t)SRT_BINS * v->cursor;
    for (int s = 0; s < S; s++) {                                      /* :270-297 and task_type1 :64-89 */
        const float *mk = v->mask[v->ofp][s];
        for (int c = 0; c < 2; c++) {
            float *h = td + (size_t)(2 * s + c) * SRT_FFT;
            const float *re = sp[2 * c] + row, *im = sp[2 * c + 1] + row;
            const float *m = mk + (size_t)c * T * F + (size_t)F * v->cursor;
            h[0] = re[0] * m[0];
            for (int i = 1; i < SRT_BINS; i++) {
                const float w = i < F ? m[i] : v->uw[s];
                h[g_rev[i]] = (re[i] + im[i]) * w;
                h[g_rev[SRT_FFT - i]] = (re[i] - im[i]) * w;
            }
            dht4096(h);
            for (int i = 0; i < SRT_FFT - 2048; i++) h[i] = h[i + 2048] * v->swin[i];   /* :303-309 */
        }
    }
    dht4096(fl);
    dht4096(fr);
    if (v->n_ready == v->cap_ready) {
        v->cap_ready *= 2;
        v->ready = (float *)realloc(v->ready, sizeof(float) * v->cap_ready * 1024 * C);
    }
    float *ob = v->ready + (size_t)v->n_ready * 1024 * C;              /* :311-320 */
    v->n_ready++;
    for (int i = 0; i < 1024; i++)
        for (int j = 0; j < C; j++) {
            ob[(size_t)i * C + j] = v->overlap[(size_t)j * 1024 + i] + td[(size_t)j * SRT_FFT + i];
            v->overlap[(size_t)j * 1024 + i] = td[(size_t)j * SRT_FFT + 1024 + i];
        }
    /* spectral analysis of the new frame (:322-349) */
    const float *hb[2] = {fl, fr};
    for (int c = 0; c < 2; c++) {
        float *re = sp[2 * c] + row, *im = sp[2 * c + 1] + row;
        float *mg = v->mag + (size_t)c * T * F + (size_t)F * v->cursor;
        re[0] = hb[c][0] * 2.0f;
        mg[0] = fabsf(re[0]) * (float)SRT_FFT;
        for (int i = 1; i < SRT_BINS; i++) {
            re[i] = hb[c][i] + hb[c][SRT_FFT - i];
            im[i] = hb[c][i] - hb[c][SRT_FFT - i];
            if (i < F) mg[i] = hypotf(re[i], im[i]) * (float)SRT_FFT;
        }
    }
    if (++v->cursor >= T) {                                            /* :351-371 */
        v->ofp = !v->ofp;
        for (int s = 0; s < S; s++) srt_oracle_unet(v->coeff[s], F, T, 1, 1, v->mag, v->mask[!v->ofp][s], NULL);
        v->cursor = 0;
    }
    v->need = SRT_HOP;
    free(td);
}

/* Spleeter4StemsProcessSamples (Spleeter4Stems.c:512-582); components: 2S planar pointers */
void srt_oracle_vst_process(void *h, const float *inL, const float *inR, int n, float *const *components)
{
    orc_vst *v = (orc_vst *)h;
    const int C = 2 * v->S, want = n;
    while (n > 0) {
        const int c = v->need < n ? v->need : n;
        for (int i = 0; i < c; i++) {
            v->in[0][(v->in_pos + i) & (SRT_FFT - 1)] = inL[i];
            v->in[1][(v->in_pos + i) & (SRT_FFT - 1)] = inR[i];
        }
        inL += c; inR += c; n -= c;
        v->in_pos = (v->in_pos + c) & (SRT_FFT - 1);
        v->need -= c;
        if (v->need == 0) vst_hop(v);
    }
    int done = 0;
    while (v->n_ready > 0 && done < want) {
        const int c = (1024 - v->read_off) < (want - done) ? (1024 - v->read_off) : (want - done);
        for (int i = 0; i < c; i++)
            for (int j = 0; j < C; j++) components[j][done + i] = v->ready[(size_t)(v->read_off + i) * C + j];
        done += c;
        v->read_off += c;
        if (v->read_off == 1024) {
            v->read_off = 0;
            v->n_ready--;
            memmove(v->ready, v->ready + (size_t)1024 * C, sizeof(float) * (size_t)v->n_ready * 1024 * C);
        }
    }
}

/* =========================================================================================
 * Sample-rate conversion in front of the path.
 *
 * main.c:264-274 resamples the decoded file to 44.1 kHz with JamesDSPOfflineResampling (main.c:209-224) =
 * libsamplerate's src_simple() (samplerate.c:427-441, end_of_input = 1) on the sinc interpolator of
 * Executable/libsamplerate/src_sinc.c with the coefficient table the host supplies (`decompressedCoefficients`,
 * src_sinc.c:141-143: 22438 floats, half length 22436, 491 table steps per input sample).
 *
 * Restated without the streaming ring buffer: samples are addressed by absolute index (zero outside the input),
 * and only the INTEGER bookkeeping of the buffer (b_current / b_end / b_real_end, prepare_data src_sinc.c:1102-1172)
 * is carried along, because it decides at which output frame the converter stops (src_sinc.c:312-328, 467-482:
 * mono stops on `>`, stereo on `>=`, both comparing against the end of the data in BUFFER coordinates).
 * Per output frame (src_sinc.c:330-347 / 484-500, calc_output_single :218-271, calc_output_stereo :366-420):
 *     inc   = lrint(index_inc * min(ratio, 1) * 4096)                    12-bit fixed point
 *     start = lrint(frac * index_inc * min(ratio, 1) * 4096)             frac = fractional input position
 *     left  = sum over the taps at table positions start + j*inc (far -> near) of x[b - j]
 *     right = sum over the taps at table positions inc - start + j*inc (far -> near) of x[b + 1 + j]
 *     y = (float)(min(ratio, 1) * (left + right))       taps interpolated linearly between table entries, sums in double
 * then frac += 1/ratio, carried into b with fmod_one (common.h:137-145).
 * ========================================================================================= */
typedef struct {
    const float *x;
    long n_floats;      /* n_in * ch */
    int ch;
} rs_src;

static inline double rs_sample(const rs_src *s, long i)
{
    return (i >= 0 && i < s->n_floats) ? (double)s->x[i] : 0.0;
}

static inline double rs_tap(const float *coeffs, int32_t fidx)
{
    const double fraction = (double)(fidx & 4095) * (1.0 / 4096.0);
    const int k = fidx >> 12;
    return coeffs[k] + fraction * (coeffs[k + 1] - coeffs[k]);     /* float difference, double interpolation */
}

static double rs_fmod_one(double v)
{
    const double r = v - (double)lrint(v);
    return r < 0.0 ? r + 1.0 : r;
}

/* Returns the number of output frames written (<= n_out); frames beyond stay untouched (the caller's buffer is
 * zero-filled in main.c:267-268).  Returns -1 for a ratio libsamplerate rejects (samplerate.c:144). */
long srt_oracle_resample(const float *in, long n_in, int ch, double ratio, const float *coeffs, int half_len, int index_inc,
                         float *out, long n_out)
{
    if (ratio < 1.0 / 256.0 || ratio > 256.0 || (ch != 1 && ch != 2)) return -1;
    const rs_src src = {in, n_in * ch, ch};
    /* buffer geometry (sinc_set_converter, src_sinc.c:150-153) */
    long b_len = 3 * lrint((half_len + 2.0) / index_inc * 256.0 + 1);
    if (b_len < 4096) b_len = 4096;
    b_len = b_len * ch + 1;
    double count = (half_len + 2.0) / index_inc;
    if (ratio < 1.0) count /= ratio;
    const long half = ch * (lrint(count) + 1);
    long b_cur = 0, b_end = 0, b_real_end = -1, in_used = 0;
    const long in_count = n_in * ch;
    long abs_cur = 0;                      /* absolute float index of the sample at b_cur */
    double frac = 0.0;
    const double terminate = 1.0 / ratio + 1e-20;
    const double scale_inc = index_inc * (ratio < 1.0 ? ratio : 1.0);
    const int32_t inc = (int32_t)lrint(scale_inc * 4096.0);
    const int32_t max_idx = (int32_t)half_len << 12;
    long gen = 0;
    while (gen < n_out) {
        long in_hand = (b_end - b_cur + b_len) % b_len;
        if (in_hand <= half) {
            /* ---- prepare_data, indices only (src_sinc.c:1102-1172) */
            if (b_real_end < 0 && in != NULL) {
                long len;
                if (b_cur == 0) {
                    len = b_len - 2 * half;
                    b_cur = b_end = half;
                } else if (b_end + half + ch < b_len) {
                    len = b_len - b_cur - half;
                    if (len < 0) len = 0;
                } else {
                    len = b_end - b_cur;
                    b_cur = half;
                    b_end = b_cur + len;
                    len = b_len - b_cur - half;
                    if (len < 0) len = 0;
                }
                if (len > in_count - in_used) len = in_count - in_used;
                len -= len % ch;
                b_end += len;
                in_used += len;
                if (in_used == in_count && b_end - b_cur < 2 * half) {      /* end_of_input is always set by src_simple */
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
        const int32_t start = (int32_t)lrint(frac * scale_inc * 4096.0);
        for (int c = 0; c < ch; c++) {
            /* left half: far tap first */
            int32_t fidx = start;
            int32_t n = (max_idx - fidx) / inc;
            fidx += n * inc;
            long di = abs_cur - (long)ch * n + c;
            double left = 0.0;
            do {
                left += rs_tap(coeffs, fidx) * rs_sample(&src, di);
                fidx -= inc;
                di += ch;
            } while (fidx >= 0);
            /* right half */
            fidx = inc - start;
            n = (max_idx - fidx) / inc;
            fidx += n * inc;
            di = abs_cur + (long)ch * (1 + n) + c;
            double right = 0.0;
            do {
                right += rs_tap(coeffs, fidx) * rs_sample(&src, di);
                fidx -= inc;
                di -= ch;
            } while (fidx > 0);
            out[gen * ch + c] = (float)((scale_inc / index_inc) * (left + right));
        }
        gen++;
        frac += 1.0 / ratio;
        const double rem = rs_fmod_one(frac);
        const long step = lrint(frac - rem);
        b_cur = (b_cur + ch * step) % b_len;
        abs_cur += ch * step;
        frac = rem;
    }
    return gen;
}
