// host_model.cpp — TEST INFRASTRUCTURE.  A CPU model of what the GPU gather-GEMM kernels do with
// the product's own k-block tables and packed weights (spleeterrt_b200/csrc/srt_plan.*): TMA box
// fetches with zero fill outside the tensor, 128B-swizzled weight blocks, tile decomposition,
// space-to-depth / phase-scatter epilogues.  It lets the CPU test-suite prove the host logic
// (tables, packing, layouts) against the oracle without a GPU.  It is never part of the product.
#include <cmath>
#include <cstring>
#include <vector>

#include "../spleeterrt_b200/csrc/srt_plan.h"

using namespace srt;

static bool g_split = false;   // two-term weights (weights not exactly representable in TF32)
extern "C" void srt_host_model_set_split(int on) { g_split = on != 0; }
static bool g_comp = false;    // compensated precision: sources rounded to TF32 + bf16 residual tensors contracted by extra k-blocks
static bool g_comp_drop = false;   // (control experiment) sources rounded, compensation blocks ignored
static int g_lo_fmt = LO_BF16;     // residual format asked for: LO_BF16, or LO_FP8 (layers with 64 residual channels stay bf16)
extern "C" void srt_host_model_set_comp(int on) { g_comp = on != 0; g_comp_drop = on == 2; g_lo_fmt = on == 3 ? LO_FP8 : LO_BF16; }
static bool g_fuse = true;     // decoder layers with 4 * cout <= 256 fuse their four output parities into N (the product's default)
extern "C" void srt_host_model_set_fuse(int on) { g_fuse = on != 0; }
static int g_min_ctas = 0;      // build_plans(min_ctas): narrower N tiles for small grids (what a small-batch context uses)
extern "C" void srt_host_model_set_min_ctas(int v) { g_min_ctas = v; }

static float act_apply(int act, float x)
{
    switch (act) {
    case ACT_LEAKY: return x >= 0.f ? x : 0.2f * x;
    case ACT_RELU: return x >= 0.f ? x : 0.f;
    case ACT_ELU_CLAMP: return x >= 0.f ? x : (x < -15.f ? -1.f : std::exp(x) - 1.f);
    case ACT_ELU: return x >= 0.f ? x : std::exp(x) - 1.f;
    default: return x;
    }
}


// What the producing epilogue stores in compensated mode: hi = tf32(a) in place, residual bf16(a - hi) in the layer's
// residual tensor (encoder: same layout; decoder: channel q's block of the concatenated [skip | up] tensor).
// `lo` holds the operand values the compensation MMAs read: bf16(a - hi), or e5m2(4 (a - hi)) in the 8-bit format.
static void split_sources(std::vector<std::vector<float>>& src, int nsrc, const SrcDesc* sd, int H, int W, const SrcDesc& lo_sd, int lo_fmt,
                          std::vector<float>& lo)
{
    lo.assign((size_t)H * W * lo_sd.C, 0.f);
    int coff = 0;
    for (int q = 0; q < nsrc; q++) {
        const int C = sd[q].C;
        for (size_t px = 0; px < (size_t)H * W; px++)
            for (int c = 0; c < C; c++) {
                float& a = src[q][px * C + c];
                const float hi = round_tf32(a);
                lo[px * lo_sd.C + coff + c] = (lo_fmt == LO_FP8 || lo_fmt == LO_FP8N) ? e5m2_to_float(e5m2_rn(4.0f * (a - hi))) : bf16_to_float(bf16_rn(a - hi));
                a = hi;
            }
        coff += C;
    }
}
// weight of a compensation block as the MMA reads it
static float lo_weight(const KBlock& kb, const float* wb, int n, int j)
{
    if (kb.part & kPartLo8n) return e5m2_to_float(reinterpret_cast<const uint8_t*>(wb)[swz64_index8(n, j)]);
    if (kb.part & kPartLo8) return e5m2_to_float(reinterpret_cast<const uint8_t*>(wb)[swz128_index8(n, j)]);
    return bf16_to_float(reinterpret_cast<const uint16_t*>(wb)[swz128_index16(n, j)]);
}

// src0/src1: planar [C][H][W] fp32 in the *reference's* layout:
//   encoder: src0 = activated input of the layer at resolution (2*Hs, 2*Ws)
//   decoder: src0 = skip, src1 = previous decoder output, both at (Hs, Ws); layer up1 has src0 only
// out: planar [cout][Hout][Wout]; encoder: raw (conv + bias) when want_act == 0, else act(scale*raw+offset)
//      decoder: scale*act(v + bias) + offset
extern "C" int srt_host_model_layer(int T, int F, int plan_index, const float* coeff, int act, const float* src0, const float* src1,
                                    float* out, int want_act)
{
    std::vector<LayerPlan> plans = build_plans(NetGeom{T, F}, 1, g_split, 1, g_min_ctas, g_comp ? 0x3ffu : 0u, g_fuse, g_lo_fmt);
    if (plan_index < 0 || plan_index >= (int)plans.size()) return -1;
    const LayerPlan& L = plans[plan_index];
    if (L.comp != g_comp) return -5;
    const CoeffLayout cl = coeff_layout();
    // ---- device layouts of the sources -------------------------------------------------
    std::vector<std::vector<float>> src(L.nsrc);
    const int H = L.Hs, W = L.Ws;
    if (!L.transposed) {
        const int cin = L.cin, Hi = 2 * H, Wi = 2 * W;
        src[0].assign((size_t)H * W * 4 * cin, 0.f);
        for (int c = 0; c < cin; c++)
            for (int y = 0; y < Hi; y++)
                for (int x = 0; x < Wi; x++)
                    src[0][(((size_t)(y / 2) * W + x / 2) * 4 + (y & 1) * 2 + (x & 1)) * cin + c] = src0[((size_t)c * Hi + y) * Wi + x];
    } else {
        const float* in[2] = {src0, src1};
        for (int q = 0; q < L.nsrc; q++) {
            const int C = L.src[q].C;
            src[q].assign((size_t)H * W * C, 0.f);
            for (int c = 0; c < C; c++)
                for (int y = 0; y < H; y++)
                    for (int x = 0; x < W; x++) src[q][((size_t)y * W + x) * C + c] = in[q][((size_t)c * H + y) * W + x];
        }
    }
    std::vector<float> lo;
    if (L.comp) split_sources(src, L.nsrc, L.src, H, W, L.lo_src, L.lo_fmt, lo);
    // ---- weights and epilogue vectors ----------------------------------------------------
    std::vector<float> wpk(L.w_floats_per_stem);
    pack_layer(L, coeff, wpk.data());
    const float* bias = coeff + (L.transposed ? cl.up_b[L.index - 5] : cl.down_b[L.index + 1]);
    const float* bn = coeff + (L.transposed ? cl.up_bn[L.index - 5] : cl.down_bn[L.index + 1]);
    const bool has_bn = L.transposed || L.index < 4;
    const int Hout = L.transposed ? 2 * H : H, Wout = L.transposed ? 2 * W : W;
    // ---- tiles ------------------------------------------------------------------------------
    const int tiles_x = (W + L.tw - 1) / L.tw, tiles_y = (H + L.th - 1) / L.th;
    std::vector<float> acc(L.n_tile);
    for (int ph = 0; ph < L.phases; ph++) {
        const size_t nkb = L.kb[ph].size();
        for (int nt = 0; nt < L.n_tiles; nt++)
            for (int ty = 0; ty < tiles_y; ty++)
                for (int tx = 0; tx < tiles_x; tx++)
                    for (int m = 0; m < kTileM; m++) {
                        const int x = m % L.tw, y = (m / L.tw) % L.th, nn = m / (L.tw * L.th);
                        const int X = tx * L.tw + x, Y = ty * L.th + y;
                        if (X >= W || Y >= H || nn >= 1) continue;   // masked rows of the tile
                        std::fill(acc.begin(), acc.end(), 0.f);
                        for (size_t k = 0; k < nkb; k++) {
                            const KBlock kb = L.kb[ph][k];
                            const int yy = Y + kb.dy, xx = X + kb.dx;
                            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;   // TMA zero fill
                            const float* wb = &wpk[L.w_phase_off[ph] + ((size_t)nt * nkb + k) * L.n_tile * kKB];
                            if (kb.part & kPartLo) {                                // compensation block: bf16 residuals x bf16 weights
                                const int width = kb_channels(kb);
                                if (!L.comp || kb.src != kSrcLo || kb.c_off < 0 || kb.c_off + width > L.lo_src.C) return -2;
                                if (((kb.part & kPartLo8) != 0) != (L.lo_fmt == LO_FP8)) return -7;
                                if (g_comp_drop) continue;
                                const float* a = &lo[((size_t)yy * W + xx) * L.lo_src.C + kb.c_off];
                                for (int n = 0; n < L.n_tile; n++) {
                                    float s = 0.f;
                                    for (int j = 0; j < width; j++) s += a[j] * lo_weight(kb, wb, n, j);
                                    acc[n] += s;
                                }
                                continue;
                            }
                            const int C = L.src[kb.src].C;
                            if (kb.c_off < 0 || kb.c_off + kKB > C) return -2;      // box must stay inside the channel dim
                            const float* a = &src[kb.src][((size_t)yy * W + xx) * C + kb.c_off];
                            for (int n = 0; n < L.n_tile; n++) {
                                float s = 0.f;
                                for (int j = 0; j < kKB; j++) s += a[j] * wb[swz128_index(n, j)];
                                acc[n] += s;
                            }
                        }
                        for (int n = 0; n < L.n_tile; n++) {
                            int o = nt * L.n_tile + n;
                            int phase = ph;
                            if (L.fused) { phase = o / L.cout; o %= L.cout; }      // column = parity * cout + channel
                            float v = acc[n] + bias[o];
                            if (L.transposed) {
                                v = bn[L.cout + o] * act_apply(act, v) + bn[o];
                                const int oy = 2 * Y + (phase >> 1), ox = 2 * X + (phase & 1);
                                out[((size_t)o * Hout + oy) * Wout + ox] = v;
                            } else {
                                if (want_act && has_bn) v = act_apply(act, bn[L.cout + o] * v + bn[o]);
                                out[((size_t)o * Hout + Y) * Wout + X] = v;
                            }
                        }
                    }
    }
    return 0;
}

extern "C" int srt_host_model_plan_info(int T, int F, int n_img, int plan_index, int* info /* tw,th,nb,n_tile,n_tiles,phases,nkb0..3 */)
{
    std::vector<LayerPlan> plans = build_plans(NetGeom{T, F}, n_img, false, 1, g_min_ctas, 0u, g_fuse);
    if (plan_index < 0 || plan_index >= (int)plans.size()) return -1;
    const LayerPlan& L = plans[plan_index];
    info[0] = L.tw; info[1] = L.th; info[2] = L.nb; info[3] = L.n_tile; info[4] = L.n_tiles; info[5] = L.phases;
    for (int p = 0; p < 4; p++) info[6 + p] = p < L.phases ? (int)L.kb[p].size() : 0;
    return 0;
}

extern "C" float srt_host_model_round_tf32(float x) { return round_tf32(x); }
extern "C" int srt_host_model_e5m2(float x) { return e5m2_rn(x); }
extern "C" float srt_host_model_e5m2_value(int b) { return e5m2_to_float((uint8_t)b); }

// ---- row-patch ("v2") form of down2 / down3 / up4 / up5 --------------------------------------
extern "C" int srt_host_model_row_layer(int T, int F, int plan_index, const float* coeff, int act, const float* src0, const float* src1,
                                        float* out, int want_act)
{
    if (!row_plan_supported(plan_index)) return -1;
    const RowPlan L = build_row_plan(NetGeom{T, F}, plan_index, g_split, g_comp, g_lo_fmt);
    const CoeffLayout cl = coeff_layout();
    const int H = L.Hs, W = L.Ws;
    std::vector<std::vector<float>> src(L.nsrc);
    if (!L.transposed) {
        const int cin = L.cin, Hi = 2 * H, Wi = 2 * W;
        src[0].assign((size_t)H * W * 4 * cin, 0.f);
        for (int c = 0; c < cin; c++)
            for (int y = 0; y < Hi; y++)
                for (int x = 0; x < Wi; x++)
                    src[0][(((size_t)(y / 2) * W + x / 2) * 4 + (y & 1) * 2 + (x & 1)) * cin + c] = src0[((size_t)c * Hi + y) * Wi + x];
    } else {
        const float* in[2] = {src0, src1};
        for (int q = 0; q < L.nsrc; q++) {
            const int C = L.src[q].C;
            src[q].assign((size_t)H * W * C, 0.f);
            for (int c = 0; c < C; c++)
                for (int y = 0; y < H; y++)
                    for (int x = 0; x < W; x++) src[q][((size_t)y * W + x) * C + c] = in[q][((size_t)c * H + y) * W + x];
        }
    }
    std::vector<float> lo;
    if (L.comp) split_sources(src, L.nsrc, L.src, H, W, L.lo_src, L.lo_fmt, lo);
    std::vector<float> wpk(L.w_floats_per_stem);
    pack_row_layer(L, coeff, wpk.data());
    const float* bias = coeff + (L.transposed ? cl.up_b[L.index - 5] : cl.down_b[L.index + 1]);
    const float* bn = coeff + (L.transposed ? cl.up_bn[L.index - 5] : cl.down_bn[L.index + 1]);
    const int Hout = L.transposed ? 2 * H : H, Wout = L.transposed ? 2 * W : W;
    // K steps the MMA issuer skips (kPartSkipShift) must hold nothing but zero weights, and the model below honours the mask
    for (size_t k = 0; k < L.kb.size(); k++) {
        const int steps = kb_ksteps(L.kb[k]), width = kb_channels(L.kb[k]), per = width / steps, skip = kb_skip_mask(L.kb[k]);
        const bool is_lo = (L.kb[k].part & kPartLo) != 0;
        for (int q = 0; q < steps; q++)
            if (skip & (1 << q))
                for (int n = 0; n < L.N; n++)
                    for (int j = q * per; j < (q + 1) * per; j++) {
                        const float* wb = &wpk[k * (size_t)L.N * kKB];
                        const float v = is_lo ? lo_weight(L.kb[k], wb, n, j) : wb[swz128_index(n, j)];
                        if (v != 0.0f) return -6;
                    }
    }
    std::vector<float> acc(L.N);
    // CTA tiles: R rows x 128 columns; the patch is rows y0-1 .. y0+R, columns x0-1 .. x0+134
    for (int y0 = 0; y0 < H; y0 += L.R)
        for (int x0 = 0; x0 < W; x0 += kTileM)
            for (int r = 0; r < L.R; r++)
                for (int m = 0; m < kTileM; m++) {
                    const int Y = y0 + r, X = x0 + m;
                    if (Y >= H || X >= W) continue;
                    std::fill(acc.begin(), acc.end(), 0.f);
                    for (const RowChunk& ch : L.chunks)
                        for (int k = ch.kb0; k < ch.kb0 + ch.nkb; k++) {
                            const KBlock kb = L.kb[k];
                            if (kb.src != ch.src || kb.c_off != ch.c_off) return -3;
                            // window position inside the patch must exist: rows r+dy+1 in [0, R+2), cols m+dx+1 in [0, 136)
                            if (r + kb.dy + 1 < 0 || r + kb.dy + 1 >= L.R + 2 || m + kb.dx + 1 < 0 || m + kb.dx + 1 >= kPatchW) return -4;
                            const int yy = Y + kb.dy, xx = X + kb.dx;
                            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                            const float* wb = &wpk[(size_t)k * L.N * kKB];
                            if (kb.part & kPartLo) {
                                const int width = kb_channels(kb);
                                if (!L.comp || ch.src != kSrcLo || kb.c_off + width > L.lo_src.C) return -2;
                                if (((kb.part & kPartLo8) != 0) != (L.lo_fmt == LO_FP8 || L.lo_fmt == LO_FP8N)) return -7;
                                if (((kb.part & kPartLo8n) != 0) != (L.lo_fmt == LO_FP8N)) return -7;
                                if (g_comp_drop) continue;
                                const float* a = &lo[((size_t)yy * W + xx) * L.lo_src.C + kb.c_off];
                                for (int n = 0; n < L.N; n++) {
                                    float s = 0.f;
                                    for (int j = 0; j < width; j++)
                                        if (!(kb_skip_mask(kb) & (1 << (j / (width / kb_ksteps(kb)))))) s += a[j] * lo_weight(kb, wb, n, j);
                                    acc[n] += s;
                                }
                                continue;
                            }
                            const int C = L.src[kb.src].C;
                            if (kb.c_off + kKB > C) return -2;
                            const float* a = &src[kb.src][((size_t)yy * W + xx) * C + kb.c_off];
                            for (int n = 0; n < L.N; n++) {
                                float s = 0.f;
                                for (int j = 0; j < kKB; j++)
                                    if (!(kb_skip_mask(kb) & (1 << (j / 8)))) s += a[j] * wb[swz128_index(n, j)];
                                acc[n] += s;
                            }
                        }
                    for (int n = 0; n < L.N; n++) {
                        const int ph = L.transposed ? n / L.cout : 0, o = L.transposed ? n % L.cout : n;
                        float v = acc[n] + bias[o];
                        if (L.transposed) {
                            v = bn[L.cout + o] * act_apply(act, v) + bn[o];
                            out[((size_t)o * Hout + 2 * Y + (ph >> 1)) * Wout + 2 * X + (ph & 1)] = v;
                        } else {
                            if (want_act) v = act_apply(act, bn[L.cout + o] * v + bn[o]);
                            out[((size_t)o * Hout + Y) * Wout + X] = v;
                        }
                    }
                }
    return 0;
}

// ---- down1 on the tensor cores: S2D magnitude (8 channels), 9 taps, several stems fused into N ----------
// mag: planar [2][T][F]; coeffs: nstems blobs; out: [nstems][16][T/2][F/2] raw conv + bias
extern "C" int srt_host_model_down1(int T, int F, const float* const* coeffs, int nstems, const float* mag, float* out)
{
    const Down1Plan L = build_down1_plan(NetGeom{T, F}, g_split);
    const CoeffLayout cl = coeff_layout();
    const int H = L.Hs, W = L.Ws, N = 16 * nstems;
    std::vector<float> src[2] = {std::vector<float>((size_t)H * W * 8, 0.f), std::vector<float>((size_t)H * W * 8, 0.f)};
    for (int c = 0; c < 2; c++)
        for (int t = 0; t < T; t++)
            for (int f = 0; f < F; f++) {
                const float m = mag[((size_t)c * T + t) * F + f], hi = round_tf32(m);
                src[0][mag_s2d_index(T, F, t, f) * 2 + c] = hi;
                src[1][mag_s2d_index(T, F, t, f) * 2 + c] = m - hi;
            }
    std::vector<float> wpk(L.kb.size() * (size_t)N * kKB1);
    pack_down1(L, coeffs, nstems, wpk.data());
    for (int Y = 0; Y < H; Y++)
        for (int X = 0; X < W; X++)
            for (int n = 0; n < N; n++) {
                float acc = 0.f;
                for (size_t k = 0; k < L.kb.size(); k++) {
                    const int yy = Y + L.kb[k].dy, xx = X + L.kb[k].dx;
                    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                    for (int j = 0; j < kKB1; j++) acc += src[L.kb[k].src][((size_t)yy * W + xx) * 8 + j] * wpk[k * (size_t)N * kKB1 + swz32_index(n, j)];
                }
                const int s = n / 16, o = n % 16;
                out[(((size_t)s * 16 + o) * H + Y) * W + X] = acc + coeffs[s][cl.down_b[0] + o];
            }
    return 0;
}

// ---- up6 as up6_tc_kernel computes it: GEMM over the PACKED weights, then the 25-value gather ---------------------------
// e1, u5: planar [16][H][W] (H = T/2, W = F/2); out: [T][F] = scale * act(tconv + bias) + offset.  lo8 = the residual term
// through e5m2 operands (one K = 32 contraction), else through fp32 residuals (TF32 operands).  The packed blocks are read back
// through the hardware's definition of SWIZZLE_32B (byte address bit 4 ^= bit 7), not through the packer's index helpers.
static inline size_t unswz32(size_t byte) { return byte ^ (((byte >> 7) & 1) << 4); }
static inline float trunc_tf32_host(float x)
{
    uint32_t u;
    std::memcpy(&u, &x, 4);
    u &= 0xffffe000u;
    std::memcpy(&x, &u, 4);
    return x;
}
extern "C" int srt_host_model_up6(int T, int F, const float* coeff, int act, const float* e1, const float* u5, float* out, int lo8)
{
    const CoeffLayout cl = coeff_layout();
    const int H = T / 2, W = F / 2;
    std::vector<float> pk(kUp6PackFloats);
    pack_up6_weights(coeff + cl.up_w[5], pk.data());
    const uint8_t* bytes = reinterpret_cast<const uint8_t*>(pk.data());
    auto w_tf32 = [&](int b, int term, int tap, int j) {
        float v;
        std::memcpy(&v, bytes + (size_t)(b * 2 + term) * 1024 + unswz32((size_t)tap * 32 + j * 4), 4);
        return v;
    };
    auto w_e5m2 = [&](int tap, int ch) { return e5m2_to_float(bytes[(size_t)8 * 1024 + unswz32((size_t)tap * 32 + ch)]); };
    std::vector<float> G((size_t)H * W * 32, 0.f);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            float a[32];
            for (int c = 0; c < 16; c++) {
                a[c] = e1[((size_t)c * H + y) * W + x];
                a[16 + c] = u5[((size_t)c * H + y) * W + x];
            }
            for (int tap = 0; tap < 32; tap++) {
                double acc = 0.0;
                for (int cin = 0; cin < 32; cin++) {
                    const int b = (cin >> 4) * 2 + ((cin >> 3) & 1), j = cin & 7;
                    const float hi = trunc_tf32_host(a[cin]), lo = a[cin] - hi;
                    acc += (double)hi * w_tf32(b, 0, tap, j);
                    if (g_split) acc += (double)hi * w_tf32(b, 1, tap, j);
                    if (lo8) acc += (double)e5m2_to_float(e5m2_rn(4.0f * lo)) * w_e5m2(tap, cin);
                    else acc += (double)trunc_tf32_host(lo) * w_tf32(b, 0, tap, j);
                }
                G[((size_t)y * W + x) * 32 + tap] = (float)acc;
            }
        }
    for (int tap = 25; tap < 32; tap++)
        for (size_t px = 0; px < (size_t)H * W; px++)
            if (G[px * 32 + tap] != 0.f) return -2;   // the padding taps must stay zero
    const float bias = coeff[cl.up_b[5]], offset = coeff[cl.up_bn[5]], scale = coeff[cl.up_bn[5] + 1];
    for (int yo = 0; yo < H; yo++)
        for (int X = 0; X < W; X++)
            for (int po = 0; po < 2; po++)
                for (int qo = 0; qo < 2; qo++) {
                    float o = 0.f;
                    for (int dy = -1; dy <= 1; dy++)
                        for (int dx = -1; dx <= 1; dx++) {
                            const int kh = po + 1 - 2 * dy, kw = qo + 1 - 2 * dx;
                            if (kh < 0 || kh > 4 || kw < 0 || kw > 4) continue;
                            const int yy = yo + dy, xx = X + dx;
                            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
                            o += G[((size_t)yy * W + xx) * 32 + kh * 5 + kw];
                        }
                    out[(size_t)(2 * yo + po) * F + 2 * X + qo] = scale * act_apply(act, o + bias) + offset;
                }
    return 0;
}
