// srt_plan.cpp — see srt_plan.h.  Plain host C++.
#include "srt_plan.h"

#include <cmath>
#include <cstring>

namespace srt {

static const int kEnc[7] = {2, 16, 32, 64, 128, 256, 512};
static const int kDecIn[6] = {512, 512, 256, 128, 64, 32};
static const int kDecOut[6] = {256, 128, 64, 32, 16, 1};

CoeffLayout coeff_layout()
{
    CoeffLayout L{};
    size_t p = 0;
    for (int i = 0; i < 6; i++) {
        L.down_w[i] = p; p += (size_t)25 * kEnc[i] * kEnc[i + 1];
        L.down_b[i] = p; p += kEnc[i + 1];
        L.down_bn[i] = p;
        if (i < 5) p += 2 * kEnc[i + 1];
    }
    for (int i = 0; i < 6; i++) {
        L.up_w[i] = p; p += (size_t)25 * kDecIn[i] * kDecOut[i];
        L.up_b[i] = p; p += kDecOut[i];
        L.up_bn[i] = p; p += 2 * kDecOut[i];
    }
    L.w7 = p; p += 32;
    L.b7 = p; p += 2;
    return L;
}

float round_tf32(float x)
{
    uint32_t u;
    std::memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;   // inf / nan untouched
    u = (u + 0x1000u) & ~0x1fffu;                      // round to nearest, ties away (cvt.rna.tf32)
    float r;
    std::memcpy(&r, &u, 4);
    return r;
}

uint16_t bf16_rn(float x)
{
    uint32_t u;
    std::memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return (uint16_t)(u >> 16);   // inf / nan
    u += 0x7fffu + ((u >> 16) & 1u);                                     // round to nearest even
    return (uint16_t)(u >> 16);
}
float bf16_to_float(uint16_t h)
{
    const uint32_t u = (uint32_t)h << 16;
    float r;
    std::memcpy(&r, &u, 4);
    return r;
}

// float -> e5m2 = the top byte of the IEEE half, rounded to nearest even on the dropped 8 bits; saturates to the largest finite value
uint8_t e5m2_rn(float x)
{
    if (x != x) return 0x7f;
    const float ax = x < 0 ? -x : x;
    const uint8_t sign = x < 0 || (x == 0 && 1.0f / x < 0) ? 0x80 : 0x00;
    if (ax >= 57344.0f) return sign | 0x7b;                  // max finite e5m2 (satfinite)
    // scale into the half format by hand: e5m2 has exponent bias 15, 2 mantissa bits, subnormals down to 2^-16
    if (ax < 7.62939453125e-06f) return sign;                 // < 2^-17: rounds to zero (2^-17 itself ties to even = 0)
    int e;
    const float m = std::frexp(ax, &e);                       // ax = m 2^e, m in [0.5, 1)
    int exp = e - 1;                                          // ax = (2m) 2^(e-1), 2m in [1, 2)
    float frac = 2.0f * m;
    int q;                                                    // quantised magnitude in units of the format's ulp
    if (exp < -14) {                                          // subnormal: ulp 2^-16
        const float u = ax * 65536.0f;                        // in units of 2^-16: < 4
        q = (int)u;
        const float r = u - (float)q;
        if (r > 0.5f || (r == 0.5f && (q & 1))) q++;
        return sign | (uint8_t)q;                             // q == 4 becomes the smallest normal (0x04): correct carry
    }
    const float u = (frac - 1.0f) * 4.0f;                     // mantissa in quarters: [0, 4)
    q = (int)u;
    const float r = u - (float)q;
    if (r > 0.5f || (r == 0.5f && (q & 1))) q++;
    if (q == 4) { q = 0; exp++; }
    if (exp > 15) return sign | 0x7b;
    return sign | (uint8_t)(((exp + 15) << 2) | q);
}
float e5m2_to_float(uint8_t b)
{
    const int e = (b >> 2) & 0x1f, m = b & 3;
    float v;
    if (e == 0) v = (float)m * 1.52587890625e-05f;            // m 2^-16
    else if (e == 31) v = m ? NAN : INFINITY;
    else v = std::ldexp(1.0f + 0.25f * (float)m, e - 15);
    return (b & 0x80) ? -v : v;
}

int layer_lo_format(int index, int want, bool row_patch)
{
    if (want != LO_FP8) return want;
    // residual channels per pixel: encoder 4 * cin (space-to-depth), decoder cin ([skip | up]); down2 = 64, up5 = 64
    const int C = index < 5 ? 4 * kEnc[index + 1] : kDecIn[index - 5];
    if (C % kKBlo8 == 0) return LO_FP8;
    return (row_patch && C == 64) ? LO_FP8N : LO_BF16;
}

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

static void choose_tile(int W, int H, int n_img, int& tw, int& th, int& nb)
{
    long best = -1;
    for (int w = 128; w >= 1; w >>= 1)
        for (int h = 128 / w; h >= 1; h >>= 1) {
            int b = 128 / (w * h);
            long cost = (long)ceil_div(W, w) * w * ceil_div(H, h) * h * ceil_div(n_img, b) * b;
            if (best < 0 || cost < best) { best = cost; tw = w; th = h; nb = b; }
        }
}

// stride-2 tap kh of the encoder -> (S2D row offset, row parity):  input row 2*oh + kh - 1
static void enc_tap(int k, int& d, int& par)
{
    static const int dd[5] = {-1, 0, 0, 1, 1}, pp[5] = {1, 0, 1, 0, 1};
    d = dd[k];
    par = pp[k];
}

// transposed-conv taps for output parity `par`: output o = 2h + kh - 1 (im2col_dilated.c:57-58)
static int dec_taps(int par, int kh[3], int dy[3])
{
    if (par == 0) { kh[0] = 1; dy[0] = 0; kh[1] = 3; dy[1] = -1; return 2; }
    kh[0] = 0; dy[0] = 1; kh[1] = 2; dy[1] = 0; kh[2] = 4; dy[2] = -1;
    return 3;
}

// Encoder, space-to-depth source [py][px][cin]: the (kh, kw) tap that channel `c` of the S2D pixel at offset (dy, dx) carries
// (input row 2*oh + kh - 1 = S2D row oh + dy, parity py), or false if none.
static bool enc_slab_tap(int cin, int c, int dy, int dx, int& ci, int& kh, int& kw)
{
    const int pp = c / cin, py = pp >> 1, px = pp & 1;
    ci = c % cin;
    for (kh = 0; kh < 5; kh++)
        for (kw = 0; kw < 5; kw++) {
            int d1, p1, d2, p2;
            enc_tap(kh, d1, p1);
            enc_tap(kw, d2, p2);
            if (d1 == dy && p1 == py && d2 == dx && p2 == px) return true;
        }
    return false;
}

static std::vector<int32_t> kelem_offsets(const std::vector<KBlock>& kb)
{
    std::vector<int32_t> off(kb.size());
    int32_t o = 0;
    for (size_t k = 0; k < kb.size(); k++) { off[k] = o; o += kb_channels(kb[k]); }
    return off;
}

// transposed conv: which kernel row serves output parity `par` at input offset d (o = 2h + kh - 1); -1 = none
int dec_kh(int par, int d)
{
    if (par == 0) return d == 0 ? 1 : (d == -1 ? 3 : -1);
    return d == 1 ? 0 : (d == 0 ? 2 : 4);
}

// duplicate every k-block (and its k-elements) as a part-1 twin right after it
template <class KE>
static void split_kblocks(std::vector<KBlock>& kb, std::vector<KE>& ke, int per)
{
    std::vector<KBlock> kb2;
    std::vector<KE> ke2;
    for (size_t k = 0; k < kb.size(); k++)
        for (int part = 0; part < 2; part++) {
            KBlock b = kb[k];
            b.part = (int8_t)part;
            kb2.push_back(b);
            ke2.insert(ke2.end(), ke.begin() + k * per, ke.begin() + (k + 1) * per);
        }
    kb.swap(kb2);
    ke.swap(ke2);
}

bool weights_tf32_exact(const float* coeff)
{
    const CoeffLayout cl = coeff_layout();
    for (int i = 0; i < 6; i++) {
        const size_t n = (size_t)25 * kEnc[i] * kEnc[i + 1];
        for (size_t j = 0; j < n; j++)
            if (round_tf32(coeff[cl.down_w[i] + j]) != coeff[cl.down_w[i] + j]) return false;
    }
    for (int i = 0; i < 5; i++) {
        const size_t n = (size_t)25 * kDecIn[i] * kDecOut[i];
        for (size_t j = 0; j < n; j++)
            if (round_tf32(coeff[cl.up_w[i] + j]) != coeff[cl.up_w[i] + j]) return false;
    }
    return true;
}

std::vector<LayerPlan> build_plans(NetGeom g, int n_img, bool split_weights, int n_stems, int min_ctas, unsigned comp_mask, bool fuse_phases, int lo_fmt)
{
    std::vector<LayerPlan> plans;
    auto narrow = [&](LayerPlan& L) {
        if (min_ctas <= 0) return;
        const long tiles = (long)ceil_div(L.Ws, L.tw) * ceil_div(L.Hs, L.th) * ceil_div(n_img, L.nb) * L.phases * (n_stems > 0 ? n_stems : 1);
        while (L.n_tile > 64 && tiles * L.n_tiles < min_ctas) {
            L.n_tile /= 2;
            L.n_tiles *= 2;
        }
    };
    // ---- encoder: down2 .. down6 --------------------------------------------------------
    for (int i = 1; i <= 5; i++) {
        LayerPlan L{};
        L.index = i - 1;
        L.transposed = false;
        L.cin = kEnc[i];
        L.cout = kEnc[i + 1];
        L.Hs = g.T >> (i + 1);
        L.Ws = g.F >> (i + 1);
        L.phases = 1;
        L.nsrc = 1;
        L.src[0] = SrcDesc{4 * L.cin, L.Ws, L.Hs};
        L.n_tile = L.cout > 256 ? 256 : L.cout;
        L.n_tiles = L.cout / L.n_tile;
        choose_tile(L.Ws, L.Hs, n_img, L.tw, L.th, L.nb);
        narrow(L);
        if (L.cin >= 32) {
            for (int kh = 0; kh < 5; kh++)
                for (int kw = 0; kw < 5; kw++) {
                    int dy, py, dx, px;
                    enc_tap(kh, dy, py);
                    enc_tap(kw, dx, px);
                    for (int c0 = 0; c0 < L.cin; c0 += kKB) {
                        L.kb[0].push_back(KBlock{0, (int8_t)dy, (int8_t)dx, 0, (py * 2 + px) * L.cin + c0});
                        for (int j = 0; j < kKB; j++) L.kelem[0].push_back(KElem{c0 + j, (int8_t)kh, (int8_t)kw});
                    }
                }
        } else {
            // cin == 16: the two column parities of one S2D pixel are 32 contiguous channels, so
            // taps (kw=1,kw=2) and (kw=3,kw=4) pair up; kw=0 pairs with a zero slab.
            for (int kh = 0; kh < 5; kh++) {
                int dy, py;
                enc_tap(kh, dy, py);
                for (int xg = -1; xg <= 1; xg++) {
                    const int kwA = 2 * xg + 1, kwB = 2 * xg + 2;   // px = 0 slab, px = 1 slab
                    L.kb[0].push_back(KBlock{0, (int8_t)dy, (int8_t)xg, 0, py * 2 * L.cin});
                    for (int j = 0; j < kKB; j++) {
                        const int kw = j < 16 ? kwA : kwB;
                        const int c = j & 15;
                        L.kelem[0].push_back(kw < 0 ? KElem{-1, 0, 0} : KElem{c, (int8_t)kh, (int8_t)kw});
                    }
                }
            }
        }
        plans.push_back(L);
    }
    // ---- decoder: up1 .. up5 ------------------------------------------------------------
    for (int d = 0; d < 5; d++) {
        LayerPlan L{};
        L.index = 5 + d;
        L.transposed = true;
        L.cin = kDecIn[d];
        L.cout = kDecOut[d];
        L.Hs = g.T >> (6 - d);
        L.Ws = g.F >> (6 - d);
        L.phases = 4;
        if (d == 0) {
            L.nsrc = 1;
            L.src[0] = SrcDesc{512, L.Ws, L.Hs};
        } else {
            L.nsrc = 2;
            L.src[0] = SrcDesc{L.cin / 2, L.Ws, L.Hs};   // skip (raw conv + bias)
            L.src[1] = SrcDesc{L.cin / 2, L.Ws, L.Hs};   // previous decoder output
        }
        L.n_tile = L.cout;
        L.n_tiles = 1;
        choose_tile(L.Ws, L.Hs, n_img, L.tw, L.th, L.nb);
        // fused parities (see LayerPlan::fused) where N = 4 * cout fits one MMA and the grid still fills the SMs
        const long fused_ctas = (long)ceil_div(L.Ws, L.tw) * ceil_div(L.Hs, L.th) * ceil_div(n_img, L.nb) * (n_stems > 0 ? n_stems : 1);
        if (fuse_phases && 4 * L.cout <= 256 && (min_ctas <= 0 || fused_ctas >= min_ctas)) {
            L.fused = true;
            L.phases = 1;
            L.n_tile = 4 * L.cout;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++)
                    for (int s = 0; s < L.nsrc; s++)
                        for (int c0 = 0; c0 < L.src[s].C; c0 += kKB) {
                            L.kb[0].push_back(KBlock{(int8_t)s, (int8_t)dy, (int8_t)dx, 0, c0});
                            const int base = (s ? L.src[0].C : 0) + c0;
                            for (int j = 0; j < kKB; j++) L.kelem[0].push_back(KElem{base + j, -1, -1});   // (kh, kw) follow from the column's parity
                        }
            plans.push_back(L);
            continue;
        }
        narrow(L);
        for (int po = 0; po < 2; po++)
            for (int qo = 0; qo < 2; qo++) {
                const int p = po * 2 + qo;
                int khs[3], dys[3], kws[3], dxs[3];
                const int nh = dec_taps(po, khs, dys), nw = dec_taps(qo, kws, dxs);
                for (int a = 0; a < nh; a++)
                    for (int b = 0; b < nw; b++)
                        for (int s = 0; s < L.nsrc; s++)
                            for (int c0 = 0; c0 < L.src[s].C; c0 += kKB) {
                                L.kb[p].push_back(KBlock{(int8_t)s, (int8_t)dys[a], (int8_t)dxs[b], 0, c0});
                                const int base = (s ? L.src[0].C : 0) + c0;
                                for (int j = 0; j < kKB; j++)
                                    L.kelem[p].push_back(KElem{base + j, (int8_t)khs[a], (int8_t)kws[b]});
                            }
            }
        plans.push_back(L);
    }
    if (split_weights)
        for (auto& L : plans)
            for (int p = 0; p < L.phases; p++) split_kblocks(L.kb[p], L.kelem[p], kKB);
    // ---- compensation blocks (after the main term: small contributions are added last) ------------
    for (auto& L : plans) {
        L.comp = (comp_mask >> L.index) & 1u;
        L.lo_fmt = L.comp ? layer_lo_format(L.index, lo_fmt) : LO_NONE;
        if (!L.comp) continue;
        const int kKBlo = L.lo_fmt == LO_FP8 ? kKBlo8 : srt::kKBlo;                 // channels per compensation block
        const int8_t kPartLo = (int8_t)(L.lo_fmt == LO_FP8 ? (srt::kPartLo | kPartLo8) : srt::kPartLo);
        if (!L.transposed) {
            L.lo_src = SrcDesc{4 * L.cin, L.Ws, L.Hs};
            for (int c_off = 0; c_off < 4 * L.cin; c_off += kKBlo)
                for (int dy = -1; dy <= 1; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        std::vector<KElem> el(kKBlo, KElem{-1, 0, 0});
                        bool any = false;
                        for (int j = 0; j < kKBlo; j++) {
                            int ci, kh, kw;
                            if (enc_slab_tap(L.cin, c_off + j, dy, dx, ci, kh, kw)) { el[j] = KElem{ci, (int8_t)kh, (int8_t)kw}; any = true; }
                        }
                        if (!any) continue;
                        L.kb[0].push_back(KBlock{(int8_t)kSrcLo, (int8_t)dy, (int8_t)dx, (int8_t)kPartLo, c_off});
                        L.kelem[0].insert(L.kelem[0].end(), el.begin(), el.end());
                    }
        } else if (L.fused) {
            L.lo_src = SrcDesc{L.cin, L.Ws, L.Hs};
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++)
                    for (int c0 = 0; c0 < L.cin; c0 += kKBlo) {
                        L.kb[0].push_back(KBlock{(int8_t)kSrcLo, (int8_t)dy, (int8_t)dx, (int8_t)kPartLo, c0});
                        for (int j = 0; j < kKBlo; j++) L.kelem[0].push_back(KElem{c0 + j, -1, -1});
                    }
        } else {
            L.lo_src = SrcDesc{L.cin, L.Ws, L.Hs};      // [skip residual | up residual] = the reference's concatenated channel order
            for (int po = 0; po < 2; po++)
                for (int qo = 0; qo < 2; qo++) {
                    const int p = po * 2 + qo;
                    int khs[3], dys[3], kws[3], dxs[3];
                    const int nh = dec_taps(po, khs, dys), nw = dec_taps(qo, kws, dxs);
                    for (int a = 0; a < nh; a++)
                        for (int b = 0; b < nw; b++)
                            for (int c0 = 0; c0 < L.cin; c0 += kKBlo) {
                                L.kb[p].push_back(KBlock{(int8_t)kSrcLo, (int8_t)dys[a], (int8_t)dxs[b], (int8_t)kPartLo, c0});
                                for (int j = 0; j < kKBlo; j++) L.kelem[p].push_back(KElem{c0 + j, (int8_t)khs[a], (int8_t)kws[b]});
                            }
                }
        }
    }
    for (auto& L : plans) {
        for (int p = 0; p < L.phases; p++) L.ke_off[p] = kelem_offsets(L.kb[p]);
        size_t off = 0;
        for (int p = 0; p < L.phases; p++) {
            L.w_phase_off[p] = off;
            off += (size_t)L.n_tiles * L.kb[p].size() * L.n_tile * kKB;
        }
        L.w_floats_per_stem = off;
    }
    return plans;
}

void pack_layer(const LayerPlan& L, const float* coeff, float* out)
{
    const CoeffLayout cl = coeff_layout();
    const float* w = coeff + (L.transposed ? cl.up_w[L.index - 5] : cl.down_w[L.index + 1]);
    for (int p = 0; p < L.phases; p++) {
        const size_t nkb = L.kb[p].size();
        for (int nt = 0; nt < L.n_tiles; nt++)
            for (size_t kb = 0; kb < nkb; kb++) {
                float* blk = out + L.w_phase_off[p] + ((size_t)nt * nkb + kb) * L.n_tile * kKB;
                const bool lo = (L.kb[p][kb].part & kPartLo) != 0;     // [n_tile][64] bf16 in the same bytes ...
                const bool lo8 = (L.kb[p][kb].part & kPartLo8) != 0;   // ... or [n_tile][128] e5m2
                const int width = kb_channels(L.kb[p][kb]);
                for (int n = 0; n < L.n_tile; n++) {
                    int o = nt * L.n_tile + n;
                    int fkh = 0, fkw = 0;
                    if (L.fused) {          // column = parity * cout + channel; the tap follows from the parity and the k-block's offset
                        const int ph = o / L.cout;
                        o %= L.cout;
                        fkh = dec_kh(ph >> 1, L.kb[p][kb].dy);
                        fkw = dec_kh(ph & 1, L.kb[p][kb].dx);
                    }
                    for (int j = 0; j < width; j++) {
                        KElem e = L.kelem[p][L.ke_off[p][kb] + j];
                        if (L.fused) {
                            if (fkh < 0 || fkw < 0) e.cin = -1;
                            e.kh = (int8_t)fkh; e.kw = (int8_t)fkw;
                        }
                        float v = 0.0f;
                        if (e.cin >= 0) {
                            const size_t idx = L.transposed
                                ? (((size_t)e.cin * L.cout + o) * 5 + e.kh) * 5 + e.kw      // [I][O][kh][kw]
                                : (((size_t)o * L.cin + e.cin) * 5 + e.kh) * 5 + e.kw;     // [O][I][kh][kw]
                            v = lo ? w[idx] : weight_part(w[idx], L.kb[p][kb].part);
                        }
                        if (lo8) reinterpret_cast<uint8_t*>(blk)[swz128_index8(n, j)] = e5m2_rn(0.25f * v);
                        else if (lo) reinterpret_cast<uint16_t*>(blk)[swz128_index16(n, j)] = bf16_rn(v);
                        else blk[swz128_index(n, j)] = v;
                    }
                }
            }
    }
}

// ------------------------------------------------------------------------------------------
// row-patch plans
// ------------------------------------------------------------------------------------------
bool row_plan_supported(int layer_index) { return layer_index == 0 || layer_index == 1 || layer_index == 8 || layer_index == 9; }


RowPlan build_row_plan(NetGeom g, int layer_index, bool split_weights, bool comp, int lo_fmt)
{
    RowPlan L{};
    L.index = layer_index;
    if (layer_index < 5) {
        const int i = layer_index + 1;          // down{i+1}: cin = kEnc[i]
        L.transposed = false;
        L.cin = kEnc[i];
        L.cout = kEnc[i + 1];
        L.Hs = g.T >> (i + 1);
        L.Ws = g.F >> (i + 1);
        L.N = L.cout;
        L.phases = 1;
        L.nsrc = 1;
        L.src[0] = SrcDesc{4 * L.cin, L.Ws, L.Hs};
        // chunks = 32-channel slabs of the S2D pixel [py][px][cin]
        for (int c_off = 0; c_off < 4 * L.cin; c_off += kKB) {
            RowChunk ch{0, c_off, (int32_t)L.kb.size(), 0};
            // which (kh, kw) land in this slab, and at which S2D offset
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    std::vector<KElemP> el(kKB);
                    bool any = false;
                    for (int j = 0; j < kKB; j++) {
                        KElemP e{-1, {-1, -1, -1, -1}, {-1, -1, -1, -1}};
                        int ci, kh, kw;
                        if (enc_slab_tap(L.cin, c_off + j, dy, dx, ci, kh, kw)) { e.cin = ci; e.kh[0] = (int8_t)kh; e.kw[0] = (int8_t)kw; any = true; }
                        el[j] = e;
                    }
                    if (!any) continue;
                    L.kb.push_back(KBlock{0, (int8_t)dy, (int8_t)dx, 0, c_off});
                    L.kelem.insert(L.kelem.end(), el.begin(), el.end());
                    ch.nkb++;
                }
            L.chunks.push_back(ch);
        }
    } else {
        const int d = layer_index - 5;          // up{d+1}
        L.transposed = true;
        L.cin = kDecIn[d];
        L.cout = kDecOut[d];
        L.Hs = g.T >> (6 - d);
        L.Ws = g.F >> (6 - d);
        L.N = 4 * L.cout;
        L.phases = 4;
        L.nsrc = d == 0 ? 1 : 2;
        L.src[0] = SrcDesc{d == 0 ? 512 : L.cin / 2, L.Ws, L.Hs};
        L.src[1] = L.src[0];
        for (int sidx = 0; sidx < L.nsrc; sidx++)
            for (int c_off = 0; c_off < L.src[sidx].C; c_off += kKB) {
                RowChunk ch{(int8_t)sidx, c_off, (int32_t)L.kb.size(), 0};
                for (int dy = -1; dy <= 1; dy++)
                    for (int dx = -1; dx <= 1; dx++) {
                        L.kb.push_back(KBlock{(int8_t)sidx, (int8_t)dy, (int8_t)dx, 0, c_off});
                        for (int j = 0; j < kKB; j++) {
                            KElemP e{(sidx ? L.src[0].C : 0) + c_off + j, {-1, -1, -1, -1}, {-1, -1, -1, -1}};
                            for (int ph = 0; ph < 4; ph++) {
                                const int kh = dec_kh(ph >> 1, dy), kw = dec_kh(ph & 1, dx);
                                if (kh >= 0 && kw >= 0) { e.kh[ph] = (int8_t)kh; e.kw[ph] = (int8_t)kw; }
                            }
                            L.kelem.push_back(e);
                        }
                        ch.nkb++;
                    }
                L.chunks.push_back(ch);
            }
    }
    if (split_weights) {
        // twins stay inside their chunk: chunk k-block ranges double
        split_kblocks(L.kb, L.kelem, kKB);
        for (auto& ch : L.chunks) { ch.kb0 *= 2; ch.nkb *= 2; }
    }
    // compensation chunks: 64-channel slabs of the bf16 residual tensor, same taps, bf16 weights (see build_plans)
    L.comp = comp;
    L.lo_fmt = comp ? layer_lo_format(layer_index, lo_fmt, true) : LO_NONE;
    if (comp) {
        const int Clo = L.transposed ? L.cin : 4 * L.cin;
        L.lo_src = SrcDesc{Clo, L.Ws, L.Hs};
        const int kKBlo = L.lo_fmt == LO_FP8 ? kKBlo8 : srt::kKBlo;                     // LO_FP8N and bf16: 64 channels per block
        const int8_t kPartLo = (int8_t)(L.lo_fmt == LO_FP8 ? (srt::kPartLo | kPartLo8)
                                        : L.lo_fmt == LO_FP8N ? (srt::kPartLo | kPartLo8 | kPartLo8n) : srt::kPartLo);
        for (int c_off = 0; c_off < Clo; c_off += kKBlo) {
            RowChunk ch{(int8_t)kSrcLo, c_off, (int32_t)L.kb.size(), 0};
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    std::vector<KElemP> el(kKBlo, KElemP{-1, {-1, -1, -1, -1}, {-1, -1, -1, -1}});
                    bool any = false;
                    for (int j = 0; j < kKBlo; j++) {
                        if (L.transposed) {
                            el[j].cin = c_off + j;
                            for (int ph = 0; ph < 4; ph++) {
                                const int kh = dec_kh(ph >> 1, dy), kw = dec_kh(ph & 1, dx);
                                if (kh >= 0 && kw >= 0) { el[j].kh[ph] = (int8_t)kh; el[j].kw[ph] = (int8_t)kw; }
                            }
                            any = true;
                        } else {
                            int ci, kh, kw;
                            if (enc_slab_tap(L.cin, c_off + j, dy, dx, ci, kh, kw)) { el[j].cin = ci; el[j].kh[0] = (int8_t)kh; el[j].kw[0] = (int8_t)kw; any = true; }
                        }
                    }
                    if (!any) continue;
                    L.kb.push_back(KBlock{(int8_t)kSrcLo, (int8_t)dy, (int8_t)dx, (int8_t)kPartLo, c_off});
                    L.kelem.insert(L.kelem.end(), el.begin(), el.end());
                    ch.nkb++;
                }
            L.chunks.push_back(ch);
        }
    }
    L.ke_off = kelem_offsets(L.kb);
    // K steps without any weight (see kPartSkipShift): quarter q of k-block k = elements [q * width / 4, (q + 1) * width / 4)
    for (size_t k = 0; k < L.kb.size(); k++) {
        const int steps = kb_ksteps(L.kb[k]), width = kb_channels(L.kb[k]), per = width / steps;
        int skip = 0;
        for (int q = 0; q < steps; q++) {
            bool any = false;
            for (int j = q * per; j < (q + 1) * per; j++) {
                const KElemP& e = L.kelem[L.ke_off[k] + j];
                for (int ph = 0; ph < 4; ph++) any = any || (e.cin >= 0 && e.kh[ph] >= 0);
            }
            if (!any) skip |= 1 << q;
        }
        if (skip != (1 << steps) - 1) L.kb[k].part = (int8_t)((L.kb[k].part & 0xf) | (skip << kPartSkipShift));
    }
    L.R = row_plan_R(L.N);
    L.w_floats_per_stem = L.kb.size() * (size_t)L.N * kKB;
    return L;
}

void pack_row_layer(const RowPlan& L, const float* coeff, float* out)
{
    const CoeffLayout cl = coeff_layout();
    const float* w = coeff + (L.transposed ? cl.up_w[L.index - 5] : cl.down_w[L.index + 1]);
    for (size_t kb = 0; kb < L.kb.size(); kb++) {
        float* blk = out + kb * (size_t)L.N * kKB;
        const bool lo = (L.kb[kb].part & kPartLo) != 0, lo8 = (L.kb[kb].part & kPartLo8) != 0;
        const int width = kb_channels(L.kb[kb]);
        for (int n = 0; n < L.N; n++) {
            const int ph = L.transposed ? n / L.cout : 0, o = L.transposed ? n % L.cout : n;
            for (int j = 0; j < width; j++) {
                const KElemP& e = L.kelem[L.ke_off[kb] + j];
                float v = 0.0f;
                if (e.cin >= 0 && e.kh[ph] >= 0) {
                    const size_t idx = L.transposed ? (((size_t)e.cin * L.cout + o) * 5 + e.kh[ph]) * 5 + e.kw[ph]
                                                    : (((size_t)o * L.cin + e.cin) * 5 + e.kh[ph]) * 5 + e.kw[ph];
                    v = lo ? w[idx] : weight_part(w[idx], L.kb[kb].part);
                }
                if (lo8 && (L.kb[kb].part & kPartLo8n)) reinterpret_cast<uint8_t*>(blk)[swz64_index8(n, j)] = e5m2_rn(0.25f * v);
                else if (lo8) reinterpret_cast<uint8_t*>(blk)[swz128_index8(n, j)] = e5m2_rn(0.25f * v);
                else if (lo) reinterpret_cast<uint16_t*>(blk)[swz128_index16(n, j)] = bf16_rn(v);
                else blk[swz128_index(n, j)] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// down1 (tensor-core form)
// ------------------------------------------------------------------------------------------
Down1Plan build_down1_plan(NetGeom g, bool split_weights)
{
    Down1Plan L{};
    L.Hs = g.T / 2;
    L.Ws = g.F / 2;
    // two slabs with the same taps and weights: source 0 = tf32(mag), source 1 = mag - tf32(mag)
    for (int part = 0; part < 2; part++)
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            L.kb.push_back(KBlock{(int8_t)part, (int8_t)dy, (int8_t)dx, 0, 0});   // src = magnitude part (hi / lo)
            for (int j = 0; j < kKB1; j++) {
                const int py = j >> 2, px = (j >> 1) & 1, c = j & 1;
                KElemP e{-1, {-1, -1, -1, -1}, {-1, -1, -1, -1}};
                for (int kh = 0; kh < 5; kh++)
                    for (int kw = 0; kw < 5; kw++) {
                        int d1, p1, d2, p2;
                        enc_tap(kh, d1, p1);
                        enc_tap(kw, d2, p2);
                        if (d1 == dy && p1 == py && d2 == dx && p2 == px) { e.cin = c; e.kh[0] = (int8_t)kh; e.kw[0] = (int8_t)kw; }
                    }
                L.kelem.push_back(e);
            }
        }
    if (split_weights) {
        // keep the (hi-magnitude taps | lo-magnitude taps) chunk structure: split each half separately
        split_kblocks(L.kb, L.kelem, kKB1);
    }
    return L;
}

void pack_down1(const Down1Plan& L, const float* const* coeffs, int nstems, float* out)
{
    const CoeffLayout cl = coeff_layout();
    const int N = 16 * nstems;
    for (size_t kb = 0; kb < L.kb.size(); kb++)
        for (int n = 0; n < N; n++) {
            const float* w = coeffs[n / 16] + cl.down_w[0];
            const int o = n % 16;
            for (int j = 0; j < kKB1; j++) {
                const KElemP& e = L.kelem[kb * kKB1 + j];
                float v = 0.0f;
                if (e.cin >= 0) v = weight_part(w[(((size_t)o * 2 + e.cin) * 5 + e.kh[0]) * 5 + e.kw[0]], L.kb[kb].part);
                out[kb * (size_t)N * kKB1 + swz32_index(n, j)] = v;
            }
        }
}

void pack_up6_weights(const float* w6, float* out)
{
    for (int i = 0; i < kUp6PackFloats; i++) out[i] = 0.0f;
    for (int b = 0; b < 4; b++)
        for (int term = 0; term < 2; term++)
            for (int tap = 0; tap < 25; tap++)
                for (int j = 0; j < 8; j++) {
                    const int cin = (b >> 1) * 16 + (b & 1) * 8 + j;
                    out[(b * 2 + term) * 256 + swz32_index(tap, j)] = weight_part(w6[cin * 25 + tap], term);
                }
    uint8_t* blk = reinterpret_cast<uint8_t*>(out + 8 * 256);
    for (int tap = 0; tap < 25; tap++)
        for (int j = 0; j < 32; j++) blk[swz32_index8(tap, j)] = e5m2_rn(0.25f * w6[j * 25 + tap]);
}

}  // namespace srt
