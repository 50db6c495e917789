// srt_conv_rp.cu — persistent "row-patch" tcgen05 kernel for the small-N layers (down2, down3, up4, up5).
//
// The generic kernel (srt_conv_tc.cu) fetches a fresh 16 KB activation tile per tap, which makes
// layers with few output channels L2->SM-bandwidth bound (ncu, profiles/r1a_*: up5 at 59 TF/s).
// Here a tile is R output rows x 128 columns of one image and, per 32-channel slab, ONE
// (R+2) x 136-pixel patch is loaded with a single TMA box; all taps of the slab are MMAs whose
// A-operand descriptors point at shifted windows of that patch:
//        start = patch + (r + dy + 1) * row_pitch + (dx + 1) * 128 B (+ 32 B per K step)
// Rows of the patch are 136 pixels * 128 B = 17 KB apart (a multiple of the 1 KB swizzle atom);
// the +-1 pixel column shift moves the start by 128 B inside the atom.  Measured on B200
// (tools/rp_probe.py): the 128B swizzle is a function of the absolute shared-memory address, so the
// descriptor's base-offset field stays 0 for such windows (setting it to (addr>>7)&7 is wrong).
// Decoder layers fuse the four output parities into N = 4*cout.
//
// One persistent CTA per SM walks tiles blockIdx.x, +gridDim.x, ... (stem-major order, so all SMs
// stream the same stem's weights out of L2).  Five pipelines run concurrently:
//   warp 8  patch producer   : TMA boxes into a 2-deep patch ring            (patch_full/empty)
//   warp 9  weight producer  : pre-swizzled [N][32] blocks into a deep ring   (w_full/empty)
//   warp 10 MMA issuer       : tcgen05.mma.kind::tf32 into one of 2 TMEM accumulator sets
//   warps 0-7 epilogue       : tcgen05.ld -> bias/BN/act -> global, overlapped with the next tile's MMAs
// The first non-persistent version of this kernel was latency-bound (refill chains of ~3 us per
// weight block / patch, profiles/r1c_*); deep prefetch across tile boundaries removes those bubbles.
#include "srt_epilogue.cuh"
#include "srt_kernels.cuh"
#include "srt_ptx.cuh"

namespace srt {

// threads per CTA = (EPW epilogue warps + 3 control warps) * 32; EPW = 8, or 16 where shared memory allows (down1)
constexpr int kRpMaxChunks = 12, kRpMaxKB = 112, kRpMaxWStages = 12;   // up4 with two-term weights + compensation: 6 chunks, 90 k-blocks

struct RpHeader {
    uint64_t patch_full[2], patch_empty[2];
    uint64_t acc_full[2], acc_empty[2];
    uint64_t w_full[kRpMaxWStages], w_empty[kRpMaxWStages];
    uint32_t tmem_base, pad;
    RowChunk chunks[kRpMaxChunks];
    KBlock kb[kRpMaxKB];
};

template <int N, int R, int WS, int KB, int EPW>
constexpr size_t rp_smem_bytes() { return sizeof(RpHeader) + 1024 + (size_t)2 * (R + 2) * kPatchW * KB * 4 + (size_t)WS * N * KB * 4 + EPW * 2048; }   // + per-warp store staging

struct RpTile {
    int s, n, x0, y0;
};
__device__ __forceinline__ RpTile rp_tile(const RowConvParams& p, int t, int R)
{
    RpTile o;
    const int tx = t % p.tiles_x;
    t /= p.tiles_x;
    const int ty = t % p.tiles_y;
    t /= p.tiles_y;
    const int b = t % p.ep.Bv;
    o.s = t / p.ep.Bv;
    o.n = o.s * p.ep.B + b;
    o.x0 = tx * kTileM;
    o.y0 = ty * R;
    return o;
}

// The MMAs of one k-block, straight-line: K step outer, row inner, so consecutive MMAs hit different accumulators (dependent MMAs on
// one accumulator cost ~100 cycles each, independent ones ~45-60: tools/mma_probe.cu).  The issuing thread is bound by instruction
// latency (~43 cycles per small-N MMA), so what is issued is decided at compile time: LO = compensation block (64 bf16 residuals per
// pixel in the same 128-byte rows: K = 16 per step, kind::f16), SKIP = K steps whose weights are all zero (kPartSkipShift).
// Returns with `first` cleared once something was issued.
template <int N, int R, int KSTEPS, int LO, int SKIP>   // LO: 0 = TF32 main term, 1 = bf16 residuals, 2 = e5m2 residuals
__device__ __forceinline__ void rp_issue(uint32_t acc, uint32_t a_lo, uint32_t b_lo, uint32_t row_pitch16, uint32_t idesc, uint32_t desc_hi, bool& first)
{
    bool fresh = first;
#pragma unroll
    for (int kk = 0; kk < KSTEPS; kk++) {
        if (SKIP & (1 << kk)) continue;
        const uint32_t accum = fresh ? 0u : 1u;
        fresh = false;
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (LO == 2) ptx::mma_f8_ss_lo(acc + (uint32_t)(r * N), a_lo + (uint32_t)r * row_pitch16 + (uint32_t)(kk * 2), b_lo + (uint32_t)(kk * 2), idesc, accum, desc_hi);
            else if (LO == 1) ptx::mma_bf16_ss_lo(acc + (uint32_t)(r * N), a_lo + (uint32_t)r * row_pitch16 + (uint32_t)(kk * 2), b_lo + (uint32_t)(kk * 2), idesc, accum, desc_hi);
            else ptx::mma_tf32_ss_lo(acc + (uint32_t)(r * N), a_lo + (uint32_t)r * row_pitch16 + (uint32_t)(kk * 2), b_lo + (uint32_t)(kk * 2), idesc, accum, desc_hi);
        }
    }
    first = fresh;
}
template <int N, int R, int KSTEPS, int LO>
__device__ __forceinline__ void rp_issue_masked(int skip, uint32_t acc, uint32_t a_lo, uint32_t b_lo, uint32_t row_pitch16, uint32_t idesc, uint32_t desc_hi,
                                                bool& first)
{
    switch (skip) {   // the masks the plans produce: a slab spanning both column parities loses its first or second half, or three quarters
    case 0x0: rp_issue<N, R, KSTEPS, LO, 0x0>(acc, a_lo, b_lo, row_pitch16, idesc, desc_hi, first); break;
    case 0x3: rp_issue<N, R, KSTEPS, LO, 0x3>(acc, a_lo, b_lo, row_pitch16, idesc, desc_hi, first); break;
    case 0xc: rp_issue<N, R, KSTEPS, LO, 0xc>(acc, a_lo, b_lo, row_pitch16, idesc, desc_hi, first); break;
    case 0x5: rp_issue<N, R, KSTEPS, LO, 0x5>(acc, a_lo, b_lo, row_pitch16, idesc, desc_hi, first); break;
    case 0xa: rp_issue<N, R, KSTEPS, LO, 0xa>(acc, a_lo, b_lo, row_pitch16, idesc, desc_hi, first); break;
    case 0x7: rp_issue<N, R, KSTEPS, LO, 0x7>(acc, a_lo, b_lo, row_pitch16, idesc, desc_hi, first); break;
    case 0xb: rp_issue<N, R, KSTEPS, LO, 0xb>(acc, a_lo, b_lo, row_pitch16, idesc, desc_hi, first); break;
    case 0xd: rp_issue<N, R, KSTEPS, LO, 0xd>(acc, a_lo, b_lo, row_pitch16, idesc, desc_hi, first); break;
    case 0xe: rp_issue<N, R, KSTEPS, LO, 0xe>(acc, a_lo, b_lo, row_pitch16, idesc, desc_hi, first); break;
    default:  rp_issue<N, R, KSTEPS, LO, 0x0>(acc, a_lo, b_lo, row_pitch16, idesc, desc_hi, first); break;   // any other mask: issue everything (zeros are harmless)
    }
}

template <int N, int R, int WS, int KB, int EPW>
__global__ void __launch_bounds__((EPW + 3) * 32, 1) conv_rp_kernel(const __grid_constant__ RowConvParams p)
{
    constexpr int kRpThreads = (EPW + 3) * 32;
    constexpr int kWarpPatch = EPW, kWarpWeights = EPW + 1, kWarpMma = EPW + 2;   // control warps sit above the epilogue warps
    extern __shared__ uint8_t smem_raw[];
    RpHeader* hdr = reinterpret_cast<RpHeader*>(smem_raw);
    const uint32_t patch_base = (ptx::smem_u32(smem_raw) + (uint32_t)sizeof(RpHeader) + 1023u) & ~1023u;
    uint8_t* patch = smem_raw + (patch_base - ptx::smem_u32(smem_raw));
    constexpr int kRowBytes = KB * 4;                 // 128 (SWIZZLE_128B) or 32 (SWIZZLE_32B)
    constexpr int kRowPitch = kPatchW * kRowBytes;    // multiple of the 8-row swizzle atom
    constexpr int kPatchBytes = (R + 2) * kRowPitch;
    constexpr int kWBytes = N * kRowBytes;
    constexpr uint32_t kDescHi = KB == 32 ? ptx::kDescHiSw128 : ptx::kDescHiSw32;
    constexpr int kAccCols = R * N;
    constexpr int kTmemCols = 2 * kAccCols <= 32 ? 32 : 2 * kAccCols <= 64 ? 64 : 2 * kAccCols <= 128 ? 128 : 2 * kAccCols <= 256 ? 256 : 512;
    static_assert(2 * kAccCols <= 512, "two accumulator sets must fit TMEM");
    uint8_t* wring = patch + 2 * kPatchBytes;
    float4* stage_all = reinterpret_cast<float4*>(wring + (size_t)WS * kWBytes);   // EPW epilogue warps x 2 KB (store16_warp)
    const uint32_t wring_base = patch_base + 2 * kPatchBytes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles = p.tiles_x * p.tiles_y * p.ep.Bv * p.ep.S;

    for (int i = threadIdx.x; i < p.n_chunks; i += kRpThreads) hdr->chunks[i] = p.chunks[i];
    for (int i = threadIdx.x; i < p.nkb; i += kRpThreads) hdr->kb[i] = p.kb[i];
    if (warp == kWarpPatch && lane == 0) {
        ptx::tma_prefetch_desc(&p.tmap[0]);
        ptx::tma_prefetch_desc(&p.tmap[1]);
        ptx::tma_prefetch_desc(&p.tmap[2]);
        for (int i = 0; i < 2; i++) {
            ptx::mbar_init(&hdr->patch_full[i], 1);
            ptx::mbar_init(&hdr->patch_empty[i], 1);
            ptx::mbar_init(&hdr->acc_full[i], 1);
            ptx::mbar_init(&hdr->acc_empty[i], EPW);   // one arrival per epilogue warp
        }
        for (int i = 0; i < WS; i++) {
            ptx::mbar_init(&hdr->w_full[i], 1);
            ptx::mbar_init(&hdr->w_empty[i], 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == kWarpMma) ptx::tmem_alloc<kTmemCols>(&hdr->tmem_base);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_d = hdr->tmem_base;

    // Role -> warp mapping: the SM's issue arbiter favours higher warp ids, so the latency-critical single-thread
    // roles (MMA issuer, producers) sit above the 8 epilogue warps instead of being starved by them.
    if (warp == kWarpPatch) {
        // ===== patch producer ==================================================================
        if (ptx::elect_one()) {   // not `lane == 0`: see srt_ptx.cuh (straight-line UTCHMMA / UTMALDG issue)
            int ps = 0;
            uint32_t pph = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const RpTile tl = rp_tile(p, t, R);
                for (int ci = 0; ci < p.n_chunks; ci++) {
                    const RowChunk ch = hdr->chunks[ci];
                    ptx::mbar_wait(&hdr->patch_empty[ps], pph ^ 1);
                    if (p.dbg & 4) { ptx::mbar_arrive(&hdr->patch_full[ps]); }
                    else {
                    // a chunk of 64-channel e5m2 residuals (kPartLo8n) has 64-byte rows: half the patch bytes
                    const bool narrow = KB == 32 && (hdr->kb[ch.kb0].part & kPartLo8n);
                    ptx::mbar_arrive_expect_tx(&hdr->patch_full[ps], narrow ? kPatchBytes / 2 : kPatchBytes);
                    ptx::tma_load_4d(patch + (size_t)ps * kPatchBytes, &p.tmap[ch.src], &hdr->patch_full[ps], ch.c_off, tl.x0 - 1, tl.y0 - 1, p.src_img0 + tl.n);
                    }
                    if (++ps == 2) { ps = 0; pph ^= 1; }
                }
            }
        }
    } else if (warp == kWarpWeights) {
        // ===== weight producer: the same nkb blocks per tile, streamed ahead across tiles ========
        if (ptx::elect_one()) {   // not `lane == 0`: see srt_ptx.cuh (straight-line UTCHMMA / UTMALDG issue)
            int ws = 0;
            uint32_t wph = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int s = t / (p.tiles_x * p.tiles_y * p.ep.Bv);
                const float* wsrc = p.w + (size_t)s * p.w_stem_stride;
                for (int k = 0; k < p.nkb; k++) {
                    ptx::mbar_wait(&hdr->w_empty[ws], wph ^ 1);
                    if (p.dbg & 8) { ptx::mbar_arrive(&hdr->w_full[ws]); }
                    else {
                    ptx::mbar_arrive_expect_tx(&hdr->w_full[ws], kWBytes);
                    ptx::bulk_load_1d(wring + (size_t)ws * kWBytes, wsrc + (size_t)k * N * KB, kWBytes, &hdr->w_full[ws]);
                    }
                    if (++ws == WS) { ws = 0; wph ^= 1; }
                }
            }
        }
    } else if (warp == kWarpMma) {
        // ===== MMA issuer ==========================================================================
        if (ptx::elect_one()) {
            constexpr uint32_t idesc = ptx::umma_idesc_tf32(kTileM, N), idesc_lo = ptx::umma_idesc_bf16(kTileM, N);
            int ws = 0, ps = 0, as = 0;
            uint32_t wph = 0, pph = 0, aph = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                ptx::mbar_wait(&hdr->acc_empty[as], aph ^ 1);     // epilogue has drained this accumulator set
                ptx::tc_fence_after();
                const uint32_t acc = tmem_d + (uint32_t)(as * kAccCols);
                bool first = true;
                for (int ci = 0; ci < p.n_chunks; ci++) {
                    const RowChunk ch = hdr->chunks[ci];
                    ptx::mbar_wait(&hdr->patch_full[ps], pph);
                    const uint32_t p_lo = ptx::umma_desc_lo(patch_base + (uint32_t)ps * kPatchBytes);
                    for (int k = ch.kb0; k < ch.kb0 + ch.nkb; k++) {
                        const KBlock kb = hdr->kb[k];
                        // descriptor low word of the tap's window for row r = 0, K step 0 (16-byte units); rows of a narrow chunk are 64 bytes
                        const bool narrow = KB == 32 && (kb.part & kPartLo8n);
                        const uint32_t pitch16 = narrow ? (uint32_t)(kRowPitch >> 5) : (uint32_t)(kRowPitch >> 4);
                        const uint32_t a_lo = p_lo + (uint32_t)(kb.dy + 1) * pitch16 + (uint32_t)(kb.dx + 1) * (narrow ? (uint32_t)(kRowBytes >> 5) : (uint32_t)(kRowBytes >> 4));
                        const uint32_t b_lo = ptx::umma_desc_lo(wring_base + (uint32_t)ws * kWBytes);
                        ptx::mbar_wait(&hdr->w_full[ws], wph);
                        ptx::tc_fence_after();
                        if (!(p.dbg & 2)) {
                            const int skip = KB == 32 ? kb_skip_mask(kb) : 0;
                            const bool lo = KB == 32 && (kb.part & kPartLo), lo8 = lo && (kb.part & kPartLo8);
                            if (narrow) {                                // 64 e5m2 residuals per pixel: two K = 32 steps over 64-byte rows (SWIZZLE_64B)
                                if (skip == 0) rp_issue<N, R, 2, 2, 0>(acc, a_lo, b_lo, pitch16, idesc_lo, ptx::kDescHiSw64, first);
                                else if (skip == 1) rp_issue<N, R, 2, 2, 1>(acc, a_lo, b_lo, pitch16, idesc_lo, ptx::kDescHiSw64, first);
                                else rp_issue<N, R, 2, 2, 2>(acc, a_lo, b_lo, pitch16, idesc_lo, ptx::kDescHiSw64, first);
                            } else if (skip == 0 && !(p.dbg & 256)) {    // the common case without a jump table in the issuing thread's path
                                if (lo8) rp_issue<N, R, KB / 8, 2, 0>(acc, a_lo, b_lo, pitch16, idesc_lo, kDescHi, first);
                                else if (lo) rp_issue<N, R, KB / 8, 1, 0>(acc, a_lo, b_lo, pitch16, idesc_lo, kDescHi, first);
                                else rp_issue<N, R, KB / 8, 0, 0>(acc, a_lo, b_lo, pitch16, idesc, kDescHi, first);
                            } else if (lo8)
                                rp_issue_masked<N, R, KB / 8, 2>(skip, acc, a_lo, b_lo, pitch16, idesc_lo, kDescHi, first);
                            else if (lo)
                                rp_issue_masked<N, R, KB / 8, 1>(skip, acc, a_lo, b_lo, pitch16, idesc_lo, kDescHi, first);
                            else
                                rp_issue_masked<N, R, KB / 8, 0>(skip, acc, a_lo, b_lo, pitch16, idesc, kDescHi, first);
                        }
                        ptx::mma_commit(&hdr->w_empty[ws]);
                        if (++ws == WS) { ws = 0; wph ^= 1; }
                    }
                    ptx::mma_commit(&hdr->patch_empty[ps]);
                    if (++ps == 2) { ps = 0; pph ^= 1; }
                }
                ptx::mma_commit(&hdr->acc_full[as]);
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
    } else {
        // ===== epilogue: warps 0..EPW-1; warp%4 selects the TMEM lane quarter, warp/4 the share of the tile's
        // (row, 16-column chunk) units.  EPW = 8: two warps per quarter split the columns in two contiguous halves, so
        // each thread writes whole 128-byte lines (fused decoder layers: both column parities of one output row;
        // encoder: half of the pixel's channels).  EPW = 16: four warps per quarter take contiguous runs of units
        // (down1: one tile row each) - the epilogue is instruction-latency bound (ncu r1m: issue slots 50 % busy
        // with 2 epilogue warps per scheduler), so twice the warps is what speeds it up.
        const int q = warp & 3, part = warp >> 2;
        const int m = q * 32 + lane;
        constexpr int kParts = EPW / 4;
        constexpr int kChunksPerRow = N / 16;
        constexpr int kUnits = R * kChunksPerRow;                    // 16-column chunks per tile and lane
        constexpr bool kSplitCols = (kParts == 2 && N >= 32);        // the original column-half mapping
        int as = 0;
        uint32_t aph = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const RpTile tl = rp_tile(p, t, R);
            const int X = tl.x0 + m;
            ptx::mbar_wait(&hdr->acc_full[as], aph);
            ptx::tc_fence_after();
            const uint32_t acc = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kAccCols);
            float4* stage = stage_all + warp * 128;
            constexpr int kPerPart = kSplitCols ? kUnits / 2 : (kUnits + kParts - 1) / kParts;
#pragma unroll 1
            for (int i = 0; i < kPerPart; i++) {
                int r, c0;
                if (kSplitCols) {            // unit i of this warp: row i / (chunks per half), chunk inside the warp's column half
                    r = i / (kChunksPerRow / 2);
                    c0 = (part * (kChunksPerRow / 2) + i % (kChunksPerRow / 2)) * 16;
                } else {
                    const int u = part * kPerPart + i;
                    if (u >= kUnits) break;
                    r = u / kChunksPerRow;
                    c0 = (u % kChunksPerRow) * 16;
                }
                const int Y = tl.y0 + r;
                const bool valid = X < p.ep.Ws && Y < p.ep.Hs;
                float v[16];
                ptx::tmem_ld16(acc + (uint32_t)(r * N + c0), v);
                if (!(p.dbg & 1)) {   // warp-uniform: the stores are warp-cooperative
                    if (p.stems_per_tile > 1 || p.stem0 > 0) {   // down1: the column selects the stem (its own output image)
                        const int se = p.stem0 + c0 / p.ep.cout;
                        epilogue16(p.ep, se, se * p.ep.B + tl.n, Y, X, 0, c0 % p.ep.cout, v, stage, valid);
                    } else if (p.ep.mode == 2) epilogue16(p.ep, tl.s, tl.n, Y, X, c0 / p.ep.cout, c0 % p.ep.cout, v, stage, valid);
                    else epilogue16(p.ep, tl.s, tl.n, Y, X, 0, c0, v, stage, valid);
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&hdr->acc_empty[as]);
            if (++as == 2) { as = 0; aph ^= 1; }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kWarpMma) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<kTmemCols>(tmem_d);
    }
}

template <int N, int R, int WS, int KB, int EPW = 8>
static void launch_rp(const RowConvParams& p, cudaStream_t st)
{
    constexpr size_t smem = rp_smem_bytes<N, R, WS, KB, EPW>();
    static_assert(smem <= 232448, "must fit the 227 KB per-CTA limit");
    static LaunchState state;
    const int sms = state.prepare(conv_rp_kernel<N, R, WS, KB, EPW>, smem);
    const int n_tiles = p.tiles_x * p.tiles_y * p.ep.Bv * p.ep.S;
    dim3 grid(n_tiles < sms ? n_tiles : sms, 1, 1);
    conv_rp_kernel<N, R, WS, KB, EPW><<<grid, (EPW + 3) * 32, smem, st>>>(p);
}

bool conv_rp_fits(int n_chunks, int nkb) { return n_chunks <= kRpMaxChunks && nkb <= kRpMaxKB; }

void launch_conv_rp(const RowConvParams& p, cudaStream_t st)
{
    if (p.kb_width == 8) {   // down1: 8-channel k-blocks, up to 4 stems fused into N; small patches leave room for 16 epilogue warps
        const bool wide = !(p.dbg & 64);   // SRT_RP_DBG bit 6: the 8-warp epilogue (timing experiments)
        if (p.N == 64 && p.R == 4) { if (wide) launch_rp<64, 4, 9, 8, 16>(p, st); else launch_rp<64, 4, 9, 8>(p, st); }
        else if (p.N == 32 && p.R == 4) { if (wide) launch_rp<32, 4, 9, 8, 16>(p, st); else launch_rp<32, 4, 9, 8>(p, st); }
        else if (p.N == 16 && p.R == 4) { if (wide) launch_rp<16, 4, 9, 8, 16>(p, st); else launch_rp<16, 4, 9, 8>(p, st); }
    } else if (p.dbg & 128) {   // experiment: 16 epilogue warps paid for with a shallower weight ring
        if (p.N == 32 && p.R == 3) launch_rp<32, 3, 4, 32, 16>(p, st);
        else if (p.N == 64 && p.R == 3) launch_rp<64, 3, 2, 32, 16>(p, st);
        else if (p.N == 128 && p.R == 2) launch_rp<128, 2, 3, 32, 16>(p, st);
    } else if (p.N == 32 && p.R == 3) launch_rp<32, 3, 8, 32>(p, st);
    else if (p.N == 64 && p.R == 3) launch_rp<64, 3, 4, 32>(p, st);
    else if (p.N == 128 && p.R == 2) launch_rp<128, 2, 4, 32>(p, st);
}

}  // namespace srt
