// srt_conv_tc.cu — the dilated-Conv2D / transposed-Conv2D layers of the U-Net as an implicit
// GEMM on the 5th-generation tensor cores (sm_100a): replaces im2col_dilated_cpu + gemm and
// gemm + col2im_dilated_cpu of the reference (Executable/spleeter.c:73-78, 96-100).
//
//   D[128 pixels, N couts] = sum over k-blocks  A_kb[128, 32] * W_kb[N, 32]^T      (TF32 in, fp32 accumulate)
//
//   * A_kb is one TMA box {32 ch, tw, th, nb} of a source activation tensor fetched at the
//     whole-pixel offset (dx, dy) the k-block table prescribes; out-of-image pixels are
//     zero-filled by the TMA unit (= the reference's zero padding).  128B-swizzled, K-major.
//   * W_kb is a pre-packed, pre-swizzled [N][32] block streamed with a 1-D bulk copy.
//   * warps 0..7 = epilogue, warp 8 = TMA producer, warp 9 = tcgen05.mma issuer (accumulator lives in TMEM)
//     (epilogue (tcgen05.ld -> bias / BN / activation -> global, see srt_epilogue.cuh).
//   * ring of mbarrier-guarded stages; MT = 1 or 2 output tiles (of the same stem, phase and n-tile) per CTA.
//   * Why MT = 2: an SM takes in ~64 B/clk from L2, and a k-block brings 16 KB of activations + N*128 B of weights for
//     4 MMAs = 2N cycles, so with one tile per CTA the tensor pipe cannot exceed 128N / (16K + 128N) = 67 % (N = 256),
//     50 % (128), 33 % (64) - ncu r1p measures 55-75 %, 43-60 %, 38 %.  (Cluster multicast of the weights does not help:
//     it saves L2 reads, not SM ingress; measured slower, profiles/r1q_tc_cluster_multicast_sweep.txt.)  Two pixel tiles
//     that share one weight block halve the weight bytes per MMA: 100 % / 67 % / 40 %.
#include <cstdlib>

#include "srt_epilogue.cuh"
#include "srt_kernels.cuh"
#include "srt_ptx.cuh"

namespace srt {

constexpr int kConvThreads = 320;   // producer, MMA issuer, 8 epilogue warps
constexpr int kABytes = kTileM * kKB * 4;   // 16 KiB per stage
constexpr int kMaxKB = 512;

struct ConvSmemHeader {
    uint64_t full[8];
    uint64_t empty[8];
    uint64_t tmem_full;
    uint32_t tmem_base;
    uint32_t pad;
    KBlock kb[kMaxKB];
};

size_t conv_tc_smem_bytes(int n_tile, int mt, int* stages_out)
{
    const size_t stage = (size_t)mt * kABytes + (size_t)n_tile * kKB * 4;
    // small stages: aim for two CTAs per SM (epilogue of one overlaps the main loop of the other)
    const size_t budget = (n_tile <= 128 && mt == 1) ? 110 * 1024 : 220 * 1024;
    int stages = (int)((budget - sizeof(ConvSmemHeader) - 1024) / stage);
    if (stages > 8) stages = 8;
    if (stages < 2) stages = 2;
    if (stages_out) *stages_out = stages;
    return sizeof(ConvSmemHeader) + 1024 + stages * stage;
}

// pair_mode (MT = 2): 0 = the CTA's tiles are neighbours in the (x, y) tile grid of one image group, 1 = the same (x, y)
// tile of two consecutive image groups (layers whose image is a single tile).
template <int N_TILE, int MT>
__global__ void __launch_bounds__(kConvThreads) conv_tc_kernel(const __grid_constant__ ConvParams p, int stages, int pair_mode)
{
    extern __shared__ uint8_t smem_raw[];
    ConvSmemHeader* hdr = reinterpret_cast<ConvSmemHeader*>(smem_raw);
    const uint32_t tiles_base = (ptx::smem_u32(smem_raw) + (uint32_t)sizeof(ConvSmemHeader) + 1023u) & ~1023u;
    uint8_t* tiles = smem_raw + (tiles_base - ptx::smem_u32(smem_raw));
    constexpr int kBBytes = N_TILE * kKB * 4;
    constexpr int kStageBytes = MT * kABytes + kBBytes;
    constexpr int kTmemCols = MT * N_TILE < 32 ? 32 : MT * N_TILE;
    static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM allocation is a power of two <= 512 columns");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nt = blockIdx.y % p.n_tiles, phase = blockIdx.y / p.n_tiles;
    const int nkb = p.nkb[phase];
    // tile m of this CTA: (tx, ty) in the tile grid, tz = image group.  Out-of-range tiles (odd counts) run on zero-filled
    // boxes and store nothing.
    int s, txm[MT], tym[MT], tzm[MT];
    if (MT == 1 || pair_mode == 0) {
        s = blockIdx.z / p.tiles_n;
        for (int m = 0; m < MT; m++) {
            const int pt = blockIdx.x * MT + m;
            txm[m] = pt % p.tiles_x; tym[m] = pt / p.tiles_x; tzm[m] = blockIdx.z % p.tiles_n;
        }
    } else {
        const int groups = (p.tiles_n + MT - 1) / MT;
        s = blockIdx.z / groups;
        for (int m = 0; m < MT; m++) {
            txm[m] = blockIdx.x % p.tiles_x; tym[m] = blockIdx.x / p.tiles_x; tzm[m] = (blockIdx.z % groups) * MT + m;
        }
    }

    // ---- one-time setup -----------------------------------------------------------------
    {
        const KBlock* src = p.kb + p.kb_off[phase];
        for (int i = threadIdx.x; i < nkb; i += kConvThreads) hdr->kb[i] = src[i];
    }
    if (warp == 8 && lane == 0) {
        ptx::tma_prefetch_desc(&p.tmap[0]);
        ptx::tma_prefetch_desc(&p.tmap[1]);
        ptx::tma_prefetch_desc(&p.tmap[2]);
        for (int i = 0; i < stages; i++) {
            ptx::mbar_init(&hdr->full[i], 1);
            ptx::mbar_init(&hdr->empty[i], 1);
        }
        ptx::mbar_init(&hdr->tmem_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 9) ptx::tmem_alloc<kTmemCols>(&hdr->tmem_base);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_d = hdr->tmem_base;

    // the issue arbiter favours higher warp ids: producer and MMA issuer sit above the 8 epilogue warps
    if (warp == 8) {
        // ===== TMA producer ==============================================================
        if (ptx::elect_one()) {   // not `lane == 0`: see srt_ptx.cuh (straight-line UTCHMMA / UTMALDG issue)
            const float* wsrc = p.w + (size_t)s * p.w_stem_stride + p.w_phase_off[phase] + (size_t)nt * nkb * N_TILE * kKB;
            int stage = 0;
            uint32_t ph = 0;
            for (int k = 0; k < nkb; k++) {
                ptx::mbar_wait(&hdr->empty[stage], ph ^ 1);
                const KBlock kb = hdr->kb[k];
                uint8_t* a_dst = tiles + (size_t)stage * kStageBytes;
                ptx::mbar_arrive_expect_tx(&hdr->full[stage], kStageBytes);
#pragma unroll
                for (int m = 0; m < MT; m++)
                    ptx::tma_load_4d(a_dst + m * kABytes, &p.tmap[kb.src], &hdr->full[stage], kb.c_off, txm[m] * p.tw + kb.dx, tym[m] * p.th + kb.dy,
                                     s * p.B + tzm[m] * p.nb);
                ptx::bulk_load_1d(a_dst + MT * kABytes, wsrc + (size_t)k * N_TILE * kKB, kBBytes, &hdr->full[stage]);
                if (++stage == stages) { stage = 0; ph ^= 1; }
            }
        }
    } else if (warp == 9) {
        // ===== MMA issuer ================================================================
        if (ptx::elect_one()) {   // not `lane == 0`: see srt_ptx.cuh (straight-line UTCHMMA / UTMALDG issue)
            constexpr uint32_t idesc = ptx::umma_idesc_tf32(kTileM, N_TILE), idesc_lo = ptx::umma_idesc_bf16(kTileM, N_TILE);
            int stage = 0;
            uint32_t ph = 0;
            for (int k = 0; k < nkb; k++) {
                const int part = hdr->kb[k].part;
                const bool comp = (part & kPartLo) != 0;   // compensation block: bf16 (4 x K = 16) or e5m2 (4 x K = 32) operands over the same 128-byte rows
                ptx::mbar_wait(&hdr->full[stage], ph);
                ptx::tc_fence_after();
                const uint32_t a_lo = ptx::umma_desc_lo(tiles_base + (uint32_t)stage * kStageBytes);
                const uint32_t b_lo = a_lo + ((MT * kABytes) >> 4);
                if (comp && (part & kPartLo8)) {
#pragma unroll
                    for (int kk = 0; kk < 4; kk++)
#pragma unroll
                        for (int m = 0; m < MT; m++)
                            ptx::mma_f8_ss_lo(tmem_d + m * N_TILE, a_lo + m * (kABytes >> 4) + kk * 2, b_lo + kk * 2, idesc_lo, (kk != 0) ? 1u : (k != 0 ? 1u : 0u));
                } else if (comp) {
#pragma unroll
                    for (int kk = 0; kk < 4; kk++)
#pragma unroll
                        for (int m = 0; m < MT; m++)
                            ptx::mma_bf16_ss_lo(tmem_d + m * N_TILE, a_lo + m * (kABytes >> 4) + kk * 2, b_lo + kk * 2, idesc_lo, (kk != 0) ? 1u : (k != 0 ? 1u : 0u));
                } else
#pragma unroll
                for (int kk = 0; kk < kKB / 8; kk++)
#pragma unroll
                    for (int m = 0; m < MT; m++)   // consecutive MMAs hit different accumulators
                        ptx::mma_tf32_ss_lo(tmem_d + m * N_TILE, a_lo + m * (kABytes >> 4) + kk * 2, b_lo + kk * 2, idesc, (kk != 0) ? 1u : (k != 0 ? 1u : 0u));
                ptx::mma_commit(&hdr->empty[stage]);   // frees the stage once these MMAs retire
                if (++stage == stages) { stage = 0; ph ^= 1; }
            }
            ptx::mma_commit(&hdr->tmem_full);          // accumulator complete
        }
    } else {
        // ===== epilogue: TMEM -> registers -> bias/BN/act -> global ======================
        const int q = warp & 3, half = warp >> 2;   // TMEM lane quarter this warp may read; column half
        const int m = q * 32 + lane;
        const int x = m % p.tw, y = (m / p.tw) % p.th, nn = m / (p.tw * p.th);
        ptx::mbar_wait(&hdr->tmem_full, 0);
        ptx::tc_fence_after();
        constexpr int kHalfCols = N_TILE >= 32 ? N_TILE / 2 : N_TILE;   // contiguous column halves per warp pair
#pragma unroll 1
        for (int t = 0; t < MT; t++) {
            const int X = txm[t] * p.tw + x, Y = tym[t] * p.th + y, b = tzm[t] * p.nb + nn;
            const bool valid = X < p.Ws && Y < p.Hs && b < p.Bv;
            const int n = s * p.B + b;
#pragma unroll 1
            for (int c0 = half * kHalfCols; c0 < (half + 1) * kHalfCols && c0 < N_TILE; c0 += 16) {
                float v[16];
                ptx::tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * N_TILE + c0), v);
                // tmem_full means every MMA has finished reading the stage ring: its first 16 KB become the warps' staging buffers
                const int col = nt * N_TILE + c0;
                epilogue16(p, s, n, Y, X, p.fused ? col / p.cout : phase, p.fused ? col % p.cout : col, v, reinterpret_cast<float4*>(tiles) + warp * 128, valid);
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<kTmemCols>(tmem_d);
    }
}

template <int N_TILE, int MT>
static void launch_one(const ConvParams& p, int pair_mode, cudaStream_t st)
{
    int stages;
    const size_t smem = conv_tc_smem_bytes(N_TILE, MT, &stages);
    static LaunchState state;
    state.prepare(conv_tc_kernel<N_TILE, MT>, smem);
    const int txy = p.tiles_x * p.tiles_y;
    dim3 grid(txy, p.n_tiles * p.phases, p.S * p.tiles_n);
    if (MT > 1) {
        if (pair_mode == 0) grid.x = (txy + MT - 1) / MT;
        else grid.z = p.S * ((p.tiles_n + MT - 1) / MT);
    }
    conv_tc_kernel<N_TILE, MT><<<grid, kConvThreads, smem, st>>>(p, stages, pair_mode);
}

bool conv_tc_supported(int n_tile) { return n_tile == 16 || n_tile == 32 || n_tile == 64 || n_tile == 128 || n_tile == 256; }

void launch_conv_tc(const ConvParams& p, cudaStream_t st, int sm_count)
{
    // Two tiles per CTA where the weight stream is the larger part of the SM's intake (see the header).  Measured per layer
    // on the 32-stream bench (profiles/r1q_tc_mt_sweep.txt): down5 0.212 -> 0.198, down6 0.164 -> 0.148, up2 0.362 -> 0.341 ms,
    // but up1 0.186 -> 0.192 and down4 0.243 -> 0.260 (fewer, longer CTAs: the last wave is emptier and, for N = 128, only one
    // CTA fits per SM).  Default: the deep encoder layers (N = 256, one phase) and the N = 128 decoder layer.
    // SRT_TC_MT = minimum N that gets MT = 2 regardless of the layer (0 = never).
    static const int mt_min_n = [] { const char* e = getenv("SRT_TC_MT"); return e ? atoi(e) : -1; }();
    const bool want = mt_min_n < 0 ? ((p.n_tile == 256 && p.phases == 1) || (p.n_tile == 128 && p.phases == 4)) : (mt_min_n > 0 && p.n_tile >= mt_min_n);
    const int txy = p.tiles_x * p.tiles_y;
    int pair_mode = -1;
    // pairing halves the CTA count: only where the unpaired grid already covers the SMs
    if (want && (long)txy * p.tiles_n * p.S * p.n_tiles * p.phases >= sm_count) {
        if (txy % 2 == 0) pair_mode = 0;
        else if (txy == 1 && p.tiles_n > 1) pair_mode = 1;
    }
    if (pair_mode >= 0) {
        switch (p.n_tile) {
        case 64: launch_one<64, 2>(p, pair_mode, st); return;
        case 128: launch_one<128, 2>(p, pair_mode, st); return;
        case 256: launch_one<256, 2>(p, pair_mode, st); return;
        default: break;
        }
    }
    switch (p.n_tile) {
    case 16: launch_one<16, 1>(p, 0, st); break;
    case 32: launch_one<32, 1>(p, 0, st); break;
    case 64: launch_one<64, 1>(p, 0, st); break;
    case 128: launch_one<128, 1>(p, 0, st); break;
    case 256: launch_one<256, 1>(p, 0, st); break;
    default: break;   // unreachable: srt_create rejects plans with an N tile conv_tc_supported() does not list
    }
}

}  // namespace srt
