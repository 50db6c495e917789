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
//   * ring of mbarrier-guarded stages; one output tile per CTA.
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

size_t conv_tc_smem_bytes(int n_tile, int* stages_out)
{
    const size_t stage = kABytes + (size_t)n_tile * kKB * 4;
    // n_tile <= 128: aim for two CTAs per SM (epilogue of one overlaps the main loop of the other)
    const size_t budget = n_tile <= 128 ? 110 * 1024 : 220 * 1024;
    int stages = (int)((budget - sizeof(ConvSmemHeader) - 1024) / stage);
    if (stages > 8) stages = 8;
    if (stages < 2) stages = 2;
    if (stages_out) *stages_out = stages;
    return sizeof(ConvSmemHeader) + 1024 + stages * stage;
}

template <int N_TILE>
__global__ void __launch_bounds__(kConvThreads) conv_tc_kernel(const __grid_constant__ ConvParams p, int stages)
{
    extern __shared__ uint8_t smem_raw[];
    ConvSmemHeader* hdr = reinterpret_cast<ConvSmemHeader*>(smem_raw);
    const uint32_t tiles_base = (ptx::smem_u32(smem_raw) + (uint32_t)sizeof(ConvSmemHeader) + 1023u) & ~1023u;
    uint8_t* tiles = smem_raw + (tiles_base - ptx::smem_u32(smem_raw));
    constexpr int kBBytes = N_TILE * kKB * 4;
    constexpr int kStageBytes = kABytes + kBBytes;
    constexpr int kTmemCols = N_TILE < 32 ? 32 : N_TILE;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tx = blockIdx.x % p.tiles_x, ty = blockIdx.x / p.tiles_x;
    const int nt = blockIdx.y % p.n_tiles, phase = blockIdx.y / p.n_tiles;
    const int s = blockIdx.z / p.tiles_n, tz = blockIdx.z % p.tiles_n;
    const int nkb = p.nkb[phase];

    // ---- one-time setup -----------------------------------------------------------------
    {
        const KBlock* src = p.kb + p.kb_off[phase];
        for (int i = threadIdx.x; i < nkb; i += kConvThreads) hdr->kb[i] = src[i];
    }
    if (warp == 8 && lane == 0) {
        ptx::tma_prefetch_desc(&p.tmap[0]);
        ptx::tma_prefetch_desc(&p.tmap[1]);
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
            const int x0 = tx * p.tw, y0 = ty * p.th, n0 = s * p.B + tz * p.nb;
            const float* wsrc = p.w + (size_t)s * p.w_stem_stride + p.w_phase_off[phase] + (size_t)nt * nkb * N_TILE * kKB;
            int stage = 0;
            uint32_t ph = 0;
            for (int k = 0; k < nkb; k++) {
                ptx::mbar_wait(&hdr->empty[stage], ph ^ 1);
                const KBlock kb = hdr->kb[k];
                uint8_t* a_dst = tiles + (size_t)stage * kStageBytes;
                ptx::mbar_arrive_expect_tx(&hdr->full[stage], kStageBytes);
                ptx::tma_load_4d(a_dst, &p.tmap[kb.src], &hdr->full[stage], kb.c_off, x0 + kb.dx, y0 + kb.dy, n0);
                ptx::bulk_load_1d(a_dst + kABytes, wsrc + (size_t)k * N_TILE * kKB, kBBytes, &hdr->full[stage]);
                if (++stage == stages) { stage = 0; ph ^= 1; }
            }
        }
    } else if (warp == 9) {
        // ===== MMA issuer ================================================================
        if (ptx::elect_one()) {   // not `lane == 0`: see srt_ptx.cuh (straight-line UTCHMMA / UTMALDG issue)
            constexpr uint32_t idesc = ptx::umma_idesc_tf32(kTileM, N_TILE);
            int stage = 0;
            uint32_t ph = 0;
            for (int k = 0; k < nkb; k++) {
                ptx::mbar_wait(&hdr->full[stage], ph);
                ptx::tc_fence_after();
                const uint32_t a_lo = ptx::umma_desc_lo(tiles_base + (uint32_t)stage * kStageBytes);
                const uint32_t b_lo = a_lo + (kABytes >> 4);
#pragma unroll
                for (int kk = 0; kk < kKB / 8; kk++)
                    ptx::mma_tf32_ss_lo(tmem_d, a_lo + kk * 2, b_lo + kk * 2, idesc, (kk != 0) ? 1u : (k != 0 ? 1u : 0u));
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
        const int X = tx * p.tw + x, Y = ty * p.th + y, b = tz * p.nb + nn;
        const bool valid = X < p.Ws && Y < p.Hs && b < p.Bv;
        const int n = s * p.B + b;
        ptx::mbar_wait(&hdr->tmem_full, 0);
        ptx::tc_fence_after();
#pragma unroll 1
        constexpr int kHalfCols = N_TILE >= 32 ? N_TILE / 2 : N_TILE;   // contiguous column halves per warp pair
        for (int c0 = half * kHalfCols; c0 < (half + 1) * kHalfCols && c0 < N_TILE; c0 += 16) {
            float v[16];
            ptx::tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            // tmem_full means every MMA has finished reading the stage ring: its first 16 KB become the warps' staging buffers
            epilogue16(p, s, n, Y, X, phase, nt * N_TILE + c0, v, reinterpret_cast<float4*>(tiles) + warp * 128, valid);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<kTmemCols>(tmem_d);
    }
}

template <int N_TILE>
static void launch_one(const ConvParams& p, cudaStream_t st)
{
    int stages;
    const size_t smem = conv_tc_smem_bytes(N_TILE, &stages);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(conv_tc_kernel<N_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    dim3 grid(p.tiles_x * p.tiles_y, p.n_tiles * p.phases, p.S * p.tiles_n);
    conv_tc_kernel<N_TILE><<<grid, kConvThreads, smem, st>>>(p, stages);
}

void launch_conv_tc(const ConvParams& p, cudaStream_t st)
{
    switch (p.n_tile) {
    case 16: launch_one<16>(p, st); break;
    case 32: launch_one<32>(p, st); break;
    case 64: launch_one<64>(p, st); break;
    case 128: launch_one<128>(p, st); break;
    case 256: launch_one<256>(p, st); break;
    default: break;
    }
}

}  // namespace srt
