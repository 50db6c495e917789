// srt_conv_rp.cu — "row-patch" tcgen05 kernel for the small-N layers (down2, down3, up4, up5).
//
// The generic kernel (srt_conv_tc.cu) fetches a fresh 16 KB activation tile per tap, which makes
// layers with few output channels L2->SM-bandwidth bound (ncu, profiles/r1a_*: up5 at 59 TF/s).
// Here a CTA owns R output rows x 128 columns of one image and, per 32-channel slab, loads ONE
// (R+2) x 136-pixel patch with a single TMA box; all taps of the slab are MMAs whose A-operand
// descriptors point at shifted windows of that patch:
//        start = patch + (r + dy + 1) * row_pitch + (dx + 1) * 128 B (+ 32 B per K step)
// Rows of the patch are 136 pixels * 128 B = 17 KB apart (a multiple of the 1 KB swizzle atom);
// the +-1 pixel column shift moves the start by 128 B inside the atom.  Measured on B200
// (tools/rp_probe.py): the 128B swizzle is a function of the absolute shared-memory address, so the
// descriptor's base-offset field stays 0 for such windows (setting it to (addr>>7)&7 is wrong).  Decoder layers fuse the four output parities into N = 4*cout.
// Weights stream through a small ring of pre-swizzled [N][32] blocks, one per (slab, tap).
// Roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (shared with the generic kernel).
#include "srt_epilogue.cuh"
#include "srt_kernels.cuh"
#include "srt_ptx.cuh"

namespace srt {

constexpr int kRpThreads = 192;
constexpr int kRowPitch = kPatchW * 128;
constexpr int kRpMaxChunks = 8, kRpMaxKB = 80, kRpMaxWStages = 4;

struct RpHeader {
    uint64_t patch_full, patch_empty, tmem_full;
    uint64_t w_full[kRpMaxWStages], w_empty[kRpMaxWStages];
    uint32_t tmem_base, pad;
    RowChunk chunks[kRpMaxChunks];
    KBlock kb[kRpMaxKB];
};

template <int N, int R, int WS>
constexpr size_t rp_smem_bytes() { return sizeof(RpHeader) + 1024 + (size_t)(R + 2) * kRowPitch + (size_t)WS * N * 128; }

__device__ __forceinline__ uint64_t rp_desc(uint32_t saddr, int bo_mode)
{
    uint64_t d = ptx::umma_desc_sw128(saddr);
    if (bo_mode == 1) d |= (uint64_t)((saddr >> 7) & 7u) << 49;   // matrix base offset: phase of the start inside the 1 KB atom
    return d;
}

template <int N, int R, int WS>
__global__ void __launch_bounds__(kRpThreads) conv_rp_kernel(const __grid_constant__ RowConvParams p)
{
    extern __shared__ uint8_t smem_raw[];
    RpHeader* hdr = reinterpret_cast<RpHeader*>(smem_raw);
    const uint32_t patch_base = (ptx::smem_u32(smem_raw) + (uint32_t)sizeof(RpHeader) + 1023u) & ~1023u;
    uint8_t* patch = smem_raw + (patch_base - ptx::smem_u32(smem_raw));
    constexpr int kPatchBytes = (R + 2) * kRowPitch;
    constexpr int kWBytes = N * 128;
    constexpr int kCols = R * N;
    constexpr int kTmemCols = kCols <= 32 ? 32 : kCols <= 64 ? 64 : kCols <= 128 ? 128 : kCols <= 256 ? 256 : 512;
    uint8_t* wring = patch + kPatchBytes;
    const uint32_t wring_base = patch_base + kPatchBytes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tx = blockIdx.x % p.tiles_x, ty = blockIdx.x / p.tiles_x;
    const int s = blockIdx.z / p.ep.Bv, b = blockIdx.z % p.ep.Bv;
    const int n = s * p.ep.B + b;
    const int x0 = tx * kTileM, y0 = ty * R;

    for (int i = threadIdx.x; i < p.n_chunks; i += kRpThreads) hdr->chunks[i] = p.chunks[i];
    for (int i = threadIdx.x; i < p.nkb; i += kRpThreads) hdr->kb[i] = p.kb[i];
    if (warp == 0 && lane == 0) {
        ptx::tma_prefetch_desc(&p.tmap[0]);
        ptx::tma_prefetch_desc(&p.tmap[1]);
        ptx::mbar_init(&hdr->patch_full, 1);
        ptx::mbar_init(&hdr->patch_empty, 1);
        ptx::mbar_init(&hdr->tmem_full, 1);
        for (int i = 0; i < WS; i++) {
            ptx::mbar_init(&hdr->w_full[i], 1);
            ptx::mbar_init(&hdr->w_empty[i], 1);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) ptx::tmem_alloc<kTmemCols>(&hdr->tmem_base);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_d = hdr->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            // ===== producer: one patch per slab, one weight block per (slab, tap) ==============
            const float* wsrc = p.w + (size_t)s * p.w_stem_stride;
            int ws = 0;
            uint32_t wph = 0, pph = 0;
            for (int ci = 0; ci < p.n_chunks; ci++) {
                const RowChunk ch = hdr->chunks[ci];
                ptx::mbar_wait(&hdr->patch_empty, pph ^ 1);
                ptx::mbar_arrive_expect_tx(&hdr->patch_full, kPatchBytes);
                ptx::tma_load_4d(patch, &p.tmap[ch.src], &hdr->patch_full, ch.c_off, x0 - 1, y0 - 1, n);
                pph ^= 1;
                for (int k = ch.kb0; k < ch.kb0 + ch.nkb; k++) {
                    ptx::mbar_wait(&hdr->w_empty[ws], wph ^ 1);
                    ptx::mbar_arrive_expect_tx(&hdr->w_full[ws], kWBytes);
                    ptx::bulk_load_1d(wring + (size_t)ws * kWBytes, wsrc + (size_t)k * N * kKB, kWBytes, &hdr->w_full[ws]);
                    if (++ws == WS) { ws = 0; wph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer ==================================================================
            constexpr uint32_t idesc = ptx::umma_idesc_tf32(kTileM, N);
            int ws = 0;
            uint32_t wph = 0, pph = 0;
            bool first = true;
            for (int ci = 0; ci < p.n_chunks; ci++) {
                const RowChunk ch = hdr->chunks[ci];
                ptx::mbar_wait(&hdr->patch_full, pph);
                pph ^= 1;
                for (int k = ch.kb0; k < ch.kb0 + ch.nkb; k++) {
                    const KBlock kb = hdr->kb[k];
                    ptx::mbar_wait(&hdr->w_full[ws], wph);
                    ptx::tc_fence_after();
                    const uint32_t b_addr = wring_base + (uint32_t)ws * kWBytes;
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        const uint32_t a_addr = patch_base + (uint32_t)(r + kb.dy + 1) * kRowPitch + (uint32_t)(kb.dx + 1) * 128u;
#pragma unroll
                        for (int kk = 0; kk < kKB / 8; kk++)
                            ptx::mma_tf32_ss(tmem_d + (uint32_t)(r * N), rp_desc(a_addr + kk * 32, p.bo_mode),
                                             ptx::umma_desc_sw128(b_addr + kk * 32), idesc, (!first || kk != 0) ? 1u : 0u);
                    }
                    first = false;
                    ptx::mma_commit(&hdr->w_empty[ws]);
                    if (++ws == WS) { ws = 0; wph ^= 1; }
                }
                ptx::mma_commit(&hdr->patch_empty);
            }
            ptx::mma_commit(&hdr->tmem_full);
        }
    } else {
        // ===== epilogue ========================================================================
        const int q = warp & 3;
        const int m = q * 32 + lane;
        const int X = x0 + m;
        ptx::mbar_wait(&hdr->tmem_full, 0);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int r = 0; r < R; r++) {
            const int Y = y0 + r;
            const bool valid = X < p.ep.Ws && Y < p.ep.Hs;
#pragma unroll 1
            for (int c0 = 0; c0 < N; c0 += 16) {
                float v[16];
                ptx::tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(r * N + c0), v);
                if (valid) {
                    if (p.ep.mode == 2) epilogue16(p.ep, s, n, Y, X, c0 / p.ep.cout, c0 % p.ep.cout, v);
                    else epilogue16(p.ep, s, n, Y, X, 0, c0, v);
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<kTmemCols>(tmem_d);
    }
}

template <int N, int R, int WS>
static void launch_rp(const RowConvParams& p, cudaStream_t st)
{
    constexpr size_t smem = rp_smem_bytes<N, R, WS>();
    static_assert(smem <= 115712, "two CTAs per SM must fit");
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(conv_rp_kernel<N, R, WS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    dim3 grid(p.tiles_x * p.tiles_y, 1, p.ep.S * p.ep.Bv);
    conv_rp_kernel<N, R, WS><<<grid, kRpThreads, smem, st>>>(p);
}

void launch_conv_rp(const RowConvParams& p, cudaStream_t st)
{
    if (p.N == 32 && p.R == 3) launch_rp<32, 3, 4>(p, st);
    else if (p.N == 64 && p.R == 3) launch_rp<64, 3, 3>(p, st);
    else if (p.N == 128 && p.R == 2) launch_rp<128, 2, 2>(p, st);
}

}  // namespace srt
