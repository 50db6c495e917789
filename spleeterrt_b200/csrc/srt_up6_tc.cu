// srt_up6_tc.cu — up6 (5x5 stride-2 transposed conv, [skip1 | up5] 32 ch -> 1 ch, + act + BN;
// Executable/spleeter.c:289-294) as "GEMM, then col2im" on the tensor cores — the reference's own
// formulation (gemm + col2im_dilated_cpu, im2col_dilated.c:42-65), re-tiled for one SM:
//
//     G[pixel][tap] = sum_c X[pixel][c] * W[c][tap]          25 taps (padded to N = 32), K = 32
//     out[2h + kh - 1][2w + kw - 1] += G[(h, w)][kh*5 + kw]  each output sums 4/6/6/9 neighbours
//
// The SIMT kernel (srt_unet_simt.cu) spends 1600 FMAs per input pixel on this layer and is FP32-pipe
// bound (23 TF/s, 1.16 ms per 32-stream step).  Here the GEMM is 8 MMAs per 128 pixels and the scatter
// becomes a 25-value gather per input pixel out of shared memory, so the layer is bound by reading its
// two 16-channel inputs once (HBM).
//
// A persistent CTA walks "units" = (stem, image, column block, row chunk).  Inside a unit it walks
// input rows top to bottom; per row:
//   warp 12      TMA: 2 boxes {16 ch, 128 px} (one per source, SWIZZLE_64B, OOB = zero).  One box per 8-channel K step
//                (4 boxes of 32-byte rows) was fetch-bound: with everything but the TMA switched off the kernel
//                still took 0.77 of its 0.85 ms (profiles/r1m_up6_stage_sweep.txt) - the copy engine works per
//                contiguous fragment, and 64-byte fragments halve their number.
//   warps 8-11   split, one warp per row: hi = the tile as loaded (the MMA sees a truncated to TF32), lo = a - trunc_tf32(a)
//                (second tile).  The two inputs are fp32 and this layer feeds the mask almost directly, so
//                plain TF32 operands would double the stem error of small nets (oracle emulation,
//                DESIGN.md); hi + lo keeps fp32 accuracy.
//   warp 13      MMA: G_row = A_hi*W_hi + A_lo*W_hi (+ A_hi*W_lo for weights that are not TF32-exact),
//                N = 32, two rows interleaved on two TMEM accumulators
//   warps 0-7    TMEM -> G ring in shared memory ([4 rows][25 taps][128 px]), then the gather for the
//                previous input row (needs rows y-1, y, y+1), bias + act + BN, float2 stores.
// Column blocks overlap by one pixel on each side (the tile starts at x0 - 1), rows chunks by one row.
#include "srt_epilogue.cuh"
#include "srt_kernels.cuh"
#include "srt_plan.h"
#include "srt_ptx.cuh"

namespace srt {

constexpr int kU6Threads = 448;            // 8 epilogue + 4 split + TMA + MMA warps
constexpr int kU6Stages = 8;               // input rows in flight (the kernel is bound by the TMA -> split -> MMA -> release
                                           // latency chain per ring slot: 2 / 3 / 4 stages ran 1.40 / 1.16 / 0.85 ms, profiles/r1m_up6_stage_sweep.txt).
                                           // 4 fit with fp32 residual tiles (2 x 16 KB per row), 6 with the 8-bit ones (16 + 4 KB); beyond 6 nothing moves
                                           // (profiles/r2_up6_sweep.txt)
constexpr int kU6Lo8Bytes = 128 * 32;      // residual tile of a row in the 8-bit form: [128 px][E1 16 ch | U5 16 ch] e5m2, SWIZZLE_32B
constexpr int kU6AccSlots = 8;             // TMEM accumulators (32 columns each)
constexpr int kU6RowBytes = 2 * 128 * 64;  // 2 boxes x 128 pixels x 16 channels fp32 = 16 KB
constexpr int kU6BoxBytes = 128 * 64;
constexpr int kU6GSlot = 25 * 128;         // floats per G row
// Two forms of the epilogue loop share the shared memory differently (Up6TcParams.pair):
//   single rows: 4 G rows (3 gathered from + 1 written), 160 KB of operand rings = 8 stages x (16 + 4 KB) or 5 x (16 + 16 KB)
//   row pairs:   6 G rows (4 + 2),                      128 KB                  = 6 stages             or 4
__host__ __device__ constexpr int u6_g_slots(int pair) { return pair ? 6 : 4; }
__host__ __device__ constexpr int u6_ring_bytes(int pair) { return (pair ? 128 : 160) * 1024; }

struct U6Header {
    uint64_t a_full[kU6Stages], a_ready[kU6Stages], a_empty[kU6Stages];
    uint64_t acc_full[kU6AccSlots], acc_empty[kU6AccSlots];
    uint64_t w_full, w_idle;       // weights of the current stem landed / every MMA that read the previous stem's weights retired
    uint32_t tmem_base, pad;
};

static size_t up6_tc_smem_bytes(int S, int pair)
{
    (void)S;   // only the current stem's weights are resident (9 KB): the space of the other stems buys ring stages
    return sizeof(U6Header) + 1024 + (size_t)u6_ring_bytes(pair) + (size_t)kUp6TcWFloatsPerStem * 4 + (size_t)u6_g_slots(pair) * kU6GSlot * 4;
}

struct U6Unit {
    int s, n, x0, r0, r1;
};
__device__ __forceinline__ U6Unit u6_unit(const Up6TcParams& p, int u)
{
    U6Unit o;
    const int tx = u % p.blocks_x;
    u /= p.blocks_x;
    const int rc = u % p.chunks;
    u /= p.chunks;
    const int b = u % p.Bv;
    o.s = u / p.Bv;
    o.n = o.s * p.B + b;
    o.x0 = tx * p.bw;
    o.r0 = rc * p.rows_per_unit;
    o.r1 = min(p.T / 2, o.r0 + p.rows_per_unit);
    return o;
}

__global__ void __launch_bounds__(kU6Threads, 1) up6_tc_kernel(const __grid_constant__ Up6TcParams p)
{
    extern __shared__ uint8_t smem_raw[];
    U6Header* hdr = reinterpret_cast<U6Header*>(smem_raw);
    const uint32_t a_base = (ptx::smem_u32(smem_raw) + (uint32_t)sizeof(U6Header) + 1023u) & ~1023u;
    uint8_t* a_raw = smem_raw + (a_base - ptx::smem_u32(smem_raw));   // [stage][source][128][16] fp32, hi after the split
    const uint32_t lo_stage = p.lo8 ? (uint32_t)kU6Lo8Bytes : (uint32_t)kU6RowBytes;
    uint8_t* a_lo = a_raw + p.stages * kU6RowBytes;                   // residuals: same layout in fp32, or one [128][32] e5m2 tile per row
    const uint32_t lo_base = a_base + p.stages * kU6RowBytes;
    float* wsm = reinterpret_cast<float*>(a_raw + u6_ring_bytes(p.pair));       // [box][term][32][8] pre-swizzled (+ the e5m2 block), current stem
    const uint32_t w_base = a_base + u6_ring_bytes(p.pair);
    float* G = wsm + (size_t)kUp6TcWFloatsPerStem;                           // [u6_g_slots][25][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = p.T / 2, W = p.F / 2;
    const int n_units = p.blocks_x * p.chunks * p.Bv * p.S;

    if (warp == 12 && lane == 0) {
        ptx::tma_prefetch_desc(&p.tmap[0]);
        ptx::tma_prefetch_desc(&p.tmap[1]);
        for (int i = 0; i < kU6Stages; i++) {
            ptx::mbar_init(&hdr->a_full[i], 1);
            ptx::mbar_init(&hdr->a_ready[i], 1);    // the split warp that owns the row
            ptx::mbar_init(&hdr->a_empty[i], 1);
        }
        for (int i = 0; i < kU6AccSlots; i++) {
            ptx::mbar_init(&hdr->acc_full[i], 1);
            ptx::mbar_init(&hdr->acc_empty[i], 8);  // one arrival per epilogue warp
        }
        ptx::mbar_init(&hdr->w_full, 1);
        ptx::mbar_init(&hdr->w_idle, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 13) ptx::tmem_alloc<kU6AccSlots * 32>(&hdr->tmem_base);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_d = hdr->tmem_base;

    if (warp == 12) {
        // ===== TMA producer =========================================================================
        if (ptx::elect_one()) {   // not `lane == 0`: see srt_ptx.cuh (straight-line UTCHMMA / UTMALDG issue)
            int st = 0;
            uint32_t ph = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const U6Unit t = u6_unit(p, u);
                for (int y = t.r0 - 1; y <= t.r1; y++) {
                    // the ring holds only kU6Stages rows (~2 us of HBM latency would cap it at one row per ~700 cycles):
                    // pull the rows further ahead into L2 first
                    if (p.prefetch_rows > 0 && y + p.prefetch_rows <= t.r1) {
#pragma unroll
                        for (int b = 0; b < 2; b++) ptx::tma_prefetch_4d(&p.tmap[b], 0, t.x0 - 1, y + p.prefetch_rows, t.n);
                    }
                    ptx::mbar_wait(&hdr->a_empty[st], ph ^ 1);
                    if (p.dbg & 4) ptx::mbar_arrive(&hdr->a_full[st]);
                    else {
                        ptx::mbar_arrive_expect_tx(&hdr->a_full[st], kU6RowBytes);
                        uint8_t* dst = a_raw + (size_t)st * kU6RowBytes;
#pragma unroll
                        for (int b = 0; b < 2; b++)
                            ptx::tma_load_4d(dst + b * kU6BoxBytes, &p.tmap[b], &hdr->a_full[st], 0, t.x0 - 1, y, t.n);
                    }
                    if (++st == p.stages) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp >= 8 && warp < 12) {
        // ===== split: hi (in place) / lo ============================================================
        // One warp per ROW (row k of this CTA goes to split warp k & 3): the fixed latencies of a row - the wait for the TMA, the
        // proxy fence, the arrival - then overlap across four rows instead of adding up once per row (with the four warps sharing
        // every row this stage alone took 0.26 ms of the kernel's 0.73: profiles/r2_up6_sweep.txt).
        const int wsplit = warp - 8;
        int st = 0, k = 0;
        uint32_t ph = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const U6Unit t = u6_unit(p, u);
            for (int y = t.r0 - 1; y <= t.r1; y++, k++) {
                if ((k & 3) == wsplit) {
                    ptx::mbar_wait(&hdr->a_full[st], ph);
                    float4* raw = reinterpret_cast<float4*>(a_raw + (size_t)st * kU6RowBytes);
                    float4* lo = reinterpret_cast<float4*>(a_lo + (size_t)st * lo_stage);
                    if (p.lo8 && !(p.dbg & 8)) {
                        // 8-bit residuals: element e of the raw stage is (source e >> 9, pixel (e & 511) >> 2, 16-byte chunk e & 3 of the pixel's
                        // 64-byte SWIZZLE_64B row, i.e. channels 4 c .. 4 c + 3 with c = chunk ^ ((px >> 1) & 3)); its four residuals go, as
                        // e5m2(4 x), to bytes 16 source + 4 c of the pixel's 32-byte row of the SWIZZLE_32B tile (16-byte chunk ^ ((px >> 2) & 1))
                        uint32_t* lo8 = reinterpret_cast<uint32_t*>(lo);
#pragma unroll 8
                        for (int i = 0; i < 32; i++) {
                            const int e = lane + 32 * i;
                            const float4 a = raw[e];
                            const int box = e >> 9, px = (e & 511) >> 2, c = (e & 3) ^ ((px >> 1) & 3);
                            const uint32_t w = ptx::pack_e5m2x4(4.0f * (a.x - ptx::trunc_tf32(a.x)), 4.0f * (a.y - ptx::trunc_tf32(a.y)),
                                                                4.0f * (a.z - ptx::trunc_tf32(a.z)), 4.0f * (a.w - ptx::trunc_tf32(a.w)));
                            lo8[px * 8 + ((box ^ ((px >> 2) & 1)) << 2) + c] = w;
                        }
                    } else if (!(p.dbg & 8)) {
                        // fp32 residuals.  hi = the raw tile itself: the tensor core reads only the TF32 bits of an fp32 operand (sign,
                        // exponent, 10 mantissa bits), i.e. a truncated to TF32.  lo = a - trunc(a) is exact in fp32 and < 2^-10 |a|, so
                        // hi + tf32(lo) carries >= 20 mantissa bits.  Not rewriting hi saves a third of this stage's shared-memory traffic.
#pragma unroll 8
                        for (int i = 0; i < 32; i++) {
                            const float4 a = raw[lane + 32 * i];
                            lo[lane + 32 * i] = make_float4(a.x - ptx::trunc_tf32(a.x), a.y - ptx::trunc_tf32(a.y), a.z - ptx::trunc_tf32(a.z),
                                                            a.w - ptx::trunc_tf32(a.w));
                        }
                    }
                    ptx::fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&hdr->a_ready[st]);
                }
                if (++st == p.stages) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 13) {
        // ===== MMA issuer: rows in pairs, so consecutive MMAs hit different accumulators ============
        if (ptx::elect_one()) {   // not `lane == 0`: see srt_ptx.cuh (straight-line UTCHMMA / UTMALDG issue)
            constexpr uint32_t idesc = ptx::umma_idesc_tf32(kTileM, 32);
            int st = 0, as = 0, cur_s = -1;
            uint32_t ph = 0, aph = 0, wph = 0, iph = 0;
            const uint32_t w_lo = ptx::umma_desc_lo(w_base);
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                const U6Unit t = u6_unit(p, u);
                if (t.s != cur_s) {
                    // units are stem-major, so a CTA changes stem at most S - 1 times: drain the MMAs that read the old
                    // weights, then bulk-copy the new stem's 8 KB (async proxy -> async proxy, completes on w_full)
                    if (cur_s >= 0) {
                        ptx::mma_commit(&hdr->w_idle);
                        ptx::mbar_wait(&hdr->w_idle, iph);
                        iph ^= 1;
                    }
                    ptx::mbar_arrive_expect_tx(&hdr->w_full, kUp6TcWFloatsPerStem * 4);
                    ptx::bulk_load_1d(wsm, p.w + (size_t)t.s * kUp6TcWFloatsPerStem, kUp6TcWFloatsPerStem * 4, &hdr->w_full);
                    ptx::mbar_wait(&hdr->w_full, wph);
                    wph ^= 1;
                    cur_s = t.s;
                }
                for (int y = t.r0 - 1; y <= t.r1; y += 2) {   // the row count r1 - r0 + 2 is even
                    int stj[2], asj[2];
                    for (int j = 0; j < 2; j++) {
                        stj[j] = st; asj[j] = as;
                        ptx::mbar_wait(&hdr->a_ready[st], ph);
                        ptx::mbar_wait(&hdr->acc_empty[as], aph ^ 1);
                        if (++st == p.stages) { st = 0; ph ^= 1; }
                        if (++as == p.acc_slots) { as = 0; aph ^= 1; }
                    }
                    ptx::tc_fence_after();
                    // terms: (A_hi, W_hi), (A_lo, W_hi), and (A_hi, W_lo) when the weights carry a residual
                    for (int term = 0; term < ((p.dbg & 2) ? 0 : 1 + p.w_terms); term++) {
                        if (term == 1 && p.lo8) {
                            // the residual term in 8 bits: one K = 32 MMA per row over [E1 | U5] (both operands SWIZZLE_32B, e5m2)
                            const uint32_t b8 = w_lo + (uint32_t)((8 * 1024) >> 4);
#pragma unroll
                            for (int j = 0; j < 2; j++)
                                ptx::mma_f8_ss_lo(tmem_d + (uint32_t)(asj[j] * 32), ptx::umma_desc_lo(lo_base + (uint32_t)stj[j] * kU6Lo8Bytes), b8,
                                                  ptx::umma_idesc_bf16(kTileM, 32), 1u, ptx::kDescHiSw32);
                            continue;
                        }
                        const uint32_t abase = (term == 1) ? lo_base : a_base;
                        const uint32_t wterm = (term == 2) ? 1u : 0u;
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            const uint32_t b_lo = w_lo + (uint32_t)(((b * 2 + wterm) * 1024) >> 4);
#pragma unroll
                            for (int j = 0; j < 2; j++) {
                                // K step b: source b >> 1, channels 8 (b & 1) .. +7 = bytes 32 (b & 1) of the 64-byte swizzled rows
                                const uint32_t a_lo_d = ptx::umma_desc_lo(abase + (uint32_t)stj[j] * kU6RowBytes + (uint32_t)(b >> 1) * kU6BoxBytes) + (uint32_t)(b & 1) * 2;
                                ptx::mma_tf32_ss_ab(tmem_d + (uint32_t)(asj[j] * 32), a_lo_d, ptx::kDescHiSw64, b_lo, ptx::kDescHiSw32, idesc, (term | b) ? 1u : 0u);
                            }
                        }
                    }
                    for (int j = 0; j < 2; j++) {
                        ptx::mma_commit(&hdr->a_empty[stj[j]]);
                        ptx::mma_commit(&hdr->acc_full[asj[j]]);
                    }
                }
            }
        }
    } else {
        // ===== epilogue warps 0..7 =================================================================
        // Software-pipelined per PAIR of input rows (y, y + 1):  issue the TMEM loads of both rows  |  gather + store the output
        // rows of input rows y - 3 and y - 2 from G rows y-4 .. y-1 (complete and published by the previous barrier) while the
        // loads are in flight  |  wait for the loads, write G[y], G[y + 1]  |  barrier.  The gathers read four ring slots and the
        // stores go to the other two, so one 256-thread barrier per row pair is enough.  (The eight warps run in lockstep, so the
        // loop period is the latency of one iteration, ~550 cycles with every other stage switched off: profiles/r2_up6_sweep.txt.)
        const int q = warp & 3, half = warp >> 2;
        const int m = q * 32 + lane;              // TMEM lane = pixel of the tile
        const int gm = threadIdx.x & 127;         // gather: pixel
        const int po = threadIdx.x >> 7;          // gather: output row parity
        int as = 0;
        uint32_t aph = 0;
        int grow = 0;                             // rows this CTA has pushed through the G ring
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const U6Unit t = u6_unit(p, u);
            const float bias = p.bias[t.s], sc = p.bn_scale[t.s], of = p.bn_offset[t.s];
            const int act = p.act[t.s];
            const int X = t.x0 - 1 + gm;
            const bool col_ok = gm >= 1 && gm <= p.bw && X < W && !(p.dbg & 1);
            // out(2 yo + po, 2 X + qo) = sum_{dy, dx} G[yo + dy][(po + 1 - 2 dy) * 5 + (qo + 1 - 2 dx)][X + dx];
            // `c` = ring index of input row yo
            auto gather = [&](int yo, int c) {
                float o0 = 0.0f, o1 = 0.0f;
#pragma unroll
                for (int dy = -1; dy <= 1; dy++) {
                    if (po == 0 && dy == 1) continue;      // kh = -1
                    const int kh = po + 1 - 2 * dy;
                    const float* gr = G + (size_t)(p.pair ? (unsigned)(c + dy) % 6u : (unsigned)(c + dy) & 3u) * kU6GSlot + (kh * 5) * 128 + gm;
                    o0 += gr[1 * 128] + gr[3 * 128 - 1];                       // qo = 0: kw = 1 (dx 0), 3 (dx -1)
                    o1 += gr[0 * 128 + 1] + gr[2 * 128] + gr[4 * 128 - 1];     // qo = 1: kw = 0 (dx +1), 2 (dx 0), 4 (dx -1)
                }
                o0 += bias;
                o1 += bias;
                float2 r;
                if (act == ACT_ELU_CLAMP) { r.x = act_fast<ACT_ELU_CLAMP>(o0); r.y = act_fast<ACT_ELU_CLAMP>(o1); }
                else if (act == ACT_RELU) { r.x = act_fast<ACT_RELU>(o0); r.y = act_fast<ACT_RELU>(o1); }
                else if (act == ACT_ELU) { r.x = act_fast<ACT_ELU>(o0); r.y = act_fast<ACT_ELU>(o1); }
                else { r.x = apply_act(act, o0); r.y = apply_act(act, o1); }
                r.x = sc * r.x + of;
                r.y = sc * r.y + of;
                *reinterpret_cast<float2*>(p.out + ((size_t)t.n * p.T + 2 * yo + po) * p.F + 2 * X) = r;
            };
            if (p.pair) {
                // two input rows per iteration (the MMA warp issues them as a pair): both accumulators are loaded at once, the gathers of
                // output-row pairs y - 3 and y - 2 (G rows y-4 .. y-1) run under the loads, and ONE 256-thread barrier publishes G[y], G[y+1]
                for (int y = t.r0 - 1; y <= t.r1; y += 2, grow += 2) {   // the row count r1 - r0 + 2 is even
                    const int as1 = as + 1;                               // acc_slots is even: a pair never wraps
                    ptx::mbar_wait(&hdr->acc_full[as], aph);
                    ptx::mbar_wait(&hdr->acc_full[as1], aph);
                    ptx::tc_fence_after();
                    uint32_t v0[16], v1[16];
                    ptx::tmem_ld16_issue(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 32 + half * 16), v0);
                    ptx::tmem_ld16_issue(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(as1 * 32 + half * 16), v1);
                    if (col_ok) {
                        if (y - 3 >= t.r0) gather(y - 3, grow - 3);
                        if (y - 2 >= t.r0) gather(y - 2, grow - 2);
                    }
                    ptx::tmem_ld16_wait(v0);
                    ptx::tmem_ld16_wait(v1);
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        ptx::mbar_arrive(&hdr->acc_empty[as]);
                        ptx::mbar_arrive(&hdr->acc_empty[as1]);
                    }
                    as += 2;
                    if (as == p.acc_slots) { as = 0; aph ^= 1; }
                    float* gs0 = G + (size_t)((unsigned)grow % 6u) * kU6GSlot;
                    float* gs1 = G + (size_t)((unsigned)(grow + 1) % 6u) * kU6GSlot;
#pragma unroll
                    for (int i = 0; i < 16; i++)
                        if (half * 16 + i < 25) {
                            gs0[(half * 16 + i) * 128 + m] = __uint_as_float(v0[i]);
                            gs1[(half * 16 + i) * 128 + m] = __uint_as_float(v1[i]);
                        }
                    asm volatile("bar.sync 1, 256;\n" ::: "memory");
                }
                if (col_ok) {   // flush: the last output rows of the unit (G rows r1-3 .. r1)
                    if (t.r1 - 2 >= t.r0) gather(t.r1 - 2, grow - 3);
                    if (t.r1 - 1 >= t.r0) gather(t.r1 - 1, grow - 2);
                }
            } else {
                // one input row per iteration: issue the TMEM load of row y  |  gather + store the output rows of input row y - 2 from G rows
                // y-3, y-2, y-1 while that load is in flight  |  wait for the load, write G[y]  |  barrier
                for (int y = t.r0 - 1; y <= t.r1; y++, grow++) {
                    ptx::mbar_wait(&hdr->acc_full[as], aph);
                    ptx::tc_fence_after();
                    uint32_t v[16];
                    ptx::tmem_ld16_issue(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 32 + half * 16), v);
                    if (y - 2 >= t.r0 && col_ok) gather(y - 2, grow - 2);
                    ptx::tmem_ld16_wait(v);
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&hdr->acc_empty[as]);
                    if (++as == p.acc_slots) { as = 0; aph ^= 1; }
                    float* gs = G + (size_t)(grow & 3) * kU6GSlot;
#pragma unroll
                    for (int i = 0; i < 16; i++)
                        if (half * 16 + i < 25) gs[(half * 16 + i) * 128 + m] = __uint_as_float(v[i]);
                    asm volatile("bar.sync 1, 256;\n" ::: "memory");
                }
                if (col_ok) gather(t.r1 - 1, grow - 2);   // flush: the last output rows of the unit (G rows r1-2, r1-1, r1)
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 13) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<kU6AccSlots * 32>(tmem_d);
    }
}

void launch_up6_tc(const Up6TcParams& p, cudaStream_t st)
{
    const size_t smem = up6_tc_smem_bytes(p.S, 0);   // the larger of the two forms
    static LaunchState state;
    const int sms = state.prepare(up6_tc_kernel, smem);
    const int n_units = p.blocks_x * p.chunks * p.Bv * p.S;
    up6_tc_kernel<<<n_units < sms ? n_units : sms, kU6Threads, smem, st>>>(p);
}

bool up6_tc_fits(int S) { return up6_tc_smem_bytes(S, 0) <= 232448 && up6_tc_smem_bytes(S, 1) <= 232448; }

}  // namespace srt
