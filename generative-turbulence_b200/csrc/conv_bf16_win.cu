// 3x3x3 convolution over halo grids on a CTA pair with ROW-WINDOW operand reuse ("win" kernel).
//
// In the halo-grid layout every filter tap is a constant row shift, and the tcgen05 shared-memory descriptors
// swizzle by ABSOLUTE shared-memory address (measured on B200 with the weight-gradient kernel: a matrix may start
// at any row of a 128B-swizzled tile, base offset 0).  So for one kx the nine (ky, kz) taps are nine views of ONE
// window of 128 + 2*Zp + 2 consecutive rows: the activation tile is brought into shared memory 3 times per
// channel chunk instead of 9 (kz-folded kernels) or 27 (per-tap kernel), every output row of the 128-row tile is
// valid (the folded kernels keep 120 of 128), and the epilogue needs no cross-lane shifts.  The weights
// ([Cout][27*Cin], the per-tap kernel's layout) are split over the two CTAs of a cta_group::2 pair and stay
// resident (27*Cin*Cout bytes per CTA <= ~116 KB: 64->64, 128->32, 32->{32,64,128}).  Optional: GroupNorm moments,
// the block's 1x1 residual projection on the centre tap (spare TMEM columns), halo rows stored (input gradients).
// Roles per CTA: warp 0 TMA producer, warp 1 TMEM owner (+ MMA issue in the leader), warps 2-9 epilogue.
// Replaces nn.Conv3d(3, padding_mode="replicate") (+ res_conv) of reference ddpm.py:164,188.
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

using namespace tdb;
using bf16 = __nv_bfloat16;

namespace {

constexpr int BM = 128;
constexpr int THREADS = 320;
constexpr int MAX_STAGES = 12;

struct WinParams {
    int64_t rows;
    int Xp, Yp, Zp;
    FastDiv by_vox, by_z, by_y;
    int Cin, chunks;
    int stages;
    int win_rows;    // rows per activation window (multiple of 8, >= 128 + 2*Zp + 2)
    int tmem_half;   // TMEM columns of one accumulator stage
    int ld_out;
    int G;
    int num_super;   // pairs of 128-row tiles
    int n_tiles;     // N tiles of COUT output channels (streamed-weight variant: Cout_total / COUT)
    int num_items;   // num_super * n_tiles work items, N tile outermost
    int cout_total;
    int all_rows;
    int proj;
    int ld_outp;
    int add2;        // 1: a 1x1 convolution of a SECOND input (same channel count) accumulates into the same output tile
};

__device__ __forceinline__ bool interior_row(int64_t p, const WinParams& P, int& b) {
    if (p < 0 || p >= P.rows) return false;
    uint32_t bb, r, q, zp, xp, yp;
    P.by_vox.divmod((uint32_t)p, bb, r);
    P.by_z.divmod(r, q, zp);
    P.by_y.divmod(q, xp, yp);
    b = (int)bb;
    return xp >= 1u && xp <= (uint32_t)(P.Xp - 2) && yp >= 1u && yp <= (uint32_t)(P.Yp - 2) && zp >= 1u &&
           zp <= (uint32_t)(P.Zp - 2);
}

// RES = 1: the CTA's half of ALL weights is resident (one N tile).  RES = 0: N tiles of COUT channels; the nine
// (ky, kz) weight tiles of the current (kx, channel chunk) travel with the activation window in every stage.
template <int COUT, int KC, int RES>
__global__ void __launch_bounds__(THREADS, 1)
conv3d_bf16_win_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_g,
                       const float* __restrict__ bias, bf16* __restrict__ out,
                       double* __restrict__ gn_stats, const float* __restrict__ bias_p, bf16* __restrict__ out_p,
                       const WinParams P) {
    constexpr int NH = COUT / 2;             // weight rows each CTA contributes to the pair's MMA
    constexpr int NCH = COUT / 16;           // 16-column epilogue chunks
    constexpr int CH_PER_WARP = NCH / 2;
    constexpr uint32_t ROWB = KC * 2;        // bytes per operand row = swizzle span
    constexpr uint32_t bh_bytes = NH * ROWB;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 6];
    constexpr bool resident = RES != 0;
    __shared__ __align__(16) float s_biasp[resident ? COUT : 512];
    __shared__ __align__(16) float s_bias[resident ? COUT : 512];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t rank = ptx::cluster_ctarank();
    const int cluster_id = blockIdx.x / 2, n_clusters = gridDim.x / 2;
    const uint32_t full_bar = ptx::smem_u32(&bars[0]);                    // used in the leader
    const uint32_t empty_bar = ptx::smem_u32(&bars[MAX_STAGES]);          // per CTA (multicast commit)
    const uint32_t acc_full = ptx::smem_u32(&bars[2 * MAX_STAGES]);       // [2] per CTA (multicast commit)
    const uint32_t acc_empty = ptx::smem_u32(&bars[2 * MAX_STAGES + 2]);  // [2] used in the leader, 16 arrivals
    const uint32_t b_full = ptx::smem_u32(&bars[2 * MAX_STAGES + 4]);     // used in the leader
    const uint32_t p_full = ptx::smem_u32(&bars[2 * MAX_STAGES + 5]);     // used in the leader (projection weights)
    const int chunks = P.chunks;
    const int n_b = resident ? 27 * chunks : 0;         // resident weight tiles (tap, channel chunk) of this CTA
    const uint32_t b_region = (uint32_t)n_b * bh_bytes;
    const uint32_t p_base_addr = smem_base + b_region;
    const uint32_t p_region = (resident && (P.proj || P.add2)) ? (uint32_t)chunks * bh_bytes : 0u;
    const uint32_t stage_base = (smem_base + b_region + p_region + 1023u) & ~1023u;
    const uint32_t win_bytes = (uint32_t)P.win_rows * ROWB;
    // streamed weights: nine (ky, kz) tiles per stage, plus a slot for the 1x1 projection tile (filled when kx == 1)
    const uint32_t stage_bytes = win_bytes + (resident ? 0u : (9u + (P.proj ? 1u : 0u)) * bh_bytes);

    for (int i = threadIdx.x; i < P.cout_total; i += THREADS) s_bias[i] = bias ? bias[i] : 0.0f;
    for (int i = threadIdx.x; i < P.cout_total; i += THREADS) s_biasp[i] = (P.proj && bias_p) ? bias_p[i] : 0.0f;
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_b);
        if (P.add2) ptx::prefetch_tensormap(&map_g);
        for (int s = 0; s < P.stages; ++s) {
            ptx::mbar_init(full_bar + 8 * s, 2);   // leader's arm (expect_tx for both CTAs' bytes) + peer's arrival
            ptx::mbar_init(empty_bar + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(acc_full + 8 * s, 1);
            ptx::mbar_init(acc_empty + 8 * s, 16);  // 8 epilogue warps in each CTA
        }
        ptx::mbar_init(b_full, 2);
        ptx::mbar_init(p_full, 2);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc_2sm(ptx::smem_u32(&tmem_base_slot), (uint32_t)(2 * P.tmem_half));
        ptx::tmem_relinquish_2sm();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();  // both CTAs: barriers initialised, TMEM allocated
    ptx::tc_fence_after();
    const uint32_t tmem_d = tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): resident half-weights once, then one row window per (kx, channel chunk);
        // every load signals the LEADER's barrier =====
        if (resident && ptx::elect_one()) {
            const uint32_t b_full_l = ptx::leader_addr(b_full);
            for (int i = 0; i < n_b; ++i)
                ptx::tma_load_3d_2sm(smem_base + (uint32_t)i * bh_bytes, &map_b, b_full_l, (i % chunks) * KC, (int)rank * NH, i / chunks);
            if (rank == 0) ptx::mbar_arrive_expect_tx(b_full, 2u * b_region);
            else ptx::mbar_arrive_remote(b_full, 0);
        }
        if (resident && (P.proj || P.add2) && ptx::elect_one()) {
            {
                const uint32_t p_full_l = ptx::leader_addr(p_full);
                for (int ch = 0; ch < chunks; ++ch)
                    ptx::tma_load_3d_2sm(p_base_addr + (uint32_t)ch * bh_bytes, &map_p, p_full_l, ch * KC, (int)rank * NH, 0);
                if (rank == 0) ptx::mbar_arrive_expect_tx(p_full, 2u * p_region);
                else ptx::mbar_arrive_remote(p_full, 0);
            }
        }
        __syncwarp();
        const int yz = P.Yp * P.Zp;
        uint32_t s = 0, ph = 1;
        for (int w = cluster_id; w < P.num_items; w += n_clusters) {
            const int nt = w / P.num_super, st = w - nt * P.num_super;
            const int tile = 2 * st + (int)rank;
            const int q0 = tile * BM - P.Zp - 1;  // first row of the kx = 1 window (rows outside the grid are zero-filled)
            for (int kx = 0; kx < 3; ++kx) {
                const int row = q0 + (kx - 1) * yz;
                for (int ch = 0; ch < chunks; ++ch) {
                    ptx::mbar_wait(empty_bar + 8 * s, ph);
                    if (ptx::elect_one()) {
                        const uint32_t full_l = ptx::leader_addr(full_bar + 8 * s);
                        ptx::tma_load_2d_2sm(stage_base + s * stage_bytes, &map_a, full_l, ch * KC, row);
                        if (!resident) {
                            // the nine (ky, kz) weight tiles of this (kx, channel chunk): ONE box [9 taps][NH rows][KC] (nine
                            // 4-8 KB boxes cost nine TMA instructions and nine L2 request trains per stage)
                            ptx::tma_load_3d_2sm(stage_base + s * stage_bytes + win_bytes, &map_b, full_l, ch * KC,
                                                 nt * COUT + (int)rank * NH, kx * 9);
                            if (P.proj && kx == 1)
                                ptx::tma_load_3d_2sm(stage_base + s * stage_bytes + win_bytes + 9u * bh_bytes, &map_p, full_l, ch * KC,
                                                     nt * COUT + (int)rank * NH, 0);
                        }
                        const uint32_t tx = resident ? win_bytes : win_bytes + (9u + ((P.proj && kx == 1) ? 1u : 0u)) * bh_bytes;
                        if (rank == 0) ptx::mbar_arrive_expect_tx(full_bar + 8 * s, 2u * tx);
                        else ptx::mbar_arrive_remote(full_bar + 8 * s, 0);
                    }
                    __syncwarp();
                    if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
                }
            }
            if (P.add2) {
                // second input: the tile's own 128 rows (no shift), one stage per channel chunk (+ its 1x1 weight tile)
                for (int ch = 0; ch < chunks; ++ch) {
                    ptx::mbar_wait(empty_bar + 8 * s, ph);
                    if (ptx::elect_one()) {
                        const uint32_t full_l = ptx::leader_addr(full_bar + 8 * s);
                        ptx::tma_load_2d_2sm(stage_base + s * stage_bytes, &map_g, full_l, ch * KC, tile * BM);
                        if (!resident)
                            ptx::tma_load_3d_2sm(stage_base + s * stage_bytes + win_bytes, &map_p, full_l, ch * KC,
                                                 nt * COUT + (int)rank * NH, 0);
                        const uint32_t tx = (uint32_t)BM * ROWB + (resident ? 0u : bh_bytes);
                        if (rank == 0) ptx::mbar_arrive_expect_tx(full_bar + 8 * s, 2u * tx);
                        else ptx::mbar_arrive_remote(full_bar + 8 * s, 0);
                    }
                    __syncwarp();
                    if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: leader CTA only; every MMA is 256 rows (128 per CTA) x COUT x 16 =====
        if (rank == 0) {
            const uint32_t idesc = ptx::umma_idesc_bf16(2 * BM, (uint32_t)COUT);
            const uint64_t desc0 = ptx::umma_smem_desc(0, ROWB);
            const uint64_t a_base = desc0 | (uint64_t)((stage_base & 0x3FFFFu) >> 4);
            const uint64_t b_base = desc0 | (uint64_t)((smem_base & 0x3FFFFu) >> 4);
            const uint64_t p_base = desc0 | (uint64_t)((p_base_addr & 0x3FFFFu) >> 4);
            const uint32_t st_step = stage_bytes >> 4;
            constexpr uint32_t b_step = bh_bytes >> 4, row16 = ROWB >> 4;
            const uint32_t zrow16 = (uint32_t)P.Zp * row16;  // one y step = Zp rows
            if (resident) ptx::mbar_wait(b_full, 0);
            if (resident && (P.proj || P.add2)) ptx::mbar_wait(p_full, 0);
            ptx::tc_fence_after();
            uint32_t s = 0, ph = 0;
            int local = 0;
            for (int w = cluster_id; w < P.num_items; w += n_clusters, ++local) {
                const int as = local & 1;
                const uint32_t aph = (uint32_t)(local >> 1) & 1u;
                ptx::mbar_wait(acc_empty + 8 * as, aph ^ 1u);  // both CTAs' epilogues have drained this stage
                ptx::tc_fence_after();
                const uint32_t d_addr = tmem_d + (uint32_t)(as * P.tmem_half);
                for (int kx = 0; kx < 3; ++kx) {
                    for (int ch = 0; ch < chunks; ++ch) {
                        ptx::mbar_wait(full_bar + 8 * s, ph);
                        ptx::tc_fence_after();
                        if (ptx::elect_one()) {
                            const uint64_t a_st = a_base + (uint64_t)(s * st_step);
                            const uint64_t b_st = resident ? b_base + (uint64_t)((uint32_t)(kx * 9 * chunks + ch) * b_step)
                                                           : a_st + (uint64_t)(win_bytes >> 4);
                            const uint32_t b_tap = resident ? (uint32_t)chunks * b_step : b_step;  // tile stride between taps
                            const uint32_t first = (kx | ch) == 0 ? 0u : 1u;
#pragma unroll
                            for (int t = 0; t < 9; ++t) {
                                // tap (ky, kz) = (t / 3, t % 3): the window viewed from row ky*Zp + kz
                                const uint64_t a_t = a_st + (uint64_t)((uint32_t)(t / 3) * zrow16 + (uint32_t)(t % 3) * row16);
                                const uint64_t b_t = b_st + (uint64_t)((uint32_t)t * b_tap);
#pragma unroll
                                for (int k = 0; k < KC / 16; ++k)
                                    ptx::umma_f16_2sm(d_addr, a_t + (uint64_t)(2 * k), b_t + (uint64_t)(2 * k), idesc,
                                                      (t | k) != 0 ? 1u : first);
                            }
                            if (P.proj && kx == 1) {
                                // centre tap: the same rows also feed the 1x1 projection (TMEM columns behind the conv's)
                                const uint64_t a_t = a_st + (uint64_t)(zrow16 + row16);
                                const uint64_t p_t = resident ? p_base + (uint64_t)((uint32_t)ch * b_step) : b_st + (uint64_t)(9u * b_step);
#pragma unroll
                                for (int k = 0; k < KC / 16; ++k)
                                    ptx::umma_f16_2sm(d_addr + (uint32_t)COUT, a_t + (uint64_t)(2 * k), p_t + (uint64_t)(2 * k), idesc,
                                                      (uint32_t)((ch | k) != 0));
                            }
                            ptx::umma_commit_2sm_mc(empty_bar + 8 * s, (uint16_t)0x3);  // frees the slot in both CTAs
                        }
                        __syncwarp();
                        if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
                    }
                }
                if (P.add2) {
                    for (int ch = 0; ch < chunks; ++ch) {
                        ptx::mbar_wait(full_bar + 8 * s, ph);
                        ptx::tc_fence_after();
                        if (ptx::elect_one()) {
                            const uint64_t a_st = a_base + (uint64_t)(s * st_step);
                            const uint64_t p_t = resident ? p_base + (uint64_t)((uint32_t)ch * b_step) : a_st + (uint64_t)(win_bytes >> 4);
#pragma unroll
                            for (int k = 0; k < KC / 16; ++k)
                                ptx::umma_f16_2sm(d_addr, a_st + (uint64_t)(2 * k), p_t + (uint64_t)(2 * k), idesc, 1u);
                            ptx::umma_commit_2sm_mc(empty_bar + 8 * s, (uint16_t)0x3);
                        }
                        __syncwarp();
                        if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
                    }
                }
                if (ptx::elect_one()) ptx::umma_commit_2sm_mc(acc_full + 8 * as, (uint16_t)0x3);
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue: 8 warps, lane group lg = warp % 4, column half = (warp - 2) / 4 =====
        const int lg = warp % 4;
        const int half = (warp - 2) / 4;
        const bool do_stats = gn_stats != nullptr;
        float st_s[CH_PER_WARP][8], st_q[CH_PER_WARP][8];  // GroupNorm partials per column pair
#pragma unroll
        for (int a = 0; a < CH_PER_WARP; ++a)
#pragma unroll
            for (int j = 0; j < 8; ++j) st_s[a][j] = st_q[a][j] = 0.0f;
        int st_b = -1, st_nt = 0;

        auto flush_stats = [&]() {
            const int cpg = P.cout_total / P.G;  // even (checked on the host)
#pragma unroll
            for (int a = 0; a < CH_PER_WARP; ++a) {
                const int cidx = 2 * a + half;
                double gs = 0.0, gq = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    gs += (double)st_s[a][j];
                    gq += (double)st_q[a][j];
                    st_s[a][j] = st_q[a][j] = 0.0f;
                    const int col_end = st_nt * COUT + cidx * 16 + 2 * j + 2;
                    if (col_end % cpg == 0 || j == 7) {
                        const double ws = warp_sum(gs), wq = warp_sum(gq);
                        if (lane == 0) {
                            const int g = (col_end - 1) / cpg;
                            atomicAdd(gn_stats + ((int64_t)st_b * P.G + g) * 2, ws);
                            atomicAdd(gn_stats + ((int64_t)st_b * P.G + g) * 2 + 1, wq);
                        }
                        gs = gq = 0.0;
                    }
                }
            }
        };

        int local = 0;
        for (int w = cluster_id; w < P.num_items; w += n_clusters, ++local) {
            const int nt = w / P.num_super, st = w - nt * P.num_super;
            const int tile = 2 * st + (int)rank;
            const int n0 = nt * COUT;
            const int as = local & 1;
            const uint32_t aph = (uint32_t)(local >> 1) & 1u;
            const int64_t p = (int64_t)tile * BM + 32 * lg + lane;
            int b = 0;
            const bool inter = interior_row(p, P, b);
            const bool valid = P.all_rows ? (p < P.rows) : inter;
            if (do_stats) {
                // valid rows of one warp share one sample (a sample boundary is two halo planes wide)
                const unsigned vmask = __ballot_sync(0xffffffffu, valid);
                if (vmask) {
                    const int b_warp = __shfl_sync(0xffffffffu, b, __ffs(vmask) - 1);
                    if (b_warp != st_b || nt != st_nt) {
                        if (st_b >= 0) flush_stats();
                        st_b = b_warp;
                        st_nt = nt;
                    }
                }
            }
            ptx::mbar_wait(acc_full + 8 * as, aph);
            ptx::tc_fence_after();
            const uint32_t t_row = tmem_d + (uint32_t)(as * P.tmem_half) + ((uint32_t)(lg * 32) << 16);
            bf16* orow = out + p * P.ld_out + n0;
#pragma unroll
            for (int a = 0; a < CH_PER_WARP; ++a) {
                const int c = (2 * a + half) * 16;
                uint32_t r0[16], r3[16];
                ptx::tmem_ld_x16(t_row + (uint32_t)c, r0);
                if (P.proj) ptx::tmem_ld_x16(t_row + (uint32_t)(COUT + c), r3);
                ptx::tmem_ld_wait();
                if (P.proj && valid) {
                    uint4 lo, hi;
                    __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&lo);
                    __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&hi);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        h0[j] = __floats2bfloat162_rn(__uint_as_float(r3[2 * j]) + s_biasp[n0 + c + 2 * j],
                                                      __uint_as_float(r3[2 * j + 1]) + s_biasp[n0 + c + 2 * j + 1]);
                        h1[j] = __floats2bfloat162_rn(__uint_as_float(r3[8 + 2 * j]) + s_biasp[n0 + c + 8 + 2 * j],
                                                      __uint_as_float(r3[8 + 2 * j + 1]) + s_biasp[n0 + c + 8 + 2 * j + 1]);
                    }
                    bf16* prow = out_p + p * P.ld_outp + n0;
                    ptx::st_global_32B(prow + c, lo, hi);
                }
                if (valid) {
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r0[j]) + s_bias[n0 + c + j];
                    uint4 lo, hi;
                    __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&lo);
                    __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&hi);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        h0[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                        h1[j] = __floats2bfloat162_rn(v[8 + 2 * j], v[8 + 2 * j + 1]);
                    }
                    ptx::st_global_32B(orow + c, lo, hi);
                    if (do_stats) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            st_s[a][j] += v[2 * j] + v[2 * j + 1];
                            st_q[a][j] = fmaf(v[2 * j], v[2 * j], fmaf(v[2 * j + 1], v[2 * j + 1], st_q[a][j]));
                        }
                    }
                }
            }
            // this warp has finished reading the accumulator stage
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) ptx::mbar_arrive(acc_empty + 8 * as);
                else ptx::mbar_arrive_remote(acc_empty + 8 * as, 0);  // the leader's MMA warp owns the accumulator ring
            }
        }
        if (do_stats && st_b >= 0) flush_stats();
    }

    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();  // the peer may still signal / be signalled until both are here
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_2sm(tmem_d, (uint32_t)(2 * P.tmem_half));
    }
}

int g_num_sms_win = 0;

template <int COUT, int KC, int RES>
int launch_win(const CUtensorMap& map_a, const CUtensorMap& map_b, const CUtensorMap& map_p, const CUtensorMap& map_g, const float* bias, bf16* out,
               double* gn_stats, const float* bias_p, bf16* out_p, const WinParams& P, size_t smem, cudaStream_t stream) {
    auto kern = conv3d_bf16_win_kernel<COUT, KC, RES>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_conv3d_bf16_win: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    int grid = 2 * P.num_items;
    const int cap = g_num_sms_win & ~1;
    if (grid > cap) grid = cap;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, map_a, map_b, map_p, map_g, bias, out, gn_stats, bias_p, out_p, P);
    TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_conv3d_bf16_win: launch: %s", cudaGetErrorString(e));
    TDB_CHECK_LAUNCH("tdb_conv3d_bf16_win");
    return 0;
}

}  // namespace

namespace {
// in2 / ld_in2 / w2: optional second input (Cin channels) whose 1x1 convolution with w2 [Cout][Cin] is added to the output
int win_impl(const void* in, int ld_in, const void* w, const float* bias, void* out, int ld_out, int B, int X, int Y, int Z, int Cin,
             int Cout, double* gn_stats, int G, unsigned flags, const void* w_proj, const float* bias_proj, void* out_proj,
             int ld_outp, const void* in2, int ld_in2, const void* w2, void* stream) {
    TDB_REQUIRE(in && w && out, TDB_E_BADARG, "tdb_conv3d_bf16_win: null pointer");
    TDB_REQUIRE(Cin % 32 == 0 && (Cout == 32 || Cout == 64 || (Cout % 128 == 0 && Cout <= 512)) && ld_in % 8 == 0 && ld_out % 8 == 0,
                TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_win: need Cin %% 32 == 0 and Cout in {32,64,128k<=512} (Cin=%d Cout=%d)", Cin, Cout);
    TDB_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)w & 15) == 0, TDB_E_UNSUPPORTED,
                "tdb_conv3d_bf16_win: pointers must be 16-byte aligned");
    TDB_REQUIRE(!gn_stats || (G >= 1 && Cout % G == 0 && (Cout / G) % 2 == 0), TDB_E_UNSUPPORTED,
                "tdb_conv3d_bf16_win: fused GroupNorm moments need an even number of channels per group");
    Grid3 g(B, X, Y, Z);
    TDB_REQUIRE(g.rows < (1ll << 31) - (1 << 20), TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_win: too many rows");
    if (g_num_sms_win == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms_win, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms_win <= 0) g_num_sms_win = 148;
    }
    const int KC = Cin % 64 == 0 ? 64 : 32;
    WinParams P;
    P.rows = g.rows;
    P.Xp = g.Xp; P.Yp = g.Yp; P.Zp = g.Zp;
    P.by_vox = FastDiv((uint32_t)g.vox_p);
    P.by_z = FastDiv((uint32_t)g.Zp);
    P.by_y = FastDiv((uint32_t)g.Yp);
    P.Cin = Cin;
    P.chunks = Cin / KC;
    P.win_rows = (BM + 2 * g.Zp + 2 + 7) & ~7;
    TDB_REQUIRE(P.win_rows <= 256, TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_win: Z + 2 = %d is too wide for one TMA box", g.Zp);
    P.proj = w_proj != nullptr ? 1 : 0;
    P.ld_outp = ld_outp;
    P.add2 = in2 != nullptr ? 1 : 0;
    TDB_REQUIRE(!P.add2 || (!P.proj && w2 && ld_in2 % 8 == 0 && ((uintptr_t)in2 & 15) == 0 && ((uintptr_t)w2 & 15) == 0), TDB_E_UNSUPPORTED,
                "tdb_conv3d_bf16_win: the added 1x1 convolution needs aligned buffers and excludes the fused projection");
    TDB_REQUIRE(!P.proj || (out_proj && ld_outp % 8 == 0 && ((uintptr_t)w_proj & 15) == 0 && ((uintptr_t)out_proj & 15) == 0 &&
                            !(flags & TDB_CONV_ALL_ROWS)),
                TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_win: the fused projection needs aligned buffers and no ALL_ROWS");
    int tile_n = Cout > 128 ? 128 : Cout;  // output channels per N tile
    // tiny grids (the bottleneck level): with 128-channel N tiles the whole layer is a few dozen work items, one per CTA
    // pair, each a serial walk over all of K (49 us at 512 -> 512 whatever the batch).  64-channel tiles double the
    // items while they still fit one wave, and their 54 KB stages pipeline four deep instead of two.
    if (Cout % 128 == 0 && Cout >= 256 && KC == 64 && ceil_div(g.rows, 2 * BM) * (Cout / 128) * 2 <= (g_num_sms_win / 2))
        tile_n = 64;
    P.n_tiles = Cout / tile_n;
    P.cout_total = Cout;
    const int bh_bytes = (tile_n / 2) * KC * 2;
    const int win_bytes = P.win_rows * KC * 2;
    const int budget = 221 * 1024;
    const int resident_bytes = (27 + ((P.proj || P.add2) ? 1 : 0)) * P.chunks * bh_bytes;
    // all weights resident when they leave room for two windows; otherwise (Cin % 64 == 0 only) the nine weight
    // tiles of a (kx, channel chunk) stream with each window and Cout is walked in N tiles of 128
    const bool res = P.n_tiles == 1 && resident_bytes + 2048 + 2 * win_bytes <= budget;
    TDB_REQUIRE(res || (KC == 64 && (tile_n == 128 || tile_n == 64)), TDB_E_UNSUPPORTED,
                "tdb_conv3d_bf16_win: streamed weights need Cin %% 64 == 0 and Cout %% 128 == 0 (Cin=%d Cout=%d)", Cin, Cout);
    const int fixed_bytes = res ? resident_bytes : 0;
    const int stage_bytes = win_bytes + (res ? 0 : (9 + (P.proj ? 1 : 0)) * bh_bytes);
    int stages = (budget - fixed_bytes - 2048) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    TDB_REQUIRE(stages >= 2, TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_win: two stages of %d bytes do not fit", stage_bytes);
    P.stages = stages;
    int half = 32;
    while (half < tile_n * (P.proj ? 2 : 1)) half *= 2;
    P.tmem_half = half;
    P.ld_out = ld_out;
    P.G = gn_stats ? G : 0;
    P.num_super = (int)ceil_div(g.rows, 2 * BM);
    P.num_items = P.num_super * P.n_tiles;
    P.all_rows = (flags & TDB_CONV_ALL_ROWS) ? 1 : 0;
    TDB_REQUIRE(!(P.all_rows && gn_stats), TDB_E_BADARG, "tdb_conv3d_bf16_win: fused moments are not available with ALL_ROWS");

    CUtensorMap map_a, map_b, map_p, map_g;
    TDB_REQUIRE(encode_fn() != nullptr, TDB_E_NODEVICE, "tdb_conv3d_bf16_win: cuTensorMapEncodeTiled unavailable (no driver)");
    TDB_REQUIRE(make_map_2d_bf16(&map_a, in, (uint64_t)Cin, (uint64_t)g.rows, (uint64_t)ld_in, (uint32_t)KC, (uint32_t)P.win_rows),
                TDB_E_BADARG, "tdb_conv3d_bf16_win: tensor map (activations) rejected");
    {
        // weights [Cout][27][Cin]; one box = the Cout/2 rows of a CTA for one (tap, channel chunk)
        const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, 27};
        const uint64_t strides[2] = {27ull * Cin, (uint64_t)Cin};
        const uint32_t box[3] = {(uint32_t)KC, (uint32_t)(tile_n / 2), res ? 1u : 9u};  // streamed: the nine taps of a kx at once
        TDB_REQUIRE(make_map_bf16(&map_b, w, 3, dims, strides, box), TDB_E_BADARG, "tdb_conv3d_bf16_win: tensor map (weights) rejected");
    }
    {
        const void* wp = P.proj ? w_proj : (P.add2 ? w2 : w);
        const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, 1};
        const uint64_t strides[2] = {(uint64_t)((P.proj || P.add2) ? Cin : 27 * Cin), (uint64_t)Cin * Cout};
        const uint32_t box[3] = {(uint32_t)KC, (uint32_t)(tile_n / 2), 1};
        TDB_REQUIRE(make_map_bf16(&map_p, wp, 3, dims, strides, box), TDB_E_BADARG, "tdb_conv3d_bf16_win: tensor map (projection) rejected");
    }
    // second input: the 128 rows of a tile
    TDB_REQUIRE(make_map_2d_bf16(&map_g, P.add2 ? in2 : in, (uint64_t)Cin, (uint64_t)g.rows, (uint64_t)(P.add2 ? ld_in2 : ld_in), (uint32_t)KC,
                                 (uint32_t)BM),
                TDB_E_BADARG, "tdb_conv3d_bf16_win: tensor map (second input) rejected");
    const size_t smem = (size_t)fixed_bytes + 1024 + (size_t)stages * stage_bytes + 1024;
    cudaStream_t s = (cudaStream_t)stream;
    bf16* o = (bf16*)out;
    bf16* op = (bf16*)out_proj;
#define TDB_WIN_CASE(CO, K, R) \
    if (tile_n == CO && KC == K && (int)res == R) return launch_win<CO, K, R>(map_a, map_b, map_p, map_g, bias, o, gn_stats, bias_proj, op, P, smem, s)
    TDB_WIN_CASE(32, 64, 1);
    TDB_WIN_CASE(64, 64, 1);
    TDB_WIN_CASE(128, 64, 1);
    TDB_WIN_CASE(32, 32, 1);
    TDB_WIN_CASE(64, 32, 1);
    TDB_WIN_CASE(128, 32, 1);
    TDB_WIN_CASE(128, 64, 0);
    TDB_WIN_CASE(64, 64, 0);
#undef TDB_WIN_CASE
    tdb::set_error("tdb_conv3d_bf16_win: no kernel for Cout=%d KC=%d", Cout, KC);
    return TDB_E_UNSUPPORTED;
}
}  // namespace

extern "C" int tdb_conv3d_bf16_win(const void* in, int ld_in, const void* w, const float* bias, void* out, int ld_out, int B, int X,
                                   int Y, int Z, int Cin, int Cout, double* gn_stats, int G, unsigned flags, const void* w_proj,
                                   const float* bias_proj, void* out_proj, int ld_outp, void* stream) {
    return win_impl(in, ld_in, w, bias, out, ld_out, B, X, Y, Z, Cin, Cout, gn_stats, G, flags, w_proj, bias_proj, out_proj, ld_outp,
                    nullptr, 0, nullptr, stream);
}

extern "C" int tdb_conv3d_bf16_win_add1x1(const void* in, int ld_in, const void* w, const float* bias, void* out, int ld_out, int B,
                                          int X, int Y, int Z, int Cin, int Cout, unsigned flags, const void* in2, int ld_in2,
                                          const void* w2, void* stream) {
    TDB_REQUIRE(in2 && w2, TDB_E_BADARG, "tdb_conv3d_bf16_win_add1x1: null pointer");
    return win_impl(in, ld_in, w, bias, out, ld_out, B, X, Y, Z, Cin, Cout, nullptr, 0, flags, nullptr, nullptr, nullptr, 0, in2, ld_in2,
                    w2, stream);
}
