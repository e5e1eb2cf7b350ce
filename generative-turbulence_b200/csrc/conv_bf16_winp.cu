// 32 -> 32 channel 3x3x3 convolution on a CTA pair: the kz-folded row-window kernel (conv_bf16_winz.cu) over PAIRED ROWS.
//
// A 32-channel bf16 halo grid has 64-byte rows, and the TMA / L2 fabric moves 128-byte lines: fetching 64-byte rows is
// bound by the row-request rate (measured: 17.6 B/clk/SM against 41 B/clk/SM for 128-byte rows; the kernel ran at the
// same 0.14 ms floor with its MMAs, TMEM loads and stores knocked out).  With the exact channel pitch (ld = 32) two
// consecutive grid rows are one contiguous 128-byte "super-row", so this kernel fetches super-rows (half the requests)
// into 128B-swizzled windows and lets K select the parity: for the K-major A operand the first two K = 16 steps of a
// super-row are the 32 channels of the EVEN grid row, the last two those of the ODD row - each parity accumulates into
// its own TMEM columns against the same weight tile.  Row shifts become super-row shifts (Zp and Yp*Zp must be even),
// and the +-1 grid-row shift of the folded kz taps mixes the parities in the epilogue:
//     out[2m]   = D0_odd[m-1] + D1_even[m] + D2_odd[m]   + bias
//     out[2m+1] = D0_even[m]  + D1_odd[m]  + D2_even[m+1] + bias
// Tiles: 128 super-rows advancing by 126 (252 output rows).  Weights: folded layout [96][9*32], resident (27 KB per CTA).
// Replaces nn.Conv3d(32, 32, 3, padding_mode="replicate") of the full-resolution blocks (reference ddpm.py:164).
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

using namespace tdb;
using bf16 = __nv_bfloat16;

namespace {

constexpr int BM = 128;        // super-rows per tile
constexpr int SR_OUT = 126;    // valid super-rows per tile (252 output rows)
constexpr int THREADS = 320;
constexpr int MAX_STAGES = 12;
constexpr int COUT = 32, CIN = 32;
constexpr int NF = 3 * COUT, NH = NF / 2;      // folded N and the weight rows staged by each CTA
constexpr uint32_t ROWA = 128;                 // bytes per super-row = swizzle span of A
constexpr uint32_t ROWB = CIN * 2;             // bytes per weight row = swizzle span of B (64)
constexpr uint32_t bh_bytes = NH * ROWB;       // one (kx, ky) weight tile of a CTA

struct WinpParams {
    int64_t rows;       // grid rows (even)
    int Xp, Yp, Zp;
    FastDiv by_vox, by_z, by_y;
    int stages;
    int win_rows;       // super-rows per activation window (multiple of 8, >= 128 + Zp)
    int tmem_half;      // TMEM columns of one accumulator stage (even + odd parity)
    int ld_out;
    int G;
    int num_super;      // pairs of tiles
    int all_rows;
    int tma_out;        // 1: the tile is staged in shared memory and written by one TMA store (needs ld_out == 32)
};

__device__ __forceinline__ bool interior_row(int64_t p, const WinpParams& P, int& b) {
    if (p < 0 || p >= P.rows) return false;
    uint32_t bb, r, q, zp, xp, yp;
    P.by_vox.divmod((uint32_t)p, bb, r);
    P.by_z.divmod(r, q, zp);
    P.by_y.divmod(q, xp, yp);
    b = (int)bb;
    return xp >= 1u && xp <= (uint32_t)(P.Xp - 2) && yp >= 1u && yp <= (uint32_t)(P.Yp - 2) && zp >= 1u &&
           zp <= (uint32_t)(P.Zp - 2);
}

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 256;" ::: "memory"); }  // the 8 epilogue warps

__global__ void __launch_bounds__(THREADS, 1)
conv3d_bf16_winp_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const __grid_constant__ CUtensorMap map_o, const float* __restrict__ bias, bf16* __restrict__ out, double* __restrict__ gn_stats, const WinpParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 6];
    __shared__ __align__(16) float s_bias[COUT];
    __shared__ __align__(16) float s_xch[2][2][4][COUT];  // [parity of the exchange][0: odd kz=0 of lane 31, 1: even kz=2 of lane 0][lane group]
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t rank = ptx::cluster_ctarank();
    const int cluster_id = blockIdx.x / 2, n_clusters = gridDim.x / 2;
    const uint32_t full_bar = ptx::smem_u32(&bars[0]);                    // used in the leader
    const uint32_t empty_bar = ptx::smem_u32(&bars[MAX_STAGES]);          // per CTA (multicast commit)
    const uint32_t acc_full = ptx::smem_u32(&bars[2 * MAX_STAGES]);       // [2] per CTA (multicast commit)
    const uint32_t acc_empty = ptx::smem_u32(&bars[2 * MAX_STAGES + 2]);  // [2] used in the leader, 16 arrivals
    const uint32_t b_full = ptx::smem_u32(&bars[2 * MAX_STAGES + 4]);     // used in the leader
    constexpr uint32_t b_region = 9u * bh_bytes;
    const uint32_t stage_base = (smem_base + b_region + 1023u) & ~1023u;
    const uint32_t stage_bytes = (uint32_t)P.win_rows * ROWA;
    const uint32_t out_stage = (stage_base + (uint32_t)P.stages * stage_bytes + 1023u) & ~1023u;  // [2][128 super-rows][128 B], swizzled
    const int zh = P.Zp / 2, yzh = (P.Yp * P.Zp) / 2;  // row shifts in super-rows

    for (int i = threadIdx.x; i < COUT; i += THREADS) s_bias[i] = bias ? bias[i] : 0.0f;
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_b);
        if (P.tma_out) ptx::prefetch_tensormap(&map_o);
        for (int s = 0; s < P.stages; ++s) {
            ptx::mbar_init(full_bar + 8 * s, 2);   // leader's arm (expect_tx for both CTAs' bytes) + peer's arrival
            ptx::mbar_init(empty_bar + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(acc_full + 8 * s, 1);
            ptx::mbar_init(acc_empty + 8 * s, 16);  // 8 epilogue warps in each CTA
        }
        ptx::mbar_init(b_full, 2);
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
        // ===== TMA producer (both CTAs): resident half-weights once, then one super-row window per kx =====
        if (ptx::elect_one()) {
            const uint32_t b_full_l = ptx::leader_addr(b_full);
            for (int i = 0; i < 9; ++i)
                ptx::tma_load_3d_2sm(smem_base + (uint32_t)i * bh_bytes, &map_b, b_full_l, 0, (int)rank * NH, i);
            if (rank == 0) ptx::mbar_arrive_expect_tx(b_full, 2u * b_region);
            else ptx::mbar_arrive_remote(b_full, 0);
        }
        __syncwarp();
        uint32_t s = 0, ph = 1;
        for (int w = cluster_id; w < P.num_super; w += n_clusters) {
            const int tile = 2 * w + (int)rank;
            const int q0 = tile * SR_OUT - 1 - zh;  // first super-row of the kx = 1 window (outside the grid: zero-filled)
            for (int kx = 0; kx < 3; ++kx) {
                ptx::mbar_wait(empty_bar + 8 * s, ph);
                if (ptx::elect_one()) {
                    const uint32_t full_l = ptx::leader_addr(full_bar + 8 * s);
                    ptx::tma_load_2d_2sm(stage_base + s * stage_bytes, &map_a, full_l, 0, q0 + (kx - 1) * yzh);
                    if (rank == 0) ptx::mbar_arrive_expect_tx(full_bar + 8 * s, 2u * stage_bytes);
                    else ptx::mbar_arrive_remote(full_bar + 8 * s, 0);
                }
                __syncwarp();
                if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: leader CTA only; every MMA is 256 super-rows (128 per CTA) x 96 x 16 =====
        if (rank == 0) {
            const uint32_t idesc = ptx::umma_idesc_bf16(2 * BM, (uint32_t)NF);
            const uint64_t a_base = ptx::umma_smem_desc(0, ROWA) | (uint64_t)((stage_base & 0x3FFFFu) >> 4);
            const uint64_t b_base = ptx::umma_smem_desc(0, ROWB) | (uint64_t)((smem_base & 0x3FFFFu) >> 4);
            const uint32_t st_step = stage_bytes >> 4;
            constexpr uint32_t b_step = bh_bytes >> 4, row16 = ROWA >> 4;
            const uint32_t zrow16 = (uint32_t)zh * row16;  // one y step = Zp/2 super-rows
            ptx::mbar_wait(b_full, 0);
            ptx::tc_fence_after();
            uint32_t s = 0, ph = 0;
            int local = 0;
            for (int w = cluster_id; w < P.num_super; w += n_clusters, ++local) {
                const int as = local & 1;
                const uint32_t aph = (uint32_t)(local >> 1) & 1u;
                ptx::mbar_wait(acc_empty + 8 * as, aph ^ 1u);  // both CTAs' epilogues have drained this stage
                ptx::tc_fence_after();
                const uint32_t d_addr = tmem_d + (uint32_t)(as * P.tmem_half);
                for (int kx = 0; kx < 3; ++kx) {
                    ptx::mbar_wait(full_bar + 8 * s, ph);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
                        const uint64_t a_st = a_base + (uint64_t)(s * st_step);
                        const uint64_t b_st = b_base + (uint64_t)((uint32_t)(kx * 3) * b_step);
                        const uint32_t first = kx == 0 ? 0u : 1u;
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
                            const uint64_t a_t = a_st + (uint64_t)((uint32_t)ky * zrow16);  // the window viewed from super-row ky*Zp/2
                            const uint64_t b_t = b_st + (uint64_t)((uint32_t)ky * b_step);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                // K steps 0,1 = the 32 channels of the even grid row, 2,3 = those of the odd row: the parity
                                // picks the accumulator, the weight tile (K = 32) is the same
                                ptx::umma_f16_2sm(d_addr + (uint32_t)((k >> 1) * NF), a_t + (uint64_t)(2 * k), b_t + (uint64_t)(2 * (k & 1)), idesc,
                                                  (ky | (k & 1)) != 0 ? 1u : first);
                            }
                        }
                        ptx::umma_commit_2sm_mc(empty_bar + 8 * s, (uint16_t)0x3);  // frees the slot in both CTAs
                    }
                    __syncwarp();
                    if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
                }
                if (ptx::elect_one()) ptx::umma_commit_2sm_mc(acc_full + 8 * as, (uint16_t)0x3);
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue: 8 warps, lane group lg = warp % 4 (super-rows 32*lg + lane of the tile), column half = (warp - 2) / 4 =====
        const int lg = warp % 4;
        const int c = ((warp - 2) / 4) * 16;  // this warp's 16 output channels
        const bool do_stats = gn_stats != nullptr;
        float st_s[8], st_q[8];  // GroupNorm partials per column pair
#pragma unroll
        for (int j = 0; j < 8; ++j) st_s[j] = st_q[j] = 0.0f;
        int st_b = -1;
        uint32_t xpar = 0;

        auto flush_stats = [&]() {
            const int cpg = COUT / P.G;  // even (checked on the host)
            double gs = 0.0, gq = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                gs += (double)st_s[j];
                gq += (double)st_q[j];
                st_s[j] = st_q[j] = 0.0f;
                const int col_end = c + 2 * j + 2;
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
        };

        int local = 0, prev_tile = 0;
        for (int w = cluster_id; w < P.num_super; w += n_clusters, ++local, xpar ^= 1u) {
            const int tile = 2 * w + (int)rank;
            const int as = local & 1;
            const uint32_t aph = (uint32_t)(local >> 1) & 1u;
            const int m = 32 * lg + lane;                              // super-row of the tile
            const int64_t p0 = 2 * ((int64_t)tile * SR_OUT - 1 + m);   // even grid row of the super-row; the odd one is p0 + 1
            const bool own = m >= 1 && m <= SR_OUT;                    // super-rows 0 and 127 belong to the neighbouring tiles
            int b0 = 0, b1 = 0;
            const bool in0 = own && interior_row(p0, P, b0), in1 = own && interior_row(p0 + 1, P, b1);
            const bool v0 = P.all_rows ? (own && p0 >= 0 && p0 < P.rows) : in0;
            const bool v1 = P.all_rows ? (own && p0 + 1 >= 0 && p0 + 1 < P.rows) : in1;
            if (do_stats) {
                // valid rows of one warp share one sample (a sample boundary is two halo planes wide)
                const unsigned vmask = __ballot_sync(0xffffffffu, v0 || v1);
                if (vmask) {
                    const int b_warp = __shfl_sync(0xffffffffu, v0 ? b0 : b1, __ffs(vmask) - 1);
                    if (b_warp != st_b) {
                        if (st_b >= 0) flush_stats();
                        st_b = b_warp;
                    }
                }
            }
            ptx::mbar_wait(acc_full + 8 * as, aph);
            ptx::tc_fence_after();
            const uint32_t t_row = tmem_d + (uint32_t)(as * P.tmem_half) + ((uint32_t)(lg * 32) << 16);
            uint32_t e0[16], e1[16], e2[16], o0[16], o1[16], o2[16];  // kz = 0, 1, 2 partials of the even / odd grid row
            ptx::tmem_ld_x16(t_row + (uint32_t)c, e0);
            ptx::tmem_ld_x16(t_row + (uint32_t)(COUT + c), e1);
            ptx::tmem_ld_x16(t_row + (uint32_t)(2 * COUT + c), e2);
            ptx::tmem_ld_x16(t_row + (uint32_t)(NF + c), o0);
            ptx::tmem_ld_x16(t_row + (uint32_t)(NF + COUT + c), o1);
            ptx::tmem_ld_x16(t_row + (uint32_t)(NF + 2 * COUT + c), o2);
            ptx::tmem_ld_wait();
            // this warp has finished reading the accumulator stage
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) ptx::mbar_arrive(acc_empty + 8 * as);
                else ptx::mbar_arrive_remote(acc_empty + 8 * as, 0);  // the leader's MMA warp owns the accumulator ring
            }
            // edge super-rows for the neighbouring lane groups: odd kz=0 partial of lane 31 (next group's lane 0 needs it for its
            // even row) and even kz=2 partial of lane 0 (previous group's lane 31 needs it for its odd row)
            if (lane == 31) {
#pragma unroll
                for (int j = 0; j < 16; ++j) s_xch[xpar][0][lg][c + j] = __uint_as_float(o0[j]);
            }
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) s_xch[xpar][1][lg][c + j] = __uint_as_float(e2[j]);
            }
            const bool issuer = P.tma_out && threadIdx.x == 64;
            if (issuer) ptx::bulk_wait_read_all();  // the store that last read this tile's staging buffer (two tiles ago) is done with it
            epi_barrier();  // also: every warp has staged (and fenced) the previous tile
            if (issuer && local > 0) {
                ptx::tma_store_2d(&map_o, out_stage + (uint32_t)((local - 1) & 1) * (BM * ROWA), 0, prev_tile * SR_OUT);
                ptx::bulk_commit_group();
            }
            prev_tile = tile;
            float ve[16], vo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float up = __shfl_up_sync(0xffffffffu, __uint_as_float(o0[j]), 1);    // kz = 0 partial of grid row 2m-1
                float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(e2[j]), 1);  // kz = 2 partial of grid row 2m+2
                if (lane == 0 && lg > 0) up = s_xch[xpar][0][lg - 1][c + j];
                if (lane == 31 && lg < 3) dn = s_xch[xpar][1][lg + 1][c + j];
                const float bj = s_bias[c + j];
                ve[j] = (up + __uint_as_float(e1[j])) + (__uint_as_float(o2[j]) + bj);   // out[2m]   = D0_odd[m-1] + D1_even[m] + D2_odd[m]
                vo[j] = (__uint_as_float(e0[j]) + __uint_as_float(o1[j])) + (dn + bj);   // out[2m+1] = D0_even[m]  + D1_odd[m]  + D2_even[m+1]
            }
            auto stage_row = [&](const float (&v)[16], int odd, bool valid) {
                // super-row m of the tile -> staging row m - 1 (the store box starts at the first owned super-row); 16-byte chunk
                // j of the 128-byte line sits at j ^ (row % 8) (CU_TENSOR_MAP_SWIZZLE_128B), which also spreads the warp over all banks
                uint4 lo = make_uint4(0u, 0u, 0u, 0u), hi = lo;  // rows that are not stored stay zero (halo rows of a convolution output)
                if (valid) {
                    __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&lo);
                    __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&hi);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        h0[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                        h1[j] = __floats2bfloat162_rn(v[8 + 2 * j], v[8 + 2 * j + 1]);
                    }
                    if (do_stats) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            st_s[j] += v[2 * j] + v[2 * j + 1];
                            st_q[j] = fmaf(v[2 * j], v[2 * j], fmaf(v[2 * j + 1], v[2 * j + 1], st_q[j]));
                        }
                    }
                }
                const uint32_t r = (uint32_t)(m - 1);
                const uint32_t j0 = (uint32_t)(odd * 4 + c / 8);
                const uint32_t row_addr = out_stage + (uint32_t)(local & 1) * (BM * ROWA) + r * ROWA;
                ptx::st_shared_v4(row_addr + ((j0 ^ (r & 7u)) << 4), lo);
                ptx::st_shared_v4(row_addr + (((j0 + 1u) ^ (r & 7u)) << 4), hi);
            };
            if (P.tma_out) {
                if (own) {
                    stage_row(ve, 0, v0);
                    stage_row(vo, 1, v1);
                }
                ptx::fence_proxy_async();
                continue;
            }
            auto store_row = [&](const float (&v)[16], int64_t p, bool valid) {
                if (!valid) return;
                uint4 lo, hi;
                __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&lo);
                __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&hi);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    h0[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                    h1[j] = __floats2bfloat162_rn(v[8 + 2 * j], v[8 + 2 * j + 1]);
                }
                bf16* orow = out + p * P.ld_out + c;
                ptx::st_global_32B(orow, lo, hi);
                if (do_stats) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        st_s[j] += v[2 * j] + v[2 * j + 1];
                        st_q[j] = fmaf(v[2 * j], v[2 * j], fmaf(v[2 * j + 1], v[2 * j + 1], st_q[j]));
                    }
                }
            };
            store_row(ve, p0, v0);
            store_row(vo, p0 + 1, v1);
        }
        if (P.tma_out) {
            epi_barrier();  // the last tile is staged
            if (threadIdx.x == 64 && local > 0) {
                ptx::tma_store_2d(&map_o, out_stage + (uint32_t)((local - 1) & 1) * (BM * ROWA), 0, prev_tile * SR_OUT);
                ptx::bulk_commit_group();
                ptx::bulk_wait_all();
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

int g_num_sms_winp = 0;

}  // namespace

extern "C" int tdb_conv3d_bf16_winp(const void* in, int ld_in, const void* w_fold, const float* bias, void* out, int ld_out, int B,
                                    int X, int Y, int Z, int Cin, int Cout, double* gn_stats, int G, unsigned flags, void* stream) {
    TDB_REQUIRE(in && w_fold && out, TDB_E_BADARG, "tdb_conv3d_bf16_winp: null pointer");
    TDB_REQUIRE(Cin == CIN && Cout == COUT && ld_in == CIN && ld_out % 8 == 0, TDB_E_UNSUPPORTED,
                "tdb_conv3d_bf16_winp: 32 -> 32 channels with an input pitch of exactly 32 only (Cin=%d Cout=%d ld_in=%d)", Cin, Cout, ld_in);
    TDB_REQUIRE(((uintptr_t)in & 127) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)w_fold & 15) == 0, TDB_E_UNSUPPORTED,
                "tdb_conv3d_bf16_winp: input must be 128-byte aligned, output / weights 16-byte aligned");
    TDB_REQUIRE(!gn_stats || (G >= 1 && Cout % G == 0 && (Cout / G) % 2 == 0), TDB_E_UNSUPPORTED,
                "tdb_conv3d_bf16_winp: fused GroupNorm moments need an even number of channels per group");
    Grid3 g(B, X, Y, Z);
    TDB_REQUIRE(g.Zp % 2 == 0, TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_winp: Z + 2 = %d must be even (row pairs must not straddle a z line)", g.Zp);
    TDB_REQUIRE(g.rows < (1ll << 31) - (1 << 20), TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_winp: too many rows");
    if (g_num_sms_winp == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms_winp, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms_winp <= 0) g_num_sms_winp = 148;
    }
    WinpParams P;
    P.rows = g.rows;
    P.Xp = g.Xp; P.Yp = g.Yp; P.Zp = g.Zp;
    P.by_vox = FastDiv((uint32_t)g.vox_p);
    P.by_z = FastDiv((uint32_t)g.Zp);
    P.by_y = FastDiv((uint32_t)g.Yp);
    P.win_rows = (BM + g.Zp + 7) & ~7;  // 128 super-rows + Zp/2 on either side
    TDB_REQUIRE(P.win_rows <= 256, TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_winp: Z + 2 = %d is too wide for one TMA box", g.Zp);
    const int resident = 9 * (int)bh_bytes;
    const int stage_bytes = P.win_rows * (int)ROWA;
    const int budget = 216 * 1024;
    const int out_stage_bytes = 2 * BM * (int)ROWA + 1024;  // staged output tiles (TMA store)
    int stages = (budget - resident - 2048 - out_stage_bytes) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    TDB_REQUIRE(stages >= 2, TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_winp: two windows of %d bytes do not fit", stage_bytes);
    P.stages = stages;
    P.tmem_half = 256;  // 2 x 96 columns (even + odd parity), power of two
    P.ld_out = ld_out;
    P.G = gn_stats ? G : 0;
    const int64_t super_rows = g.rows / 2;  // rows is even: Zp is
    P.num_super = (int)ceil_div(ceil_div(super_rows, SR_OUT), 2);
    P.all_rows = (flags & TDB_CONV_ALL_ROWS) ? 1 : 0;
    TDB_REQUIRE(!(P.all_rows && gn_stats), TDB_E_BADARG, "tdb_conv3d_bf16_winp: fused moments are not available with ALL_ROWS");

    CUtensorMap map_a, map_b, map_o;
    TDB_REQUIRE(encode_fn() != nullptr, TDB_E_NODEVICE, "tdb_conv3d_bf16_winp: cuTensorMapEncodeTiled unavailable (no driver)");
    // output with the exact pitch: [rows / 2][64] super-rows, one box = the 126 owned super-rows of a tile
    P.tma_out = (ld_out == COUT && ((uintptr_t)out & 127) == 0) ? 1 : 0;
    if (P.tma_out)
        TDB_REQUIRE(make_map_2d_bf16(&map_o, out, 64, (uint64_t)super_rows, 64, 64, (uint32_t)SR_OUT), TDB_E_BADARG,
                    "tdb_conv3d_bf16_winp: tensor map (output) rejected");
    else
        map_o = CUtensorMap{};
    // activations viewed as [rows / 2][64]: one super-row = two consecutive 32-channel grid rows = one 128-byte line
    TDB_REQUIRE(make_map_2d_bf16(&map_a, in, 64, (uint64_t)super_rows, 64, 64, (uint32_t)P.win_rows), TDB_E_BADARG,
                "tdb_conv3d_bf16_winp: tensor map (activations) rejected");
    {
        // folded weights [96][9*32]: row = kz*32 + co, column = (kx*3 + ky)*32 + ci; one box = the 48 rows of one CTA
        const uint64_t dims[3] = {(uint64_t)CIN, (uint64_t)NF, 9};
        const uint64_t strides[2] = {9ull * CIN, (uint64_t)CIN};
        const uint32_t box[3] = {(uint32_t)CIN, (uint32_t)NH, 1};
        TDB_REQUIRE(make_map_bf16(&map_b, w_fold, 3, dims, strides, box), TDB_E_BADARG, "tdb_conv3d_bf16_winp: tensor map (weights) rejected");
    }
    const size_t smem = (size_t)resident + 1024 + (size_t)stages * stage_bytes + 1024 + out_stage_bytes;
    auto kern = conv3d_bf16_winp_kernel;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_conv3d_bf16_winp: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    int grid = 2 * P.num_super;
    const int cap = g_num_sms_winp & ~1;
    if (grid > cap) grid = cap;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, map_a, map_b, map_o, bias, (bf16*)out, gn_stats, P);
    TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_conv3d_bf16_winp: launch: %s", cudaGetErrorString(e));
    TDB_CHECK_LAUNCH("tdb_conv3d_bf16_winp");
    return 0;
}
