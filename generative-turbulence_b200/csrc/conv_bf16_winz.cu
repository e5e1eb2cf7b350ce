// Row-window CTA-pair convolution with kz folded into N ("winz"): the narrow layers (Cout = 32, also 64).
//
// conv_bf16_win.cu stages one shared-memory window per kx and views it nine times; its MMAs are N = Cout wide, and for
// Cout = 32 a cta_group::2 MMA costs ~50 cycles whatever its width (measured; alternating accumulators does not
// help), so those layers ran faster on the kz-folded kernels (N = 3*Cout) - which, however, fetch nine activation
// tiles per channel chunk and are bound by the L2 -> SM traffic (10 TB/s for 128->32 at 194x50x50).  This kernel
// combines both: one window per kx (3 fetches per chunk), three ky views of it, and kz folded into N:
//     D[row][kz*Cout + co] = sum_{kx,ky,ci} X[row + (kx-1)*Yp*Zp + (ky-1)*Zp][ci] * W[co][ci][kx][ky][kz]
//     out[row][co]         = D[row-1][0*Cout+co] + D[row][1*Cout+co] + D[row+1][2*Cout+co] + bias
// The +-1 row shift of the epilogue is a lane shift (warp shuffles) plus an exchange of the two edge rows of every
// warp through shared memory; tiles are 128 consecutive rows advancing by 126 (rows 0 and 127 of a tile only feed
// their neighbours).  Weights: the folded layout of tdb_conv3d_bf16_fold2 ([3*Cout][9*Cin]), resident, split over
// the pair.  Optional GroupNorm moments, fused 1x1 projection (centre view), halo rows stored (input gradients).
// Replaces nn.Conv3d(3, padding_mode="replicate") (+ res_conv) of reference ddpm.py:164,188.
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

using namespace tdb;
using bf16 = __nv_bfloat16;

namespace {

constexpr int BM = 128;
constexpr int ROWS_OUT = 126;  // valid output rows per tile
constexpr int THREADS = 320;
constexpr int MAX_STAGES = 12;

struct WinzParams {
    int64_t rows;
    int Xp, Yp, Zp;
    FastDiv by_vox, by_z, by_y;
    int Cin, chunks;
    int stages;
    int win_rows;    // rows per activation window (multiple of 8, >= 128 + 2*Zp)
    int tmem_half;   // TMEM columns of one accumulator stage
    int ld_out;
    int G;
    int num_super;   // pairs of tiles
    int all_rows;
    int proj;
    int ld_outp;
};

__device__ __forceinline__ bool interior_row(int64_t p, const WinzParams& P, int& b) {
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

template <int COUT, int KC>
__global__ void __launch_bounds__(THREADS, 1)
conv3d_bf16_winz_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const __grid_constant__ CUtensorMap map_p, const float* __restrict__ bias, bf16* __restrict__ out,
                        double* __restrict__ gn_stats, const float* __restrict__ bias_p, bf16* __restrict__ out_p,
                        const WinzParams P) {
    constexpr int NF = 3 * COUT, NH = NF / 2;  // folded N and the half of its weight rows staged by each CTA
    constexpr int PH = COUT / 2;               // projection weight rows per CTA
    constexpr int NCH = COUT / 16;             // 16-column epilogue chunks
    constexpr int CH_PER_WARP = NCH / 2;
    constexpr uint32_t ROWB = KC * 2;          // bytes per operand row = swizzle span
    constexpr uint32_t bh_bytes = NH * ROWB, ph_bytes = PH * ROWB;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 6];
    __shared__ __align__(16) float s_biasp[COUT];
    __shared__ __align__(16) float s_bias[COUT];
    __shared__ __align__(16) float s_xch[2][2][4][COUT];  // [parity][0: kz=0 row of lane 31, 1: kz=2 row of lane 0][lane group]
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
    const int n_b = 9 * chunks;  // resident folded weight tiles ((kx, ky), channel chunk) of this CTA
    const uint32_t b_region = (uint32_t)n_b * bh_bytes;
    const uint32_t p_base_addr = smem_base + b_region;
    const uint32_t p_region = P.proj ? (uint32_t)chunks * ph_bytes : 0u;
    const uint32_t stage_base = (smem_base + b_region + p_region + 1023u) & ~1023u;
    const uint32_t stage_bytes = (uint32_t)P.win_rows * ROWB;

    for (int i = threadIdx.x; i < COUT; i += THREADS) {
        s_bias[i] = bias ? bias[i] : 0.0f;
        s_biasp[i] = (P.proj && bias_p) ? bias_p[i] : 0.0f;
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_b);
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
        if (ptx::elect_one()) {
            const uint32_t b_full_l = ptx::leader_addr(b_full);
            for (int i = 0; i < n_b; ++i)
                ptx::tma_load_3d_2sm(smem_base + (uint32_t)i * bh_bytes, &map_b, b_full_l, (i % chunks) * KC, (int)rank * NH, i / chunks);
            if (rank == 0) ptx::mbar_arrive_expect_tx(b_full, 2u * b_region);
            else ptx::mbar_arrive_remote(b_full, 0);
            if (P.proj) {
                const uint32_t p_full_l = ptx::leader_addr(p_full);
                for (int ch = 0; ch < chunks; ++ch)
                    ptx::tma_load_3d_2sm(p_base_addr + (uint32_t)ch * ph_bytes, &map_p, p_full_l, ch * KC, (int)rank * PH, 0);
                if (rank == 0) ptx::mbar_arrive_expect_tx(p_full, 2u * p_region);
                else ptx::mbar_arrive_remote(p_full, 0);
            }
        }
        __syncwarp();
        const int yz = P.Yp * P.Zp;
        uint32_t s = 0, ph = 1;
        for (int w = cluster_id; w < P.num_super; w += n_clusters) {
            const int tile = 2 * w + (int)rank;
            const int q0 = tile * ROWS_OUT - 1 - P.Zp;  // first row of the kx = 1 window (rows outside the grid are zero-filled)
            for (int kx = 0; kx < 3; ++kx) {
                const int row = q0 + (kx - 1) * yz;
                for (int ch = 0; ch < chunks; ++ch) {
                    ptx::mbar_wait(empty_bar + 8 * s, ph);
                    if (ptx::elect_one()) {
                        const uint32_t full_l = ptx::leader_addr(full_bar + 8 * s);
                        ptx::tma_load_2d_2sm(stage_base + s * stage_bytes, &map_a, full_l, ch * KC, row);
                        if (rank == 0) ptx::mbar_arrive_expect_tx(full_bar + 8 * s, 2u * stage_bytes);
                        else ptx::mbar_arrive_remote(full_bar + 8 * s, 0);
                    }
                    __syncwarp();
                    if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: leader CTA only; every MMA is 256 rows (128 per CTA) x 3*COUT x 16 =====
        if (rank == 0) {
            const uint32_t idesc = ptx::umma_idesc_bf16(2 * BM, (uint32_t)NF);
            const uint32_t idesc_p = ptx::umma_idesc_bf16(2 * BM, (uint32_t)COUT);
            const uint64_t desc0 = ptx::umma_smem_desc(0, ROWB);
            const uint64_t a_base = desc0 | (uint64_t)((stage_base & 0x3FFFFu) >> 4);
            const uint64_t b_base = desc0 | (uint64_t)((smem_base & 0x3FFFFu) >> 4);
            const uint64_t p_base = desc0 | (uint64_t)((p_base_addr & 0x3FFFFu) >> 4);
            const uint32_t st_step = stage_bytes >> 4;
            constexpr uint32_t b_step = bh_bytes >> 4, p_step = ph_bytes >> 4, row16 = ROWB >> 4;
            const uint32_t zrow16 = (uint32_t)P.Zp * row16;  // one y step = Zp rows
            ptx::mbar_wait(b_full, 0);
            if (P.proj) ptx::mbar_wait(p_full, 0);
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
                    for (int ch = 0; ch < chunks; ++ch) {
                        ptx::mbar_wait(full_bar + 8 * s, ph);
                        ptx::tc_fence_after();
                        if (ptx::elect_one()) {
                            const uint64_t a_st = a_base + (uint64_t)(s * st_step);
                            const uint64_t b_st = b_base + (uint64_t)((uint32_t)(kx * 3 * chunks + ch) * b_step);
                            const uint32_t first = (kx | ch) == 0 ? 0u : 1u;
#pragma unroll
                            for (int ky = 0; ky < 3; ++ky) {
                                const uint64_t a_t = a_st + (uint64_t)((uint32_t)ky * zrow16);  // the window viewed from row ky*Zp
                                const uint64_t b_t = b_st + (uint64_t)((uint32_t)(ky * chunks) * b_step);
#pragma unroll
                                for (int k = 0; k < KC / 16; ++k)
                                    ptx::umma_f16_2sm(d_addr, a_t + (uint64_t)(2 * k), b_t + (uint64_t)(2 * k), idesc, (ky | k) != 0 ? 1u : first);
                            }
                            if (P.proj && kx == 1) {
                                // centre view (kx = ky = 1): the same rows also feed the 1x1 projection (columns behind the folded ones)
                                const uint64_t a_t = a_st + (uint64_t)zrow16;
#pragma unroll
                                for (int k = 0; k < KC / 16; ++k)
                                    ptx::umma_f16_2sm(d_addr + (uint32_t)NF, a_t + (uint64_t)(2 * k), p_base + (uint64_t)((uint32_t)ch * p_step + 2 * k),
                                                      idesc_p, (uint32_t)((ch | k) != 0));
                            }
                            ptx::umma_commit_2sm_mc(empty_bar + 8 * s, (uint16_t)0x3);  // frees the slot in both CTAs
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
        // ===== epilogue: 8 warps, lane group lg = warp % 4 (tile rows 32*lg + lane), column half = (warp - 2) / 4 =====
        const int lg = warp % 4;
        const int half = (warp - 2) / 4;
        const bool do_stats = gn_stats != nullptr;
        float st_s[CH_PER_WARP][8], st_q[CH_PER_WARP][8];  // GroupNorm partials per column pair
#pragma unroll
        for (int a = 0; a < CH_PER_WARP; ++a)
#pragma unroll
            for (int j = 0; j < 8; ++j) st_s[a][j] = st_q[a][j] = 0.0f;
        int st_b = -1;
        uint32_t xpar = 0;

        auto flush_stats = [&]() {
            const int cpg = COUT / P.G;  // even (checked on the host)
#pragma unroll
            for (int a = 0; a < CH_PER_WARP; ++a) {
                const int cidx = 2 * a + half;
                double gs = 0.0, gq = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    gs += (double)st_s[a][j];
                    gq += (double)st_q[a][j];
                    st_s[a][j] = st_q[a][j] = 0.0f;
                    const int col_end = cidx * 16 + 2 * j + 2;
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
        for (int w = cluster_id; w < P.num_super; w += n_clusters, ++local) {
            const int tile = 2 * w + (int)rank;
            const int as = local & 1;
            const uint32_t aph = (uint32_t)(local >> 1) & 1u;
            const int m = 32 * lg + lane;                              // row of the tile
            const int64_t p = (int64_t)tile * ROWS_OUT - 1 + m;        // row of the grid
            const bool own = m >= 1 && m <= ROWS_OUT;                  // rows 0 and 127 belong to the neighbouring tiles
            int b = 0;
            const bool inter = own && interior_row(p, P, b);
            const bool valid = P.all_rows ? (own && p >= 0 && p < P.rows) : inter;
            if (do_stats) {
                // valid rows of one warp share one sample (a sample boundary is two halo planes wide)
                const unsigned vmask = __ballot_sync(0xffffffffu, valid);
                if (vmask) {
                    const int b_warp = __shfl_sync(0xffffffffu, b, __ffs(vmask) - 1);
                    if (b_warp != st_b) {
                        if (st_b >= 0) flush_stats();
                        st_b = b_warp;
                    }
                }
            }
            ptx::mbar_wait(acc_full + 8 * as, aph);
            ptx::tc_fence_after();
            const uint32_t t_row = tmem_d + (uint32_t)(as * P.tmem_half) + ((uint32_t)(lg * 32) << 16);
            bf16* orow = out + p * P.ld_out;
#pragma unroll
            for (int a = 0; a < CH_PER_WARP; ++a, xpar ^= 1u) {
                const int c = (2 * a + half) * 16;
                uint32_t r0[16], r1[16], r2[16], r3[16];
                ptx::tmem_ld_x16(t_row + (uint32_t)c, r0);
                ptx::tmem_ld_x16(t_row + (uint32_t)(COUT + c), r1);
                ptx::tmem_ld_x16(t_row + (uint32_t)(2 * COUT + c), r2);
                if (P.proj) ptx::tmem_ld_x16(t_row + (uint32_t)(NF + c), r3);
                ptx::tmem_ld_wait();
                // edge rows of this warp for its neighbours: kz = 0 partial of lane 31 (needed by lane 0 of the next lane
                // group) and kz = 2 partial of lane 0 (needed by lane 31 of the previous lane group)
                if (lane == 31) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) s_xch[xpar][0][lg][c + j] = __uint_as_float(r0[j]);
                }
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) s_xch[xpar][1][lg][c + j] = __uint_as_float(r2[j]);
                }
                epi_barrier();
                if (P.proj && valid) {
                    // projection rows are unshifted: lane i holds output row i
                    uint4 lo, hi;
                    __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&lo);
                    __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&hi);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        h0[j] = __floats2bfloat162_rn(__uint_as_float(r3[2 * j]) + s_biasp[c + 2 * j],
                                                      __uint_as_float(r3[2 * j + 1]) + s_biasp[c + 2 * j + 1]);
                        h1[j] = __floats2bfloat162_rn(__uint_as_float(r3[8 + 2 * j]) + s_biasp[c + 8 + 2 * j],
                                                      __uint_as_float(r3[8 + 2 * j + 1]) + s_biasp[c + 8 + 2 * j + 1]);
                    }
                    bf16* prow = out_p + p * P.ld_outp;
                    ptx::st_global_32B(prow + c, lo, hi);
                }
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float up = __shfl_up_sync(0xffffffffu, __uint_as_float(r0[j]), 1);    // kz = 0 partial of row m-1
                    float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[j]), 1);  // kz = 2 partial of row m+1
                    if (lane == 0 && lg > 0) up = s_xch[xpar][0][lg - 1][c + j];
                    if (lane == 31 && lg < 3) dn = s_xch[xpar][1][lg + 1][c + j];
                    v[j] = (up + __uint_as_float(r1[j])) + (dn + s_bias[c + j]);
                }
                if (valid) {
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

int g_num_sms_winz = 0;

template <int COUT, int KC>
int launch_winz(const CUtensorMap& map_a, const CUtensorMap& map_b, const CUtensorMap& map_p, const float* bias, bf16* out,
                double* gn_stats, const float* bias_p, bf16* out_p, const WinzParams& P, size_t smem, cudaStream_t stream) {
    auto kern = conv3d_bf16_winz_kernel<COUT, KC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_conv3d_bf16_winz: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    int grid = 2 * P.num_super;
    const int cap = g_num_sms_winz & ~1;
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
    e = cudaLaunchKernelEx(&cfg, kern, map_a, map_b, map_p, bias, out, gn_stats, bias_p, out_p, P);
    TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_conv3d_bf16_winz: launch: %s", cudaGetErrorString(e));
    TDB_CHECK_LAUNCH("tdb_conv3d_bf16_winz");
    return 0;
}

}  // namespace

extern "C" int tdb_conv3d_bf16_winz(const void* in, int ld_in, const void* w_fold, const float* bias, void* out, int ld_out, int B,
                                    int X, int Y, int Z, int Cin, int Cout, double* gn_stats, int G, unsigned flags,
                                    const void* w_proj, const float* bias_proj, void* out_proj, int ld_outp, void* stream) {
    TDB_REQUIRE(in && w_fold && out, TDB_E_BADARG, "tdb_conv3d_bf16_winz: null pointer");
    TDB_REQUIRE(Cin % 32 == 0 && (Cout == 32 || Cout == 64) && ld_in % 8 == 0 && ld_out % 8 == 0, TDB_E_UNSUPPORTED,
                "tdb_conv3d_bf16_winz: need Cin %% 32 == 0 and Cout in {32,64} (Cin=%d Cout=%d)", Cin, Cout);
    TDB_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)w_fold & 15) == 0, TDB_E_UNSUPPORTED,
                "tdb_conv3d_bf16_winz: pointers must be 16-byte aligned");
    TDB_REQUIRE(!gn_stats || (G >= 1 && Cout % G == 0 && (Cout / G) % 2 == 0), TDB_E_UNSUPPORTED,
                "tdb_conv3d_bf16_winz: fused GroupNorm moments need an even number of channels per group");
    Grid3 g(B, X, Y, Z);
    TDB_REQUIRE(g.rows < (1ll << 31) - (1 << 20), TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_winz: too many rows");
    if (g_num_sms_winz == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms_winz, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms_winz <= 0) g_num_sms_winz = 148;
    }
    const int KC = Cin % 64 == 0 ? 64 : 32;
    WinzParams P;
    P.rows = g.rows;
    P.Xp = g.Xp; P.Yp = g.Yp; P.Zp = g.Zp;
    P.by_vox = FastDiv((uint32_t)g.vox_p);
    P.by_z = FastDiv((uint32_t)g.Zp);
    P.by_y = FastDiv((uint32_t)g.Yp);
    P.Cin = Cin;
    P.chunks = Cin / KC;
    P.win_rows = (BM + 2 * g.Zp + 7) & ~7;
    TDB_REQUIRE(P.win_rows <= 256, TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_winz: Z + 2 = %d is too wide for one TMA box", g.Zp);
    P.proj = w_proj != nullptr ? 1 : 0;
    P.ld_outp = ld_outp;
    TDB_REQUIRE(!P.proj || (out_proj && ld_outp % 8 == 0 && ((uintptr_t)w_proj & 15) == 0 && ((uintptr_t)out_proj & 15) == 0 &&
                            !(flags & TDB_CONV_ALL_ROWS)),
                TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_winz: the fused projection needs aligned buffers and no ALL_ROWS");
    const int NF = 3 * Cout;
    const int bh_bytes = (NF / 2) * KC * 2, ph_bytes = (Cout / 2) * KC * 2;
    const int resident = 9 * P.chunks * bh_bytes + (P.proj ? P.chunks * ph_bytes : 0);
    const int stage_bytes = P.win_rows * KC * 2;
    const int budget = 221 * 1024;
    int stages = (budget - resident - 2048) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    TDB_REQUIRE(stages >= 2, TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_winz: weights (%d bytes per CTA) do not leave room for two windows", resident);
    P.stages = stages;
    int half = 32;
    while (half < NF + (P.proj ? Cout : 0)) half *= 2;
    P.tmem_half = half;
    P.ld_out = ld_out;
    P.G = gn_stats ? G : 0;
    P.num_super = (int)ceil_div(ceil_div(g.rows, ROWS_OUT), 2);
    P.all_rows = (flags & TDB_CONV_ALL_ROWS) ? 1 : 0;
    TDB_REQUIRE(!(P.all_rows && gn_stats), TDB_E_BADARG, "tdb_conv3d_bf16_winz: fused moments are not available with ALL_ROWS");

    CUtensorMap map_a, map_b, map_p;
    TDB_REQUIRE(encode_fn() != nullptr, TDB_E_NODEVICE, "tdb_conv3d_bf16_winz: cuTensorMapEncodeTiled unavailable (no driver)");
    TDB_REQUIRE(make_map_2d_bf16(&map_a, in, (uint64_t)Cin, (uint64_t)g.rows, (uint64_t)ld_in, (uint32_t)KC, (uint32_t)P.win_rows),
                TDB_E_BADARG, "tdb_conv3d_bf16_winz: tensor map (activations) rejected");
    {
        // folded weights [3*Cout][9*Cin]: row = kz*Cout + co, column = (kx*3 + ky)*Cin + ci; one box = the rows of one CTA
        const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)NF, 9};
        const uint64_t strides[2] = {9ull * Cin, (uint64_t)Cin};
        const uint32_t box[3] = {(uint32_t)KC, (uint32_t)(NF / 2), 1};
        TDB_REQUIRE(make_map_bf16(&map_b, w_fold, 3, dims, strides, box), TDB_E_BADARG, "tdb_conv3d_bf16_winz: tensor map (weights) rejected");
    }
    {
        const void* wp = P.proj ? w_proj : w_fold;
        const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)(P.proj ? Cout : NF), 1};
        const uint64_t strides[2] = {(uint64_t)(P.proj ? Cin : 9 * Cin), (uint64_t)Cin * (P.proj ? Cout : NF)};
        const uint32_t box[3] = {(uint32_t)KC, (uint32_t)(Cout / 2), 1};
        TDB_REQUIRE(make_map_bf16(&map_p, wp, 3, dims, strides, box), TDB_E_BADARG, "tdb_conv3d_bf16_winz: tensor map (projection) rejected");
    }
    const size_t smem = (size_t)resident + 1024 + (size_t)stages * stage_bytes + 1024;
    cudaStream_t s = (cudaStream_t)stream;
    bf16* o = (bf16*)out;
    bf16* op = (bf16*)out_proj;
#define TDB_WINZ_CASE(CO, K) \
    if (Cout == CO && KC == K) return launch_winz<CO, K>(map_a, map_b, map_p, bias, o, gn_stats, bias_proj, op, P, smem, s)
    TDB_WINZ_CASE(32, 64);
    TDB_WINZ_CASE(64, 64);
    TDB_WINZ_CASE(32, 32);
    TDB_WINZ_CASE(64, 32);
#undef TDB_WINZ_CASE
    tdb::set_error("tdb_conv3d_bf16_winz: no kernel for Cout=%d KC=%d", Cout, KC);
    return TDB_E_UNSUPPORTED;
}
