// cta_group::2 variant of the kz-folded convolution (see conv_bf16_fold.cu for the folding itself).
//
// The 1-CTA kernel is bound by the shared-memory ingest of an SM (~60 B/clk): per K step it must receive a 16 KB
// activation tile AND the 3*Cout x 64 weight chunk (24 KB for Cout = 64) for only 384 MMA cycles.  Here two
// CTAs of a cluster form ONE 256-row MMA (tcgen05.mma.cta_group::2): each CTA stages its own 128 activation rows
// and only HALF of the weight rows (the hardware shares the halves), which halves the weight ingest - and makes
// the per-CTA half of the folded weights (<= 110.6 KB for 64->64 and 128->32) RESIDENT in shared memory, so that
// only activations stream.  Roles per CTA: warp 0 TMA producer (loads signal the LEADER's barrier), warp 1
// TMEM alloc (+ MMA issue in the leader), warps 2-9 epilogue on the CTA's own TMEM lanes.
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

using namespace tdb;
using bf16 = __nv_bfloat16;

namespace {

constexpr int BM = 128;
constexpr int ROWS_WARP = 30;
constexpr int ROWS_OUT = 4 * ROWS_WARP;
constexpr int THREADS = 320;
constexpr int MAX_STAGES = 12;
constexpr int KC = 64;

struct FoldParams {
    int64_t rows;
    uint32_t vox_p;
    int Xp, Yp, Zp;
    FastDiv by_vox, by_z, by_y;
    int Cin;
    int stages;
    int pad_rows;
    int b_resident;
    int tmem_half;
    int ld_out;
    int G;
    int num_super;   // pairs of 120-row tiles
    int n_tiles;     // N tiles of COUT output channels (Cout_total / COUT)
    int num_items;   // num_super * n_tiles work items, N tile outermost
    int cout_total;
    int all_rows;
    int proj;        // 1: also compute the block's 1x1 residual projection of the SAME input (centre tap) into out_p
    int ld_outp;
};

__device__ __forceinline__ bool interior_row(int64_t p, const FoldParams& P, int& b) {
    if (p < 0 || p >= P.rows) return false;
    uint32_t bb, r, q, zp, xp, yp;
    P.by_vox.divmod((uint32_t)p, bb, r);
    P.by_z.divmod(r, q, zp);
    P.by_y.divmod(q, xp, yp);
    b = (int)bb;
    return xp >= 1u && xp <= (uint32_t)(P.Xp - 2) && yp >= 1u && yp <= (uint32_t)(P.Yp - 2) && zp >= 1u &&
           zp <= (uint32_t)(P.Zp - 2);
}

template <int COUT, int RES_T>
__global__ void __launch_bounds__(THREADS, 1)
conv3d_bf16_fold2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         const __grid_constant__ CUtensorMap map_p, const float* __restrict__ bias, bf16* __restrict__ out,
                         double* __restrict__ gn_stats, const float* __restrict__ bias_p, bf16* __restrict__ out_p,
                         const FoldParams P) {
    constexpr int NF = 3 * COUT, NH = NF / 2;  // folded N of one N tile, and the half staged by each CTA
    constexpr int NSUB = NF > 256 ? 2 : 1;     // MMAs per K step (N <= 256 each)
    constexpr int NSUBN = NF / NSUB;           // N of one MMA
    constexpr int PIECE = NSUBN / 2;           // weight rows each CTA contributes to one MMA
    constexpr int ACC_STAGES = NF > 256 ? 1 : 2;  // TMEM holds 512 columns: 384-wide accumulators cannot be double-buffered
    constexpr int NCH = COUT / 16;
    constexpr int CH_PER_WARP = (NCH + 1) / 2;
    constexpr bool resident = RES_T != 0;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 6];
    __shared__ __align__(16) float s_biasp[64];
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float s_bias[512];

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t rank = ptx::cluster_ctarank();
    const int cluster_id = blockIdx.x / 2, n_clusters = gridDim.x / 2;
    const uint32_t full_bar = ptx::smem_u32(&bars[0]);                    // used in the leader
    const uint32_t empty_bar = ptx::smem_u32(&bars[MAX_STAGES]);          // per CTA (multicast commit)
    const uint32_t acc_full = ptx::smem_u32(&bars[2 * MAX_STAGES]);       // [2] per CTA (multicast commit)
    const uint32_t acc_empty = ptx::smem_u32(&bars[2 * MAX_STAGES + 2]);  // [2] used in the leader, 16 arrivals
    const uint32_t b_full = ptx::smem_u32(&bars[2 * MAX_STAGES + 4]);     // used in the leader
    const uint32_t p_full = ptx::smem_u32(&bars[2 * MAX_STAGES + 5]);     // used in the leader (projection weights)
    const int chunks = P.Cin / KC;
    const int k_iters = 9 * chunks;
    constexpr uint32_t a_bytes = BM * KC * 2, bh_bytes = NH * KC * 2;
    const uint32_t b_region = resident ? (uint32_t)k_iters * bh_bytes : 0u;
    constexpr uint32_t stage_bytes = a_bytes + (resident ? 0u : bh_bytes);
    // fused 1x1 projection (COUT <= 64): each CTA keeps COUT/2 rows x Cin of the projection weights resident
    constexpr uint32_t ph_bytes = (COUT / 2) * KC * 2;
    const uint32_t p_base_addr = smem_base + b_region;
    const uint32_t p_region = P.proj ? (uint32_t)chunks * ph_bytes : 0u;
    const uint32_t stage_base = smem_base + b_region + p_region;

    for (int i = threadIdx.x; i < P.cout_total; i += THREADS) s_bias[i] = bias ? bias[i] : 0.0f;
    if (P.proj)
        for (int i = threadIdx.x; i < COUT; i += THREADS) s_biasp[i % 64] = bias_p ? bias_p[i] : 0.0f;
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
        ptx::tmem_alloc_2sm(ptx::smem_u32(&tmem_base_slot), (uint32_t)(ACC_STAGES * P.tmem_half));
        ptx::tmem_relinquish_2sm();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();  // both CTAs: barriers initialised, TMEM allocated
    ptx::tc_fence_after();
    const uint32_t tmem_d = tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own activation rows + own half of the weight rows; every load signals
        // the LEADER's full barrier =====
        const uint32_t b_full_l = ptx::leader_addr(b_full);
        if (resident && ptx::elect_one()) {
            for (int kc = 0; kc < k_iters; ++kc)
                ptx::tma_load_3d_2sm(smem_base + kc * bh_bytes, &map_b, b_full_l, (kc % chunks) * KC, (int)rank * NH, kc / chunks);
            if (rank == 0) ptx::mbar_arrive_expect_tx(b_full, 2u * (uint32_t)k_iters * bh_bytes);
            else ptx::mbar_arrive_remote(b_full, 0);
        }
        if (P.proj && ptx::elect_one()) {
            const uint32_t p_full_l = ptx::leader_addr(p_full);
            for (int ch = 0; ch < chunks; ++ch)
                ptx::tma_load_3d_2sm(p_base_addr + ch * ph_bytes, &map_p, p_full_l, ch * KC, (int)rank * (COUT / 2), 0);
            if (rank == 0) ptx::mbar_arrive_expect_tx(p_full, 2u * (uint32_t)chunks * ph_bytes);
            else ptx::mbar_arrive_remote(p_full, 0);
        }
        __syncwarp();
        const int yz = P.Yp * P.Zp;
        uint32_t s = 0, ph = 1;
        for (int w = cluster_id; w < P.num_items; w += n_clusters) {
            const int nt = w / P.num_super, st = w - nt * P.num_super;
            const int tile = 2 * st + (int)rank;
            const int q0 = tile * ROWS_OUT - 1 + P.pad_rows;
            for (int t9 = 0; t9 < 9; ++t9) {
                const int row = q0 + (t9 / 3 - 1) * yz + (t9 % 3 - 1) * P.Zp;
                for (int ch = 0; ch < chunks; ++ch) {
                    ptx::mbar_wait(empty_bar + 8 * s, ph);
                    if (ptx::elect_one()) {
                        const uint32_t a_dst = stage_base + s * stage_bytes;
                        const uint32_t full_l = ptx::leader_addr(full_bar + 8 * s);
                        ptx::tma_load_4d_2sm(a_dst, &map_a, full_l, ch * KC, row, 0, 0);
                        if (!resident) {
                            // for MMA j this CTA supplies folded-weight rows [j*NSUBN + rank*PIECE, +PIECE) of N tile nt
#pragma unroll
                            for (int j = 0; j < NSUB; ++j)
                                ptx::tma_load_3d_2sm(a_dst + a_bytes + j * (PIECE * KC * 2), &map_b, full_l, ch * KC,
                                                     nt * NF + j * NSUBN + (int)rank * PIECE, t9);
                        }
                        if (rank == 0) ptx::mbar_arrive_expect_tx(full_bar + 8 * s, 2u * stage_bytes);
                        else ptx::mbar_arrive_remote(full_bar + 8 * s, 0);
                    }
                    __syncwarp();
                    if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: leader CTA only; one 256 x NF x 16 MMA spans both SMs =====
        if (rank == 0) {
            const uint32_t idesc = ptx::umma_idesc_bf16(2 * BM, (uint32_t)NSUBN);
            const uint64_t desc0 = ptx::umma_smem_desc(0, (uint32_t)KC * 2u);
            constexpr uint32_t b_step = bh_bytes >> 4, st_step = stage_bytes >> 4;
            const uint64_t a_base = desc0 | (uint64_t)((stage_base & 0x3FFFFu) >> 4);
            const uint64_t b_base = desc0 | (uint64_t)(((resident ? smem_base : stage_base + a_bytes) & 0x3FFFFu) >> 4);
            if (resident) {
                ptx::mbar_wait(b_full, 0);
                ptx::tc_fence_after();
            }
            const uint32_t idesc_p = ptx::umma_idesc_bf16(2 * BM, (uint32_t)(COUT < 128 ? COUT : 64));
            const uint64_t p_base = desc0 | (uint64_t)((p_base_addr & 0x3FFFFu) >> 4);
            if (P.proj) {
                ptx::mbar_wait(p_full, 0);
                ptx::tc_fence_after();
            }
            uint32_t s = 0, ph = 0;
            int local = 0;
            for (int w = cluster_id; w < P.num_items; w += n_clusters, ++local) {
                const int as = ACC_STAGES == 2 ? (local & 1) : 0;
                const uint32_t aph = (uint32_t)(ACC_STAGES == 2 ? (local >> 1) : local) & 1u;
                ptx::mbar_wait(acc_empty + 8 * as, aph ^ 1u);  // both CTAs' epilogues have drained this stage
                ptx::tc_fence_after();
                const uint32_t d_addr = tmem_d + (uint32_t)(as * P.tmem_half);
                for (int i = 0; i < k_iters; ++i) {
                    ptx::mbar_wait(full_bar + 8 * s, ph);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
                        const uint64_t a_st = a_base + (uint64_t)(s * st_step);
                        const uint64_t b_st = resident ? b_base + (uint64_t)((uint32_t)i * b_step) : b_base + (uint64_t)(s * st_step);
#pragma unroll
                        for (int j = 0; j < NSUB; ++j)
#pragma unroll
                            for (int k = 0; k < KC / 16; ++k)
                                ptx::umma_f16_2sm(d_addr + (uint32_t)(j * NSUBN), a_st + (uint64_t)(2 * k),
                                                  b_st + (uint64_t)(j * ((PIECE * KC * 2) >> 4) + 2 * k), idesc, (uint32_t)((i | k) != 0));
                        if (P.proj && i >= 4 * chunks && i < 5 * chunks) {
                            // centre tap (kx = ky = 1): the same activation tile also feeds the 1x1 projection,
                            // accumulated over the channel chunks into the TMEM columns behind the folded ones
                            const int ch = i - 4 * chunks;
#pragma unroll
                            for (int k = 0; k < KC / 16; ++k)
                                ptx::umma_f16_2sm(d_addr + (uint32_t)NF, a_st + (uint64_t)(2 * k),
                                                  p_base + (uint64_t)((uint32_t)ch * (ph_bytes >> 4) + 2 * k), idesc_p, (uint32_t)((ch | k) != 0));
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
        // ===== epilogue: 8 warps, lane group lg = warp % 4, column half = (warp - 2) / 4 =====
        const int lg = warp % 4;
        const int half = (warp - 2) / 4;
        const bool do_stats = gn_stats != nullptr;
        // per-thread GroupNorm partials: column PAIRS of this warp's chunks, across all tiles of one sample
        float st_s[CH_PER_WARP][8], st_q[CH_PER_WARP][8];
#pragma unroll
        for (int a = 0; a < CH_PER_WARP; ++a)
#pragma unroll
            for (int j = 0; j < 8; ++j) st_s[a][j] = st_q[a][j] = 0.0f;
        int st_b = -1, st_nt = 0;

        auto flush_stats = [&]() {
            // pairs -> groups (cpg even), warp reduce in double, one atomic per (warp, group, moment)
            const int cpg = P.cout_total / P.G;
#pragma unroll
            for (int a = 0; a < CH_PER_WARP; ++a) {
                const int cidx = 2 * a + half;  // chunk index of this warp
                if (cidx < NCH) {
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
            }
        };

        int local = 0;
        for (int w = cluster_id; w < P.num_items; w += n_clusters, ++local) {
            const int nt = w / P.num_super, st = w - nt * P.num_super;
            const int tile = 2 * st + (int)rank;
            const int n0 = nt * COUT;
            const int as = ACC_STAGES == 2 ? (local & 1) : 0;
            const uint32_t aph = (uint32_t)(ACC_STAGES == 2 ? (local >> 1) : local) & 1u;
            const int64_t p = (int64_t)tile * ROWS_OUT + ROWS_WARP * lg - 1 + lane;
            int b = 0;
            const bool inter = lane >= 1 && lane <= ROWS_WARP && interior_row(p, P, b);
            const bool valid = P.all_rows ? (lane >= 1 && lane <= ROWS_WARP && p >= 0 && p < P.rows) : inter;
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
                if (c < COUT) {
                    uint32_t r0[16], r1[16], r2[16], r3[16];
                    ptx::tmem_ld_x16(t_row + (uint32_t)c, r0);
                    ptx::tmem_ld_x16(t_row + (uint32_t)(COUT + c), r1);
                    ptx::tmem_ld_x16(t_row + (uint32_t)(2 * COUT + c), r2);
                    if (P.proj) ptx::tmem_ld_x16(t_row + (uint32_t)(NF + c), r3);
                    ptx::tmem_ld_wait();
                    if (P.proj && valid) {
                        // projection rows are unshifted: lane i holds output row i
                        uint4 lo, hi;
                        __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&lo);
                        __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&hi);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            h0[j] = __floats2bfloat162_rn(__uint_as_float(r3[2 * j]) + s_biasp[(c + 2 * j) % 64],
                                                          __uint_as_float(r3[2 * j + 1]) + s_biasp[(c + 2 * j + 1) % 64]);
                            h1[j] = __floats2bfloat162_rn(__uint_as_float(r3[8 + 2 * j]) + s_biasp[(c + 8 + 2 * j) % 64],
                                                          __uint_as_float(r3[8 + 2 * j + 1]) + s_biasp[(c + 8 + 2 * j + 1) % 64]);
                        }
                        bf16* prow = out_p + p * P.ld_outp;
                        ptx::st_global_32B(prow + c, lo, hi);
                    }
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(r0[j]), 1);    // Y_0 of row i-1
                        const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[j]), 1);  // Y_2 of row i+1
                        v[j] = (up + __uint_as_float(r1[j])) + (dn + s_bias[n0 + c + j]);
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
        ptx::tmem_dealloc_2sm(tmem_d, (uint32_t)(ACC_STAGES * P.tmem_half));
    }
}

int g_num_sms2 = 0;

template <int COUT, int RES_T>
int launch_fold2(const CUtensorMap& map_a, const CUtensorMap& map_b, const CUtensorMap& map_p, const float* bias, bf16* out,
                 double* gn_stats, const float* bias_p, bf16* out_p, const FoldParams& P, size_t smem, cudaStream_t stream) {
    auto kern = conv3d_bf16_fold2_kernel<COUT, RES_T>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_conv3d_bf16_fold2: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    int grid = 2 * P.num_items;
    const int cap = g_num_sms2 & ~1;
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
    TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_conv3d_bf16_fold2: launch: %s", cudaGetErrorString(e));
    TDB_CHECK_LAUNCH("tdb_conv3d_bf16_fold2");
    return 0;
}

}  // namespace

// Same contract as tdb_conv3d_bf16_fold (include/turbdiff_b200.h); Cin % 64 == 0 and Cout in {32, 64} or a multiple of
// 128 (<= 512), which is processed as N tiles of 128 channels: w_fold rows are then ordered [n tile][kz][co in tile].
extern "C" int tdb_conv3d_bf16_fold2(const void* in, int ld_in, int pad_rows, const void* w_fold, const float* bias, void* out,
                                     int ld_out, int B, int X, int Y, int Z, int Cin, int Cout, double* gn_stats, int G,
                                     unsigned flags, const void* w_proj, const float* bias_proj, void* out_proj, int ld_outp,
                                     void* stream) {
    TDB_REQUIRE(in && w_fold && out, TDB_E_BADARG, "tdb_conv3d_bf16_fold2: null pointer");
    TDB_REQUIRE(Cin % 64 == 0 && (Cout == 32 || Cout == 64 || (Cout % 128 == 0 && Cout <= 512)) && ld_in % 8 == 0 && ld_out % 8 == 0,
                TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_fold2: need Cin %% 64 == 0 and Cout in {32,64,128k<=512} (Cin=%d Cout=%d)", Cin, Cout);
    TDB_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)w_fold & 15) == 0, TDB_E_UNSUPPORTED,
                "tdb_conv3d_bf16_fold2: pointers must be 16-byte aligned");
    TDB_REQUIRE(!gn_stats || (G >= 1 && Cout % G == 0 && (Cout / G) % 2 == 0), TDB_E_UNSUPPORTED,
                "tdb_conv3d_bf16_fold2: fused GroupNorm moments need an even number of channels per group");
    Grid3 g(B, X, Y, Z);
    TDB_REQUIRE(g.rows + 2ll * pad_rows < (1ll << 31) - 4096, TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_fold2: too many rows");
    TDB_REQUIRE(pad_rows >= g.Yp * g.Zp + 2 * g.Zp + 256, TDB_E_BADARG, "tdb_conv3d_bf16_fold2: pad_rows=%d too small", pad_rows);
    if (g_num_sms2 == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms2, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms2 <= 0) g_num_sms2 = 148;
    }
    FoldParams P;
    P.rows = g.rows;
    P.vox_p = (uint32_t)g.vox_p;
    P.Xp = g.Xp; P.Yp = g.Yp; P.Zp = g.Zp;
    P.by_vox = FastDiv((uint32_t)g.vox_p);
    P.by_z = FastDiv((uint32_t)g.Zp);
    P.by_y = FastDiv((uint32_t)g.Yp);
    P.Cin = Cin;
    P.pad_rows = pad_rows;
    const int tile_n = Cout >= 128 ? 128 : Cout;  // output channels per N tile
    const int NF = 3 * tile_n, NH = NF / 2;
    P.n_tiles = Cout / tile_n;
    P.cout_total = Cout;
    const int a_bytes = BM * KC * 2, bh_bytes = NH * KC * 2;
    const int k_iters = 9 * (Cin / KC);
    const int budget = 221 * 1024;
    P.proj = w_proj != nullptr ? 1 : 0;
    P.ld_outp = ld_outp;
    TDB_REQUIRE(!P.proj || (Cout <= 64 && out_proj && ld_outp % 8 == 0 && ((uintptr_t)w_proj & 15) == 0 && ((uintptr_t)out_proj & 15) == 0 &&
                            !(flags & TDB_CONV_ALL_ROWS)),
                TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_fold2: the fused projection needs Cout <= 64 and aligned buffers");
    const int proj_bytes = P.proj ? (Cin / KC) * (Cout / 2) * KC * 2 : 0;
    P.b_resident = (P.n_tiles == 1 && (int64_t)k_iters * bh_bytes + proj_bytes <= 116 * 1024) ? 1 : 0;
    const int resident_bytes = (P.b_resident ? k_iters * bh_bytes : 0) + proj_bytes;
    const int unit = a_bytes + (P.b_resident ? 0 : bh_bytes);
    int stages = (budget - resident_bytes) / unit;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    TDB_REQUIRE(stages >= 2, TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_fold2: tiles do not fit in shared memory");
    P.stages = stages;
    int half = 32;
    while (half < NF) half *= 2;
    P.tmem_half = half;
    P.ld_out = ld_out;
    P.G = gn_stats ? G : 0;
    P.num_super = (int)ceil_div(g.rows, 2 * ROWS_OUT);
    P.num_items = P.num_super * P.n_tiles;
    P.all_rows = (flags & TDB_CONV_ALL_ROWS) ? 1 : 0;
    TDB_REQUIRE(!(P.all_rows && gn_stats), TDB_E_BADARG, "tdb_conv3d_bf16_fold2: fused moments are not available with ALL_ROWS");

    CUtensorMap map_a, map_b, map_p;
    TDB_REQUIRE(encode_fn() != nullptr, TDB_E_NODEVICE, "tdb_conv3d_bf16_fold2: cuTensorMapEncodeTiled unavailable (no driver)");
    {
        const bf16* base = (const bf16*)in - (int64_t)pad_rows * ld_in;
        const uint64_t total_rows = (uint64_t)g.rows + 2ull * pad_rows;
        const uint64_t dims[4] = {(uint64_t)Cin, total_rows - 90 - 2ull * g.Zp, 4, 3};
        const uint64_t strides[3] = {(uint64_t)ld_in, 30ull * ld_in, (uint64_t)g.Zp * ld_in};
        const uint32_t box[4] = {(uint32_t)KC, 32, 4, 1};
        TDB_REQUIRE(make_map_bf16(&map_a, base, 4, dims, strides, box), TDB_E_BADARG, "tdb_conv3d_bf16_fold2: tensor map (activations) rejected");
    }
    {
        // folded weights [n_tiles * 3 * tile_n][9 * Cin]; one box = the rows one CTA contributes to one MMA
        const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)NF * P.n_tiles, 9};
        const uint64_t strides[2] = {9ull * Cin, (uint64_t)Cin};
        const uint32_t box[3] = {(uint32_t)KC, (uint32_t)(NF > 256 ? NF / 4 : NH), 1};
        TDB_REQUIRE(make_map_bf16(&map_b, w_fold, 3, dims, strides, box), TDB_E_BADARG, "tdb_conv3d_bf16_fold2: tensor map (weights) rejected");
    }
    {
        // projection weights [Cout][Cin] (the 1x1 convolution's ordinary layout); one box = the COUT/2 rows of a CTA
        const void* wp = P.proj ? w_proj : w_fold;
        const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)(P.proj ? Cout : NF), 1};
        const uint64_t strides[2] = {(uint64_t)(P.proj ? Cin : 9 * Cin), (uint64_t)Cin * (P.proj ? Cout : NF)};
        const uint32_t box[3] = {(uint32_t)KC, (uint32_t)(Cout >= 128 ? 64 : Cout / 2), 1};
        TDB_REQUIRE(make_map_bf16(&map_p, wp, 3, dims, strides, box), TDB_E_BADARG, "tdb_conv3d_bf16_fold2: tensor map (projection) rejected");
    }
    const size_t smem = (size_t)resident_bytes + (size_t)stages * unit + 1024;
    cudaStream_t s = (cudaStream_t)stream;
    bf16* o = (bf16*)out;
    bf16* op = (bf16*)out_proj;
    if (Cout == 32) {
        if (P.b_resident) return launch_fold2<32, 1>(map_a, map_b, map_p, bias, o, gn_stats, bias_proj, op, P, smem, s);
        return launch_fold2<32, 0>(map_a, map_b, map_p, bias, o, gn_stats, bias_proj, op, P, smem, s);
    }
    if (Cout >= 128) return launch_fold2<128, 0>(map_a, map_b, map_p, bias, o, gn_stats, bias_proj, op, P, smem, s);
    if (P.b_resident) return launch_fold2<64, 1>(map_a, map_b, map_p, bias, o, gn_stats, bias_proj, op, P, smem, s);
    return launch_fold2<64, 0>(map_a, map_b, map_p, bias, o, gn_stats, bias_proj, op, P, smem, s);
}
