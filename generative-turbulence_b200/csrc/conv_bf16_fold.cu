// kz-folded, persistent bf16 implicit-GEMM 3x3x3 convolution for the narrow layers (Cout in
// {16,32,64}), which carry 60 % of the network's FLOPs at full resolution and are L2->SMEM bound in
// the plain per-tap kernel (conv_bf16_tc.cu fetches every input row 27 times for only Cout MACs each).
//
// Fold the fastest filter axis into the GEMM N dimension.  With rows linearised over the halo grid,
//     out[p] = sum_kz Y_kz[p + kz - 1],      Y_kz[q] = sum_{kx,ky,ci} W[kx,ky,kz][ci] * in[q + s(kx,ky)][ci]
//     s(kx,ky) = (kx-1)*Yp*Zp + (ky-1)*Zp
// so one GEMM with N' = 3*Cout (column kz*Cout+co) and K' = 9*Cin produces all three Y_kz from only
// NINE row-shifted A tiles (3x less activation traffic, 3x wider MMAs so the SMEM operand bandwidth
// per MMA cycle stays below 128 B/clk), and the kz shift becomes a +-1 LANE shift of accumulator rows
// in the epilogue (two warp shuffles per value).  To keep the shift inside a warp, the 128-row A tile
// consists of FOUR 32-row groups that overlap by two rows: TMEM lane group w holds rows
// q = tile*120 + 30*w - 1 + lane, and lanes 1..30 of each warp produce outputs (120 rows per tile).
// ONE 4-D TMA box fetches a whole pipeline stage of A: the tensor map views the activation matrix as
// [ky (pitch Zp rows)][w (pitch 30 rows)][r][c], so the box {KC, 32, 4, TY} lands as TY consecutive
// canonical 128-row K-major tiles (TY = 3 ky taps per stage when shared memory allows, else 1).  The
// overlapping view cannot rely on TMA's out-of-bounds fill, so the caller guarantees `pad_rows`
// readable rows before and after the halo grid (never used by a stored output).
//
// Persistent CTAs (one per SM) walk tiles round-robin.  TMEM holds two accumulator stages so the
// epilogue of tile j overlaps the MMAs of tile j+1.  When the folded weights fit (<= 112 KB: the
// 32->32 layers) they are loaded into shared memory ONCE per CTA and only activations stream.
// GroupNorm moments are accumulated in registers across all tiles of the CTA and flushed with one
// double atomicAdd per (warp, group) at the end (or when the sample index changes).
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 TMEM alloc + MMA issue, warps 2-9 epilogue
// (two warps per TMEM lane group, alternating 16-column chunks: two warps per scheduler hide the
// TMEM-load / shuffle latencies).
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

using namespace tdb;
using bf16 = __nv_bfloat16;

namespace {

constexpr int BM = 128;
constexpr int ROWS_WARP = 30;             // output rows per lane group
constexpr int ROWS_OUT = 4 * ROWS_WARP;   // output rows per tile
constexpr int THREADS = 320;
constexpr int MAX_STAGES = 12;

struct FoldParams {
    int64_t rows;
    uint32_t vox_p;
    int Xp, Yp, Zp;
    FastDiv by_vox, by_z, by_y;
    int Cin;
    int KC;          // channels per K chunk
    int stages;
    int TY;          // ky taps per pipeline stage (1 or 3)
    int a_bytes;     // one 128-row A tile (1024-aligned)
    int b_bytes;     // one B chunk (1024-aligned)
    int pad_rows;    // rows of padding in front of the halo grid (tensor-map row 0 = grid row -pad_rows)
    int b_resident;  // 1: all 9*Cin/KC B chunks live in smem for the whole kernel
    int tmem_half;   // columns per accumulator stage
    int ld_out;
    int G;           // groups for fused GroupNorm moments (0 = off); (Cout/G) even
    int num_tiles;
    int all_rows;    // 1: store halo rows too (input-gradient use)
};

// interior test of a linear halo-grid row; returns the sample index through b
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

// KC_T / TY_T / RES_T > 0 (>= 0 for RES_T) specialise the K chunk, taps per stage and weight residency at
// compile time (the hot shapes of the shapes config); -1 reads them from the launch parameters.
template <int COUT, int KC_T, int TY_T, int RES_T>
__global__ void __launch_bounds__(THREADS, 1)
conv3d_bf16_fold_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const float* __restrict__ bias, bf16* __restrict__ out, double* __restrict__ gn_stats,
                        const FoldParams P) {
    constexpr int NF = 3 * COUT;
    constexpr int NCH = COUT / 16;           // 16-column chunks per kz block
    constexpr int CH_PER_WARP = (NCH + 1) / 2;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 5];
    __shared__ uint32_t tmem_base_slot;
    __shared__ __align__(16) float s_bias[COUT];

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t full_bar = ptx::smem_u32(&bars[0]);
    const uint32_t empty_bar = ptx::smem_u32(&bars[MAX_STAGES]);
    const uint32_t acc_full = ptx::smem_u32(&bars[2 * MAX_STAGES]);       // [2]
    const uint32_t acc_empty = ptx::smem_u32(&bars[2 * MAX_STAGES + 2]);  // [2]
    const uint32_t b_full = ptx::smem_u32(&bars[2 * MAX_STAGES + 4]);
    const int KC = KC_T > 0 ? KC_T : P.KC;
    const int TY = TY_T > 0 ? TY_T : P.TY;
    const bool resident = RES_T >= 0 ? (RES_T != 0) : (P.b_resident != 0);
    const int chunks = P.Cin / KC;
    const int n_bchunks = 9 * chunks;           // B chunks of KC channels
    const int k_iters = n_bchunks / TY;         // pipeline stages per tile, walked as (kx, [ky], ch)
    const uint32_t a_bytes = (uint32_t)(BM * KC * 2), b_bytes = (uint32_t)(NF * KC * 2);
    const uint32_t b_region = resident ? (uint32_t)n_bchunks * b_bytes : 0u;
    const uint32_t stage_bytes = (uint32_t)TY * (a_bytes + (resident ? 0u : b_bytes));
    const uint32_t stage_base = smem_base + b_region;

    for (int i = threadIdx.x; i < COUT; i += THREADS) s_bias[i] = bias ? bias[i] : 0.0f;
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_b);
        for (int s = 0; s < P.stages; ++s) {
            ptx::mbar_init(full_bar + 8 * s, 1);
            ptx::mbar_init(empty_bar + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(acc_full + 8 * s, 1);
            ptx::mbar_init(acc_empty + 8 * s, 8);  // one arrival per epilogue warp
        }
        ptx::mbar_init(b_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), (uint32_t)(2 * P.tmem_half));
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_d = tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer: the whole warp walks the (warp-uniform) loop, one elected lane issues =====
        // Resident weights: chunk (tap t9 = kx*3+ky, ch) lives in slot ((kx*chunks + ch)*3 + ky) for TY = 3 and
        // slot (t9*chunks + ch) for TY = 1, i.e. always slot = stage*TY + ty in the order the MMA warp walks.
        if (resident && ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(b_full, (uint32_t)n_bchunks * b_bytes);
            for (int t9 = 0; t9 < 9; ++t9)
                for (int ch = 0; ch < chunks; ++ch) {
                    const int slot = TY == 3 ? ((t9 / 3) * chunks + ch) * 3 + t9 % 3 : t9 * chunks + ch;
                    ptx::tma_load_3d(smem_base + slot * b_bytes, &map_b, b_full, ch * KC, 0, t9);
                }
        }
        __syncwarp();
        const int yz = P.Yp * P.Zp;
        const uint32_t tx = (uint32_t)TY * (a_bytes + (resident ? 0u : b_bytes));
        const int ny = 3 / TY;  // ky steps walked by the producer (1 when a stage holds all three)
        uint32_t s = 0, ph = 1;   // ring position and the parity of "slot is free"
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
            const int q0 = tile * ROWS_OUT - 1 + P.pad_rows;
            for (int kx = 0; kx < 3; ++kx) {
                for (int kyi = 0; kyi < ny; ++kyi) {
                    const int row = q0 + (kx - 1) * yz + (kyi - 1) * P.Zp;
                    for (int ch = 0; ch < chunks; ++ch) {
                        ptx::mbar_wait(empty_bar + 8 * s, ph);
                        if (ptx::elect_one()) {
                            const uint32_t a_dst = stage_base + s * stage_bytes;
                            ptx::mbar_arrive_expect_tx(full_bar + 8 * s, tx);
                            // {c, r, w, ky}: TY x (4 groups of 32 rows overlapping by two)
                            ptx::tma_load_4d(a_dst, &map_a, full_bar + 8 * s, ch * KC, row, 0, 0);
                            if (!resident)
                                ptx::tma_load_3d(a_dst + TY * a_bytes, &map_b, full_bar + 8 * s, ch * KC, 0, kx * 3 + kyi);
                        }
                        __syncwarp();
                        if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: warp-uniform flat loop over the stages of a tile, one elected lane issues =====
        const uint32_t idesc = ptx::umma_idesc_bf16(BM, (uint32_t)NF);
        const int kk = KC / 16;
        // descriptors are the template below plus a 14-bit start address (>> 4): all steps are additive
        const uint64_t desc0 = ptx::umma_smem_desc(0, (uint32_t)KC * 2u);
        const uint32_t a_step = a_bytes >> 4, b_step = b_bytes >> 4, st_step = stage_bytes >> 4;
        const uint64_t a_base = desc0 | (uint64_t)((stage_base & 0x3FFFFu) >> 4);
        const uint64_t b_base = desc0 | (uint64_t)(((resident ? smem_base : stage_base + TY * a_bytes) & 0x3FFFFu) >> 4);
        if (resident) {
            ptx::mbar_wait(b_full, 0);
            ptx::tc_fence_after();
        }
        uint32_t s = 0, ph = 0;
        int local = 0;
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++local) {
            const int as = local & 1;
            const uint32_t aph = (uint32_t)(local >> 1) & 1u;
            ptx::mbar_wait(acc_empty + 8 * as, aph ^ 1u);  // epilogue has drained this accumulator stage
            ptx::tc_fence_after();
            const uint32_t d_addr = tmem_d + (uint32_t)(as * P.tmem_half);
            for (int i = 0; i < k_iters; ++i) {
                ptx::mbar_wait(full_bar + 8 * s, ph);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint64_t a_st = a_base + (uint64_t)(s * st_step);
                    const uint64_t b_st = resident ? b_base + (uint64_t)((uint32_t)(i * TY) * b_step) : b_base + (uint64_t)(s * st_step);
#pragma unroll
                    for (int ty = 0; ty < (TY_T > 0 ? TY_T : 3); ++ty) {
                        if (ty < TY) {
#pragma unroll
                            for (int k = 0; k < (KC_T > 0 ? KC_T / 16 : 4); ++k) {
                                if (k < kk)
                                    ptx::umma_f16(d_addr, a_st + (uint64_t)(ty * a_step + 2 * k), b_st + (uint64_t)(ty * b_step + 2 * k), idesc,
                                                  (uint32_t)((i | ty | k) != 0));
                            }
                        }
                    }
                    ptx::umma_commit(empty_bar + 8 * s);
                }
                __syncwarp();
                if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
            }
            if (ptx::elect_one()) ptx::umma_commit(acc_full + 8 * as);
            __syncwarp();
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
        int st_b = -1;

        auto flush_stats = [&]() {
            // pairs -> groups (cpg even), warp reduce in double, one atomic per (warp, group, moment)
            const int cpg = COUT / P.G;
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
            }
        };

        int local = 0;
        for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x, ++local) {
            const int as = local & 1;
            const uint32_t aph = (uint32_t)(local >> 1) & 1u;
            const int64_t p = (int64_t)tile * ROWS_OUT + ROWS_WARP * lg - 1 + lane;
            int b = 0;
            const bool inter = lane >= 1 && lane <= ROWS_WARP && interior_row(p, P, b);
            const bool valid = P.all_rows ? (lane >= 1 && lane <= ROWS_WARP && p >= 0 && p < P.rows) : inter;
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
            for (int a = 0; a < CH_PER_WARP; ++a) {
                const int c = (2 * a + half) * 16;
                if (c < COUT) {
                    uint32_t r0[16], r1[16], r2[16];
                    ptx::tmem_ld_x16(t_row + (uint32_t)c, r0);
                    ptx::tmem_ld_x16(t_row + (uint32_t)(COUT + c), r1);
                    ptx::tmem_ld_x16(t_row + (uint32_t)(2 * COUT + c), r2);
                    ptx::tmem_ld_wait();
                    float v[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(r0[j]), 1);    // Y_0 of row i-1
                        const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[j]), 1);  // Y_2 of row i+1
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
            }
            // this warp has finished reading the accumulator stage
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(acc_empty + 8 * as);
        }
        if (do_stats && st_b >= 0) flush_stats();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_d, (uint32_t)(2 * P.tmem_half));
    }
}

int g_num_sms = 0;

template <int COUT, int KC_T, int TY_T, int RES_T>
int launch_fold_as(const CUtensorMap& map_a, const CUtensorMap& map_b, const float* bias, bf16* out, double* gn_stats,
                   const FoldParams& P, size_t smem, cudaStream_t stream) {
    auto kern = conv3d_bf16_fold_kernel<COUT, KC_T, TY_T, RES_T>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_conv3d_bf16_fold: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    const int grid = P.num_tiles < g_num_sms ? P.num_tiles : g_num_sms;
    kern<<<grid, THREADS, smem, stream>>>(map_a, map_b, bias, out, gn_stats, P);
    TDB_CHECK_LAUNCH("tdb_conv3d_bf16_fold");
    return 0;
}

// hot shapes of the shapes config get fully specialised kernels; everything else the generic one
template <int COUT>
int launch_fold(const CUtensorMap& map_a, const CUtensorMap& map_b, const float* bias, bf16* out, double* gn_stats,
                const FoldParams& P, size_t smem, cudaStream_t stream) {
    if (P.KC == 64 && P.TY == 1 && !P.b_resident)
        return launch_fold_as<COUT, 64, 1, 0>(map_a, map_b, bias, out, gn_stats, P, smem, stream);
    if (P.KC == 32 && P.TY == 3 && P.b_resident)
        return launch_fold_as<COUT, 32, 3, 1>(map_a, map_b, bias, out, gn_stats, P, smem, stream);
    return launch_fold_as<COUT, -1, -1, -1>(map_a, map_b, bias, out, gn_stats, P, smem, stream);
}

}  // namespace

extern "C" int tdb_conv3d_bf16_fold(const void* in, int ld_in, int pad_rows, const void* w_fold, const float* bias, void* out,
                                    int ld_out, int B, int X, int Y, int Z, int Cin, int Cout, double* gn_stats, int G,
                                    unsigned flags, void* stream) {
    TDB_REQUIRE(in && w_fold && out, TDB_E_BADARG, "tdb_conv3d_bf16_fold: null pointer");
    TDB_REQUIRE(Cin % 16 == 0 && (Cout == 16 || Cout == 32 || Cout == 64) && ld_in % 8 == 0 && ld_out % 8 == 0, TDB_E_UNSUPPORTED,
                "tdb_conv3d_bf16_fold: need Cin %% 16 == 0 and Cout in {16,32,64} (Cin=%d Cout=%d)", Cin, Cout);
    TDB_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)w_fold & 15) == 0, TDB_E_UNSUPPORTED,
                "tdb_conv3d_bf16_fold: pointers must be 16-byte aligned");
    TDB_REQUIRE(!gn_stats || (G >= 1 && Cout % G == 0 && (Cout / G) % 2 == 0), TDB_E_UNSUPPORTED,
                "tdb_conv3d_bf16_fold: fused GroupNorm moments need an even number of channels per group");
    Grid3 g(B, X, Y, Z);
    TDB_REQUIRE(g.rows + 2ll * pad_rows < (1ll << 31) - 4096, TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_fold: too many rows for 32-bit TMA coordinates");
    TDB_REQUIRE(pad_rows >= g.Yp * g.Zp + 2 * g.Zp + 256, TDB_E_BADARG,
                "tdb_conv3d_bf16_fold: pad_rows=%d, need >= Yp*Zp + 2*Zp + 256 = %d readable rows around the grid", pad_rows,
                g.Yp * g.Zp + 2 * g.Zp + 256);
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }

    FoldParams P;
    P.rows = g.rows;
    P.vox_p = (uint32_t)g.vox_p;
    P.Xp = g.Xp; P.Yp = g.Yp; P.Zp = g.Zp;
    P.by_vox = FastDiv((uint32_t)g.vox_p);
    P.by_z = FastDiv((uint32_t)g.Zp);
    P.by_y = FastDiv((uint32_t)g.Yp);
    P.Cin = Cin;
    P.KC = Cin % 64 == 0 ? 64 : (Cin % 32 == 0 ? 32 : 16);
    P.pad_rows = pad_rows;
    const int NF = 3 * Cout;
    P.a_bytes = BM * P.KC * 2;   // multiples of the swizzle atom (8 rows): tiles pack densely
    P.b_bytes = NF * P.KC * 2;
    const int n_bchunks = 9 * (Cin / P.KC);
    const int budget = 221 * 1024;  // dynamic smem we allow ourselves (227 KB max, minus static + alignment slack)
    P.b_resident = (int64_t)n_bchunks * P.b_bytes <= 112 * 1024 ? 1 : 0;
    const int resident_bytes = P.b_resident ? n_bchunks * P.b_bytes : 0;
    const int unit = P.a_bytes + (P.b_resident ? 0 : P.b_bytes);
    // three ky taps per stage (one TMA box, 3x fewer barrier round trips) when >= 4 such stages fit
    P.TY = (budget - resident_bytes) / (3 * unit) >= 4 ? 3 : 1;
    int stages = (budget - resident_bytes) / (P.TY * unit);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    TDB_REQUIRE(stages >= 2, TDB_E_UNSUPPORTED, "tdb_conv3d_bf16_fold: tiles do not fit in shared memory");
    P.stages = stages;
    int half = 32;
    while (half < NF) half *= 2;
    P.tmem_half = half;
    P.ld_out = ld_out;
    P.G = gn_stats ? G : 0;
    P.num_tiles = (int)ceil_div(g.rows, ROWS_OUT);
    P.all_rows = (flags & TDB_CONV_ALL_ROWS) ? 1 : 0;
    TDB_REQUIRE(!(P.all_rows && gn_stats), TDB_E_BADARG, "tdb_conv3d_bf16_fold: fused moments are not available with ALL_ROWS");

    CUtensorMap map_a, map_b;
    TDB_REQUIRE(encode_fn() != nullptr, TDB_E_NODEVICE, "tdb_conv3d_bf16_fold: cuTensorMapEncodeTiled unavailable (no driver)");
    {
        // A: [ky][w][r][c] overlapping view of the padded activation matrix (row 0 = grid row -pad_rows)
        const bf16* base = (const bf16*)in - (int64_t)pad_rows * ld_in;
        const uint64_t total_rows = (uint64_t)g.rows + 2ull * pad_rows;
        const uint64_t dims[4] = {(uint64_t)Cin, total_rows - 90 - 2ull * g.Zp, 4, 3};
        const uint64_t strides[3] = {(uint64_t)ld_in, 30ull * ld_in, (uint64_t)g.Zp * ld_in};
        const uint32_t box[4] = {(uint32_t)P.KC, 32, 4, (uint32_t)P.TY};
        TDB_REQUIRE(make_map_bf16(&map_a, base, 4, dims, strides, box), TDB_E_BADARG,
                    "tdb_conv3d_bf16_fold: tensor map (activations) rejected");
    }
    {
        // B: [tap9][n'][ci] view of the folded weights [3*Cout][9*Cin]
        const uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)NF, 9};
        const uint64_t strides[2] = {9ull * Cin, (uint64_t)Cin};
        const uint32_t box[3] = {(uint32_t)P.KC, (uint32_t)NF, (uint32_t)(P.b_resident ? 1 : P.TY)};
        TDB_REQUIRE(make_map_bf16(&map_b, w_fold, 3, dims, strides, box), TDB_E_BADARG,
                    "tdb_conv3d_bf16_fold: tensor map (weights) rejected");
    }

    const size_t smem = (size_t)resident_bytes + (size_t)stages * P.TY * unit + 1024;
    cudaStream_t s = (cudaStream_t)stream;
    if (Cout == 16) return launch_fold<16>(map_a, map_b, bias, (bf16*)out, gn_stats, P, smem, s);
    if (Cout == 32) return launch_fold<32>(map_a, map_b, bias, (bf16*)out, gn_stats, P, smem, s);
    return launch_fold<64>(map_a, map_b, bias, (bf16*)out, gn_stats, P, smem, s);
}
