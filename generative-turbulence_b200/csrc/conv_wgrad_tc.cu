// Weight gradient of the 3x3x3 / 1x1x1 convolutions on the 5th-generation tensor cores (bf16 path).
//
//   dW[tap][ci][co] = sum_{rows r} X[r + delta(tap)][ci] * dY[r][co]        (dY is ZERO on halo rows)
//
// replaces torch.autograd's conv weight gradient behind nn.Conv3d (reference ddpm.py:164,188 through
// loss.backward(), ddpm.py:874-882).  Both operands are the halo grids themselves, [rows][channels] with the
// channels contiguous: for the GEMM  D[M = (tap, ci)][N = co] += A[M][K = rows] * B[N][K]^T  that is the
// "MN-major" operand form of tcgen05.mma, so the TMA tiles (128-byte swizzled [rows][64 ch] boxes, the same
// boxes the forward convolution loads) feed the tensor core directly - no transpose pass.
//
//   * one M block (128 TMEM lanes) = 128/KC "windows" (tap, 64- or 32-channel chunk of Cin) stacked along M
//     through the descriptor's leading-dimension stride; N = a Cout chunk (<= 256); fp32 accumulators of up to
//     512/N M-blocks stay in TMEM for the whole row range of the CTA;
//   * kz sharing: the three kz taps of one (kx, ky) read the same rows shifted by one, so a window of R+8 rows is
//     loaded once and used three times by starting the matrix descriptor 0/1/2 rows into it (the swizzle is a
//     function of the absolute shared-memory address, so the descriptor keeps base offset 0) - 3x less ingest;
//   * split over the row range (grid.x) and over (M-block group, Cout chunk) types (grid.y); every CTA adds its
//     accumulators into dW with fp32 vector reductions.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM owner), 2..5 = epilogue.
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"
#include <cstdlib>

using namespace tdb;
using bf16 = __nv_bfloat16;

namespace {

constexpr int WG_THREADS = 192;
constexpr int WG_MAX_STAGES = 8;
constexpr int WG_MAX_TILES = 32;   // X windows per stage
constexpr int WG_MAX_MB = 16;      // M-blocks per CTA

struct WgParams {
    float* dw;
    int rows, Cin, Cout, ntaps;
    int yz_p, z_p;
    int n_ci;         // Cin / KC
    int n_win;        // windows: (shared ? ntaps/3 : ntaps) * n_ci   (ntaps == 1: n_ci)
    int n_shift;      // 3: kz shared inside a window, 1: one window per tap
    int n_mblocks;    // ceil(n_win / UPB) * n_shift
    int mb_per_cta;
    int n_slots;      // window groups resident per stage (upper bound over CTA types)
    int nc, n_cc;     // Cout chunk per MMA / number of chunks
    int R, win_rows;  // rows per stage / rows per X window (R or R + 8)
    int rows_per_split;
    int stages;
    int tmem_cols;
};

// MN-major shared-memory matrix descriptor: `lbo` = byte stride between swizzle atoms along M/N,
// `sbo` = byte stride between 8-row groups along K, `row_bytes` = swizzle span; matrix base offset 0.
__device__ __forceinline__ uint64_t mn_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : 4ull;
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= layout << 61;
    return d;
}

template <int KC>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, const WgParams P) {
    constexpr int UPB = 128 / KC;           // windows stacked along M in one M-block
    constexpr uint32_t XROW = KC * 2;       // bytes per window row = swizzle span of A
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bars[2 * WG_MAX_STAGES + 1];
    __shared__ uint32_t tmem_base_slot;
    __shared__ int s_col[WG_MAX_TILES], s_delta[WG_MAX_TILES];  // TMA coordinates of the X windows of this CTA
    __shared__ __align__(8) uint64_t s_adesc[WG_MAX_MB];        // per M-block: A descriptor in stage 0

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int cc = blockIdx.y % P.n_cc, tg = blockIdx.y / P.n_cc;
    const int mb0 = tg * P.mb_per_cta, mb1 = min(P.n_mblocks, mb0 + P.mb_per_cta), nmb = mb1 - mb0;
    const int wg0 = mb0 / P.n_shift, wg1 = (mb1 - 1) / P.n_shift, nslots = wg1 - wg0 + 1;
    const int ntiles = nslots * UPB;
    const uint32_t win_bytes = (uint32_t)P.win_rows * XROW;
    const uint32_t dy_row = P.nc >= 64 ? 128u : 64u;           // swizzle span of B
    const uint32_t dy_sub = (uint32_t)P.R * dy_row;            // one [R][64 | 32] box
    const uint32_t n_dy = P.nc >= 64 ? (uint32_t)P.nc / 64u : 1u;
    const uint32_t x_bytes = (uint32_t)P.n_slots * UPB * win_bytes;
    const uint32_t stage_bytes = x_bytes + n_dy * dy_sub;
    const uint32_t full_bar = ptx::smem_u32(&bars[0]);
    const uint32_t empty_bar = ptx::smem_u32(&bars[WG_MAX_STAGES]);
    const uint32_t done_bar = ptx::smem_u32(&bars[2 * WG_MAX_STAGES]);
    const int r_begin = blockIdx.x * P.rows_per_split;
    const int r_end = min(P.rows, r_begin + P.rows_per_split);
    const int n_iter = r_end > r_begin ? (r_end - r_begin + P.R - 1) / P.R : 0;

    if (threadIdx.x < ntiles) {
        const int j = threadIdx.x;
        const int w = min(wg0 * UPB + j, P.n_win - 1);  // windows past the end repeat the last one (discarded later)
        const int txy = w / P.n_ci, ch = w % P.n_ci;
        int delta = 0;
        if (P.ntaps == 27) {
            if (P.n_shift == 3) delta = (txy / 3 - 1) * P.yz_p + (txy % 3 - 1) * P.z_p - 1;
            else delta = (txy / 9 - 1) * P.yz_p + ((txy / 3) % 3 - 1) * P.z_p + (txy % 3 - 1);
        }
        s_col[j] = ch * KC;
        s_delta[j] = delta;
    }
    if (threadIdx.x < nmb) {
        const int mbi = mb0 + threadIdx.x;
        const int wg = mbi / P.n_shift, sh = mbi % P.n_shift;
        // kz shift = start the matrix `sh` rows into the window.  The swizzle is a function of the absolute
        // shared-memory address (measured on B200: a non-zero matrix base offset gives wrong results, zero is exact).
        const uint32_t a_off = (uint32_t)(wg - wg0) * UPB * win_bytes + (uint32_t)sh * XROW;
        s_adesc[threadIdx.x] = mn_desc(smem_base + a_off, win_bytes, 8u * XROW, XROW);
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_x);
        ptx::prefetch_tensormap(&map_dy);
        for (int s = 0; s < P.stages; ++s) {
            ptx::mbar_init(full_bar + 8 * s, 1);
            ptx::mbar_init(empty_bar + 8 * s, 1);
        }
        ptx::mbar_init(done_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), (uint32_t)P.tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_d = tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        const uint32_t tx = (uint32_t)ntiles * win_bytes + n_dy * dy_sub;
        uint32_t s = 0, ph = 1;
        for (int it = 0; it < n_iter; ++it) {
            const int r0 = r_begin + it * P.R;
            ptx::mbar_wait(empty_bar + 8 * s, ph);
            if (ptx::elect_one()) {
                const uint32_t dst = smem_base + s * stage_bytes;
                ptx::mbar_arrive_expect_tx(full_bar + 8 * s, tx);
                for (int j = 0; j < ntiles; ++j)
                    ptx::tma_load_2d(dst + (uint32_t)j * win_bytes, &map_x, full_bar + 8 * s, s_col[j], r0 + s_delta[j]);
                for (uint32_t q = 0; q < n_dy; ++q)
                    ptx::tma_load_2d(dst + x_bytes + q * dy_sub, &map_dy, full_bar + 8 * s, cc * P.nc + (int)q * 64, r0);
            }
            __syncwarp();
            if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // instruction descriptor: D = f32, A = B = bf16, both MN-major (bits 15, 16), M = 128, N = nc
        const uint32_t idesc = ptx::umma_idesc_bf16(128, (uint32_t)P.nc) | (1u << 15) | (1u << 16);
        const uint32_t sbo_b = 8u * dy_row;
        const uint64_t b_desc0 = mn_desc(smem_base + x_bytes, dy_sub, sbo_b, dy_row);
        const uint64_t ka = (uint64_t)((2u * 8u * XROW) >> 4), kb = (uint64_t)((2u * sbo_b) >> 4);  // 16 rows along K
        const uint32_t stage16 = stage_bytes >> 4;
        uint32_t s = 0, ph = 0;
        for (int it = 0; it < n_iter; ++it) {
            ptx::mbar_wait(full_bar + 8 * s, ph);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint64_t so = (uint64_t)(s * stage16);
                const uint64_t bd = b_desc0 + so;
                const uint32_t acc = it > 0 ? 1u : 0u;
                uint32_t d = tmem_d;
                const int ksteps = P.R / 16;
                for (int m = 0; m < nmb; ++m, d += (uint32_t)P.nc) {
                    const uint64_t ad = s_adesc[m] + so;
                    ptx::umma_f16(d, ad, bd, idesc, acc);
                    for (int k = 1; k < ksteps; ++k) ptx::umma_f16(d, ad + (uint64_t)k * ka, bd + (uint64_t)k * kb, idesc, 1u);
                }
                ptx::umma_commit(empty_bar + 8 * s);
            }
            __syncwarp();
            if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
        }
        if (ptx::elect_one()) ptx::umma_commit(done_bar);
        __syncwarp();
    } else if (n_iter > 0) {
        // ===== epilogue: TMEM -> fp32 vector reductions into dW =====
        const int lg = warp % 4;
        ptx::mbar_wait(done_bar, 0);
        ptx::tc_fence_after();
        const int m_lane = lg * 32 + lane;              // row of the M-block
        const int j = m_lane / KC, ci_l = m_lane % KC;  // window inside the block, channel inside the window
        for (int m = 0; m < nmb; ++m) {
            const int mbi = mb0 + m;
            const int wg = mbi / P.n_shift, sh = mbi % P.n_shift;
            const int w = wg * UPB + j;
            const bool valid = w < P.n_win;
            const int wc = valid ? w : 0;
            const int txy = wc / P.n_ci, ch = wc % P.n_ci;
            const int tap = P.n_shift == 3 ? txy * 3 + sh : txy;
            float* drow = P.dw + ((int64_t)tap * P.Cin + ch * KC + ci_l) * P.Cout + cc * P.nc;
            const uint32_t t_row = tmem_d + ((uint32_t)(lg * 32) << 16) + (uint32_t)(m * P.nc);
            for (int c = 0; c < P.nc; c += 16) {
                uint32_t r[16];
                ptx::tmem_ld_x16(t_row + (uint32_t)c, r);
                ptx::tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int q = 0; q < 16; q += 4)
                        atomicAdd(reinterpret_cast<float4*>(drow + c + q),
                                  make_float4(__uint_as_float(r[q]), __uint_as_float(r[q + 1]), __uint_as_float(r[q + 2]),
                                              __uint_as_float(r[q + 3])));
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_d, (uint32_t)P.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// kz on N ("kzn" form, Cout in {32, 64}): with N = Cout an MMA reads 4 KB of A and only Cout*32 bytes of B from shared
// memory for Cout/2 cycles of math - the narrow full-resolution layers were bound by shared-memory reads and by the TMA
// ingest of their many activation windows (0.23 - 0.57 of the tensor peak).  Here the three kz taps sit on the N side:
//     dW[(kx,ky), kz][ci][co] = sum_r' X[r' + delta(kx,ky)][ci] * dY[r' - (kz - 1)][co]
// so one activation window per (kx, ky, channel chunk) serves all three kz, and the B operand is ONE window of the output
// gradient (rows r0 - 1 .. r0 + R + 1) seen as three swizzle atoms that start one row apart: the MN-major descriptor's
// leading-dimension stride is a single row (the swizzle is a function of the absolute shared-memory address, as with
// the row-shifted A views above).  N = 3*Cout (96 / 192): 3x the math per activation byte staged and read.
struct WgKznParams {
    float* dw;
    int rows, Cin, Cout;
    int yz_p, z_p;
    int n_ci, n_win;       // windows = 9 * n_ci, stacked UPB per M-block
    int n_mblocks;
    int R, win_rows_b;     // rows per stage; rows of the dY window (R + 8)
    int stages;
    int n_types;
    int type_cta0[17];     // first CTA of each type (prefix sums)
    int type_mb0[17];      // first M-block of each type
    int type_rps[16];      // rows per CTA of each type
    int mb_max;
};

template <int KC>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_kzn_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy, const WgKznParams P) {
    constexpr int UPB = 128 / KC;
    constexpr uint32_t XROW = KC * 2;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bars[2 * WG_MAX_STAGES + 1];
    __shared__ uint32_t tmem_base_slot;
    __shared__ int s_col[WG_MAX_TILES], s_delta[WG_MAX_TILES];

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    int type = 0;
    while (type + 1 < P.n_types && (int)blockIdx.x >= P.type_cta0[type + 1]) ++type;
    const int split = (int)blockIdx.x - P.type_cta0[type];
    const int mb0 = P.type_mb0[type], nmb = P.type_mb0[type + 1] - mb0;
    const int ntiles = nmb * UPB;
    const int N = 3 * P.Cout;
    const uint32_t win_bytes = (uint32_t)P.R * XROW;
    const uint32_t dy_row = (uint32_t)P.Cout * 2u;            // 64 or 128 bytes: swizzle span of B
    const uint32_t x_bytes = (uint32_t)P.mb_max * UPB * win_bytes;
    const uint32_t stage_bytes = x_bytes + (uint32_t)P.win_rows_b * dy_row;
    const uint32_t full_bar = ptx::smem_u32(&bars[0]);
    const uint32_t empty_bar = ptx::smem_u32(&bars[WG_MAX_STAGES]);
    const uint32_t done_bar = ptx::smem_u32(&bars[2 * WG_MAX_STAGES]);
    const int rps = P.type_rps[type];
    // r' runs over [-1, rows + 1): the dY rows one before / after a stage's range belong to the kz = 2 / kz = 0 taps
    const int r_begin = split * rps - P.R, r_stop = min(P.rows + P.R, r_begin + rps);
    const int n_iter = r_stop > r_begin ? (r_stop - r_begin + P.R - 1) / P.R : 0;

    if (threadIdx.x < ntiles) {
        const int j = threadIdx.x;
        const int w = min(mb0 * UPB + j, P.n_win - 1);  // windows past the end repeat the last one (discarded later)
        const int txy = w / P.n_ci, ch = w % P.n_ci;
        s_col[j] = ch * KC;
        s_delta[j] = (txy / 3 - 1) * P.yz_p + (txy % 3 - 1) * P.z_p;
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_x);
        ptx::prefetch_tensormap(&map_dy);
        for (int s = 0; s < P.stages; ++s) {
            ptx::mbar_init(full_bar + 8 * s, 1);
            ptx::mbar_init(empty_bar + 8 * s, 1);
        }
        ptx::mbar_init(done_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), 512u);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_d = tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        const uint32_t tx = (uint32_t)ntiles * win_bytes + (uint32_t)P.win_rows_b * dy_row;
        uint32_t s = 0, ph = 1;
        for (int it = 0; it < n_iter; ++it) {
            const int r0 = r_begin + it * P.R;
            ptx::mbar_wait(empty_bar + 8 * s, ph);
            if (ptx::elect_one()) {
                const uint32_t dst = smem_base + s * stage_bytes;
                ptx::mbar_arrive_expect_tx(full_bar + 8 * s, tx);
                for (int j = 0; j < ntiles; ++j)
                    ptx::tma_load_2d(dst + (uint32_t)j * win_bytes, &map_x, full_bar + 8 * s, s_col[j], r0 + s_delta[j]);
                ptx::tma_load_2d(dst + x_bytes, &map_dy, full_bar + 8 * s, 0, r0 - 1);
            }
            __syncwarp();
            if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: per stage and M-block two K = 16 MMAs of 128 x 3*Cout =====
        const uint32_t idesc = ptx::umma_idesc_bf16(128, (uint32_t)N) | (1u << 15) | (1u << 16);
        const uint32_t sbo_b = 8u * dy_row;
        // B: three atoms (kz = 2, 1, 0) one ROW apart inside the dY window
        const uint64_t b_desc0 = mn_desc(smem_base + x_bytes, dy_row, sbo_b, dy_row);
        const uint64_t a_desc0 = mn_desc(smem_base, win_bytes, 8u * XROW, XROW);
        const uint64_t ka = (uint64_t)((2u * 8u * XROW) >> 4), kb = (uint64_t)((2u * sbo_b) >> 4);  // 16 rows along K
        const uint64_t a_mb = (uint64_t)((UPB * win_bytes) >> 4);
        const uint32_t stage16 = stage_bytes >> 4;
        const int ksteps = P.R / 16;
        uint32_t s = 0, ph = 0;
        for (int it = 0; it < n_iter; ++it) {
            ptx::mbar_wait(full_bar + 8 * s, ph);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint64_t so = (uint64_t)(s * stage16);
                const uint64_t bd = b_desc0 + so;
                uint32_t d = tmem_d;
                for (int m = 0; m < nmb; ++m, d += (uint32_t)N) {
                    const uint64_t ad = a_desc0 + so + (uint64_t)m * a_mb;
                    for (int k = 0; k < ksteps; ++k)
                        ptx::umma_f16(d, ad + (uint64_t)k * ka, bd + (uint64_t)k * kb, idesc, (it | k) ? 1u : 0u);
                }
                ptx::umma_commit(empty_bar + 8 * s);
            }
            __syncwarp();
            if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
        }
        if (ptx::elect_one()) ptx::umma_commit(done_bar);
        __syncwarp();
    } else if (n_iter > 0) {
        // ===== epilogue: TMEM -> fp32 vector reductions into dW; column block j of an accumulator is kz = 2 - j =====
        const int lg = warp % 4;
        ptx::mbar_wait(done_bar, 0);
        ptx::tc_fence_after();
        const int m_lane = lg * 32 + lane;
        const int j = m_lane / KC, ci_l = m_lane % KC;
        for (int m = 0; m < nmb; ++m) {
            const int w = (mb0 + m) * UPB + j;
            const bool valid = w < P.n_win;
            const int wc = valid ? w : 0;
            const int txy = wc / P.n_ci, ch = wc % P.n_ci;
            const uint32_t t_row = tmem_d + ((uint32_t)(lg * 32) << 16) + (uint32_t)(m * N);
            for (int kb3 = 0; kb3 < 3; ++kb3) {
                const int tap = txy * 3 + (2 - kb3);
                float* drow = P.dw + ((int64_t)tap * P.Cin + ch * KC + ci_l) * P.Cout;
                for (int c = 0; c < P.Cout; c += 16) {
                    uint32_t r[16];
                    ptx::tmem_ld_x16(t_row + (uint32_t)(kb3 * P.Cout + c), r);
                    ptx::tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int q = 0; q < 16; q += 4)
                            atomicAdd(reinterpret_cast<float4*>(drow + c + q),
                                      make_float4(__uint_as_float(r[q]), __uint_as_float(r[q + 1]), __uint_as_float(r[q + 2]),
                                                  __uint_as_float(r[q + 3])));
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_d, 512u);
    }
}

int launch_wgrad_kzn(const void* in, int ld_in, const void* d_out, int ld_do, float* dw, const Grid3& g, int Cin, int Cout,
                     cudaStream_t stream) {
    const int KC = Cin % 64 == 0 ? 64 : 32;
    const int UPB = 128 / KC;
    WgKznParams P;
    P.dw = dw;
    P.rows = (int)g.rows;
    P.Cin = Cin; P.Cout = Cout;
    P.yz_p = g.Yp * g.Zp; P.z_p = g.Zp;
    P.n_ci = Cin / KC;
    P.n_win = 9 * P.n_ci;
    P.n_mblocks = (int)ceil_div(P.n_win, UPB);
    // Stage geometry (measured on B200, profiles/r02_wgrad_kzn_sweep.txt): a stage costs ~0.6 us whatever it holds, so the
    // stages are as tall as shared memory allows with >= 3 of them; all CTA types hold the same number of M-blocks and
    // sweep the rows in step (the windows then come from DRAM once and from L2 afterwards).
    const int n_mblocks = (int)ceil_div(9 * (Cin / KC), UPB);
    const bool many = n_mblocks >= 16;  // 256 -> 64: two M-blocks per CTA halve the dY re-reads
    P.R = getenv("TDB_WGRAD_R") ? atoi(getenv("TDB_WGRAD_R")) : (many ? 64 : (Cout == 32 ? 192 : 128));
    P.win_rows_b = Cout == 32 ? P.R + 16 : P.R + 8;  // R + 2 rows are used; the window stays a multiple of 1024 bytes
    const int N = 3 * Cout;
    int mb = many ? 2 : 1;
    if (getenv("TDB_WGRAD_MB") && atoi(getenv("TDB_WGRAD_MB")) <= 512 / N) mb = atoi(getenv("TDB_WGRAD_MB"));
    const int win_bytes = P.R * KC * 2, dyw_bytes = P.win_rows_b * Cout * 2;
    while (mb > 1 && (mb * UPB > WG_MAX_TILES || 3 * (mb * UPB * win_bytes + dyw_bytes) > 200 * 1024)) --mb;
    if (mb > P.n_mblocks) mb = P.n_mblocks;
    P.n_types = (int)ceil_div(P.n_mblocks, mb);
    if (P.n_types > 16) return -1;  // caller falls back to the M-stacked form
    mb = (int)ceil_div(P.n_mblocks, P.n_types);
    P.mb_max = mb;
    const int stage_bytes = mb * UPB * win_bytes + dyw_bytes;
    int stages = (200 * 1024) / stage_bytes;
    if (stages > WG_MAX_STAGES) stages = WG_MAX_STAGES;
    if (stages < 2) return -1;
    P.stages = stages;
    // M-blocks per type (balanced), CTAs per type in proportion to its M-blocks, each CTA an equal share of the rows
    int base = P.n_mblocks / P.n_types, rem = P.n_mblocks % P.n_types;
    int ctas_total = 148 < P.n_types ? P.n_types : 148;
    P.type_mb0[0] = 0;
    P.type_cta0[0] = 0;
    const int64_t span = g.rows + 2 * P.R;  // r' in [-R, rows + R)
    for (int t = 0; t < P.n_types; ++t) {
        const int nb = base + (t < rem ? 1 : 0);
        P.type_mb0[t + 1] = P.type_mb0[t] + nb;
        int64_t ctas = ((int64_t)ctas_total * nb) / P.n_mblocks;
        if (ctas < 1) ctas = 1;
        int64_t rps = ceil_div(ceil_div(span, ctas), P.R) * P.R;
        if (rps < 4 * P.R) rps = 4 * P.R;
        ctas = ceil_div(span, rps);
        P.type_rps[t] = (int)rps;
        P.type_cta0[t + 1] = P.type_cta0[t] + (int)ctas;
    }
    CUtensorMap map_x, map_dy;
    if (encode_fn() == nullptr) return -1;
    if (!make_map_2d_bf16(&map_x, in, (uint64_t)Cin, (uint64_t)g.rows, (uint64_t)ld_in, (uint32_t)KC, (uint32_t)P.R)) return -1;
    if (!make_map_2d_bf16(&map_dy, d_out, (uint64_t)Cout, (uint64_t)g.rows, (uint64_t)ld_do, (uint32_t)Cout, (uint32_t)P.win_rows_b))
        return -1;
    const size_t smem = (size_t)stages * stage_bytes + 1024;
    auto kern = KC == 64 ? conv_wgrad_kzn_kernel<64> : conv_wgrad_kzn_kernel<32>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    kern<<<(unsigned)P.type_cta0[P.n_types], WG_THREADS, smem, stream>>>(map_x, map_dy, P);
    return 0;
}

}  // namespace

// mode: bit 0 (TDB_WGRAD_SHARE_KZ) = share one row window between the three kz taps (row-shifted descriptors);
// bit 1 (TDB_WGRAD_KZ_ON_N) = Cout in {32, 64}: the three kz taps on the N side (one dY window, atoms one row apart).
extern "C" int tdb_conv3d_wgrad_tc(const void* in, int ld_in, const void* d_out, int ld_do, float* dw, int B, int X, int Y, int Z,
                                   int Cin, int Cout, int ntaps, unsigned mode, void* stream) {
    TDB_REQUIRE(in && d_out && dw, TDB_E_BADARG, "tdb_conv3d_wgrad_tc: null pointer");
    TDB_REQUIRE(ntaps == 1 || ntaps == 27, TDB_E_BADARG, "tdb_conv3d_wgrad_tc: ntaps must be 1 or 27");
    TDB_REQUIRE(Cin % 32 == 0 && Cout % 32 == 0 && (Cout <= 256 ? (Cout == 32 || Cout % 64 == 0) : Cout % 256 == 0) &&
                    ld_in % 8 == 0 && ld_do % 8 == 0,
                TDB_E_UNSUPPORTED, "tdb_conv3d_wgrad_tc: unsupported channels Cin=%d Cout=%d", Cin, Cout);
    TDB_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)d_out & 15) == 0 && ((uintptr_t)dw & 15) == 0, TDB_E_UNSUPPORTED,
                "tdb_conv3d_wgrad_tc: pointers must be 16-byte aligned");
    Grid3 g(B, X, Y, Z);
    TDB_REQUIRE(g.rows < (1ll << 31) - 65536, TDB_E_UNSUPPORTED, "tdb_conv3d_wgrad_tc: too many rows for 32-bit TMA coordinates");
    if ((mode & TDB_WGRAD_KZ_ON_N) && ntaps == 27 && (Cout == 32 || Cout == 64)) {
        if (launch_wgrad_kzn(in, ld_in, d_out, ld_do, dw, g, Cin, Cout, (cudaStream_t)stream) == 0) {
            TDB_CHECK_LAUNCH("tdb_conv3d_wgrad_tc");
            return 0;
        }
    }
    const int KC = Cin % 64 == 0 ? 64 : 32;
    const int UPB = 128 / KC;
    WgParams P;
    P.dw = dw;
    P.rows = (int)g.rows;
    P.Cin = Cin; P.Cout = Cout; P.ntaps = ntaps;
    P.yz_p = g.Yp * g.Zp; P.z_p = g.Zp;
    P.n_ci = Cin / KC;
    const bool shared = ntaps == 27 && (mode & 1u);
    P.n_shift = shared ? 3 : 1;
    P.n_win = (shared ? 9 : ntaps) * P.n_ci;
    P.n_mblocks = (int)ceil_div(P.n_win, UPB) * P.n_shift;
    P.nc = Cout <= 256 ? Cout : 256;
    P.n_cc = Cout / P.nc;
    // rows per stage (R / 16 MMAs of K = 16 per M-block): measured on B200, 64-row stages beat 32-row ones by 10 - 30 % on
    // the wide layers (every stage costs ~0.6 us whatever it holds), 96 / 128 rows leave too few stages in shared memory
    P.R = getenv("TDB_WGRAD_R1") ? atoi(getenv("TDB_WGRAD_R1")) : 64;
    P.win_rows = shared ? P.R + 8 : P.R;
    const int win_bytes = P.win_rows * KC * 2;
    const int dy_bytes = P.R * P.nc * 2;
    // M-blocks per CTA: TMEM columns (512), the descriptor tables, and >= 3 stages in ~200 KB of shared memory
    int mb = 512 / P.nc;
    if (mb > WG_MAX_MB) mb = WG_MAX_MB;
    if (mb > P.n_mblocks) mb = P.n_mblocks;
    auto slots_for = [&](int m) { return (m + P.n_shift - 2) / P.n_shift + 1; };  // worst-case window groups spanned
    auto stage_for = [&](int m) { return slots_for(m) * UPB * win_bytes + dy_bytes; };
    while (mb > 1 && (slots_for(mb) * UPB > WG_MAX_TILES || 3 * stage_for(mb) > 200 * 1024)) --mb;
    // enough CTA types x row splits for one wave
    const int n_groups = (int)ceil_div(P.n_mblocks, mb);
    mb = (int)ceil_div(P.n_mblocks, n_groups);  // balance the groups
    P.mb_per_cta = mb;
    P.n_slots = P.n_shift == 1 ? mb : slots_for(mb);
    const int stage_bytes = P.n_slots * UPB * win_bytes + dy_bytes;
    int stages = (200 * 1024) / stage_bytes;
    if (stages > WG_MAX_STAGES) stages = WG_MAX_STAGES;
    TDB_REQUIRE(stages >= 2, TDB_E_UNSUPPORTED, "tdb_conv3d_wgrad_tc: stage of %d bytes does not fit", stage_bytes);
    P.stages = stages;
    int cols = 32;
    while (cols < mb * P.nc) cols *= 2;
    P.tmem_cols = cols;
    const int types = n_groups * P.n_cc;
    int64_t splits = 148 / types;
    if (splits < 1) splits = 1;
    int64_t rps = ceil_div(ceil_div(g.rows, splits), P.R) * P.R;
    if (rps < 4 * P.R) rps = 4 * P.R;
    P.rows_per_split = (int)rps;
    splits = ceil_div(g.rows, rps);

    CUtensorMap map_x, map_dy;
    TDB_REQUIRE(encode_fn() != nullptr, TDB_E_NODEVICE, "tdb_conv3d_wgrad_tc: cuTensorMapEncodeTiled unavailable (no driver)");
    TDB_REQUIRE(make_map_2d_bf16(&map_x, in, (uint64_t)Cin, (uint64_t)g.rows, (uint64_t)ld_in, (uint32_t)KC, (uint32_t)P.win_rows),
                TDB_E_BADARG, "tdb_conv3d_wgrad_tc: tensor map (activations) rejected");
    TDB_REQUIRE(make_map_2d_bf16(&map_dy, d_out, (uint64_t)Cout, (uint64_t)g.rows, (uint64_t)ld_do, (uint32_t)(P.nc >= 64 ? 64 : 32),
                                 (uint32_t)P.R),
                TDB_E_BADARG, "tdb_conv3d_wgrad_tc: tensor map (output gradient) rejected");
    const size_t smem = (size_t)stages * stage_bytes + 1024;
    auto kern = KC == 64 ? conv_wgrad_tc_kernel<64> : conv_wgrad_tc_kernel<32>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_conv3d_wgrad_tc: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    dim3 grid((unsigned)splits, (unsigned)types);
    kern<<<grid, WG_THREADS, smem, (cudaStream_t)stream>>>(map_x, map_dy, P);
    TDB_CHECK_LAUNCH("tdb_conv3d_wgrad_tc");
    return 0;
}
