// bf16 implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// GEMM view (see conv_f32.cu for the row-shift identity over the halo grid):
//     M = rows of the haloed, linearised voxel grid (128-row tiles)
//     N = Cout (tile BN <= 256)          K = ntaps * Cin, walked as (tap, KC-channel chunk)
//     A tile (tap, chunk) = rows [p0 + delta(tap), +128) x channels [c0, c0+KC) of the input
//                           = ONE 2-D TMA box of the [rows][ld] activation matrix (rows that
//                             fall outside the buffer are zero-filled by TMA; they only feed
//                             halo outputs, which are never stored);
//     B tile             = rows [n0, n0+BN) x k [tap*Cin + c0, +KC) of the packed weights.
// Both tiles land in shared memory K-major with the hardware swizzle whose span equals the row
// (KC*2 bytes in {32,64,128}), which is exactly the canonical UMMA K-major layout, so the MMA
// warp only builds descriptors.  Accumulators live in TMEM (128 lanes x BN fp32 columns).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one
// elected lane), warps 2-5 = epilogue (TMEM lane group = warp % 4): tcgen05.ld -> +bias ->
// optional GroupNorm partial moments (from the fp32 accumulators) -> bf16 -> global.
// Two CTAs fit per SM for the narrow layers, so one CTA's epilogue overlaps the other's MMAs.
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

using namespace tdb;
using bf16 = __nv_bfloat16;

namespace {

constexpr int BM = 128;
constexpr int THREADS = 192;
constexpr int MAX_STAGES = 8;

struct ConvParams {
    int64_t rows;       // B * vox_p
    int64_t vox_p;
    int Xp, Yp, Zp;
    int Cin, Cout, ntaps;
    int KC;             // channels per K chunk (16/32/64)
    int BN;             // N tile
    int stages;
    int a_bytes, b_bytes;  // per-stage tile sizes (1024-aligned)
    int tmem_cols;
    int ld_out;
    int G;              // groups for fused GroupNorm moments (0 = off)
    int all_rows;       // 1: store halo rows too (input-gradient use)
    int mc;             // 1: launched as 2-CTA clusters along M, the weight tile is fetched once per cluster (TMA multicast)
    int splits;         // split-K factor (gridDim.z); > 1: fp32 partial sums are atomically added to `scratch`
};

__device__ __forceinline__ bool row_is_interior(int64_t p, const ConvParams& P, int& b) {
    b = (int)(p / P.vox_p);
    int64_t r = p - (int64_t)b * P.vox_p;
    const int zp = (int)(r % P.Zp);
    r /= P.Zp;
    const int yp = (int)(r % P.Yp);
    const int xp = (int)(r / P.Yp);
    return xp >= 1 && xp <= P.Xp - 2 && yp >= 1 && yp <= P.Yp - 2 && zp >= 1 && zp <= P.Zp - 2;
}

__global__ void __launch_bounds__(THREADS)
conv3d_bf16_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const float* __restrict__ bias, bf16* __restrict__ out, double* __restrict__ gn_stats,
                      float* __restrict__ scratch, const ConvParams P) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment (swizzle-128B atoms) is established by hand; the launcher over-allocates.
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bars[2 * MAX_STAGES + 1];
    __shared__ uint32_t tmem_base_slot;

    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int64_t p0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * P.BN;
    const uint32_t stage_bytes = (uint32_t)(P.a_bytes + P.b_bytes);
    const uint32_t full_bar = ptx::smem_u32(&bars[0]);
    const uint32_t empty_bar = ptx::smem_u32(&bars[MAX_STAGES]);
    const uint32_t accum_bar = ptx::smem_u32(&bars[2 * MAX_STAGES]);
    const int chunks = P.Cin / P.KC;
    // split-K: this CTA walks k-steps [it_begin, it_end) of the ntaps*chunks (tap, channel-chunk) sequence
    const int k_total = P.ntaps * chunks;
    const int k_per = (k_total + P.splits - 1) / P.splits;
    const int it_begin = (int)blockIdx.z * k_per;
    const int it_end = min(k_total, it_begin + k_per);
    const int k_iters = it_end - it_begin;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&map_a);
        ptx::prefetch_tensormap(&map_b);
        for (int s = 0; s < P.stages; ++s) {
            ptx::mbar_init(full_bar + 8 * s, 1);
            ptx::mbar_init(empty_bar + 8 * s, P.mc ? 2 : 1);  // with multicast both CTAs' MMAs must release a slot
        }
        ptx::mbar_init(accum_bar, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(ptx::smem_u32(&tmem_base_slot), (uint32_t)P.tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (P.mc) ptx::cluster_sync();  // the peer's barriers are initialised before anything can arrive on them
    ptx::tc_fence_after();
    const uint32_t tmem_d = tmem_base_slot;

    if (warp == 0) {
        // ===== TMA producer: warp-uniform loop, one elected lane issues =====
        const int yz = P.Yp * P.Zp;
        const uint32_t tx = (uint32_t)((BM + P.BN) * P.KC * 2);
        uint32_t s = 0, ph = 1;
        int tap = it_begin / chunks, ch = it_begin % chunks;
        for (int it = it_begin; it < it_end; ++it) {
            int64_t delta = 0;
            if (P.ntaps == 27) delta = (int64_t)(tap / 9 - 1) * yz + (int64_t)((tap / 3) % 3 - 1) * P.Zp + (tap % 3 - 1);
            const int row = (int)(p0 + delta);
            ptx::mbar_wait(empty_bar + 8 * s, ph);
            if (ptx::elect_one()) {
                const uint32_t a_dst = smem_base + s * stage_bytes;
                ptx::mbar_arrive_expect_tx(full_bar + 8 * s, tx);
                ptx::tma_load_2d(a_dst, &map_a, full_bar + 8 * s, ch * P.KC, row);
                if (P.mc) {
                    // each CTA of the pair fetches one half of the weight tile and multicasts it to both
                    const uint32_t half_rows = (uint32_t)P.BN / 2, rank = ptx::cluster_ctarank();
                    ptx::tma_load_2d_mc(a_dst + P.a_bytes + rank * half_rows * (uint32_t)P.KC * 2u, &map_b, full_bar + 8 * s,
                                        tap * P.Cin + ch * P.KC, n0 + (int)(rank * half_rows), (uint16_t)0x3);
                } else {
                    ptx::tma_load_2d(a_dst + P.a_bytes, &map_b, full_bar + 8 * s, tap * P.Cin + ch * P.KC, n0);
                }
            }
            __syncwarp();
            if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
            if (++ch == chunks) { ch = 0; ++tap; }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: warp-uniform loop, one elected lane issues =====
        const uint32_t idesc = ptx::umma_idesc_bf16(BM, (uint32_t)P.BN);
        const uint32_t row_bytes = (uint32_t)P.KC * 2u;
        const uint64_t desc0 = ptx::umma_smem_desc(0, row_bytes);
        const int kk = P.KC / 16;
        uint32_t s = 0, ph = 0, accum = 0;
        for (int it = 0; it < k_iters; ++it) {
            ptx::mbar_wait(full_bar + 8 * s, ph);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint32_t a_src = smem_base + s * stage_bytes;
                const uint64_t a_desc = desc0 | (uint64_t)((a_src & 0x3FFFFu) >> 4);
                const uint64_t b_desc = desc0 | (uint64_t)(((a_src + P.a_bytes) & 0x3FFFFu) >> 4);
                for (int k = 0; k < kk; ++k) {
                    // advance 16 elements (32 bytes) along K inside the swizzle span: +2 in 16-byte units
                    ptx::umma_f16(tmem_d, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, accum);
                    accum = 1;
                }
                // smem slot reusable once these MMAs retire (in BOTH CTAs when the weight tile is shared)
                if (P.mc) ptx::umma_commit_mc(empty_bar + 8 * s, (uint16_t)0x3);
                else ptx::umma_commit(empty_bar + 8 * s);
            }
            accum = 1;
            __syncwarp();
            if (++s == (uint32_t)P.stages) { s = 0; ph ^= 1u; }
        }
        if (ptx::elect_one()) ptx::umma_commit(accum_bar);  // accumulator complete
        __syncwarp();
    } else {
        // ===== epilogue: TMEM -> registers -> global =====
        const int lg = warp % 4;  // TMEM lane group this warp may access
        ptx::mbar_wait(accum_bar, 0);
        ptx::tc_fence_after();
        const int64_t p = p0 + lg * 32 + lane;
        int b = 0;
        const bool interior = p < P.rows && row_is_interior(p, P, b);
        const uint32_t t_row = tmem_d + ((uint32_t)(lg * 32) << 16);
        const bool store = P.all_rows ? (p < P.rows) : interior;
        bf16* orow = out + p * P.ld_out + n0;
        const bool do_stats = gn_stats != nullptr;
        const int cpg = do_stats ? P.Cout / P.G : 1;
        // interior rows of one warp always belong to one sample (a sample boundary is >= 2 halo planes wide)
        const unsigned int_mask = __ballot_sync(0xffffffffu, interior);
        const int b_warp = __shfl_sync(0xffffffffu, b, int_mask ? __ffs(int_mask) - 1 : 0);
        float gs = 0.0f, gss = 0.0f;
        for (int c = 0; c < P.BN; c += 16) {
            uint32_t r[16];
            ptx::tmem_ld_x16(t_row + (uint32_t)c, r);
            ptx::tmem_ld_wait();
            if (scratch) {
                // split-K partial: fp32 vector reductions into the [rows][Cout] scratch (bias etc. in the finalize pass)
                if (p < P.rows && k_iters > 0) {
                    float* dst = scratch + p * P.Cout + n0 + c;
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        atomicAdd(reinterpret_cast<float4*>(dst + j),
                                  make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
                }
                continue;
            }
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) + (bias ? __ldg(bias + n0 + c + j) : 0.0f);
            if (store) {
                uint4 lo, hi;
                __nv_bfloat162* h0 = reinterpret_cast<__nv_bfloat162*>(&lo);
                __nv_bfloat162* h1 = reinterpret_cast<__nv_bfloat162*>(&hi);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    h0[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                    h1[j] = __floats2bfloat162_rn(v[8 + 2 * j], v[8 + 2 * j + 1]);
                }
                ptx::st_global_32B(orow + c, lo, hi);
            }
            if (do_stats) {
                // warp-uniform walk over the groups covered by these 16 columns
                const int sub = cpg >= 16 ? 16 : cpg;
                for (int j0 = 0; j0 < 16; j0 += sub) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (j >= j0 && j < j0 + sub && interior) {
                            gs += v[j];
                            gss = fmaf(v[j], v[j], gss);
                        }
                    }
                    const int col_end = n0 + c + j0 + sub;  // exclusive
                    // flush when a group is complete, or at the end of this N tile (group wider than BN)
                    if (col_end % cpg == 0 || c + j0 + sub == P.BN) {
                        const double ds = warp_sum((double)gs), dss = warp_sum((double)gss);
                        if (lane == 0 && int_mask) {
                            const int g = (col_end - 1) / cpg;
                            atomicAdd(gn_stats + ((int64_t)b_warp * P.G + g) * 2, ds);
                            atomicAdd(gn_stats + ((int64_t)b_warp * P.G + g) * 2 + 1, dss);
                        }
                        gs = gss = 0.0f;
                    }
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (P.mc) ptx::cluster_sync();  // no CTA leaves while its peer can still write into it / signal its barriers
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_d, (uint32_t)P.tmem_cols);
    }
}

// split-K finalize: out[p][c] = bf16(scratch[p][c] + bias[c]) on interior rows (all rows with all_rows)
__global__ void __launch_bounds__(256)
splitk_finalize_kernel(const float* __restrict__ scratch, const float* __restrict__ bias, bf16* __restrict__ out,
                       const ConvParams P) {
    const int chunks = P.Cout / 8;
    const int64_t total = P.rows * chunks;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = idx / chunks;
        const int c0 = (int)(idx % chunks) * 8;
        int b;
        if (!P.all_rows && !row_is_interior(p, P, b)) continue;
        const float4 a = *reinterpret_cast<const float4*>(scratch + p * P.Cout + c0);
        const float4 c = *reinterpret_cast<const float4*>(scratch + p * P.Cout + c0 + 4);
        float v[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
        if (bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += __ldg(bias + c0 + j);
        }
        Vec<bf16>::store(out + p * P.ld_out + c0, v);
    }
}

// ---- host side ---------------------------------------------------------------------------------

int pick_bn(int cout) {
    for (int bn = 256; bn >= 16; bn -= 16)
        if (cout % bn == 0) return bn;
    return 0;
}

}  // namespace

extern "C" int tdb_conv3d_bf16(const void* in, int ld_in, const void* w, const float* bias, void* out, int ld_out,
                               int B, int X, int Y, int Z, int Cin, int Cout, int ntaps, double* gn_stats, int G,
                               unsigned flags, float* splitk_scratch, void* stream) {
    TDB_REQUIRE(in && w && out, TDB_E_BADARG, "tdb_conv3d_bf16: null pointer");
    TDB_REQUIRE(ntaps == 1 || ntaps == 27, TDB_E_BADARG, "tdb_conv3d_bf16: ntaps must be 1 or 27");
    TDB_REQUIRE(Cin % 16 == 0 && Cout % 16 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0, TDB_E_UNSUPPORTED,
                "tdb_conv3d_bf16: need Cin %% 16 == 0, Cout %% 16 == 0, pitches %% 8 == 0 (Cin=%d Cout=%d)", Cin, Cout);
    TDB_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)w & 15) == 0, TDB_E_UNSUPPORTED,
                "tdb_conv3d_bf16: pointers must be 16-byte aligned");
    TDB_REQUIRE(!gn_stats || (G >= 1 && Cout % G == 0 && ((Cout / G) % 16 == 0 || 16 % (Cout / G) == 0)), TDB_E_UNSUPPORTED,
                "tdb_conv3d_bf16: fused GroupNorm moments need Cout/G to divide or be a multiple of 16");
    Grid3 g(B, X, Y, Z);
    TDB_REQUIRE(g.rows < (1ll << 31) - 4096, TDB_E_UNSUPPORTED, "tdb_conv3d_bf16: too many rows for 32-bit TMA coordinates");

    ConvParams P;
    P.rows = g.rows;
    P.vox_p = g.vox_p;
    P.Xp = g.Xp; P.Yp = g.Yp; P.Zp = g.Zp;
    P.Cin = Cin; P.Cout = Cout; P.ntaps = ntaps;
    P.KC = Cin % 64 == 0 ? 64 : (Cin % 32 == 0 ? 32 : 16);
    P.BN = pick_bn(Cout);
    TDB_REQUIRE(P.BN >= 16, TDB_E_UNSUPPORTED, "tdb_conv3d_bf16: no N tile for Cout=%d", Cout);
    auto up1k = [](int v) { return (v + 1023) & ~1023; };
    P.a_bytes = up1k(BM * P.KC * 2);
    P.b_bytes = up1k(P.BN * P.KC * 2);
    const int stage_bytes = P.a_bytes + P.b_bytes;
    // keep <= ~100 KB so that two CTAs share an SM when the tiles are narrow
    int stages = (100 * 1024) / stage_bytes;
    if (stages < 3) stages = (200 * 1024) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) stages = 2;
    P.stages = stages;
    int cols = 32;
    while (cols < P.BN) cols *= 2;
    P.tmem_cols = cols;
    P.ld_out = ld_out;
    P.G = gn_stats ? G : 0;
    P.all_rows = (flags & TDB_CONV_ALL_ROWS) ? 1 : 0;

    CUtensorMap map_a, map_b;
    TDB_REQUIRE(encode_fn() != nullptr, TDB_E_NODEVICE, "tdb_conv3d_bf16: cuTensorMapEncodeTiled unavailable (no driver)");
    TDB_REQUIRE(make_map_2d_bf16(&map_a, in, (uint64_t)Cin, (uint64_t)g.rows, (uint64_t)ld_in, (uint32_t)P.KC, BM), TDB_E_BADARG,
                "tdb_conv3d_bf16: tensor map (activations) rejected");
    // Opt-in: 2-CTA clusters along M share the weight tile (TMA multicast).  Measured on B200 (round 1) this is
    // 5-10 % SLOWER than independent CTAs: these layers are bound by the per-SM shared-memory ingest
    // (~60 B/clk/SM), which multicast does not reduce - kept for the cta_group::2 follow-up.
    P.mc = ((flags & TDB_CONV_CLUSTER_MC) && ceil_div(g.rows, BM) >= 8 && P.BN % 16 == 0) ? 1 : 0;
    TDB_REQUIRE(make_map_2d_bf16(&map_b, w, (uint64_t)ntaps * Cin, (uint64_t)Cout, (uint64_t)ntaps * Cin, (uint32_t)P.KC,
                            (uint32_t)(P.mc ? P.BN / 2 : P.BN)),
                TDB_E_BADARG, "tdb_conv3d_bf16: tensor map (weights) rejected");

    const size_t smem = (size_t)stages * stage_bytes + 1024;
    cudaError_t e = cudaFuncSetAttribute(conv3d_bf16_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_conv3d_bf16: cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
    dim3 grid((unsigned)ceil_div(g.rows, BM), (unsigned)(Cout / P.BN));
    if (P.mc) grid.x = (grid.x + 1) & ~1u;  // whole clusters; a surplus CTA computes (and stores) nothing
    // split-K for the small-M / huge-K layers of the deep levels: few output tiles, hundreds of serial k-steps
    const int ctas = (int)(grid.x * grid.y);
    const int k_total = ntaps * (Cin / P.KC);
    int splits = 1;
    if (splitk_scratch && ctas < 120 && k_total >= 32) {
        splits = (2 * 148) / ctas;
        if (splits > k_total / 8) splits = k_total / 8;
        if (splits > 16) splits = 16;
        if (splits < 1) splits = 1;
        const int k_per = (k_total + splits - 1) / splits;
        splits = (k_total + k_per - 1) / k_per;  // every split owns at least one k-step
    }
    P.splits = splits;
    cudaStream_t s = (cudaStream_t)stream;
    auto launch = [&](const float* bias_, double* stats_, float* scratch_) -> cudaError_t {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = P.mc ? 2 : 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        bf16* out_ = (bf16*)out;
        return cudaLaunchKernelEx(&cfg, conv3d_bf16_tc_kernel, map_a, map_b, bias_, out_, stats_, scratch_, P);
    };
    if (splits == 1) {
        e = launch(bias, gn_stats, nullptr);
        TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_conv3d_bf16: launch: %s", cudaGetErrorString(e));
        TDB_CHECK_LAUNCH("tdb_conv3d_bf16");
        return 0;
    }
    grid.z = (unsigned)splits;
    e = cudaMemsetAsync(splitk_scratch, 0, (size_t)g.rows * Cout * sizeof(float), s);
    TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_conv3d_bf16: cudaMemsetAsync: %s", cudaGetErrorString(e));
    e = launch(nullptr, nullptr, splitk_scratch);
    TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_conv3d_bf16: launch (split-K): %s", cudaGetErrorString(e));
    TDB_CHECK_LAUNCH("tdb_conv3d_bf16 (split-K)");
    const int64_t items = g.rows * (Cout / 8);
    int fblocks = (int)ceil_div(items, 256);
    if (fblocks > 148 * 8) fblocks = 148 * 8;
    splitk_finalize_kernel<<<fblocks, 256, 0, s>>>(splitk_scratch, bias, (bf16*)out, P);
    TDB_CHECK_LAUNCH("tdb_conv3d_bf16 (split-K finalize)");
    if (gn_stats) return tdb_gn_stats(out, ld_out, gn_stats, B, X, Y, Z, Cout, G, TDB_BF16, stream);
    return 0;
}
