// Kernel-layout copies of the convolution weights (bf16) straight from the fp32 parameters, one launch per weight.
//
// The parameters keep the reference's layout (Cout, Cin, kD, kH, kW) fp32 (nn.Conv3d, ddpm.py:164,188).  The tensor-core
// kernels want bf16 K-major matrices: per-tap  [O][tap*I + i]  (tdb_conv3d_bf16, _win) or kz-folded
// [(o/T*3 + kz)*T + o%T][(kx*3+ky)*I + i]  (_fold, _fold2, _winz, _winp), and the input-gradient convolutions want the
// same layouts of the tap-reversed transpose  W'[o = ci][i = co][tap] = W[co][ci][26 - tap].  In training every one of
// them is rebuilt every step (the optimizer just changed the parameters); as torch ops that is a strided permute copy, a
// flip (index kernel) and a cast per weight - 178 launches, ~1.7 ms of serialized device time per step, most of it on the
// critical path.  Here a block moves a 32 x 32 (co, ci) tile with all taps through shared memory: coalesced fp32 reads
// (the ci*taps run of one co is contiguous), 64-byte bf16 write segments, no intermediate tensors.
#include "common.cuh"

using namespace tdb;
using bf16 = __nv_bfloat16;

namespace {

constexpr int PT = 32;  // tile edge in both channel dimensions

template <int taps>
__device__ __forceinline__ void pack_tile(const float* __restrict__ w, bf16* __restrict__ dst, int Cout, int Cin, int folded, int tile_n,
                                          int transpose, int bx, int by) {
    extern __shared__ float sm[];  // [PT co][PT*taps + 1]
    const int pitch = PT * taps + 1;
    const int co0 = by * PT, ci0 = bx * PT;
    const int n_co = min(PT, Cout - co0), n_ci = min(PT, Cin - ci0);
    // read: for every co of the tile the contiguous run of n_ci*taps floats
    // (U loads in flight per thread: one load per trip left a single block at ~40 us whatever its size)
    constexpr int U = taps == 27 ? 9 : 1;
    const int run = n_ci * taps;
    for (int co_l = threadIdx.x / 32; co_l < n_co; co_l += blockDim.x / 32) {
        const float* src = w + ((int64_t)(co0 + co_l) * Cin + ci0) * taps;
        for (int k0 = threadIdx.x % 32; k0 < run; k0 += 32 * U) {
            float r[U];
#pragma unroll
            for (int u = 0; u < U; ++u) r[u] = k0 + 32 * u < run ? __ldg(src + k0 + 32 * u) : 0.0f;
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (k0 + 32 * u < run) sm[co_l * pitch + k0 + 32 * u] = r[u];
        }
    }
    __syncthreads();
    // logical weight Wl[o][i][tap]: forward o = co, i = ci; input gradient o = ci, i = co, taps reversed
    const int O = transpose ? Cin : Cout, I = transpose ? Cout : Cin;
    (void)O;
    const int n_o = transpose ? n_ci : n_co, n_i = transpose ? n_co : n_ci;
    const int o0 = transpose ? ci0 : co0, i0 = transpose ? co0 : ci0;
    const int lane = threadIdx.x % 32, warp = threadIdx.x / 32, nwarps = blockDim.x / 32;
    // one warp per (o, tap): lanes walk i - 32 consecutive bf16 of one dst row
    for (int job = warp; job < n_o * taps; job += nwarps) {
        const int o_l = job / taps, tap = job % taps;      // tap = logical tap index
        if (lane >= n_i) continue;
        const int src_tap = transpose ? taps - 1 - tap : tap;
        const float v = transpose ? sm[lane * pitch + o_l * taps + src_tap] : sm[o_l * pitch + lane * taps + src_tap];
        const int o = o0 + o_l, i = i0 + lane;
        int64_t row, col;
        if (folded) {  // taps == 27
            const int kz = tap % 3, kxy = tap / 3;
            row = (int64_t)((o / tile_n) * 3 + kz) * tile_n + o % tile_n;
            col = (int64_t)kxy * I + i;
            dst[row * (9 * (int64_t)I) + col] = __float2bfloat16_rn(v);
        } else {
            row = o;
            col = (int64_t)tap * I + i;
            dst[row * ((int64_t)taps * I) + col] = __float2bfloat16_rn(v);
        }
    }
}

template <int taps>
__global__ void __launch_bounds__(1024)
pack_conv_kernel(const float* __restrict__ w, bf16* __restrict__ dst, int Cout, int Cin, int folded, int tile_n, int transpose) {
    pack_tile<taps>(w, dst, Cout, Cin, folded, tile_n, transpose, blockIdx.x, blockIdx.y);
}

// Many weights in ONE launch: the job table travels as a kernel parameter (no device table, graph-capturable as is); a block
// finds its job from the running block offsets.  62 one-weight launches of 4 .. 256 blocks were ~0.7 ms of serialized,
// mostly latency-bound device time at the head of every training step.
constexpr int PACK_BATCH = 64;
struct PackJobDev {
    const float* w;
    bf16* dst;
    int Cout, Cin, taps, folded, tile_n, transpose, block0, gx;
};
struct PackBatch {
    PackJobDev j[PACK_BATCH];
    int n;
};
__global__ void __launch_bounds__(1024) pack_conv_batch_kernel(const __grid_constant__ PackBatch P) {
    int k = 0;
    for (int q = 1; q < P.n; ++q)
        if ((int)blockIdx.x >= P.j[q].block0) k = q;
    const PackJobDev& J = P.j[k];
    const int lb = (int)blockIdx.x - J.block0;
    if (J.taps == 27) pack_tile<27>(J.w, J.dst, J.Cout, J.Cin, J.folded, J.tile_n, J.transpose, lb % J.gx, lb / J.gx);
    else pack_tile<1>(J.w, J.dst, J.Cout, J.Cin, J.folded, J.tile_n, J.transpose, lb % J.gx, lb / J.gx);
}

// Weight gradient from the kernels' accumulation layout to the parameter layout: dw [taps][Cin][Cout] fp32 ->
// out (Cout, Cin, taps) fp32.  A block moves a (TCI ci x 32 co x all taps) tile through shared memory: 128-byte reads along
// co, contiguous runs of TCI * taps floats per co on the write side.
template <int taps, int TCI>
__global__ void __launch_bounds__(256) unpack_wgrad_kernel(const float* __restrict__ dw, float* __restrict__ out, int Cin, int Cout) {
    __shared__ float sm[taps * TCI][33];
    const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * TCI;
    const int n_co = min(32, Cout - co0), n_ci = min(TCI, Cin - ci0);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    for (int r = warp; r < taps * n_ci; r += 8) {  // r = t * n_ci + ci_l
        const int t = r / n_ci, ci_l = r % n_ci;
        if (lane < n_co) sm[ci_l * taps + t][lane] = dw[((int64_t)t * Cin + ci0 + ci_l) * Cout + co0 + lane];
    }
    __syncthreads();
    const int run = n_ci * taps;
    for (int co_l = warp; co_l < n_co; co_l += 8) {
        float* dst = out + ((int64_t)(co0 + co_l) * Cin + ci0) * taps;
        for (int e = lane; e < run; e += 32) {
            dst[e] = sm[e][co_l];  // row e = ci_l * taps + t: consecutive lanes read consecutive rows (33-float pitch: no conflicts)
        }
    }
}

}  // namespace

extern "C" int tdb_unpack_wgrad(const float* dw, float* out, int Cout, int Cin, int taps, void* stream) {
    TDB_REQUIRE(dw && out, TDB_E_BADARG, "tdb_unpack_wgrad: null pointer");
    TDB_REQUIRE((taps == 27 || taps == 1) && Cout >= 1 && Cin >= 1, TDB_E_BADARG, "tdb_unpack_wgrad: taps must be 27 or 1");
    if (taps == 27) {
        dim3 grid((unsigned)ceil_div(Cout, 32), (unsigned)ceil_div(Cin, 8));
        unpack_wgrad_kernel<27, 8><<<grid, 256, 0, (cudaStream_t)stream>>>(dw, out, Cin, Cout);
    } else {
        dim3 grid((unsigned)ceil_div(Cout, 32), (unsigned)ceil_div(Cin, 32));
        unpack_wgrad_kernel<1, 32><<<grid, 256, 0, (cudaStream_t)stream>>>(dw, out, Cin, Cout);
    }
    TDB_CHECK_LAUNCH("tdb_unpack_wgrad");
    return 0;
}

// w: fp32 (Cout, Cin, taps) contiguous (taps = 27 or 1).  dst: bf16, Cout*Cin*taps elements.
// folded: 0 = per-tap layout [O][taps*I]; 1 = kz-folded [3*O][9*I] in N tiles of tile_n rows (taps must be 27).
// transpose: 0 = the forward weights (O = Cout, I = Cin); 1 = the input-gradient weights W'[ci][co][26 - tap] (O = Cin, I = Cout).
extern "C" int tdb_pack_conv_weights(const float* w, void* dst, int Cout, int Cin, int taps, int folded, int tile_n, int transpose,
                                     void* stream) {
    TDB_REQUIRE(w && dst, TDB_E_BADARG, "tdb_pack_conv_weights: null pointer");
    TDB_REQUIRE((taps == 27 || taps == 1) && Cout >= 1 && Cin >= 1, TDB_E_BADARG, "tdb_pack_conv_weights: taps must be 27 or 1");
    const int O = transpose ? Cin : Cout;
    TDB_REQUIRE(!folded || (taps == 27 && tile_n >= 1 && O % tile_n == 0), TDB_E_BADARG,
                "tdb_pack_conv_weights: the kz-folded layout needs 27 taps and an N tile that divides %d", O);
    const size_t smem = (size_t)PT * (PT * taps + 1) * sizeof(float);
    dim3 grid((unsigned)ceil_div(Cin, PT), (unsigned)ceil_div(Cout, PT));
    if (taps == 27) {
        cudaError_t e = cudaFuncSetAttribute(pack_conv_kernel<27>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_pack_conv_weights: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        pack_conv_kernel<27><<<grid, 1024, smem, (cudaStream_t)stream>>>(w, (bf16*)dst, Cout, Cin, folded, tile_n, transpose);
    } else {
        pack_conv_kernel<1><<<grid, 1024, smem, (cudaStream_t)stream>>>(w, (bf16*)dst, Cout, Cin, folded, tile_n, transpose);
    }
    TDB_CHECK_LAUNCH("tdb_pack_conv_weights");
    return 0;
}

// The same for n weights at once (launches of up to 64 weights each).  jobs: host array, read during the call only.
extern "C" int tdb_pack_conv_weights_batch(const TdbPackJob* jobs, int n, void* stream) {
    TDB_REQUIRE(jobs && n >= 0, TDB_E_BADARG, "tdb_pack_conv_weights_batch: null pointer");
    const size_t smem = (size_t)PT * (PT * 27 + 1) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(pack_conv_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_pack_conv_weights_batch: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    for (int first = 0; first < n; first += PACK_BATCH) {
        PackBatch P;
        P.n = n - first < PACK_BATCH ? n - first : PACK_BATCH;
        int blocks = 0;
        for (int q = 0; q < P.n; ++q) {
            const TdbPackJob& a = jobs[first + q];
            TDB_REQUIRE(a.w && a.dst, TDB_E_BADARG, "tdb_pack_conv_weights_batch: null pointer in job %d", first + q);
            TDB_REQUIRE((a.taps == 27 || a.taps == 1) && a.Cout >= 1 && a.Cin >= 1, TDB_E_BADARG,
                        "tdb_pack_conv_weights_batch: taps must be 27 or 1 (job %d)", first + q);
            const int O = a.transpose ? a.Cin : a.Cout;
            TDB_REQUIRE(!a.folded || (a.taps == 27 && a.tile_n >= 1 && O % a.tile_n == 0), TDB_E_BADARG,
                        "tdb_pack_conv_weights_batch: the kz-folded layout needs 27 taps and an N tile that divides %d (job %d)", O, first + q);
            PackJobDev& d = P.j[q];
            d.w = a.w; d.dst = (bf16*)a.dst;
            d.Cout = a.Cout; d.Cin = a.Cin; d.taps = a.taps; d.folded = a.folded; d.tile_n = a.tile_n; d.transpose = a.transpose;
            d.block0 = blocks;
            d.gx = (int)ceil_div(a.Cin, PT);
            blocks += d.gx * (int)ceil_div(a.Cout, PT);
        }
        if (blocks == 0) continue;
        pack_conv_batch_kernel<<<(unsigned)blocks, 1024, smem, (cudaStream_t)stream>>>(P);
        TDB_CHECK_LAUNCH("tdb_pack_conv_weights_batch");
    }
    return 0;
}
