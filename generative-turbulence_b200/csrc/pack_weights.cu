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
__global__ void __launch_bounds__(1024)
pack_conv_kernel(const float* __restrict__ w, bf16* __restrict__ dst, int Cout, int Cin, int folded, int tile_n, int transpose) {
    extern __shared__ float sm[];  // [PT co][PT*taps + 1]
    const int pitch = PT * taps + 1;
    const int co0 = blockIdx.y * PT, ci0 = blockIdx.x * PT;
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

}  // namespace

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
