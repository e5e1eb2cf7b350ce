// fp32 CUDA-core convolution over halo grids: the 1e-5 parity path (exact fp32 FMA
// accumulation; TF32/bf16 tensor-core inputs cannot hold 1e-5).
//
// Because the input carries a materialised replicate halo and rows are linearised over the
// haloed grid, every filter tap is a constant row shift:
//     out[p][co] = bias[co] + sum_tap sum_ci in[p + delta(tap)][ci] * w[tap][ci][co]
//     delta = (kx-1)*Yp*Zp + (ky-1)*Zp + (kz-1)
// which turns the convolution into 27 accumulated row-shifted GEMMs with no boundary logic.
// Halo rows of `out` receive meaningless values (never read: consumers clamp to the interior).
#include "common.cuh"

using namespace tdb;

namespace {

constexpr int BM = 64, BN = 64, KC = 8, THREADS = 256;

__global__ void __launch_bounds__(THREADS)
conv3d_f32_kernel(const float* __restrict__ in, int ld_in, const float* __restrict__ w,
                  const float* __restrict__ bias, float* __restrict__ out, int ld_out, int64_t rows,
                  int yz_p, int z_p, int Cin, int Cout, int ntaps) {
    __shared__ float As[KC][BM + 4];
    __shared__ float Bs[KC][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    const int a_row = tid / 4, a_k = (tid % 4) * 2;
    const int b_k = tid / 32, b_n = (tid % 32) * 2;

    for (int tap = 0; tap < ntaps; ++tap) {
        int64_t delta = 0;
        if (ntaps == 27) delta = (int64_t)(tap / 9 - 1) * yz_p + (int64_t)((tap / 3) % 3 - 1) * z_p + (tap % 3 - 1);
        int64_t src = m0 + a_row + delta;
        src = src < 0 ? 0 : (src >= rows ? rows - 1 : src);  // only halo rows can leave the range
        const float* a_ptr = in + src * ld_in + a_k;
        const float* b_ptr = w + ((int64_t)tap * Cin + b_k) * Cout + n0 + b_n;
        const bool b_ok = n0 + b_n < Cout;
        for (int c0 = 0; c0 < Cin; c0 += KC) {
            const float2 av = *reinterpret_cast<const float2*>(a_ptr + c0);
            float2 bv = make_float2(0.0f, 0.0f);
            if (b_ok) bv = *reinterpret_cast<const float2*>(b_ptr + (int64_t)c0 * Cout);
            As[a_k][a_row] = av.x;
            As[a_k + 1][a_row] = av.y;
            Bs[b_k][b_n] = bv.x;
            Bs[b_k][b_n + 1] = bv.y;
            __syncthreads();
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w};
                const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
    const int n = n0 + tx * 4;
    if (n < Cout) {
        float bb[4] = {0.f, 0.f, 0.f, 0.f};
        if (bias) {
#pragma unroll
            for (int j = 0; j < 4; ++j) bb[j] = bias[n + j];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t m = m0 + ty * 4 + i;
            if (m < rows)
                *reinterpret_cast<float4*>(out + m * ld_out + n) =
                    make_float4(acc[i][0] + bb[0], acc[i][1] + bb[1], acc[i][2] + bb[2], acc[i][3] + bb[3]);
        }
    }
}

}  // namespace

extern "C" int tdb_conv3d_f32(const float* in, int ld_in, const float* w, const float* bias, float* out,
                              int ld_out, int B, int X, int Y, int Z, int Cin, int Cout, int ntaps,
                              void* stream) {
    TDB_REQUIRE(in && w && out, TDB_E_BADARG, "tdb_conv3d_f32: null pointer");
    TDB_REQUIRE(ntaps == 1 || ntaps == 27, TDB_E_BADARG, "tdb_conv3d_f32: ntaps must be 1 or 27");
    TDB_REQUIRE(Cin % KC == 0 && Cout % 4 == 0 && ld_in % 4 == 0 && ld_out % 4 == 0 &&
                    ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)w & 15) == 0,
                TDB_E_UNSUPPORTED, "tdb_conv3d_f32: need Cin %% 8 == 0, Cout %% 4 == 0 (Cin=%d Cout=%d)", Cin, Cout);
    Grid3 g(B, X, Y, Z);
    dim3 grid((unsigned)ceil_div(g.rows, BM), (unsigned)ceil_div(Cout, BN));
    conv3d_f32_kernel<<<grid, THREADS, 0, (cudaStream_t)stream>>>(in, ld_in, w, bias, out, ld_out, g.rows,
                                                                  g.Yp * g.Zp, g.Zp, Cin, Cout, ntaps);
    TDB_CHECK_LAUNCH("tdb_conv3d_f32");
    return 0;
}
