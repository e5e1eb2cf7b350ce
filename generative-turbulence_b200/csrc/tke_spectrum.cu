// Turbulent-kinetic-energy spectrum of a cube of the flow field on the GPU (reference:
// TurbulentKineticEnergySpectrum.forward, turbdiff/models/metrics.py:296-320, interp3 :211-267).
//
//   tke     = |u - u_mean|^2 / 2 per voxel                                   (metrics.py:298, 362-363)
//   F       = fftshift(fftn(tke))            three passes of a dense DFT along one axis each (n <= 64 per axis: 48^3 is
//                                            0.13 GFLOP per cube - a matrix DFT with double accumulation is both exact
//                                            to fp32 storage and far below the launch latency of anything smarter)
//   L       = log |F|^2
//   E(k)    = 4 pi k^2 * sum_j w_j exp(trilinear(L, k p_j + centre))         log-domain interpolation onto the sphere
//
// All tensors fp32, NCDHW like the reference.  HBM-trivial (a 48^3 cube is 0.4 MB): latency-bound, 6 launches.
#include "common.cuh"

#include <math_constants.h>

using namespace tdb;

namespace {

constexpr int kT = 256;
constexpr int kMaxN = 64;

__global__ void __launch_bounds__(kT)
tke_field_kernel(const float* __restrict__ u, const float* __restrict__ u_mean, float2* __restrict__ out, int B, int64_t nvox) {
    const int64_t total = (int64_t)B * nvox;
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
        const int64_t b = i / nvox, v = i - b * nvox;
        float acc = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float d = u[(b * 3 + c) * nvox + v] - (u_mean ? u_mean[c * nvox + v] : 0.0f);
            acc += d * d;  // (u'**2).sum(dim=-4): squares summed in channel order
        }
        out[i] = make_float2(0.5f * acc, 0.0f);
    }
}

// One DFT pass along the axis of length n in the view [outer][n][inner]; the result is stored fft-shifted along that
// axis (torch.fft.fftshift = roll by n / 2).  Twiddles exp(-2 pi i j / n) in double, accumulation in double.
__global__ void __launch_bounds__(kT)
dft_axis_kernel(const float2* __restrict__ in, float2* __restrict__ out, int64_t outer, int n, int64_t inner) {
    __shared__ double tw_c[kMaxN], tw_s[kMaxN];
    for (int j = threadIdx.x; j < n; j += kT) {
        double s, c;
        sincospi(-2.0 * (double)j / (double)n, &s, &c);
        tw_c[j] = c;
        tw_s[j] = s;
    }
    __syncthreads();
    const int64_t total = outer * n * inner;
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
        const int64_t in_idx = i % inner;
        const int f = (int)((i / inner) % n);
        const int64_t o = i / (inner * n);
        const float2* line = in + o * n * inner + in_idx;
        double re = 0.0, im = 0.0;
        int ph = 0;  // (f * x) mod n
        for (int x = 0; x < n; ++x) {
            const float2 v = line[(int64_t)x * inner];
            const double c = tw_c[ph], s = tw_s[ph];
            re = fma((double)v.x, c, fma(-(double)v.y, s, re));
            im = fma((double)v.x, s, fma((double)v.y, c, im));
            ph += f;
            if (ph >= n) ph -= n;
        }
        const int fs = (f + n / 2) % n;
        out[(o * n + fs) * inner + in_idx] = make_float2((float)re, (float)im);
    }
}

__global__ void __launch_bounds__(kT)
log_power_kernel(const float2* __restrict__ f, float* __restrict__ out, int64_t total) {
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
        const float2 v = f[i];
        const float a = sqrtf(v.x * v.x + v.y * v.y);  // tke_fft.abs()
        out[i] = logf(a * a);                          // (abs ** 2).log()
    }
}

// grid = (K, B): E[b][k] = 4 pi k^2 * sum_j w_j exp(interp3(L_b, k p_j + centre)), arithmetic in the reference's order.
__global__ void __launch_bounds__(kT)
sphere_kernel(const float* __restrict__ L, const float* __restrict__ kvals, const float* __restrict__ pts, const float* __restrict__ wts,
              int P, int n0, int n1, int n2, float* __restrict__ E, int K) {
    __shared__ float red[kT / 32];
    const int ki = blockIdx.x, b = blockIdx.y;
    const float k = kvals[ki];
    const float* g = L + (int64_t)b * n0 * n1 * n2;
    const float c0 = (float)(n0 / 2), c1 = (float)(n1 / 2), c2 = (float)(n2 / 2);
    float acc = 0.0f;
    for (int j = threadIdx.x; j < P; j += kT) {
        const float qx = k * pts[3 * j] + c0, qy = k * pts[3 * j + 1] + c1, qz = k * pts[3 * j + 2] + c2;
        const int fx = (int)floorf(qx), fy = (int)floorf(qy), fz = (int)floorf(qz);
        const int x0 = clampi(fx, 0, n0 - 1), y0 = clampi(fy, 0, n1 - 1), z0 = clampi(fz, 0, n2 - 1);
        const int x1 = clampi(fx + 1, 0, n0 - 1), y1 = clampi(fy + 1, 0, n1 - 1), z1 = clampi(fz + 1, 0, n2 - 1);
        const float wx = qx - (float)x0, wy = qy - (float)y0, wz = qz - (float)z0;  // against the CLAMPED lower index (metrics.py:252)
        auto at = [&](int x, int y, int z) { return g[((int64_t)x * n1 + y) * n2 + z]; };
        float v = (1 - wx) * (1 - wy) * (1 - wz) * at(x0, y0, z0);
        v += (1 - wx) * (1 - wy) * wz * at(x0, y0, z1);
        v += (1 - wx) * wy * (1 - wz) * at(x0, y1, z0);
        v += (1 - wx) * wy * wz * at(x0, y1, z1);
        v += wx * (1 - wy) * (1 - wz) * at(x1, y0, z0);
        v += wx * (1 - wy) * wz * at(x1, y0, z1);
        v += wx * wy * (1 - wz) * at(x1, y1, z0);
        v += wx * wy * wz * at(x1, y1, z1);
        acc = fmaf(expf(v), wts[j], acc);
    }
    acc = warp_sum(acc);
    if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < kT / 32 ? red[threadIdx.x] : 0.0f;
        v = warp_sum(v);
        if (threadIdx.x == 0) E[(int64_t)b * K + ki] = v * (4.0f * CUDART_PI_F * k * k);
    }
}

int blocks(int64_t n) {
    int64_t b = ceil_div(n, kT);
    return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

}  // namespace

extern "C" int tdb_tke_spectrum(const float* u, const float* u_mean, int B, int n0, int n1, int n2, const float* k, int K,
                                const float* points, const float* weights, int P, float* work, float* E, void* stream) {
    TDB_REQUIRE(u && k && points && weights && work && E, TDB_E_BADARG, "tdb_tke_spectrum: null pointer");
    TDB_REQUIRE(B >= 0 && K >= 0 && P >= 1 && n0 >= 1 && n1 >= 1 && n2 >= 1 && n0 <= kMaxN && n1 <= kMaxN && n2 <= kMaxN, TDB_E_UNSUPPORTED,
                "tdb_tke_spectrum: cube axes must be in [1, %d] (got %d x %d x %d)", kMaxN, n0, n1, n2);
    if (B == 0 || K == 0) return 0;
    const int64_t nvox = (int64_t)n0 * n1 * n2, total = (int64_t)B * nvox;
    cudaStream_t s = (cudaStream_t)stream;
    float2* a = reinterpret_cast<float2*>(work);
    float2* b = a + total;
    tke_field_kernel<<<blocks(total), kT, 0, s>>>(u, u_mean, a, B, nvox);
    TDB_CHECK_LAUNCH("tdb_tke_spectrum (field)");
    dft_axis_kernel<<<blocks(total), kT, 0, s>>>(a, b, (int64_t)B * n0 * n1, n2, 1);
    TDB_CHECK_LAUNCH("tdb_tke_spectrum (dft z)");
    dft_axis_kernel<<<blocks(total), kT, 0, s>>>(b, a, (int64_t)B * n0, n1, n2);
    TDB_CHECK_LAUNCH("tdb_tke_spectrum (dft y)");
    dft_axis_kernel<<<blocks(total), kT, 0, s>>>(a, b, B, n0, (int64_t)n1 * n2);
    TDB_CHECK_LAUNCH("tdb_tke_spectrum (dft x)");
    float* L = reinterpret_cast<float*>(a);
    log_power_kernel<<<blocks(total), kT, 0, s>>>(b, L, total);
    TDB_CHECK_LAUNCH("tdb_tke_spectrum (log power)");
    sphere_kernel<<<dim3((unsigned)K, (unsigned)B), kT, 0, s>>>(L, k, points, weights, P, n0, n1, n2, E, K);
    TDB_CHECK_LAUNCH("tdb_tke_spectrum (sphere)");
    return 0;
}
