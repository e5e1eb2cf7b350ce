// Backward kernels of the denoiser (training path, ddpm.py:833-882 through autograd in the reference).
// Gradients travel as halo grids in the storage type of the forward activations.  Convolution
// input-gradients are produced by the forward convolution kernels themselves (transposed, tap-reversed
// weights over a zero-halo output gradient, all rows stored); the kernels here provide the rest:
//   tdb_halo_fold            adjoint of halo materialisation: halo rows are added onto the border voxels (and zeroed)
//   tdb_pointwise_bwd_reduce per-(sample, channel) sums  A1 = sum g_u,  A2 = sum g_u * xhat
//   tdb_pointwise_bwd_apply  GroupNorm/FiLM/SiLU input gradient from g_out and the sums
//   tdb_conv3d_wgrad         weight gradient of the 3x3x3 / 1x1x1 convolutions (fp32 accumulate)
//   tdb_trilinear_bwd        transpose of the align_corners trilinear resampling (gather form)
//   tdb_attention_bwd        softmax-attention backward per (sample, head)
//   tdb_cl_nc_outer          out[c][f] = sum_{b,v} G[b,v][c] * Q[b][f][v]  (1x1 encoder/decoder weight grads)
#include <mma.h>

#include "common.cuh"

using namespace tdb;
using bf16 = __nv_bfloat16;

namespace {

constexpr int kThreads = 256;

struct RowSplit {
    FastDiv by_z, by_y;
    __device__ __forceinline__ void operator()(uint32_t r, int& xp, int& yp, int& zp) const {
        uint32_t q, zz, xx, yy;
        by_z.divmod(r, q, zz);
        by_y.divmod(q, xx, yy);
        xp = (int)xx; yp = (int)yy; zp = (int)zz;
    }
};
RowSplit make_split(const Grid3& g) {
    RowSplit s;
    s.by_z = FastDiv((uint32_t)g.Zp);
    s.by_y = FastDiv((uint32_t)g.Yp);
    return s;
}
int blocks_per_sample(int64_t items_per_sample, int B) {
    int64_t blocks = ceil_div(items_per_sample, kThreads);
    int64_t cap = (148 * 16) / (B < 1 ? 1 : B);
    if (cap < 8) cap = 8;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}
bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

// ---------------------------------------------------------------- halo fold
// haloed coordinates of the images of interior coordinate c (1..n) along one axis
__device__ __forceinline__ int axis_images(int c, int n, int (&img)[3]) {
    int k = 0;
    img[k++] = c;
    if (c == 1) img[k++] = 0;
    if (c == n) img[k++] = n + 1;
    return k;
}

template <typename T>
__device__ __forceinline__ void fold_border_voxel(T* __restrict__ g, int ld, const Grid3& gr, int b, int xp, int yp, int zp, int c0) {
    constexpr int N = Vec<T>::N;
    int ix[3], iy[3], iz[3];
    const int nx = axis_images(xp, gr.X, ix), ny = axis_images(yp, gr.Y, iy), nz = axis_images(zp, gr.Z, iz);
    float acc[N];
#pragma unroll
    for (int i = 0; i < N; ++i) acc[i] = 0.0f;
    for (int a = 0; a < nx; ++a)
        for (int bb = 0; bb < ny; ++bb)
            for (int c = 0; c < nz; ++c) {
                float v[N];
                T* src = g + ((int64_t)b * gr.vox_p + ((int64_t)ix[a] * gr.Yp + iy[bb]) * gr.Zp + iz[c]) * ld + c0;
                Vec<T>::load(src, v);
#pragma unroll
                for (int i = 0; i < N; ++i) acc[i] += v[i];
                if (a | bb | c) {  // every halo row is the image of exactly one border voxel: leave it zero
                    float z[N];
#pragma unroll
                    for (int i = 0; i < N; ++i) z[i] = 0.0f;
                    Vec<T>::store(src, z);
                }
            }
    Vec<T>::store(g + ((int64_t)b * gr.vox_p + ((int64_t)xp * gr.Yp + yp) * gr.Zp + zp) * ld + c0, acc);
}

// generic form: walks every row and keeps the border voxels (grids with an axis of length 1)
template <typename T>
__global__ void __launch_bounds__(kThreads)
halo_fold_kernel(T* __restrict__ g, int ld, Grid3 gr, RowSplit split, int chunks) {
    constexpr int N = Vec<T>::N;
    const int b = blockIdx.y;
    const int vox_step = kThreads / chunks;
    const int ch = threadIdx.x % chunks, lane_vox = threadIdx.x / chunks;
    if (lane_vox >= vox_step) return;
    for (uint32_t r = blockIdx.x * vox_step + lane_vox; r < (uint32_t)gr.vox_p; r += gridDim.x * vox_step) {
        int xp, yp, zp;
        split(r, xp, yp, zp);
        const bool interior = xp >= 1 && xp <= gr.X && yp >= 1 && yp <= gr.Y && zp >= 1 && zp <= gr.Z;
        const bool border = xp == 1 || xp == gr.X || yp == 1 || yp == gr.Y || zp == 1 || zp == gr.Z;
        if (interior && border) fold_border_voxel<T>(g, ld, gr, b, xp, yp, zp, ch * N);
    }
}

// X, Y, Z >= 2: enumerates only the border voxels - the two x faces, then the y faces without the x faces,
// then the z faces without both (12x fewer work items than rows at 192x48x48).  Every axis has at most one halo
// image, so the <= 8 image rows are loaded with independent predicated loads before anything is stored.
struct BorderEnum {
    uint32_t n_xf, n_yf, n_total;
    FastDiv by_yz, by_z, by_xz, by_xy, by_ym2;
};
__device__ __forceinline__ int halo_image(int c, int n) { return c == 1 ? 0 : (c == n ? n + 1 : -1); }

template <typename T>
__global__ void __launch_bounds__(kThreads)
halo_fold_border_kernel(T* __restrict__ g, int ld, Grid3 gr, BorderEnum e, int chunks) {
    constexpr int N = Vec<T>::N;
    const int b = blockIdx.y;
    const int vox_step = kThreads / chunks;
    const int ch = threadIdx.x % chunks, lane_vox = threadIdx.x / chunks;
    if (lane_vox >= vox_step) return;
    T* gb = g + (int64_t)b * gr.vox_p * ld + ch * N;
    for (uint32_t k = blockIdx.x * vox_step + lane_vox; k < e.n_total; k += gridDim.x * vox_step) {
        uint32_t face, rem, a, c;
        int xp, yp, zp;
        if (k < e.n_xf) {
            e.by_yz.divmod(k, face, rem);
            e.by_z.divmod(rem, a, c);
            xp = face ? gr.X : 1; yp = (int)a + 1; zp = (int)c + 1;
        } else if (k < e.n_xf + e.n_yf) {
            e.by_xz.divmod(k - e.n_xf, face, rem);
            e.by_z.divmod(rem, a, c);
            xp = (int)a + 2; yp = face ? gr.Y : 1; zp = (int)c + 1;
        } else {
            e.by_xy.divmod(k - e.n_xf - e.n_yf, face, rem);
            e.by_ym2.divmod(rem, a, c);
            xp = (int)a + 2; yp = (int)c + 2; zp = face ? gr.Z : 1;
        }
        const int xi = halo_image(xp, gr.X), yi = halo_image(yp, gr.Y), zi = halo_image(zp, gr.Z);
        float v[8][N];
        int64_t off[8];
        bool on[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int xx = (j & 4) ? xi : xp, yy = (j & 2) ? yi : yp, zz = (j & 1) ? zi : zp;
            on[j] = xx >= 0 && yy >= 0 && zz >= 0;
            off[j] = (((int64_t)xx * gr.Yp + yy) * gr.Zp + zz) * ld;
            if (on[j]) Vec<T>::load(gb + off[j], v[j]);
        }
        float acc[N], zero[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            acc[i] = v[0][i];
            zero[i] = 0.0f;
        }
#pragma unroll
        for (int j = 1; j < 8; ++j)
            if (on[j]) {
#pragma unroll
                for (int i = 0; i < N; ++i) acc[i] += v[j][i];
                Vec<T>::store(gb + off[j], zero);  // every halo row is the image of exactly one border voxel
            }
        Vec<T>::store(gb + off[0], acc);
    }
}

// ---------------------------------------------------------------- pointwise backward
template <typename T>
__device__ __forceinline__ float dsilu(float u) {
    if constexpr (sizeof(T) == 2) {
        // bf16 storage: sigmoid(u) = (1 + t)/2 with t = tanh(u/2), so silu'(u) = (1 + t + (u/2)(1 - t^2))/2 - ONE
        // special-function op (tanh.approx.f32, abs. error 2^-11, far below the bf16 rounding of the result)
        const float h = 0.5f * u;
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
        return fmaf(0.5f, fmaf(h, fmaf(-t, t, 1.0f), t), 0.5f);
    } else {
        const float s = 1.0f / (1.0f + expf(-u));
        return s * (1.0f + u * (1.0f - s));
    }
}

struct PwCoef {  // per channel: forward affine u = a*x + o; xhat = (x - mean) * rstd
    float a, o, mean, rstd, k;  // k = gamma * (scale + 1)
};

__device__ __forceinline__ PwCoef pw_coef(int b, int c, int C, int G, const double* stats, const float* gamma, const float* beta,
                                          const float* film, int film_ld, double inv_n, float eps) {
    PwCoef r{1.0f, 0.0f, 0.0f, 1.0f, 1.0f};
    if (stats) {
        const int gi = c / (C / G);
        const double mean = stats[((int64_t)b * G + gi) * 2] * inv_n;
        const double var = fma(-mean, mean, stats[((int64_t)b * G + gi) * 2 + 1] * inv_n);
        r.rstd = 1.0f / sqrtf(fmaxf((float)var, 0.0f) + eps);  // same fp32 rounding as the forward kernel
        r.mean = (float)mean;
        r.a = r.rstd * gamma[c];
        r.o = beta[c] - r.mean * r.a;
        r.k = gamma[c];
    }
    if (film) {
        const float sc = film[(int64_t)b * film_ld + c] + 1.0f;
        r.a *= sc;
        r.o = fmaf(r.o, sc, film[(int64_t)b * film_ld + C + c]);
        r.k *= sc;
    }
    return r;
}

// red[b][c] = (sum g_u, sum g_u*xhat, sum raw, sum g_out) over interior voxels (pre-zeroed): fp32 per thread (~100 voxels),
// fp32 shared atomics per CTA, one double atomic per (CTA, channel, moment).  The third sum lets the host derive the
// per-channel sum of d_raw (= the bias gradient of the convolution below) without another pass over d_raw.
// Per-channel coefficients (the only fp64 arithmetic) are computed once per block by one thread per channel; the
// streaming loop issues the 2 x U loads of U voxels before it uses any of them.
template <typename T>
__global__ void __launch_bounds__(kThreads)
pw_bwd_reduce_kernel(const T* __restrict__ g_out, int ld_g, const T* __restrict__ raw, int ld_raw,
                     const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const float* __restrict__ film, int film_ld, double* __restrict__ red, Grid3 gr, int C, int G, float eps,
                     unsigned flags, FastDiv by_z, FastDiv by_y) {
    constexpr int N = Vec<T>::N;
    constexpr int U = 4;
    extern __shared__ float sm_red[];  // [C][4] partial sums, then [C][4] = (a, o, mean, rstd)
    float* sred = sm_red;
    float* scoef = sm_red + 4 * C;
    const int b = blockIdx.y;
    {
        const double inv_n = 1.0 / ((double)(C / G) * gr.X * gr.Y * gr.Z);
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const PwCoef k = pw_coef(b, c, C, G, stats, gamma, beta, film, film_ld, inv_n, eps);
            scoef[4 * c] = k.a;
            scoef[4 * c + 1] = k.o;
            scoef[4 * c + 2] = k.mean;
            scoef[4 * c + 3] = k.rstd;
            sred[4 * c] = sred[4 * c + 1] = sred[4 * c + 2] = sred[4 * c + 3] = 0.0f;
        }
    }
    __syncthreads();
    const int chunks = C / N;
    const int vox_step = kThreads / chunks;
    const int ch = threadIdx.x % chunks, lane_vox = threadIdx.x / chunks;
    if (lane_vox < vox_step) {
        const int c0 = ch * N;
        // the loop accumulates raw moments (sum g_u, sum g_u*raw, sum raw); xhat = (raw-mean)*rstd is applied to the
        // block's partial sums at the flush, which keeps mean/rstd out of the registers of the streaming loop
        float ka[N], ko[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            ka[i] = scoef[4 * (c0 + i)];
            ko[i] = scoef[4 * (c0 + i) + 1];
        }
        const bool act = flags & TDB_PW_SILU;
        const uint32_t nvox = (uint32_t)(gr.X * gr.Y * gr.Z);
        const T* rb = raw + (int64_t)b * gr.vox_p * ld_raw + c0;
        const T* gb = g_out + (int64_t)b * gr.vox_p * ld_g + c0;
        float a1[N], a2[N], a3[N], a4[N];
#pragma unroll
        for (int i = 0; i < N; ++i) a1[i] = a2[i] = a3[i] = a4[i] = 0.0f;
        auto row_of = [&](uint32_t v) {
            uint32_t q, z, x, y;
            by_z.divmod(v, q, z);
            by_y.divmod(q, x, y);
            return ((int64_t)(x + 1) * gr.Yp + (y + 1)) * gr.Zp + (z + 1);
        };
        const uint32_t stride = gridDim.x * (uint32_t)vox_step;
        for (uint32_t v0 = blockIdx.x * (uint32_t)vox_step + lane_vox; v0 < nvox; v0 += U * stride) {
            uint4 xr[U], gv[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t v = v0 + u * stride;
                ok[u] = v < nvox;
                const int64_t r = row_of(ok[u] ? v : nvox - 1);
                xr[u] = Vec<T>::load_raw(rb + r * ld_raw);
                gv[u] = Vec<T>::load_raw(gb + r * ld_g);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!ok[u]) continue;
                float x[N], g[N];
                Vec<T>::unpack(xr[u], x);
                Vec<T>::unpack(gv[u], g);
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const float uu = fmaf(ka[i], x[i], ko[i]);
                    const float gu = act ? g[i] * dsilu<T>(uu) : g[i];
                    a1[i] += gu;
                    a2[i] = fmaf(gu, x[i], a2[i]);
                    a3[i] += x[i];
                    a4[i] += g[i];
                }
            }
        }
        // lanes of a warp that own the same channel chunk (lane % chunks) are combined by shuffles first, so that only
        // `chunks` lanes per warp touch the shared accumulators (the per-thread shared atomics used to dominate this kernel)
        const bool shuffle = chunks <= 32 && (chunks & (chunks - 1)) == 0;
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            float v1 = a1[i], v2 = a2[i], v3 = a3[i], v4 = a4[i];
            if (shuffle) {
                for (int o = 16; o >= chunks; o >>= 1) {
                    v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                    v2 += __shfl_xor_sync(0xffffffffu, v2, o);
                    v3 += __shfl_xor_sync(0xffffffffu, v3, o);
                    v4 += __shfl_xor_sync(0xffffffffu, v4, o);
                }
            }
            if (!shuffle || lane < chunks) {
                atomicAdd(&sred[4 * (c0 + i)], v1);
                atomicAdd(&sred[4 * (c0 + i) + 1], v2);
                atomicAdd(&sred[4 * (c0 + i) + 2], v3);
                atomicAdd(&sred[4 * (c0 + i) + 3], v4);
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float mean = scoef[4 * c + 2], rstd = scoef[4 * c + 3];
        const float s1 = sred[4 * c], s2 = sred[4 * c + 1], s3 = sred[4 * c + 2];
        double* dst = red + ((int64_t)b * C + c) * 4;
        atomicAdd(dst, (double)s1);
        atomicAdd(dst + 1, (double)(rstd * (s2 - mean * s1)));  // sum g_u * xhat
        atomicAdd(dst + 2, (double)s3);
        atomicAdd(dst + 3, (double)sred[4 * c + 3]);            // sum g_out (bias gradient of a residual projection)
    }
}

// d_raw = rstd * (k*g_u - m1 - xhat*m2) on interior rows, 0 on halo rows.  grp[b][g] = (m1, m2) fp32.
// Fused form (fin.red != nullptr): the group sums m1, m2 are derived from the reduce kernel's red[] in every block's
// prologue (one thread per channel, double shared-memory atomics), and gridDim.x - 1 is an EXTRA block per sample that
// streams nothing and writes the parameter gradients instead (what pw_bwd_finalize_kernel computes).  The separate
// one-block finalize launch sat on the critical path between reduce and apply with a serial chain of dependent global
// loads per (sample, group) - measured 1.3 ms per training step for its 23 launches.
struct PwFinal {
    const double* red;  // [B][C][4] = (A1, A2, Sx, Sg)
    float *colsum, *gw, *gb, *gsum, *dfilm;
    int dfilm_ld, B;
};

template <typename T>
__global__ void __launch_bounds__(kThreads)
pw_bwd_apply_kernel(const T* __restrict__ g_out, int ld_g, const T* __restrict__ raw, int ld_raw,
                    const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                    const float* __restrict__ film, int film_ld, const float* __restrict__ grp, T* __restrict__ d_raw, int ld_d,
                    Grid3 gr, int C, int G, float eps, unsigned flags, RowSplit split, int chunks, PwFinal fin) {
    constexpr int N = Vec<T>::N;
    constexpr int U = 4;
    extern __shared__ double s_apply_d[];  // fused form: [4*B*G] doubles (mean, rstd, m1, m2 per (sample, group)); then the floats
    const bool fused = fin.red != nullptr;
    const int n_d = fused ? 4 * fin.B * G : 0;
    float* s_apply = reinterpret_cast<float*>(s_apply_d + n_d);  // [C][5] = (a, o, c1, c2, c3): d_raw = c1*g_u + c2*raw + c3
    const int b = blockIdx.y;
    const int cpg = C / G;
    const double nv = (double)gr.X * gr.Y * gr.Z;
    const int n_stream = fused ? (int)gridDim.x - 1 : (int)gridDim.x;  // blocks that stream rows
    if (fused) {
        // (m1, m2) of this sample's groups - or of every sample's in the extra block of sample 0 (cross-sample sums)
        const bool extra = (int)blockIdx.x == n_stream;
        const int b_lo = (extra && b == 0) ? 0 : b, b_hi = (extra && b == 0) ? fin.B : b + 1;
        for (int i = threadIdx.x; i < 4 * fin.B * G; i += blockDim.x) s_apply_d[i] = 0.0;
        __syncthreads();
        for (int i = threadIdx.x; i < (b_hi - b_lo) * C; i += blockDim.x) {
            const int bb = b_lo + i / C, c = i % C;
            double k = (double)gamma[c];
            if (film) k *= (double)film[(int64_t)bb * film_ld + c] + 1.0;
            const double* r = fin.red + ((int64_t)bb * C + c) * 4;
            atomicAdd(&s_apply_d[4 * (bb * G + c / cpg) + 2], k * r[0]);
            atomicAdd(&s_apply_d[4 * (bb * G + c / cpg) + 3], k * r[1]);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < (b_hi - b_lo) * G; i += blockDim.x) {
            const int q = (b_lo + i / G) * G + i % G;
            const double n = (double)cpg * nv;
            const double mean = stats[(int64_t)q * 2] / n;
            const double var = fmax(stats[(int64_t)q * 2 + 1] / n - mean * mean, 0.0);
            s_apply_d[4 * q] = mean;
            s_apply_d[4 * q + 1] = 1.0 / sqrt(var + (double)eps);
            s_apply_d[4 * q + 2] /= n;
            s_apply_d[4 * q + 3] /= n;
        }
        __syncthreads();
        if (extra) {
            // parameter gradients (same arithmetic as pw_bwd_finalize_kernel): FiLM scale / shift of this sample, and in the
            // block of sample 0 the sums over samples
            for (int c = threadIdx.x; c < C; c += blockDim.x) {
                const double* r = fin.red + ((int64_t)b * C + c) * 4;
                if (fin.dfilm) {
                    fin.dfilm[(int64_t)b * fin.dfilm_ld + c] = (float)((double)gamma[c] * r[1] + (double)beta[c] * r[0]);
                    fin.dfilm[(int64_t)b * fin.dfilm_ld + C + c] = (float)r[0];
                }
                if (b != 0) continue;
                double cs = 0.0, w = 0.0, bsum = 0.0, gs = 0.0;
                for (int bb = 0; bb < fin.B; ++bb) {
                    const double* rr = fin.red + ((int64_t)bb * C + c) * 4;
                    const double* q = s_apply_d + 4 * (bb * G + c / cpg);
                    const double sc = film ? (double)film[(int64_t)bb * film_ld + c] + 1.0 : 1.0;
                    const double k = (double)gamma[c] * sc;
                    gs += rr[3];
                    cs += q[1] * (k * rr[0] - nv * q[2] - q[3] * q[1] * (rr[2] - nv * q[0]));
                    w += sc * rr[1];
                    bsum += sc * rr[0];
                }
                fin.colsum[c] = (float)cs;
                fin.gw[c] = (float)w;
                fin.gb[c] = (float)bsum;
                if (fin.gsum) fin.gsum[c] = (float)gs;
            }
            return;
        }
    }
    {
        const double inv_n = 1.0 / ((double)cpg * nv);
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const PwCoef k = pw_coef(b, c, C, G, stats, gamma, beta, film, film_ld, inv_n, eps);
            const int gi = c / cpg;
            float m1 = 0.0f, m2 = 0.0f;
            if (stats) {
                m1 = fused ? (float)s_apply_d[4 * (b * G + gi) + 2] : grp[((int64_t)b * G + gi) * 2];
                m2 = fused ? (float)s_apply_d[4 * (b * G + gi) + 3] : grp[((int64_t)b * G + gi) * 2 + 1];
            }
            const float c2 = stats ? -k.rstd * k.rstd * m2 : 0.0f;
            s_apply[5 * c] = k.a;
            s_apply[5 * c + 1] = k.o;
            s_apply[5 * c + 2] = stats ? k.rstd * k.k : 1.0f;
            s_apply[5 * c + 3] = c2;
            s_apply[5 * c + 4] = stats ? -k.rstd * m1 - c2 * k.mean : 0.0f;
        }
    }
    __syncthreads();
    const int vox_step = kThreads / chunks;
    const int ch = threadIdx.x % chunks, lane_vox = threadIdx.x / chunks;
    if (lane_vox >= vox_step) return;
    const int c0 = ch * N;
    float ka[N], ko[N], c1[N], c2[N], c3[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        ka[i] = s_apply[5 * (c0 + i)];
        ko[i] = s_apply[5 * (c0 + i) + 1];
        c1[i] = s_apply[5 * (c0 + i) + 2];
        c2[i] = s_apply[5 * (c0 + i) + 3];
        c3[i] = s_apply[5 * (c0 + i) + 4];
    }
    const bool act = flags & TDB_PW_SILU;
    const uint32_t total = (uint32_t)gr.vox_p;
    const uint32_t stride = (uint32_t)n_stream * (uint32_t)vox_step;
    const int64_t base = (int64_t)b * gr.vox_p;
    for (uint32_t r0 = blockIdx.x * (uint32_t)vox_step + lane_vox; r0 < total; r0 += U * stride) {
        uint4 xr[U], gv[U];
        bool ok[U], inter[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t r = r0 + u * stride;
            ok[u] = r < total;
            int xp, yp, zp;
            split(ok[u] ? r : total - 1, xp, yp, zp);
            inter[u] = ok[u] && xp >= 1 && xp <= gr.X && yp >= 1 && yp <= gr.Y && zp >= 1 && zp <= gr.Z;
            if (inter[u]) {
                xr[u] = Vec<T>::load_raw(raw + (base + r) * ld_raw + c0);
                gv[u] = Vec<T>::load_raw(g_out + (base + r) * ld_g + c0);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            float o[N];
            if (inter[u]) {
                float x[N], g[N];
                Vec<T>::unpack(xr[u], x);
                Vec<T>::unpack(gv[u], g);
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const float uu = fmaf(ka[i], x[i], ko[i]);
                    const float gu = act ? g[i] * dsilu<T>(uu) : g[i];
                    o[i] = fmaf(c1[i], gu, fmaf(c2[i], x[i], c3[i]));
                }
            } else {
#pragma unroll
                for (int i = 0; i < N; ++i) o[i] = 0.0f;
            }
            Vec<T>::store(d_raw + (base + r0 + u * stride) * ld_d + c0, o);
        }
    }
}

// ---------------------------------------------------------------- pointwise backward: group / parameter bookkeeping
// One block.  From red[b][c] = (A1, A2, Sx, Sg) and the forward moments: grp[b][g] = (m1, m2) for the apply pass,
// colsum[c] = sum over samples and voxels of d_raw (bias gradient of the convolution that produced raw),
// gw[c] / gb[c] = GroupNorm weight / bias gradients, dfilm[b][c], dfilm[b][C + c] = FiLM scale / shift gradients,
// gsum[c] = sum of the incoming gradient g_out (bias gradient of a 1x1 residual projection applied to the same output).
__global__ void __launch_bounds__(kThreads)
pw_bwd_finalize_kernel(const double* __restrict__ red, const double* __restrict__ stats, const float* __restrict__ gamma,
                       const float* __restrict__ beta, const float* __restrict__ film, int film_ld, float* __restrict__ grp,
                       float* __restrict__ colsum, float* __restrict__ gw, float* __restrict__ gb, float* __restrict__ gsum,
                       float* __restrict__ dfilm, int dfilm_ld, int B, int C, int G, double nv, double eps) {
    extern __shared__ double sg[];  // [B*G][4] = mean, rstd, m1, m2
    const int cpg = C / G;
    const double n = (double)cpg * nv;
    for (int i = threadIdx.x; i < B * G; i += blockDim.x) {
        const int b = i / G, g = i % G;
        const double mean = stats[(int64_t)i * 2] / n;
        const double var = fmax(stats[(int64_t)i * 2 + 1] / n - mean * mean, 0.0);
        const double rstd = 1.0 / sqrt(var + eps);
        double m1 = 0.0, m2 = 0.0;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
            double k = (double)gamma[c];
            if (film) k *= (double)film[(int64_t)b * film_ld + c] + 1.0;
            m1 += k * red[((int64_t)b * C + c) * 4];
            m2 += k * red[((int64_t)b * C + c) * 4 + 1];
        }
        m1 /= n;
        m2 /= n;
        sg[4 * i] = mean; sg[4 * i + 1] = rstd; sg[4 * i + 2] = m1; sg[4 * i + 3] = m2;
        grp[2 * i] = (float)m1;
        grp[2 * i + 1] = (float)m2;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        double cs = 0.0, w = 0.0, bsum = 0.0, gs = 0.0;
        for (int b = 0; b < B; ++b) {
            const double* r = red + ((int64_t)b * C + c) * 4;
            gs += r[3];
            const double* q = sg + 4 * (b * G + g);
            const double sc = film ? (double)film[(int64_t)b * film_ld + c] + 1.0 : 1.0;
            const double k = (double)gamma[c] * sc;
            cs += q[1] * (k * r[0] - nv * q[2] - q[3] * q[1] * (r[2] - nv * q[0]));
            w += sc * r[1];
            bsum += sc * r[0];
            if (dfilm) {
                dfilm[(int64_t)b * dfilm_ld + c] = (float)((double)gamma[c] * r[1] + (double)beta[c] * r[0]);
                dfilm[(int64_t)b * dfilm_ld + C + c] = (float)r[0];
            }
        }
        colsum[c] = (float)cs;
        gw[c] = (float)w;
        gb[c] = (float)bsum;
        if (gsum) gsum[c] = (float)gs;
    }
}

// ---------------------------------------------------------------- conv weight gradient
// dW[tap][ci][co] += sum_{p interior} in[p + delta(tap)][ci] * d_out[p][co]  (halo rows of d_out are masked).
// Block = one (tap, 64x64 tile) over a slice of rows; 4x4 register tile per thread; fp32 atomics at the end.
struct InteriorTest {
    FastDiv by_vox, by_z, by_y;
    int X, Y, Z;
    __device__ __forceinline__ bool operator()(int64_t p) const {
        uint32_t bb, r, q, zp, xp, yp;
        by_vox.divmod((uint32_t)p, bb, r);
        by_z.divmod(r, q, zp);
        by_y.divmod(q, xp, yp);
        return xp >= 1u && xp <= (uint32_t)X && yp >= 1u && yp <= (uint32_t)Y && zp >= 1u && zp <= (uint32_t)Z;
    }
};

template <typename T>
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const T* __restrict__ in, int ld_in, const T* __restrict__ d_out, int ld_do, float* __restrict__ dw,
                  int64_t rows, int yz_p, int z_p, int Cin, int Cout, int ntaps, int ci_tiles, int co_tiles, int rows_per_block,
                  InteriorTest interior) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tile = blockIdx.y;
    const int tap = tile / (ci_tiles * co_tiles);
    const int ci0 = ((tile / co_tiles) % ci_tiles) * 64;
    const int co0 = (tile % co_tiles) * 64;
    int64_t delta = 0;
    if (ntaps == 27) delta = (int64_t)(tap / 9 - 1) * yz_p + (int64_t)((tap / 3) % 3 - 1) * z_p + (tap % 3 - 1);
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int lr = tid / 16, lc = (tid % 16) * 4;  // loader: row 0..15, 4 consecutive channels
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r_end = min(rows, r_begin + rows_per_block);
    for (int64_t r0 = r_begin; r0 < r_end; r0 += 16) {
        const int64_t p = r0 + lr;
        float a[4] = {0.f, 0.f, 0.f, 0.f}, bq[4] = {0.f, 0.f, 0.f, 0.f};
        if (p < r_end && interior(p)) {
            const int64_t q = p + delta;
            if (q >= 0 && q < rows) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (ci0 + lc + i < Cin) a[i] = (float)in[q * ld_in + ci0 + lc + i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (co0 + lc + i < Cout) bq[i] = (float)d_out[p * ld_do + co0 + lc + i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            As[lr][lc + i] = a[i];
            Bs[lr][lc + i] = bq[i];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ci = ci0 + ty * 4 + i, co = co0 + tx * 4 + j;
            if (ci < Cin && co < Cout) atomicAdd(&dw[((int64_t)tap * Cin + ci) * Cout + co], acc[i][j]);
        }
}

// bf16 variant on the warp-level tensor-core path (wmma, fp32 accumulators): same tiling (64x64 per tap over a
// row slice, 8 warps = 4 (ci) x 2 (co), two 16x16 accumulators each), 64-row K chunks staged in shared memory
// with 16-byte loads.  d_out must be ZERO on halo rows and `in` must be readable for Yp*Zp+Zp+1 rows before and
// after the grid (the halo-grid workspace guarantees both), so no per-row masking is needed.
// (Stepping stone: a tcgen05 wgrad with MN-major operands is the follow-up.)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int n = valid ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(n) : "memory");
}

__global__ void __launch_bounds__(256)
conv_wgrad_wmma_kernel(const bf16* __restrict__ in, int ld_in, const bf16* __restrict__ d_out, int ld_do, float* __restrict__ dw,
                       int64_t rows, int yz_p, int z_p, int Cin, int Cout, int ntaps, int ci_tiles, int co_tiles, int rows_per_block) {
    using namespace nvcuda;
    constexpr int KR = 64, LDS = 64 + 8, STAGE = 2 * KR * LDS;
    __shared__ __align__(32) bf16 sbuf[2 * STAGE];  // 2 stages x {A [row][ci], B [row][co]}
    const int tile = blockIdx.y;
    const int tap = tile / (ci_tiles * co_tiles);
    const int ci0 = ((tile / co_tiles) % ci_tiles) * 64;
    const int co0 = (tile % co_tiles) * 64;
    int64_t delta = 0;
    if (ntaps == 27) delta = (int64_t)(tap / 9 - 1) * yz_p + (int64_t)((tap / 3) % 3 - 1) * z_p + (tap % 3 - 1);
    const int tid = threadIdx.x, warp = tid / 32;
    const int wm = warp % 4, wn = warp / 4;  // warp tile: ci [16*wm, +16), co [32*wn, +32)
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[2];
    wmma::fill_fragment(acc[0], 0.0f);
    wmma::fill_fragment(acc[1], 0.0f);
    const int lr = tid / 8, lc = (tid % 8) * 8;  // loader: rows lr and lr+32, 8 consecutive channels
    const bool a_ok = ci0 + lc < Cin, b_ok = co0 + lc < Cout;
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r_end = min(rows, r_begin + rows_per_block);

    auto prefetch = [&](int stage, int64_t r0) {
        bf16* As = sbuf + stage * STAGE;
        bf16* Bs = As + KR * LDS;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t p = r0 + lr + 32 * h;
            const bool ok = p < r_end;
            const int64_t pa = ok ? p + delta : r_begin, pb = ok ? p : r_begin;  // keep the address valid when masked
            cp_async16(As + (lr + 32 * h) * LDS + lc, in + pa * ld_in + ci0 + (a_ok ? lc : 0), ok && a_ok);
            cp_async16(Bs + (lr + 32 * h) * LDS + lc, d_out + pb * ld_do + co0 + (b_ok ? lc : 0), ok && b_ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    int stage = 0;
    if (r_begin < r_end) prefetch(0, r_begin);
    for (int64_t r0 = r_begin; r0 < r_end; r0 += KR, stage ^= 1) {
        if (r0 + KR < r_end) {
            prefetch(stage ^ 1, r0 + KR);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const bf16* As = sbuf + stage * STAGE;
        const bf16* Bs = As + KR * LDS;
#pragma unroll
        for (int k = 0; k < KR; k += 16) {
            wmma::fragment<wmma::matrix_a, 16, 16, 16, bf16, wmma::col_major> fa;  // (m = ci, k = row)
            wmma::load_matrix_sync(fa, As + k * LDS + 16 * wm, LDS);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16, wmma::row_major> fb;  // (k = row, n = co)
                wmma::load_matrix_sync(fb, Bs + k * LDS + 32 * wn + 16 * j, LDS);
                wmma::mma_sync(acc[j], fa, fb, acc[j]);
            }
        }
        __syncthreads();
    }
    // accumulators -> shared (fp32 scratch over the stage buffers) -> fp32 atomics
    float* cs = reinterpret_cast<float*>(sbuf);
    static_assert(sizeof(sbuf) >= 64 * 64 * sizeof(float), "scratch");
#pragma unroll
    for (int j = 0; j < 2; ++j) wmma::store_matrix_sync(cs + (16 * wm) * 64 + 32 * wn + 16 * j, acc[j], 64, wmma::mem_row_major);
    __syncthreads();
    for (int i = tid; i < 64 * 64; i += 256) {
        const int ci = ci0 + i / 64, co = co0 + i % 64;
        if (ci < Cin && co < Cout) atomicAdd(&dw[((int64_t)tap * Cin + ci) * Cout + co], cs[i]);
    }
}

// ---------------------------------------------------------------- trilinear backward (gather)
struct Lerp {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ Lerp axis_lerp(int o, int n_in, float scale) {
    const float src = scale * (float)o;
    Lerp r;
    r.i0 = min((int)src, n_in - 1);
    r.i1 = r.i0 + (r.i0 < n_in - 1 ? 1 : 0);
    r.l1 = src - (float)r.i0;
    r.l0 = 1.0f - r.l1;
    return r;
}
// outputs o (and their weights) that read input index i along one axis (host guarantees <= 12: scale >= 0.2)
__device__ __forceinline__ int axis_sources(int i, int n_in, int n_out, float scale, int (&oo)[12], float (&ww)[12]) {
    int k = 0;
    int lo = 0, hi = n_out - 1;
    if (scale > 0.0f) {
        lo = max(0, (int)floorf((float)(i - 1) / scale) - 1);
        hi = min(n_out - 1, (int)ceilf((float)(i + 1) / scale) + 1);
    }
    for (int o = lo; o <= hi && k < 12; ++o) {
        const Lerp l = axis_lerp(o, n_in, scale);
        float w = 0.0f;
        if (l.i0 == i) w += l.l0;
        if (l.i1 == i) w += l.l1;
        if (w != 0.0f) {
            oo[k] = o;
            ww[k] = w;
            ++k;
        }
    }
    return k;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
trilinear_bwd_kernel(const T* __restrict__ g_out, int ld_g, Grid3 go, T* __restrict__ d_in, int ld_d, Grid3 gi, int C,
                     RowSplit split, int chunks, float sx, float sy, float sz, int accumulate) {
    constexpr int N = Vec<T>::N;
    const int b = blockIdx.y;
    const int vox_step = kThreads / chunks;
    const int ch = threadIdx.x % chunks, lane_vox = threadIdx.x / chunks;
    if (lane_vox >= vox_step) return;
    const int c0 = ch * N;
    for (uint32_t r = blockIdx.x * vox_step + lane_vox; r < (uint32_t)gi.vox_p; r += gridDim.x * vox_step) {
        int xp, yp, zp;
        split(r, xp, yp, zp);
        const bool interior = xp >= 1 && xp <= gi.X && yp >= 1 && yp <= gi.Y && zp >= 1 && zp <= gi.Z;
        if (accumulate && !interior) continue;  // d_in += ...: halo rows stay as they are
        float acc[N];
#pragma unroll
        for (int i = 0; i < N; ++i) acc[i] = 0.0f;
        T* dst = d_in + ((int64_t)b * gi.vox_p + r) * ld_d + c0;
        if (interior) {
            if (accumulate) Vec<T>::load(dst, acc);
            int ox[12], oy[12], oz[12];
            float wx[12], wy[12], wz[12];
            const int nx = axis_sources(xp - 1, gi.X, go.X, sx, ox, wx);
            const int ny = axis_sources(yp - 1, gi.Y, go.Y, sy, oy, wy);
            const int nz = axis_sources(zp - 1, gi.Z, go.Z, sz, oz, wz);
            for (int a = 0; a < nx; ++a)
                for (int bb = 0; bb < ny; ++bb)
                    for (int c = 0; c < nz; ++c) {
                        const float w = wx[a] * wy[bb] * wz[c];
                        float v[N];
                        Vec<T>::load(g_out + go.row(b, ox[a], oy[bb], oz[c]) * ld_g + c0, v);
#pragma unroll
                        for (int i = 0; i < N; ++i) acc[i] = fmaf(w, v[i], acc[i]);
                    }
        }
        Vec<T>::store(dst, acc);
    }
}

// Line walker (the form the training step runs on): a thread owns one (y, z) line of the INPUT-side grid, one 16-byte
// channel chunk and a segment of x.  The forward op is out[o] = l0*in[i0] + l1*in[i1] per axis, so along y and z the
// line gathers its <= 6 source lines per axis (tables per block in shared memory), and along x it walks the output-side
// x index once, scattering every y/z-reduced value into the two running accumulators of in[i0], in[i0+1] and storing an
// input voxel as soon as the walk has passed it.  Consecutive threads are consecutive z, so at every step a block
// touches contiguous runs of memory; each source row is read by ~2x2 lines (instead of by every input voxel it touches:
// 4^3 reads per voxel for the adjoint of the 2x upsampling) and nothing is re-derived per voxel.  accumulate: d_in +=
// (halo rows untouched) - the skip connection's gradient is added in the same pass.
constexpr int TB_S = 6;  // sources per axis (scale >= 0.4: at most ceil(2 / scale) + 1)
constexpr int TB_G = 4;  // z sources fetched per group (predicated-off slots still cost issue slots)

template <typename T>
__global__ void __launch_bounds__(kThreads, 3)
trilinear_bwd_walk_kernel(const T* __restrict__ g_out, int ld_g, Grid3 go, T* __restrict__ d_in, int ld_d, Grid3 gi, int chunks,
                          int seg_len, float sx, float sy, float sz, int accumulate, FastDiv by_zp) {
    constexpr int N = Vec<T>::N;
    extern __shared__ int s_tab[];  // per haloed y, then per haloed z: [TB_S] source offsets, [TB_S] weights, count
    constexpr int ENT = 2 * TB_S + 1;
    const int b = blockIdx.y;
    for (int e = threadIdx.x; e < gi.Yp + gi.Zp; e += blockDim.x) {
        const bool is_y = e < gi.Yp;
        const int p = is_y ? e : e - gi.Yp;  // haloed coordinate
        const int n_in = is_y ? gi.Y : gi.Z, n_out = is_y ? go.Y : go.Z;
        int oo[12];
        float ww[12];
        int n = 0;
        if (p >= 1 && p <= n_in) n = axis_sources(p - 1, n_in, n_out, is_y ? sy : sz, oo, ww);
        if (n > TB_S) n = TB_S;  // excluded by the host (scale >= 0.4)
        int* t = s_tab + e * ENT;
        for (int k = 0; k < TB_S; ++k) {
            t[k] = k < n ? (oo[k] + 1) * (is_y ? go.Zp : 1) : 0;  // row offset of the source inside an x plane
            t[TB_S + k] = __float_as_int(k < n ? ww[k] : 0.0f);
        }
        t[2 * TB_S] = n;
    }
    __syncthreads();
    const int ch = threadIdx.x % chunks;
    const uint32_t col = blockIdx.x * (uint32_t)(kThreads / chunks) + threadIdx.x / chunks;
    if (threadIdx.x / chunks >= kThreads / chunks || col >= (uint32_t)(gi.Yp * gi.Zp)) return;
    uint32_t ypu, zpu;
    by_zp.divmod(col, ypu, zpu);
    const int yp = (int)ypu, zp = (int)zpu;
    const int c0 = ch * N;
    const int i_lo = blockIdx.z * seg_len, i_hi = min(gi.X, i_lo + seg_len);  // input-side x range of this segment
    const int64_t plane_d = (int64_t)gi.Yp * gi.Zp * ld_d;
    T* dline = d_in + ((int64_t)b * gi.vox_p + (int64_t)yp * gi.Zp + zp) * ld_d + c0;  // x = halo plane 0 of this line
    float zero[N];
#pragma unroll
    for (int i = 0; i < N; ++i) zero[i] = 0.0f;
    const bool interior = yp >= 1 && yp <= gi.Y && zp >= 1 && zp <= gi.Z;
    if (!interior) {
        if (!accumulate) {
            const int x_end = i_hi == gi.X ? gi.Xp : i_hi + 1;
            for (int xp = i_lo == 0 ? 0 : i_lo + 1; xp < x_end; ++xp) Vec<T>::store(dline + xp * plane_d, zero);
        }
        return;
    }
    const int* ty = s_tab + yp * ENT;
    const int* tz = s_tab + (gi.Yp + zp) * ENT;
    const int ny = ty[2 * TB_S], nz = tz[2 * TB_S];
    int zoff[TB_S];
    float wz[TB_S];
#pragma unroll
    for (int k = 0; k < TB_S; ++k) {
        zoff[k] = tz[k] * ld_g;  // element offset inside an x plane (< 2^31: checked on the host)
        wz[k] = __int_as_float(tz[TB_S + k]);
    }
    const T* gb = g_out + (int64_t)b * go.vox_p * ld_g + c0;
    const int64_t plane_g = (int64_t)go.Yp * go.Zp;
    float acc0[N], acc1[N];
#pragma unroll
    for (int i = 0; i < N; ++i) acc0[i] = acc1[i] = 0.0f;
    int cur = i_lo;  // input-side x index held by acc0 (acc1: cur + 1)
    // accumulate: the existing values of the next two voxels of the line are fetched ahead of their flush
    uint4 e0 = make_uint4(0u, 0u, 0u, 0u), e1 = e0;
    auto existing = [&](int i) { return (accumulate && i < i_hi) ? Vec<T>::load_raw(dline + (int64_t)(i + 1) * plane_d) : make_uint4(0u, 0u, 0u, 0u); };
    e0 = existing(cur);
    e1 = existing(cur + 1);
    auto flush = [&]() {
        T* dst = dline + (int64_t)(cur + 1) * plane_d;
        if (accumulate) {
            float e[N];
            Vec<T>::unpack(e0, e);
#pragma unroll
            for (int i = 0; i < N; ++i) acc0[i] += e[i];
        }
        Vec<T>::store(dst, acc0);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            acc0[i] = acc1[i];
            acc1[i] = 0.0f;
        }
        ++cur;
        e0 = e1;
        e1 = existing(cur + 1);
    };
    // output-side x range that touches [i_lo, i_hi): a little wider than needed, contributions outside are dropped
    int o_lo = 0, o_hi = go.X - 1;
    if (sx > 0.0f) {
        o_lo = max(0, (int)floorf((float)(i_lo - 1) / sx) - 1);
        o_hi = min(go.X - 1, (int)ceilf((float)i_hi / sx) + 1);
    }
    for (int ox = o_lo; ox <= o_hi; ++ox) {
        const Lerp lx = axis_lerp(ox, gi.X, sx);
        if (lx.i1 < i_lo) continue;
        if (lx.i0 >= i_hi) break;
        float s[N];
#pragma unroll
        for (int i = 0; i < N; ++i) s[i] = 0.0f;
        const T* gx = gb + (int64_t)(ox + 1) * plane_g * ld_g;
        for (int a = 0; a < ny; ++a) {
            const T* gy = gx + ty[a] * ld_g;
            const float wya = __int_as_float(ty[TB_S + a]);
#pragma unroll
            for (int k0 = 0; k0 < TB_S; k0 += TB_G) {  // TB_G loads in flight: the 2x up-sampling has 4 sources per axis almost everywhere
                if (k0 >= nz) break;
                uint4 raw[TB_G];
#pragma unroll
                for (int k = 0; k < TB_G; ++k)
                    if (k0 + k < TB_S && k0 + k < nz) raw[k] = Vec<T>::load_raw(gy + zoff[k0 + k]);
#pragma unroll
                for (int k = 0; k < TB_G; ++k)
                    if (k0 + k < TB_S && k0 + k < nz) {
                        float v[N];
                        Vec<T>::unpack(raw[k], v);
                        const float w = wya * wz[k0 + k];
#pragma unroll
                        for (int i = 0; i < N; ++i) s[i] = fmaf(w, v[i], s[i]);
                    }
            }
        }
        while (cur < lx.i0 && cur < i_hi) flush();
        // targets (i0, l0) and (i1, l1): acc0 holds cur, acc1 holds cur + 1, anything else lies outside this segment
        const float w0 = (lx.i0 == cur ? lx.l0 : 0.0f) + (lx.i1 == cur ? lx.l1 : 0.0f);
        const float w1 = lx.i1 == cur + 1 ? lx.l1 : 0.0f;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            acc0[i] = fmaf(w0, s[i], acc0[i]);
            acc1[i] = fmaf(w1, s[i], acc1[i]);
        }
    }
    while (cur < i_hi) flush();
    if (!accumulate) {
        if (i_lo == 0) Vec<T>::store(dline, zero);
        if (i_hi == gi.X) Vec<T>::store(dline + (int64_t)(gi.Xp - 1) * plane_d, zero);
    }
}

// ---------------------------------------------------------------- attention backward
// ATT_SPLIT CTAs per (sample, head); q,k,v,dO and the S x S probability / score-gradient matrices live in smem.
constexpr int ATT_SPLIT = 4;
template <typename T>
__global__ void __launch_bounds__(256)
attention_bwd_kernel(const T* __restrict__ qkv, int ld_qkv, const T* __restrict__ d_out, int ld_do, T* __restrict__ d_qkv,
                     int ld_dq, Grid3 g, int heads, int S) {
    constexpr int DH = 32;
    extern __shared__ float sm[];
    float* sq = sm;
    float* sk = sq + (size_t)S * (DH + 1);
    float* sv = sk + (size_t)S * (DH + 1);
    float* sdo = sv + (size_t)S * (DH + 1);
    float* sp = sdo + (size_t)S * (DH + 1);  // [S][S+1] probabilities
    float* sds = sp + (size_t)S * (S + 1);   // [S][S+1] dS
    // ATT_SPLIT CTAs per (sample, head): each recomputes the (cheap) probability / score-gradient rows and produces a
    // quarter of the dQ / dK / dV elements - the kernel sits on the critical path of the backward walk with B * heads CTAs
    const int bh = blockIdx.x / ATT_SPLIT, part = blockIdx.x % ATT_SPLIT;
    const int b = bh / heads, h = bh % heads;
    const int hid = heads * DH;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nwarps = blockDim.x / 32;
    const float scale = rsqrtf((float)DH);
    auto row_of = [&](int s) {
        const int z = s % g.Z, y = (s / g.Z) % g.Y, x = s / (g.Z * g.Y);
        return g.row(b, x, y, z);
    };
    for (int i = threadIdx.x; i < S * DH; i += blockDim.x) {
        const int s = i / DH, d = i % DH;
        const int64_t row = row_of(s);
        const T* p = qkv + row * ld_qkv + h * DH + d;
        sq[s * (DH + 1) + d] = (float)p[0];
        sk[s * (DH + 1) + d] = (float)p[hid];
        sv[s * (DH + 1) + d] = (float)p[2 * hid];
        sdo[s * (DH + 1) + d] = (float)d_out[row * ld_do + h * DH + d];
    }
    __syncthreads();
    // P and dS rows
    for (int i = warp; i < S; i += nwarps) {
        float mx = -INFINITY;
        for (int j = lane; j < S; j += 32) {
            float acc = 0.0f;
#pragma unroll
            for (int d = 0; d < DH; ++d) acc = fmaf(sq[i * (DH + 1) + d], sk[j * (DH + 1) + d], acc);
            acc *= scale;
            sp[i * (S + 1) + j] = acc;
            mx = fmaxf(mx, acc);
        }
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int j = lane; j < S; j += 32) {
            const float e = expf(sp[i * (S + 1) + j] - mx);
            sp[i * (S + 1) + j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        float dsum = 0.0f;
        for (int j = lane; j < S; j += 32) {
            const float p = sp[i * (S + 1) + j] * inv;
            sp[i * (S + 1) + j] = p;
            float dp = 0.0f;
#pragma unroll
            for (int d = 0; d < DH; ++d) dp = fmaf(sdo[i * (DH + 1) + d], sv[j * (DH + 1) + d], dp);
            sds[i * (S + 1) + j] = dp;
            dsum = fmaf(p, dp, dsum);
        }
        dsum = warp_sum(dsum);
        for (int j = lane; j < S; j += 32) sds[i * (S + 1) + j] = sp[i * (S + 1) + j] * (sds[i * (S + 1) + j] - dsum) * scale;
    }
    __syncthreads();
    // dQ[i] = sum_j dS[i][j] K[j];  dK[j] = sum_i dS[i][j] Q[i];  dV[j] = sum_i P[i][j] dO[i]
    const int per = (S * DH + ATT_SPLIT - 1) / ATT_SPLIT;
    for (int idx = part * per + threadIdx.x; idx < min(S * DH, (part + 1) * per); idx += blockDim.x) {
        const int s = idx / DH, d = idx % DH;
        float dq = 0.0f, dk = 0.0f, dv = 0.0f;
        for (int j = 0; j < S; ++j) {
            dq = fmaf(sds[s * (S + 1) + j], sk[j * (DH + 1) + d], dq);
            dk = fmaf(sds[j * (S + 1) + s], sq[j * (DH + 1) + d], dk);
            dv = fmaf(sp[j * (S + 1) + s], sdo[j * (DH + 1) + d], dv);
        }
        T* o = d_qkv + row_of(s) * ld_dq + h * DH + d;
        o[0] = (T)dq;
        o[hid] = (T)dk;
        o[2 * hid] = (T)dv;
    }
}

// Sequences whose S x S matrices do not fit in shared memory (S > 135): nothing quadratic is stored.  grid = (B * heads,
// row chunks).  Phase 1 (K, V resident): every CTA recomputes the softmax statistics of ALL rows - log-sum-exp and
// sum_j P_ij dP_ij, S^2 * 64 FMAs, cheap - and writes dQ for the rows of its chunk.  Phase 2 (Q, dO resident in the same
// buffers): dK and dV of the chunk's keys, one warp per key, the column of P / dS rebuilt from the row statistics.
constexpr int ATTB_WARPS = 8;
constexpr int ATTB_ROWS = 32;  // rows (phase 1) and keys (phase 2) per CTA
template <typename T>
__global__ void __launch_bounds__(ATTB_WARPS * 32)
attention_bwd_stream_kernel(const T* __restrict__ qkv, int ld_qkv, const T* __restrict__ d_out, int ld_do, T* __restrict__ d_qkv,
                            int ld_dq, Grid3 g, int heads, int S) {
    constexpr int DH = 32, P = DH + 1;
    extern __shared__ float sm[];
    float* sa = sm;                            // [S][P]  phase 1: K, phase 2: Q
    float* sb = sa + (size_t)S * P;            // [S][P]  phase 1: V, phase 2: dO
    float* lse = sb + (size_t)S * P;           // [S]
    float* dsum = lse + S;                     // [S]
    float* wbuf = dsum + S;                    // [WARPS][2][S]  probabilities / score gradients of the row (column) in flight
    float* wvec = wbuf + (size_t)ATTB_WARPS * 2 * S;  // [WARPS][2][P]  q_i, dO_i (phase 1) / k_j, v_j (phase 2)
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int hid = heads * DH;
    const int lo = blockIdx.y * ATTB_ROWS, hi = min(S, lo + ATTB_ROWS);
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const float scale = rsqrtf((float)DH);
    auto row_of = [&](int s) {
        const int z = s % g.Z, y = (s / g.Z) % g.Y, x = s / (g.Z * g.Y);
        return g.row(b, x, y, z);
    };
    for (int i = threadIdx.x; i < S * DH; i += blockDim.x) {
        const int s = i / DH, d = i % DH;
        const T* src = qkv + row_of(s) * ld_qkv + h * DH + d;
        sa[s * P + d] = (float)src[hid];
        sb[s * P + d] = (float)src[2 * hid];
    }
    __syncthreads();
    float* p = wbuf + (size_t)warp * 2 * S;
    float* ds = p + S;
    float* v0 = wvec + warp * 2 * P;
    float* v1 = v0 + P;
    for (int i = warp; i < S; i += ATTB_WARPS) {
        const int64_t row = row_of(i);
        v0[lane] = (float)qkv[row * ld_qkv + h * DH + lane];
        v1[lane] = (float)d_out[row * ld_do + h * DH + lane];
        __syncwarp();
        float mx = -INFINITY;
        for (int j = lane; j < S; j += 32) {
            float acc = 0.0f;
#pragma unroll
            for (int d = 0; d < DH; ++d) acc = fmaf(v0[d], sa[j * P + d], acc);
            acc *= scale;
            p[j] = acc;
            mx = fmaxf(mx, acc);
        }
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int j = lane; j < S; j += 32) {
            const float e = expf(p[j] - mx);
            p[j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        float dsm = 0.0f;
        for (int j = lane; j < S; j += 32) {
            const float pj = p[j] * inv;
            float dp = 0.0f;
#pragma unroll
            for (int d = 0; d < DH; ++d) dp = fmaf(v1[d], sb[j * P + d], dp);
            p[j] = pj;
            ds[j] = dp;
            dsm = fmaf(pj, dp, dsm);
        }
        dsm = warp_sum(dsm);
        if (lane == 0) {
            lse[i] = mx + logf(sum);
            dsum[i] = dsm;
        }
        if (i >= lo && i < hi) {
            for (int j = lane; j < S; j += 32) ds[j] = p[j] * (ds[j] - dsm) * scale;
            __syncwarp();
            float dq = 0.0f;
            for (int j = 0; j < S; ++j) dq = fmaf(ds[j], sa[j * P + lane], dq);
            d_qkv[row * ld_dq + h * DH + lane] = (T)dq;
        }
        __syncwarp();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S * DH; i += blockDim.x) {
        const int s = i / DH, d = i % DH;
        const int64_t row = row_of(s);
        sa[s * P + d] = (float)qkv[row * ld_qkv + h * DH + d];
        sb[s * P + d] = (float)d_out[row * ld_do + h * DH + d];
    }
    __syncthreads();
    for (int j = lo + warp; j < hi; j += ATTB_WARPS) {
        const int64_t row = row_of(j);
        v0[lane] = (float)qkv[row * ld_qkv + hid + h * DH + lane];
        v1[lane] = (float)qkv[row * ld_qkv + 2 * hid + h * DH + lane];
        __syncwarp();
        for (int i = lane; i < S; i += 32) {
            float sc = 0.0f, dp = 0.0f;
#pragma unroll
            for (int d = 0; d < DH; ++d) {
                sc = fmaf(sa[i * P + d], v0[d], sc);
                dp = fmaf(sb[i * P + d], v1[d], dp);
            }
            const float pij = expf(sc * scale - lse[i]);
            p[i] = pij;
            ds[i] = pij * (dp - dsum[i]) * scale;
        }
        __syncwarp();
        float dk = 0.0f, dv = 0.0f;
        for (int i = 0; i < S; ++i) {
            dk = fmaf(ds[i], sa[i * P + lane], dk);
            dv = fmaf(p[i], sb[i * P + lane], dv);
        }
        T* o = d_qkv + row * ld_dq + h * DH + lane;
        o[hid] = (T)dk;
        o[2 * hid] = (T)dv;
        __syncwarp();
    }
}

// ---------------------------------------------------------------- channels-last x channels-first outer product
// out[c][f] += sum_{b, v interior} G[b,v][c] * Q[b*q_bstride + f*nvox + v]   (thread = one (c,f) pair)
template <typename T>
__global__ void __launch_bounds__(kThreads)
cl_nc_outer_kernel(const T* __restrict__ G, int ld, const float* __restrict__ Q, int64_t q_bstride, float* __restrict__ out,
                   float* __restrict__ colsum, Grid3 gr, int C, int F, int vox_per_block) {
    const int b = blockIdx.y;
    const int64_t nvox = (int64_t)gr.X * gr.Y * gr.Z;
    const int64_t v_begin = (int64_t)blockIdx.x * vox_per_block;
    const int64_t v_end = min(nvox, v_begin + vox_per_block);
    if (colsum)
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float acc = 0.0f;
            for (int64_t v = v_begin; v < v_end; ++v) {
                const int z = (int)(v % gr.Z);
                const int y = (int)((v / gr.Z) % gr.Y);
                const int x = (int)(v / ((int64_t)gr.Z * gr.Y));
                acc += (float)G[gr.row(b, x, y, z) * ld + c];
            }
            atomicAdd(&colsum[c], acc);
        }
    for (int pair = threadIdx.x; pair < C * F; pair += blockDim.x) {
        const int c = pair % C, f = pair / C;
        const float* q = Q + (int64_t)b * q_bstride + (int64_t)f * nvox;
        float acc = 0.0f;
        for (int64_t v = v_begin; v < v_end; ++v) {
            const int z = (int)(v % gr.Z);
            const int y = (int)((v / gr.Z) % gr.Y);
            const int x = (int)(v / ((int64_t)gr.Z * gr.Y));
            acc = fmaf((float)G[gr.row(b, x, y, z) * ld + c], __ldg(q + v), acc);
        }
        atomicAdd(&out[(int64_t)c * F + f], acc);
    }
}

// Vector form (C/N a power of two <= 32, pitch and base 16-byte aligned): thread = (voxel lane, 16-byte channel
// chunk) with FMAX x N accumulators in registers (+ N for the optional per-channel sum of G); U voxels per trip with
// all loads issued up front; warp shuffles over the voxel lanes, shared-memory atomics over the warps, then one global
// atomicAdd per (c, f) and block.
template <typename T, int FMAX>
__global__ void __launch_bounds__(kThreads)
cl_nc_outer_vec_kernel(const T* __restrict__ G, int ld, const float* __restrict__ Q, int64_t q_bstride, float* __restrict__ out,
                       float* __restrict__ colsum, Grid3 gr, int C, int F, int chunks, int vox_per_block, FastDiv by_z, FastDiv by_y) {
    constexpr int N = Vec<T>::N;
    constexpr int U = 4;
    __shared__ float sred[32 * N * (FMAX + 1)];
    const int b = blockIdx.y;
    const uint32_t nvox = (uint32_t)(gr.X * gr.Y * gr.Z);
    const uint32_t v_begin = blockIdx.x * (uint32_t)vox_per_block;
    const uint32_t v_end = min(nvox, v_begin + (uint32_t)vox_per_block);
    const int ch = threadIdx.x % chunks, lane_vox = threadIdx.x / chunks, vox_lanes = kThreads / chunks;
    const T* gb = G + (int64_t)b * gr.vox_p * ld + ch * N;
    for (int f0 = 0; f0 < F; f0 += FMAX) {
        const int nf = min(FMAX, F - f0);
        const bool want_sum = colsum != nullptr && f0 == 0;
        for (int i = threadIdx.x; i < C * (FMAX + 1); i += kThreads) sred[i] = 0.0f;
        __syncthreads();
        float acc[FMAX + 1][N];
#pragma unroll
        for (int f = 0; f <= FMAX; ++f)
#pragma unroll
            for (int i = 0; i < N; ++i) acc[f][i] = 0.0f;
        const float* qb = Q + (int64_t)b * q_bstride + (int64_t)f0 * nvox;
        for (uint32_t v0 = v_begin + lane_vox; v0 < v_end; v0 += U * vox_lanes) {
            uint4 graw[U];
            float qv[U][FMAX];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t v = min(v0 + u * vox_lanes, v_end - 1);
                uint32_t q, z, x, y;
                by_z.divmod(v, q, z);
                by_y.divmod(q, x, y);
                graw[u] = Vec<T>::load_raw(gb + (((int64_t)(x + 1) * gr.Yp + (y + 1)) * gr.Zp + (z + 1)) * ld);
#pragma unroll
                for (int f = 0; f < FMAX; ++f) qv[u][f] = f < nf ? __ldg(qb + (int64_t)f * nvox + v) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (v0 + u * vox_lanes >= v_end) break;
                float g[N];
                Vec<T>::unpack(graw[u], g);
#pragma unroll
                for (int f = 0; f < FMAX; ++f)
#pragma unroll
                    for (int i = 0; i < N; ++i) acc[f][i] = fmaf(g[i], qv[u][f], acc[f][i]);
#pragma unroll
                for (int i = 0; i < N; ++i) acc[FMAX][i] += g[i];
            }
        }
#pragma unroll
        for (int f = 0; f <= FMAX; ++f)
#pragma unroll
            for (int i = 0; i < N; ++i) {
                float v = acc[f][i];
                for (int o = 16; o >= chunks; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if ((threadIdx.x & 31) < chunks && (f < nf || (f == FMAX && want_sum))) atomicAdd(&sred[(ch * N + i) * (FMAX + 1) + f], v);
            }
        __syncthreads();
        for (int i = threadIdx.x; i < C * (FMAX + 1); i += kThreads) {
            const int c = i / (FMAX + 1), f = i % (FMAX + 1);
            if (f < nf) atomicAdd(&out[(int64_t)c * F + f0 + f], sred[i]);
            else if (f == FMAX && want_sum) atomicAdd(&colsum[c], sred[i]);
        }
        __syncthreads();
    }
}

}  // namespace

extern "C" {

int tdb_halo_fold(void* g, int ld, int B, int X, int Y, int Z, int C, int dtype, void* stream) {
    TDB_REQUIRE(g, TDB_E_BADARG, "tdb_halo_fold: null pointer");
    const int n = dtype == TDB_BF16 ? 8 : 4;
    TDB_REQUIRE(C % n == 0 && ld % n == 0 && aligned16(g), TDB_E_UNSUPPORTED, "tdb_halo_fold: C/ld must be multiples of %d", n);
    Grid3 gr(B, X, Y, Z);
    // a block covers at most kThreads 16-byte channel vectors per voxel: wider tensors (2048 fp32 channels in the first up
    // block of a dim = 64 model) are folded in channel slices, one launch each
    const int slice = kThreads * n;
    for (int c0 = 0; c0 < C; c0 += slice) {
        const int chunks = ((C - c0) < slice ? (C - c0) : slice) / n;
        void* gs = (char*)g + (size_t)c0 * (dtype == TDB_BF16 ? 2 : 4);
        if (X >= 2 && Y >= 2 && Z >= 2) {
            BorderEnum e;
            e.n_xf = 2u * Y * Z;
            e.n_yf = 2u * (X - 2) * Z;
            e.n_total = e.n_xf + e.n_yf + 2u * (X - 2) * (Y - 2);
            auto fd = [](int v) { return FastDiv((uint32_t)(v < 1 ? 1 : v)); };
            e.by_yz = fd(Y * Z); e.by_z = fd(Z); e.by_xz = fd((X - 2) * Z); e.by_xy = fd((X - 2) * (Y - 2)); e.by_ym2 = fd(Y - 2);
            dim3 grid((unsigned)ceil_div((int64_t)e.n_total * chunks, kThreads), (unsigned)B);  // one item per thread
            if (dtype == TDB_BF16)
                halo_fold_border_kernel<bf16><<<grid, kThreads, 0, (cudaStream_t)stream>>>((bf16*)gs, ld, gr, e, chunks);
            else
                halo_fold_border_kernel<float><<<grid, kThreads, 0, (cudaStream_t)stream>>>((float*)gs, ld, gr, e, chunks);
        } else {
            dim3 grid((unsigned)blocks_per_sample(gr.vox_p * chunks, B), (unsigned)B);
            if (dtype == TDB_BF16)
                halo_fold_kernel<bf16><<<grid, kThreads, 0, (cudaStream_t)stream>>>((bf16*)gs, ld, gr, make_split(gr), chunks);
            else
                halo_fold_kernel<float><<<grid, kThreads, 0, (cudaStream_t)stream>>>((float*)gs, ld, gr, make_split(gr), chunks);
        }
        TDB_CHECK_LAUNCH("tdb_halo_fold");
    }
    return 0;
}

int tdb_pointwise_bwd_reduce(const void* g_out, int ld_g, const void* raw, int ld_raw, const double* stats, const float* gamma,
                             const float* beta, const float* film, int film_ld, double* red, int B, int X, int Y, int Z, int C,
                             int G, float eps, unsigned flags, int dtype, void* stream) {
    TDB_REQUIRE(g_out && raw && red, TDB_E_BADARG, "tdb_pointwise_bwd_reduce: null pointer");
    TDB_REQUIRE(!stats || (gamma && beta && G >= 1 && C % G == 0), TDB_E_BADARG, "tdb_pointwise_bwd_reduce: norm args");
    const int n = dtype == TDB_BF16 ? 8 : 4;
    TDB_REQUIRE(C % n == 0 && ld_g % n == 0 && ld_raw % n == 0 && C / n <= kThreads && aligned16(g_out) && aligned16(raw),
                TDB_E_UNSUPPORTED, "tdb_pointwise_bwd_reduce: C/ld must be multiples of %d and C <= %d", n, n * kThreads);
    if (G < 1) G = 1;
    Grid3 gr(B, X, Y, Z);
    // blocks per sample: every block ends with 4*C double atomics into red[] (~6.5 G/s device-wide, measured), so a block
    // streams at least ~512 KB of its two inputs; the coarse levels then run on a few dozen blocks instead of 148 per sample
    // - but on at least 8 (one trip each on the smallest grids: a lone block per sample walks its voxels serially)
    const int64_t bytes_per_sample = (int64_t)X * Y * Z * C * (dtype == TDB_BF16 ? 2 : 4) * 2;
    int64_t blocks = ceil_div(bytes_per_sample, 512 * 1024);
    const int64_t trips = ceil_div((int64_t)X * Y * Z, (int64_t)(kThreads / (C / n)) * 4);
    if (blocks < 8) blocks = trips < 8 ? trips : 8;
    // whole waves: the kernel runs two blocks per SM (128 registers), so more than one wave's worth is rounded down to a
    // multiple of it (124 blocks per sample at 32 channels were 1.7 waves)
    const int64_t per_wave = (148 * 2) / (B < 1 ? 1 : B) < 1 ? 1 : (148 * 2) / (B < 1 ? 1 : B);
    if (blocks > per_wave) blocks = (blocks / per_wave) * per_wave;
    const int64_t cap = (148 * 4) / (B < 1 ? 1 : B) < 1 ? 1 : (148 * 4) / (B < 1 ? 1 : B);
    if (blocks > cap) blocks = cap;
    dim3 grid((unsigned)blocks, (unsigned)B);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (size_t)8 * C * sizeof(float);
    if (dtype == TDB_BF16)
        pw_bwd_reduce_kernel<bf16><<<grid, kThreads, smem, s>>>((const bf16*)g_out, ld_g, (const bf16*)raw, ld_raw, stats, gamma, beta,
                                                              film, film_ld, red, gr, C, G, eps, flags, FastDiv((uint32_t)Z), FastDiv((uint32_t)Y));
    else
        pw_bwd_reduce_kernel<float><<<grid, kThreads, smem, s>>>((const float*)g_out, ld_g, (const float*)raw, ld_raw, stats, gamma,
                                                               beta, film, film_ld, red, gr, C, G, eps, flags, FastDiv((uint32_t)Z), FastDiv((uint32_t)Y));
    TDB_CHECK_LAUNCH("tdb_pointwise_bwd_reduce");
    return 0;
}

int tdb_pointwise_bwd_finalize(const double* red, const double* stats, const float* gamma, const float* beta, const float* film,
                               int film_ld, float* grp, float* colsum, float* gw, float* gb, float* gsum, float* dfilm, int dfilm_ld,
                               int B, int X, int Y, int Z, int C, int G, float eps, void* stream) {
    TDB_REQUIRE(red && stats && gamma && beta && grp && colsum && gw && gb, TDB_E_BADARG, "tdb_pointwise_bwd_finalize: null pointer");
    const size_t smem = (size_t)B * G * 4 * sizeof(double);
    TDB_REQUIRE(G >= 1 && C % G == 0 && smem <= 200 * 1024, TDB_E_UNSUPPORTED,
                "tdb_pointwise_bwd_finalize: B*G=%d groups do not fit in shared memory", B * G);
    if (smem > 40 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(pw_bwd_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_pointwise_bwd_finalize: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    pw_bwd_finalize_kernel<<<1, kThreads, smem, (cudaStream_t)stream>>>(
        red, stats, gamma, beta, film, film_ld, grp, colsum, gw, gb, gsum, dfilm, dfilm_ld, B, C, G, (double)X * Y * Z, (double)eps);
    TDB_CHECK_LAUNCH("tdb_pointwise_bwd_finalize");
    return 0;
}

static int launch_pw_bwd_apply(const char* who, const void* g_out, int ld_g, const void* raw, int ld_raw, const double* stats, const float* gamma,
                               const float* beta, const float* film, int film_ld, const float* grp, void* d_raw, int ld_d, int B, int X,
                               int Y, int Z, int C, int G, float eps, unsigned flags, int dtype, PwFinal fin, void* stream) {
    const int n = dtype == TDB_BF16 ? 8 : 4;
    TDB_REQUIRE(C % n == 0 && ld_g % n == 0 && ld_raw % n == 0 && ld_d % n == 0 && C / n <= kThreads && aligned16(g_out) &&
                    aligned16(raw) && aligned16(d_raw),
                TDB_E_UNSUPPORTED, "%s: C/ld must be multiples of %d and C <= %d", who, n, n * kThreads);
    if (G < 1) G = 1;
    Grid3 gr(B, X, Y, Z);
    const int chunks = C / n;
    int64_t blocks = ceil_div(gr.vox_p, (int64_t)(kThreads / chunks) * 4);
    const int64_t cap = (148 * 8) / (B < 1 ? 1 : B) < 1 ? 1 : (148 * 8) / (B < 1 ? 1 : B);
    if (blocks > cap) blocks = cap;
    const bool fused = fin.red != nullptr;
    dim3 grid((unsigned)(blocks + (fused ? 1 : 0)), (unsigned)B);  // fused: one extra block per sample for the parameter gradients
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (size_t)5 * C * sizeof(float) + (fused ? (size_t)4 * B * G * sizeof(double) : 0);
    TDB_REQUIRE(smem <= 200 * 1024, TDB_E_UNSUPPORTED, "%s: B*G=%d groups do not fit in shared memory", who, B * G);
    cudaError_t e = cudaSuccess;
    if (dtype == TDB_BF16) {
        if (smem > 40 * 1024) e = cudaFuncSetAttribute(pw_bwd_apply_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            pw_bwd_apply_kernel<bf16><<<grid, kThreads, smem, s>>>((const bf16*)g_out, ld_g, (const bf16*)raw, ld_raw, stats, gamma, beta, film,
                                                                 film_ld, grp, (bf16*)d_raw, ld_d, gr, C, G, eps, flags, make_split(gr), chunks,
                                                                 fin);
    } else {
        if (smem > 40 * 1024) e = cudaFuncSetAttribute(pw_bwd_apply_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            pw_bwd_apply_kernel<float><<<grid, kThreads, smem, s>>>((const float*)g_out, ld_g, (const float*)raw, ld_raw, stats, gamma, beta,
                                                                  film, film_ld, grp, (float*)d_raw, ld_d, gr, C, G, eps, flags,
                                                                  make_split(gr), chunks, fin);
    }
    TDB_REQUIRE(e == cudaSuccess, (int)e, "%s: cudaFuncSetAttribute: %s", who, cudaGetErrorString(e));
    return 0;
}

int tdb_pointwise_bwd_apply(const void* g_out, int ld_g, const void* raw, int ld_raw, const double* stats, const float* gamma,
                            const float* beta, const float* film, int film_ld, const float* grp, void* d_raw, int ld_d, int B,
                            int X, int Y, int Z, int C, int G, float eps, unsigned flags, int dtype, void* stream) {
    TDB_REQUIRE(g_out && raw && d_raw, TDB_E_BADARG, "tdb_pointwise_bwd_apply: null pointer");
    TDB_REQUIRE(!stats || (gamma && beta && grp && G >= 1 && C % G == 0), TDB_E_BADARG, "tdb_pointwise_bwd_apply: norm args");
    PwFinal fin{};
    const int rc = launch_pw_bwd_apply("tdb_pointwise_bwd_apply", g_out, ld_g, raw, ld_raw, stats, gamma, beta, film, film_ld, grp, d_raw, ld_d,
                                       B, X, Y, Z, C, G, eps, flags, dtype, fin, stream);
    if (rc) return rc;
    TDB_CHECK_LAUNCH("tdb_pointwise_bwd_apply");
    return 0;
}

int tdb_pointwise_bwd_apply_fused(const void* g_out, int ld_g, const void* raw, int ld_raw, const double* stats, const float* gamma,
                                  const float* beta, const float* film, int film_ld, const double* red, void* d_raw, int ld_d,
                                  float* colsum, float* gw, float* gb, float* gsum, float* dfilm, int dfilm_ld, int B, int X, int Y,
                                  int Z, int C, int G, float eps, unsigned flags, int dtype, void* stream) {
    TDB_REQUIRE(g_out && raw && d_raw && red && stats && gamma && beta && colsum && gw && gb, TDB_E_BADARG,
                "tdb_pointwise_bwd_apply_fused: null pointer");
    TDB_REQUIRE(G >= 1 && C % G == 0, TDB_E_BADARG, "tdb_pointwise_bwd_apply_fused: norm args");
    PwFinal fin{red, colsum, gw, gb, gsum, dfilm, dfilm_ld, B};
    const int rc = launch_pw_bwd_apply("tdb_pointwise_bwd_apply_fused", g_out, ld_g, raw, ld_raw, stats, gamma, beta, film, film_ld, nullptr,
                                       d_raw, ld_d, B, X, Y, Z, C, G, eps, flags, dtype, fin, stream);
    if (rc) return rc;
    TDB_CHECK_LAUNCH("tdb_pointwise_bwd_apply_fused");
    return 0;
}

int tdb_conv3d_wgrad(const void* in, int ld_in, const void* d_out, int ld_do, float* dw, int B, int X, int Y, int Z, int Cin,
                     int Cout, int ntaps, int dtype, unsigned flags, void* stream) {
    TDB_REQUIRE(in && d_out && dw, TDB_E_BADARG, "tdb_conv3d_wgrad: null pointer");
    TDB_REQUIRE(ntaps == 1 || ntaps == 27, TDB_E_BADARG, "tdb_conv3d_wgrad: ntaps must be 1 or 27");
    Grid3 g(B, X, Y, Z);
    TDB_REQUIRE(g.rows < (1ll << 31), TDB_E_UNSUPPORTED, "tdb_conv3d_wgrad: too many rows");
    InteriorTest it;
    it.by_vox = FastDiv((uint32_t)g.vox_p);
    it.by_z = FastDiv((uint32_t)g.Zp);
    it.by_y = FastDiv((uint32_t)g.Yp);
    it.X = X; it.Y = Y; it.Z = Z;
    const int ci_tiles = (int)ceil_div(Cin, 64), co_tiles = (int)ceil_div(Cout, 64);
    const int tiles = ntaps * ci_tiles * co_tiles;
    // enough row slices to fill the machine ~4x, at least 256 rows each
    int64_t slices = ceil_div(148 * 4, tiles);
    int64_t rows_per_block = ceil_div(g.rows, slices < 1 ? 1 : slices);
    if (rows_per_block < 256) rows_per_block = 256;
    rows_per_block = ceil_div(rows_per_block, 16) * 16;
    dim3 grid((unsigned)ceil_div(g.rows, rows_per_block), (unsigned)tiles);
    cudaStream_t s = (cudaStream_t)stream;
    // bf16 with a zero-halo output gradient: tcgen05 kernel (MN-major operands straight from the halo grids)
    if (dtype == TDB_BF16 && (flags & TDB_WGRAD_ZERO_HALO) && Cin % 32 == 0 &&
        (Cout == 32 || (Cout <= 256 && Cout % 64 == 0) || Cout % 256 == 0) && ld_in % 8 == 0 && ld_do % 8 == 0 && aligned16(in) &&
        aligned16(d_out) && aligned16(dw))
        return tdb_conv3d_wgrad_tc(in, ld_in, d_out, ld_do, dw, B, X, Y, Z, Cin, Cout, ntaps, TDB_WGRAD_SHARE_KZ | TDB_WGRAD_KZ_ON_N, stream);
    const bool tensor_path = dtype == TDB_BF16 && (flags & TDB_WGRAD_ZERO_HALO) && Cin % 8 == 0 && Cout % 8 == 0 && ld_in % 8 == 0 &&
                             ld_do % 8 == 0 && aligned16(in) && aligned16(d_out);
    if (tensor_path) {
        rows_per_block = ceil_div(rows_per_block, 64) * 64;
        grid.x = (unsigned)ceil_div(g.rows, rows_per_block);
        conv_wgrad_wmma_kernel<<<grid, 256, 0, s>>>((const bf16*)in, ld_in, (const bf16*)d_out, ld_do, dw, g.rows, g.Yp * g.Zp, g.Zp, Cin,
                                                    Cout, ntaps, ci_tiles, co_tiles, (int)rows_per_block);
    } else if (dtype == TDB_BF16)
        conv_wgrad_kernel<bf16><<<grid, 256, 0, s>>>((const bf16*)in, ld_in, (const bf16*)d_out, ld_do, dw, g.rows, g.Yp * g.Zp, g.Zp,
                                                     Cin, Cout, ntaps, ci_tiles, co_tiles, (int)rows_per_block, it);
    else
        conv_wgrad_kernel<float><<<grid, 256, 0, s>>>((const float*)in, ld_in, (const float*)d_out, ld_do, dw, g.rows, g.Yp * g.Zp,
                                                      g.Zp, Cin, Cout, ntaps, ci_tiles, co_tiles, (int)rows_per_block, it);
    TDB_CHECK_LAUNCH("tdb_conv3d_wgrad");
    return 0;
}

int tdb_trilinear_bwd(const void* g_out, int ld_g, int Xo, int Yo, int Zo, void* d_in, int ld_d, int Xi, int Yi, int Zi, int B,
                      int C, int dtype, unsigned flags, void* stream) {
    TDB_REQUIRE(g_out && d_in, TDB_E_BADARG, "tdb_trilinear_bwd: null pointer");
    const int n = dtype == TDB_BF16 ? 8 : 4;
    TDB_REQUIRE(C % n == 0 && ld_g % n == 0 && ld_d % n == 0 && C / n <= kThreads && aligned16(g_out) && aligned16(d_in),
                TDB_E_UNSUPPORTED, "tdb_trilinear_bwd: C/ld must be multiples of %d and C <= %d", n, n * kThreads);
    Grid3 gi(B, Xi, Yi, Zi), go(B, Xo, Yo, Zo);
    const int chunks = C / n;
    const int accumulate = (flags & TDB_TRIBWD_ACCUMULATE) ? 1 : 0;
    auto scale_of = [](int n_in, int n_out) { return n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.0f; };
    const float sx = scale_of(Xi, Xo), sy = scale_of(Yi, Yo), sz = scale_of(Zi, Zo);
    auto ok = [](float sc, int n_out) { return n_out == 1 || sc >= 0.2f; };
    TDB_REQUIRE(ok(sx, Xo) && ok(sy, Yo) && ok(sz, Zo), TDB_E_UNSUPPORTED, "tdb_trilinear_bwd: upsampling factor above 5 per axis");
    cudaStream_t s = (cudaStream_t)stream;
    // line walker: <= 6 sources per axis in y and z (up-sampling factor <= 2.5)
    auto few = [](float sc, int n_out) { return n_out == 1 || sc >= 0.4f; };
    if (few(sy, Yo) && few(sz, Zo) && kThreads % chunks == 0 && gi.Yp * gi.Zp < (1 << 24) && B <= 65535) {
        const int cols = kThreads / chunks;  // (y, z) lines per block
        const int n_blocks = (int)ceil_div((int64_t)gi.Yp * gi.Zp, cols);
        // x segments: enough threads to fill the machine (the lines alone are few on the coarse side of an up-sampling)
        int64_t nseg = ceil_div((int64_t)148 * 2048, (int64_t)n_blocks * kThreads * B);
        if (nseg > Xi / 4) nseg = Xi / 4;
        if (nseg < 1) nseg = 1;
        const int seg_len = (int)ceil_div(Xi, nseg);
        nseg = ceil_div(Xi, seg_len);
        dim3 grid((unsigned)n_blocks, (unsigned)B, (unsigned)nseg);
        const size_t smem = (size_t)(gi.Yp + gi.Zp) * (2 * TB_S + 1) * sizeof(int);
        TDB_REQUIRE(smem <= 40 * 1024, TDB_E_UNSUPPORTED, "tdb_trilinear_bwd: %d + %d lines exceed the source tables", gi.Yp, gi.Zp);
        const FastDiv by_zp((uint32_t)gi.Zp);
        if (dtype == TDB_BF16)
            trilinear_bwd_walk_kernel<bf16><<<grid, kThreads, smem, s>>>((const bf16*)g_out, ld_g, go, (bf16*)d_in, ld_d, gi, chunks,
                                                                          seg_len, sx, sy, sz, accumulate, by_zp);
        else
            trilinear_bwd_walk_kernel<float><<<grid, kThreads, smem, s>>>((const float*)g_out, ld_g, go, (float*)d_in, ld_d, gi, chunks,
                                                                           seg_len, sx, sy, sz, accumulate, by_zp);
        TDB_CHECK_LAUNCH("tdb_trilinear_bwd");
        return 0;
    }
    dim3 grid((unsigned)blocks_per_sample(gi.vox_p * chunks, B), (unsigned)B);
    if (dtype == TDB_BF16)
        trilinear_bwd_kernel<bf16><<<grid, kThreads, 0, s>>>((const bf16*)g_out, ld_g, go, (bf16*)d_in, ld_d, gi, C, make_split(gi), chunks,
                                                              sx, sy, sz, accumulate);
    else
        trilinear_bwd_kernel<float><<<grid, kThreads, 0, s>>>((const float*)g_out, ld_g, go, (float*)d_in, ld_d, gi, C, make_split(gi),
                                                               chunks, sx, sy, sz, accumulate);
    TDB_CHECK_LAUNCH("tdb_trilinear_bwd");
    return 0;
}

int tdb_attention_bwd(const void* qkv, int ld_qkv, const void* d_out, int ld_do, void* d_qkv, int ld_dq, int B, int X, int Y,
                      int Z, int heads, int dh, int dtype, void* stream) {
    TDB_REQUIRE(qkv && d_out && d_qkv, TDB_E_BADARG, "tdb_attention_bwd: null pointer");
    TDB_REQUIRE(dh == 32, TDB_E_UNSUPPORTED, "tdb_attention_bwd: dim_head must be 32 (got %d)", dh);
    const int S = X * Y * Z;
    const size_t smem = ((size_t)4 * S * 33 + (size_t)2 * S * (S + 1)) * sizeof(float);
    if (smem > 220 * 1024) {
        // the S x S form does not fit (S > 135): streaming form, same limit as the forward kernel's CUDA-core path and beyond
        const size_t smem2 = ((size_t)2 * S * 33 + 2 * (size_t)S + (size_t)ATTB_WARPS * 2 * S + ATTB_WARPS * 2 * 33) * sizeof(float);
        TDB_REQUIRE(smem2 <= 220 * 1024, TDB_E_UNSUPPORTED, "tdb_attention_bwd: sequence of %d voxels exceeds shared memory", S);
        Grid3 g2(B, X, Y, Z);
        cudaStream_t s2 = (cudaStream_t)stream;
        dim3 grid((unsigned)(B * heads), (unsigned)((S + ATTB_ROWS - 1) / ATTB_ROWS));
        cudaError_t e2;
        if (dtype == TDB_BF16) {
            e2 = cudaFuncSetAttribute(attention_bwd_stream_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
            if (e2 == cudaSuccess)
                attention_bwd_stream_kernel<bf16><<<grid, ATTB_WARPS * 32, smem2, s2>>>((const bf16*)qkv, ld_qkv, (const bf16*)d_out, ld_do,
                                                                                       (bf16*)d_qkv, ld_dq, g2, heads, S);
        } else {
            e2 = cudaFuncSetAttribute(attention_bwd_stream_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
            if (e2 == cudaSuccess)
                attention_bwd_stream_kernel<float><<<grid, ATTB_WARPS * 32, smem2, s2>>>((const float*)qkv, ld_qkv, (const float*)d_out, ld_do,
                                                                                        (float*)d_qkv, ld_dq, g2, heads, S);
        }
        TDB_REQUIRE(e2 == cudaSuccess, (int)e2, "tdb_attention_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e2));
        TDB_CHECK_LAUNCH("tdb_attention_bwd (streaming)");
        return 0;
    }
    Grid3 g(B, X, Y, Z);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e;
    if (dtype == TDB_BF16) {
        e = cudaFuncSetAttribute(attention_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            attention_bwd_kernel<bf16><<<B * heads * ATT_SPLIT, 256, smem, s>>>((const bf16*)qkv, ld_qkv, (const bf16*)d_out, ld_do, (bf16*)d_qkv, ld_dq,
                                                                    g, heads, S);
    } else {
        e = cudaFuncSetAttribute(attention_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            attention_bwd_kernel<float><<<B * heads * ATT_SPLIT, 256, smem, s>>>((const float*)qkv, ld_qkv, (const float*)d_out, ld_do, (float*)d_qkv,
                                                                     ld_dq, g, heads, S);
    }
    TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_attention_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    TDB_CHECK_LAUNCH("tdb_attention_bwd");
    return 0;
}

int tdb_cl_nc_outer(const void* G, int ld, const float* Q, int64_t q_bstride, float* out, float* colsum, int B, int X, int Y, int Z,
                    int C, int F, int dtype, void* stream) {
    TDB_REQUIRE(G && Q && out, TDB_E_BADARG, "tdb_cl_nc_outer: null pointer");
    Grid3 gr(B, X, Y, Z);
    cudaStream_t s = (cudaStream_t)stream;
    const int n = dtype == TDB_BF16 ? 8 : 4;
    const int chunks = C / n;
    if (C % n == 0 && ld % n == 0 && aligned16(G) && chunks >= 1 && chunks <= 32 && (chunks & (chunks - 1)) == 0 &&
        (int64_t)X * Y * Z < (1ll << 31)) {
        const int64_t nvox = (int64_t)X * Y * Z;
        int64_t blocks = (148 * 8) / (B < 1 ? 1 : B);
        if (blocks < 1) blocks = 1;
        int64_t vpb = ceil_div(nvox, blocks);
        if (vpb < 256) vpb = 256;
        dim3 grid((unsigned)ceil_div(nvox, vpb), (unsigned)B);
        const FastDiv by_z((uint32_t)Z), by_y((uint32_t)Y);
        if (dtype == TDB_BF16)
            cl_nc_outer_vec_kernel<bf16, 4><<<grid, kThreads, 0, s>>>((const bf16*)G, ld, Q, q_bstride, out, colsum, gr, C, F, chunks, (int)vpb,
                                                                     by_z, by_y);
        else
            cl_nc_outer_vec_kernel<float, 4><<<grid, kThreads, 0, s>>>((const float*)G, ld, Q, q_bstride, out, colsum, gr, C, F, chunks,
                                                                      (int)vpb, by_z, by_y);
        TDB_CHECK_LAUNCH("tdb_cl_nc_outer");
        return 0;
    }
    const int vox_per_block = 512;
    dim3 grid((unsigned)ceil_div((int64_t)X * Y * Z, vox_per_block), (unsigned)B);
    if (dtype == TDB_BF16)
        cl_nc_outer_kernel<bf16><<<grid, kThreads, 0, s>>>((const bf16*)G, ld, Q, q_bstride, out, colsum, gr, C, F, vox_per_block);
    else
        cl_nc_outer_kernel<float><<<grid, kThreads, 0, s>>>((const float*)G, ld, Q, q_bstride, out, colsum, gr, C, F, vox_per_block);
    TDB_CHECK_LAUNCH("tdb_cl_nc_outer");
    return 0;
}

}  // extern "C"
