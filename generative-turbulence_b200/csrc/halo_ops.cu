// Bandwidth-bound kernels over halo grids: input encode, output decode, GroupNorm statistics,
// the fused GroupNorm/FiLM/SiLU/residual pointwise kernel and trilinear resampling.
// Every thread moves one 16-byte channel vector of one voxel; consecutive threads walk the
// channel vectors of a voxel and then the next voxel, so global accesses are fully coalesced.
#include <cstdlib>
#include "common.cuh"

using namespace tdb;
using bf16 = __nv_bfloat16;

namespace {

constexpr int kThreads = 256;

// Row index -> haloed coordinates with launch-time fast dividers (Zp, Yp).
struct RowSplit {
    FastDiv by_z, by_y;
    __device__ __forceinline__ void operator()(uint32_t r, int& xp, int& yp, int& zp) const {
        uint32_t q, zz, xx, yy;
        by_z.divmod(r, q, zz);
        by_y.divmod(q, xx, yy);
        xp = (int)xx; yp = (int)yy; zp = (int)zz;
    }
};

// ---------------------------------------------------------------- encode_x / encode_c_local
// grid = (blocks per sample, B).  A thread owns ONE 16-byte channel vector position (its 1x1-conv
// weights live in registers for the whole kernel) and walks haloed voxels.
template <typename T, int FMAX>
__global__ void __launch_bounds__(kThreads)
encode_input_kernel(const float* __restrict__ x, const float* __restrict__ c_local,
                    const float* __restrict__ wx, const float* __restrict__ bx,
                    const float* __restrict__ wc, const float* __restrict__ bc, T* __restrict__ out,
                    int ld_out, Grid3 g, int F, int Fc, int dim, int parts, RowSplit split, int chunks, int c_begin) {
    constexpr int N = Vec<T>::N;
    const int b = blockIdx.y;
    const int vox_step = kThreads / chunks;
    const int ch = threadIdx.x % chunks, lane_vox = threadIdx.x / chunks;
    if (lane_vox >= vox_step) return;
    const int c0 = c_begin + ch * N;  // `chunks` covers only the requested half when parts selects one
    const bool c_half = c0 >= dim;
    if (c_half ? !(parts & 2) : !(parts & 1)) return;
    const int nf = c_half ? Fc : F;
    const float* w = c_half ? wc + (int64_t)(c0 - dim) * Fc : wx + (int64_t)c0 * F;
    const float* bias = c_half ? bc + (c0 - dim) : bx + c0;
    float wr[N][FMAX], br[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        br[i] = bias[i];
#pragma unroll
        for (int f = 0; f < FMAX; ++f) wr[i][f] = f < nf ? w[i * nf + f] : 0.0f;
    }
    const int64_t nvox = (int64_t)g.X * g.Y * g.Z;
    const float* src0 = c_half ? c_local : x + (int64_t)b * F * nvox;
    const uint32_t total = (uint32_t)g.vox_p;
    // U voxels per trip: all U * nf scalar loads are issued before the first is used (memory-level parallelism -
    // one load in flight per thread left the kernel latency-bound at 40 % of the HBM rate)
    constexpr int U = 4;
    const uint32_t stride = gridDim.x * vox_step;
    for (uint32_t r0 = blockIdx.x * vox_step + lane_vox; r0 < total; r0 += stride * U) {
        float in[U][FMAX];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t r = min(r0 + u * stride, total - 1);
            int xp, yp, zp;
            split(r, xp, yp, zp);
            const int xs = clampi(xp - 1, 0, g.X - 1), ys = clampi(yp - 1, 0, g.Y - 1), zs = clampi(zp - 1, 0, g.Z - 1);
            const float* src = src0 + ((int64_t)xs * g.Y + ys) * g.Z + zs;
#pragma unroll
            for (int f = 0; f < FMAX; ++f) in[u][f] = f < nf ? __ldg(src + (int64_t)f * nvox) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t r = r0 + u * stride;
            if (r >= total) break;
            float o[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                float acc = br[i];
#pragma unroll
                for (int f = 0; f < FMAX; ++f) acc = fmaf(wr[i][f], in[u][f], acc);
                o[i] = acc;
            }
            Vec<T>::store(out + ((int64_t)b * g.vox_p + r) * ld_out + c0, o);
        }
    }
}

// ---------------------------------------------------------------- decode[1]
template <typename T>
__global__ void __launch_bounds__(kThreads)
decode_output_kernel(const T* __restrict__ act, int ld, const float* __restrict__ w,
                     const float* __restrict__ bias, float* __restrict__ out, Grid3 g, int dim, int F, FastDiv by_vox, FastDiv by_z,
                     FastDiv by_y) {
    constexpr int N = Vec<T>::N;
    extern __shared__ float sw[];  // F*dim weights + F biases
    for (int i = threadIdx.x; i < F * dim; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < F; i += blockDim.x) sw[F * dim + i] = bias[i];
    __syncthreads();
    const int64_t nvox = (int64_t)g.X * g.Y * g.Z;
    const uint32_t total = (uint32_t)(g.B * nvox);
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        uint32_t bb, v, q, zz, xx, yy;
        by_vox.divmod(idx, bb, v);
        by_z.divmod(v, q, zz);
        by_y.divmod(q, xx, yy);
        const int b = (int)bb, x = (int)xx, y = (int)yy, z = (int)zz;
        const T* a = act + g.row(b, x, y, z) * ld;
        float acc[8];
#pragma unroll
        for (int f = 0; f < 8; ++f) acc[f] = f < F ? sw[F * dim + f] : 0.0f;
        for (int c0 = 0; c0 < dim; c0 += N) {
            float vv[N];
            Vec<T>::load(a + c0, vv);
#pragma unroll
            for (int f = 0; f < 8; ++f)
                if (f < F) {
#pragma unroll
                    for (int i = 0; i < N; ++i) acc[f] = fmaf(sw[f * dim + c0 + i], vv[i], acc[f]);
                }
        }
#pragma unroll
        for (int f = 0; f < 8; ++f)
            if (f < F) out[((int64_t)b * F + f) * nvox + v] = acc[f];
    }
}

// ---------------------------------------------------------------- GroupNorm statistics
// grid = (blocks per sample, B).  Each thread accumulates <= 64 voxels of one channel vector in
// fp32, widens to double, and the block merges per group in shared memory before one double
// atomicAdd per (block, group, moment).  E[x^2]-E[x]^2 is then evaluated in double by the
// consumer, which keeps the 1e-5 parity budget at 3.9 M elements per group.
template <typename T>
__global__ void __launch_bounds__(kThreads)
gn_stats_kernel(const T* __restrict__ raw, int ld, double* __restrict__ stats, Grid3 g, int C, int G,
                int vox_per_block) {
    constexpr int N = Vec<T>::N;
    extern __shared__ double sacc[];  // [G][2]
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sacc[i] = 0.0;
    __syncthreads();
    const int b = blockIdx.y;
    const int chunks = C / N;
    const int cpg = C / G;
    const int64_t nvox = (int64_t)g.X * g.Y * g.Z;
    const int64_t v_begin = (int64_t)blockIdx.x * vox_per_block;
    const int64_t v_end = min(nvox, v_begin + vox_per_block);
    const int vox_step = blockDim.x / chunks;  // host guarantees chunks <= blockDim.x
    const int lane_vox = threadIdx.x / chunks;
    const int c0 = (threadIdx.x % chunks) * N;
    float s[N], ss[N];
#pragma unroll
    for (int i = 0; i < N; ++i) s[i] = ss[i] = 0.0f;
    if (lane_vox < vox_step) {
        for (int64_t v = v_begin + lane_vox; v < v_end; v += vox_step) {
            const int z = (int)(v % g.Z);
            const int y = (int)((v / g.Z) % g.Y);
            const int x = (int)(v / ((int64_t)g.Z * g.Y));
            float vv[N];
            Vec<T>::load(raw + g.row(b, x, y, z) * ld + c0, vv);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                s[i] += vv[i];
                ss[i] = fmaf(vv[i], vv[i], ss[i]);
            }
        }
        // flush per-channel partials into per-group shared accumulators
        int gcur = c0 / cpg;
        double ds = 0.0, dss = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const int gi = (c0 + i) / cpg;
            if (gi != gcur) {
                atomicAdd(&sacc[2 * gcur], ds);
                atomicAdd(&sacc[2 * gcur + 1], dss);
                ds = dss = 0.0;
                gcur = gi;
            }
            ds += (double)s[i];
            dss += (double)ss[i];
        }
        atomicAdd(&sacc[2 * gcur], ds);
        atomicAdd(&sacc[2 * gcur + 1], dss);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(&stats[(int64_t)b * 2 * G + i], sacc[i]);
}

// ---------------------------------------------------------------- fused pointwise
template <typename T>
__device__ __forceinline__ float act_silu(float v) {
    if constexpr (sizeof(T) == 2)  // bf16 storage: one special-function op (tanh.approx), error far below the output rounding
        return silu_tanh(v);
    else
        return silu_f(v);
}

// grid = (blocks per sample, B).  A thread owns one 16-byte channel vector position: its affine
// coefficients (GroupNorm mean/rstd/gamma/beta folded with FiLM scale/shift) live in registers.
template <typename T>
__global__ void __launch_bounds__(kThreads)
pointwise_kernel(const T* __restrict__ raw, int ld_raw, const double* __restrict__ stats,
                 const float* __restrict__ gamma, const float* __restrict__ beta,
                 const float* __restrict__ film, int film_ld, const T* __restrict__ res, int ld_res,
                 T* __restrict__ out, int ld_out, Grid3 g, int C, int G, float eps, unsigned flags,
                 RowSplit split, int chunks) {
    constexpr int N = Vec<T>::N;
    extern __shared__ float s_coef[];  // [C][2]: per-channel scale / offset of this sample
    const int b = blockIdx.y;
    const int vox_step = kThreads / chunks;
    const int ch = threadIdx.x % chunks, lane_vox = threadIdx.x / chunks;
    const int c0 = ch * N;
    // the affine coefficients depend on (sample, channel) only: computed once per block (one thread per channel, the
    // group moments in double once per thread), not once per thread - on the small levels the per-thread prologue
    // (8 channels x gamma / beta / FiLM / moments) used to cost more than the rows themselves
    {
        const int cpg = C / G;
        const double inv_n = 1.0 / ((double)cpg * g.X * g.Y * g.Z);
        for (int c = threadIdx.x; c < C; c += kThreads) {
            float a = 1.0f, o = 0.0f;
            if (stats) {
                const int gi = c / cpg;
                const double mean = stats[((int64_t)b * G + gi) * 2] * inv_n;
                const double var = fma(-mean, mean, stats[((int64_t)b * G + gi) * 2 + 1] * inv_n);
                const float mean_f = (float)mean;
                const float rstd = 1.0f / sqrtf(fmaxf((float)var, 0.0f) + eps);
                a = rstd * gamma[c];
                o = beta[c] - mean_f * a;
            }
            if (film) {
                const float sc = film[(int64_t)b * film_ld + c] + 1.0f;
                const float sh = film[(int64_t)b * film_ld + C + c];
                a *= sc;
                o = fmaf(o, sc, sh);
            }
            s_coef[2 * c] = a;
            s_coef[2 * c + 1] = o;
        }
    }
    __syncthreads();
    if (lane_vox >= vox_step) return;
    float ca[N], co[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        ca[i] = s_coef[2 * (c0 + i)];
        co[i] = s_coef[2 * (c0 + i) + 1];
    }
    const bool interior_only = flags & TDB_PW_NOHALO;
    const bool act = flags & TDB_PW_SILU;
    const uint32_t total = (uint32_t)g.vox_p;
    // U rows per trip, every load issued before the first use: with one 16-byte load in flight per thread the kernel was
    // latency-bound (2048 threads x 16 B per SM = 4.8 MB in flight chip-wide against ~6.5 MB needed at the HBM rate)
    constexpr int U = 4;
    const uint32_t stride = gridDim.x * vox_step;
    const int64_t base = (int64_t)b * g.vox_p;
    for (uint32_t r0 = blockIdx.x * vox_step + lane_vox; r0 < total; r0 += stride * U) {
        uint4 rv[U], sv[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t r = r0 + u * stride;
            ok[u] = r < total;
            int xp, yp, zp;
            split(ok[u] ? r : total - 1, xp, yp, zp);
            const int xs = clampi(xp, 1, g.X), ys = clampi(yp, 1, g.Y), zs = clampi(zp, 1, g.Z);
            if (interior_only && (xs != xp || ys != yp || zs != zp)) ok[u] = false;
            const int64_t src = base + ((int64_t)xs * g.Yp + ys) * g.Zp + zs;
            rv[u] = Vec<T>::load_raw(raw + src * ld_raw + c0);
            if (res) sv[u] = Vec<T>::load_raw(res + src * ld_res + c0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            float v[N];
            Vec<T>::unpack(rv[u], v);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const float y = fmaf(ca[i], v[i], co[i]);
                v[i] = act ? act_silu<T>(y) : y;
            }
            if (res) {
                float rr[N];
                Vec<T>::unpack(sv[u], rr);
#pragma unroll
                for (int i = 0; i < N; ++i) v[i] += rr[i];
            }
            Vec<T>::store(out + (base + r0 + u * stride) * ld_out + c0, v);
        }
    }
}

// ---------------------------------------------------------------- trilinear (align_corners)
struct Lerp {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ Lerp axis_lerp(int o, int n_in, float scale) {
    // ATen: scale = (float)(n_in-1)/(n_out-1); src = scale*o; i0 = (int)src; lambda = src - i0
    const float src = scale * (float)o;
    Lerp r;
    r.i0 = min((int)src, n_in - 1);
    r.i1 = r.i0 + (r.i0 < n_in - 1 ? 1 : 0);
    r.l1 = src - (float)r.i0;
    r.l0 = 1.0f - r.l1;
    return r;
}

// grid = (blocks per sample, B)
template <typename T>
__global__ void __launch_bounds__(kThreads)
trilinear_kernel(const T* __restrict__ in, int ld_in, Grid3 gi, T* __restrict__ out, int ld_out, Grid3 go, int C,
                 RowSplit split, int chunks, float sx, float sy, float sz) {
    constexpr int N = Vec<T>::N;
    const int b = blockIdx.y;
    const int vox_step = kThreads / chunks;
    const int ch = threadIdx.x % chunks, lane_vox = threadIdx.x / chunks;
    if (lane_vox >= vox_step) return;
    const int c0 = ch * N;
    // line walker over the output grid: x/y interpolation and the four input line bases once per (x, y) line
    const uint32_t lines = (uint32_t)(go.Xp * go.Yp);
    for (uint32_t line = blockIdx.x * vox_step + lane_vox; line < lines; line += gridDim.x * vox_step) {
        uint32_t xq, yq;
        split.by_y.divmod(line, xq, yq);
        const Lerp lx = axis_lerp(clampi((int)xq - 1, 0, go.X - 1), gi.X, sx);
        const Lerp ly = axis_lerp(clampi((int)yq - 1, 0, go.Y - 1), gi.Y, sy);
        const T* base[4];
        float wxy[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xi = (k & 2) ? lx.i1 : lx.i0, yi = (k & 1) ? ly.i1 : ly.i0;
            base[k] = in + gi.row(b, xi, yi, 0) * ld_in + c0;
            wxy[k] = ((k & 2) ? lx.l1 : lx.l0) * ((k & 1) ? ly.l1 : ly.l0);
        }
        T* dst = out + ((int64_t)b * go.vox_p + (int64_t)line * go.Zp) * ld_out + c0;
        // the x/y-interpolated input planes P(iz) are kept in registers while the walk moves along z: an output is
        // l0*P(i0) + l1*P(i1), and when upsampling each P is reused by about two outputs (i0/i1 are warp-uniform)
        auto plane = [&](int iz, float (&pl)[N]) {
#pragma unroll
            for (int i = 0; i < N; ++i) pl[i] = 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float v[N];
                Vec<T>::load(base[k] + (int64_t)iz * ld_in, v);
#pragma unroll
                for (int i = 0; i < N; ++i) pl[i] = fmaf(wxy[k], v[i], pl[i]);
            }
        };
        int c_i0 = -1, c_i1 = -1;
        float p0[N], p1[N];
        for (int zp = 0; zp < go.Zp; ++zp) {
            const Lerp lz = axis_lerp(clampi(zp - 1, 0, go.Z - 1), gi.Z, sz);
            if (lz.i0 != c_i0) {
                if (lz.i0 == c_i1) {
#pragma unroll
                    for (int i = 0; i < N; ++i) p0[i] = p1[i];
                } else {
                    plane(lz.i0, p0);
                }
                c_i0 = lz.i0;
            }
            if (lz.i1 != c_i1) {
                if (lz.i1 == c_i0) {
#pragma unroll
                    for (int i = 0; i < N; ++i) p1[i] = p0[i];
                } else {
                    plane(lz.i1, p1);
                }
                c_i1 = lz.i1;
            }
            float acc[N];
#pragma unroll
            for (int i = 0; i < N; ++i) acc[i] = fmaf(lz.l0, p0[i], lz.l1 * p1[i]);
            Vec<T>::store(dst + (int64_t)zp * ld_out, acc);
        }
    }
}

// UP-sampling, two stages per output line (ncu of the walker above at 97x25x25 -> 194x50x50: issue slots 66 % busy at 34 %
// DRAM, ~110 instructions per 16-byte output vector - the z interpolation recomputed per thread and z, the x/y blend of
// four input vectors rebuilt in registers per thread).  Here a WARP owns an output (x, y) line: stage 1 blends the four
// input lines in x/y ONCE into shared memory (fp32, Zi x C), stage 2 blends two of those rows per output voxel with the z
// weights of a per-block table.  Same arithmetic in the same order as the walker (bit-identical results; the packed
// fma.rn.f32x2 / mul.rn.f32x2 round each lane like the scalar instructions), 25 % fewer instructions (134 M -> 101 M at the
// level-0 shape, B = 8): 195 -> 177 us, where it stops being issue-bound (issue slots 54 %) and sits at 2.7 TB/s of
// writes.  Lanes = CW channel vectors x (32 / CW) z positions; needs CW = C / N to divide 32.
// packed fp32 pairs (sm_100a): d = a * b + c on two lanes, each rounded like a scalar fma.rn / mul.rn
__device__ __forceinline__ void fma2(float& d0, float& d1, float a, float b0, float b1, float c0, float c1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %2};\n\tmov.b64 rb, {%3, %4};\n\tmov.b64 rc, {%5, %6};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void mul2(float& d0, float& d1, float a, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %2};\n\tmov.b64 rb, {%3, %4};\n\t"
        "mul.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a), "f"(b0), "f"(b1));
}

constexpr int TRI_WARPS = 4;
template <typename T>
__global__ void __launch_bounds__(TRI_WARPS * 32)
trilinear_up_line_kernel(const T* __restrict__ in, int ld_in, Grid3 gi, T* __restrict__ out, int ld_out, Grid3 go, int C,
                         RowSplit split, int cw, float sx, float sy, float sz) {
    constexpr int N = Vec<T>::N;
    extern __shared__ __align__(16) unsigned char tri_smem[];
    int4* ztab = reinterpret_cast<int4*>(tri_smem);                      // [Zp] (i0, i1, l0, l1)
    float* pl = reinterpret_cast<float*>(tri_smem + (size_t)go.Zp * sizeof(int4)) + (size_t)(threadIdx.x / 32) * gi.Z * C;  // [Zi][C]
    for (int zp = threadIdx.x; zp < go.Zp; zp += blockDim.x) {
        const Lerp lz = axis_lerp(clampi(zp - 1, 0, go.Z - 1), gi.Z, sz);
        ztab[zp] = make_int4(lz.i0, lz.i1, __float_as_int(lz.l0), __float_as_int(lz.l1));
    }
    __syncthreads();
    const int b = blockIdx.y;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int ch = lane % cw, zl = lane / cw, zstep = 32 / cw;
    const int c0 = ch * N;
    const uint32_t lines = (uint32_t)(go.Xp * go.Yp);
    for (uint32_t line = blockIdx.x * TRI_WARPS + warp; line < lines; line += gridDim.x * TRI_WARPS) {
        uint32_t xq, yq;
        split.by_y.divmod(line, xq, yq);
        const Lerp lx = axis_lerp(clampi((int)xq - 1, 0, go.X - 1), gi.X, sx);
        const Lerp ly = axis_lerp(clampi((int)yq - 1, 0, go.Y - 1), gi.Y, sy);
        const T* base[4];
        float wxy[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xi = (k & 2) ? lx.i1 : lx.i0, yi = (k & 1) ? ly.i1 : ly.i0;
            base[k] = in + gi.row(b, xi, yi, 0) * ld_in + c0;
            wxy[k] = ((k & 2) ? lx.l1 : lx.l0) * ((k & 1) ? ly.l1 : ly.l0);
        }
        // stage 1: x/y-interpolated input line
#pragma unroll 2
        for (int iz = zl; iz < gi.Z; iz += zstep) {
            uint4 raw[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) raw[k] = Vec<T>::load_raw(base[k] + iz * ld_in);
            float acc[N];
#pragma unroll
            for (int i = 0; i < N; ++i) acc[i] = 0.0f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float v[N];
                Vec<T>::unpack(raw[k], v);
#pragma unroll
                for (int i = 0; i < N; i += 2) fma2(acc[i], acc[i + 1], wxy[k], v[i], v[i + 1], acc[i], acc[i + 1]);
            }
            // row layout [N / 4][cw] float4: the cw lanes of a z position touch consecutive 16-byte words (no bank conflicts)
            float4* dstp = reinterpret_cast<float4*>(pl + iz * C) + ch;
#pragma unroll
            for (int q = 0; q < N / 4; ++q) dstp[q * cw] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        }
        __syncwarp();
        // stage 2: z blend
        T* dst = out + ((int64_t)b * go.vox_p + (int64_t)line * go.Zp) * ld_out + c0;
        for (int zp = zl; zp < go.Zp; zp += zstep) {
            const int4 e = ztab[zp];
            const float l0 = __int_as_float(e.z), l1 = __int_as_float(e.w);
            const float4* p0 = reinterpret_cast<const float4*>(pl + e.x * C) + ch;
            const float4* p1 = reinterpret_cast<const float4*>(pl + e.y * C) + ch;
            float acc[N];
#pragma unroll
            for (int q = 0; q < N / 4; ++q) {
                const float4 a = p0[q * cw], c = p1[q * cw];
                float m0, m1, m2, m3;
                mul2(m0, m1, l1, c.x, c.y);
                mul2(m2, m3, l1, c.z, c.w);
                fma2(acc[4 * q], acc[4 * q + 1], l0, a.x, a.y, m0, m1);
                fma2(acc[4 * q + 2], acc[4 * q + 3], l0, a.z, a.w, m2, m3);
            }
            Vec<T>::store(dst + (int64_t)zp * ld_out, acc);
        }
        __syncwarp();  // the next line's stage 1 overwrites the buffer
    }
}

// DOWN-sampling variant, grid = (blocks per sample, B).  Gather form: a thread owns one 16-byte channel vector of one OUTPUT row (haloed rows
// included: their source is the clamped interior voxel), fetches its eight source vectors and blends them - x/y first,
// then z, the order ATen's separable kernel and the fp32 parity tests use.  Consecutive threads walk the channel vectors
// of a row and then the next row (= next z), so stores are fully coalesced and the (up to 8x smaller or 8x larger)
// source is read through L1/L2; U rows per trip keep 16 independent loads in flight per thread.
template <typename T>
__global__ void __launch_bounds__(kThreads)
trilinear_gather_kernel(const T* __restrict__ in, int ld_in, Grid3 gi, T* __restrict__ out, int ld_out, Grid3 go, int C,
                 RowSplit split, int chunks, float sx, float sy, float sz) {
    constexpr int N = Vec<T>::N;
    constexpr int U = 2;
    const int b = blockIdx.y;
    const int vox_step = kThreads / chunks;
    const int ch = threadIdx.x % chunks, lane_vox = threadIdx.x / chunks;
    if (lane_vox >= vox_step) return;
    const int c0 = ch * N;
    const uint32_t total = (uint32_t)go.vox_p;
    const uint32_t stride = gridDim.x * vox_step;
    const T* in_b = in + (int64_t)b * gi.vox_p * ld_in + c0;
    T* out_b = out + (int64_t)b * go.vox_p * ld_out + c0;
    for (uint32_t r0 = blockIdx.x * vox_step + lane_vox; r0 < total; r0 += stride * U) {
        uint4 v[U][2][4];
        float wxy[U][4], wz[U][2];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t r = min(r0 + u * stride, total - 1);
            int xq, yq, zq;
            split(r, xq, yq, zq);
            const Lerp lx = axis_lerp(clampi(xq - 1, 0, go.X - 1), gi.X, sx);
            const Lerp ly = axis_lerp(clampi(yq - 1, 0, go.Y - 1), gi.Y, sy);
            const Lerp lz = axis_lerp(clampi(zq - 1, 0, go.Z - 1), gi.Z, sz);
            wz[u][0] = lz.l0;
            wz[u][1] = lz.l1;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int xi = (k & 2) ? lx.i1 : lx.i0, yi = (k & 1) ? ly.i1 : ly.i0;
                wxy[u][k] = ((k & 2) ? lx.l1 : lx.l0) * ((k & 1) ? ly.l1 : ly.l0);
                const int64_t line = ((int64_t)(xi + 1) * gi.Yp + (yi + 1)) * gi.Zp + 1;
                v[u][0][k] = Vec<T>::load_raw(in_b + (line + lz.i0) * ld_in);
                v[u][1][k] = Vec<T>::load_raw(in_b + (line + lz.i1) * ld_in);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t r = r0 + u * stride;
            if (r >= total) break;
            float p[2][N];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int i = 0; i < N; ++i) p[h][i] = 0.0f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float t[N];
                    Vec<T>::unpack(v[u][h][k], t);
#pragma unroll
                    for (int i = 0; i < N; ++i) p[h][i] = fmaf(wxy[u][k], t[i], p[h][i]);
                }
            }
            float acc[N];
#pragma unroll
            for (int i = 0; i < N; ++i) acc[i] = fmaf(wz[u][0], p[0][i], wz[u][1] * p[1][i]);
            Vec<T>::store(out_b + (int64_t)r * ld_out, acc);
        }
    }
}

RowSplit make_split(const Grid3& g) {
    RowSplit s;
    s.by_z = FastDiv((uint32_t)g.Zp);
    s.by_y = FastDiv((uint32_t)g.Yp);
    return s;
}

// blocks per sample so that the whole launch is ~16 resident waves of 148 SMs at most
int blocks_per_sample(int64_t items_per_sample, int B) {
    int64_t blocks = ceil_div(items_per_sample, kThreads);
    int64_t cap = (148 * 16) / (B < 1 ? 1 : B);
    if (cap < 8) cap = 8;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

int grid_for(int64_t work_items) {
    int64_t blocks = ceil_div(work_items, kThreads);
    const int64_t cap = 148 * 16;  // grid-stride beyond 16 resident waves of 148 SMs
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

extern "C" {

int tdb_encode_input(const float* x, const float* c_local, const float* wx, const float* bx,
                     const float* wc, const float* bc, void* out, int ld_out, int B, int F, int Fc,
                     int dim, int X, int Y, int Z, int parts, int dtype, void* stream) {
    TDB_REQUIRE(x && wx && bx && out, TDB_E_BADARG, "tdb_encode_input: null pointer");
    TDB_REQUIRE(Fc == 0 || (c_local && wc && bc), TDB_E_BADARG, "tdb_encode_input: c_local missing");
    TDB_REQUIRE(F >= 1 && F <= 8 && Fc >= 0 && Fc <= 8, TDB_E_UNSUPPORTED, "tdb_encode_input: F, Fc must be <= 8");
    const int n = dtype == TDB_BF16 ? 8 : 4;
    TDB_REQUIRE(dim % n == 0 && ld_out % n == 0 && aligned16(out), TDB_E_UNSUPPORTED,
                "tdb_encode_input: dim/ld_out must be multiples of %d and out 16B aligned", n);
    Grid3 g(B, X, Y, Z);
    const int ctot = dim + (Fc > 0 ? dim : 0);
    // parts: bit 0 = the x half, bit 1 = the c_local half; a single half launches only its own channel chunks
    const bool both = Fc > 0 && (parts & 3) == 3;
    const int c_begin = (Fc > 0 && (parts & 3) == 2) ? dim : 0;
    const int chunks = (both ? ctot : dim) / n;
    TDB_REQUIRE(g.vox_p < (1ll << 31) && chunks <= kThreads, TDB_E_UNSUPPORTED, "tdb_encode_input: grid too large for 32-bit indexing, or more than %d channel vectors of 16 bytes per voxel", kThreads);
    dim3 grid((unsigned)blocks_per_sample(g.vox_p * chunks, B), (unsigned)B);
    const RowSplit split = make_split(g);
    cudaStream_t s = (cudaStream_t)stream;
    const bool small = F <= 4 && Fc <= 4;  // u+p and the 4-d cell-type embedding: keep the weights in 32 registers
    if (dtype == TDB_BF16) {
        if (small)
            encode_input_kernel<bf16, 4><<<grid, kThreads, 0, s>>>(x, c_local, wx, bx, wc, bc, (bf16*)out, ld_out, g, F, Fc, dim, parts, split, chunks, c_begin);
        else
            encode_input_kernel<bf16, 8><<<grid, kThreads, 0, s>>>(x, c_local, wx, bx, wc, bc, (bf16*)out, ld_out, g, F, Fc, dim, parts, split, chunks, c_begin);
    } else {
        if (small)
            encode_input_kernel<float, 4><<<grid, kThreads, 0, s>>>(x, c_local, wx, bx, wc, bc, (float*)out, ld_out, g, F, Fc, dim, parts, split, chunks, c_begin);
        else
            encode_input_kernel<float, 8><<<grid, kThreads, 0, s>>>(x, c_local, wx, bx, wc, bc, (float*)out, ld_out, g, F, Fc, dim, parts, split, chunks, c_begin);
    }
    TDB_CHECK_LAUNCH("tdb_encode_input");
    return 0;
}

int tdb_decode_output(const void* act, int ld, const float* w, const float* b, float* out, int B,
                      int X, int Y, int Z, int dim, int F, int dtype, void* stream) {
    TDB_REQUIRE(act && w && b && out, TDB_E_BADARG, "tdb_decode_output: null pointer");
    const int n = dtype == TDB_BF16 ? 8 : 4;
    TDB_REQUIRE(F >= 1 && F <= 8 && dim % n == 0 && ld % n == 0 && aligned16(act), TDB_E_UNSUPPORTED,
                "tdb_decode_output: F <= 8, dim/ld multiples of %d", n);
    TDB_REQUIRE((int64_t)B * X * Y * Z < (1ll << 31), TDB_E_UNSUPPORTED, "tdb_decode_output: grid too large for 32-bit indexing");
    Grid3 g(B, X, Y, Z);
    const FastDiv by_vox((uint32_t)(X * Y * Z)), by_z((uint32_t)Z), by_y((uint32_t)Y);
    const int blocks = grid_for((int64_t)B * X * Y * Z);
    const size_t smem = (size_t)(F * dim + F) * sizeof(float);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == TDB_BF16)
        decode_output_kernel<bf16><<<blocks, kThreads, smem, s>>>((const bf16*)act, ld, w, b, out, g, dim, F, by_vox, by_z, by_y);
    else
        decode_output_kernel<float><<<blocks, kThreads, smem, s>>>((const float*)act, ld, w, b, out, g, dim, F, by_vox, by_z, by_y);
    TDB_CHECK_LAUNCH("tdb_decode_output");
    return 0;
}

int tdb_gn_stats(const void* raw, int ld, double* stats, int B, int X, int Y, int Z, int C, int G,
                 int dtype, void* stream) {
    TDB_REQUIRE(raw && stats, TDB_E_BADARG, "tdb_gn_stats: null pointer");
    const int n = dtype == TDB_BF16 ? 8 : 4;
    TDB_REQUIRE(G >= 1 && C % G == 0 && C % n == 0 && ld % n == 0 && C / n <= kThreads && aligned16(raw),
                TDB_E_UNSUPPORTED, "tdb_gn_stats: C=%d G=%d ld=%d unsupported", C, G, ld);
    Grid3 g(B, X, Y, Z);
    const int chunks = C / n;
    const int vox_step = kThreads / chunks;
    const int64_t nvox = (int64_t)X * Y * Z;
    // <= 64 voxels per thread (fp32 partial sums), fewer on small grids so that the launch still fills the SMs
    int64_t iters = ceil_div(nvox, (int64_t)vox_step * ((148 * 2) / (B < 1 ? 1 : B) + 1));
    iters = iters < 1 ? 1 : (iters > 64 ? 64 : iters);
    const int vox_per_block = vox_step * (int)iters;
    dim3 grid((unsigned)ceil_div(nvox, vox_per_block), (unsigned)B);
    const size_t smem = (size_t)2 * G * sizeof(double);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == TDB_BF16)
        gn_stats_kernel<bf16><<<grid, kThreads, smem, s>>>((const bf16*)raw, ld, stats, g, C, G, vox_per_block);
    else
        gn_stats_kernel<float><<<grid, kThreads, smem, s>>>((const float*)raw, ld, stats, g, C, G, vox_per_block);
    TDB_CHECK_LAUNCH("tdb_gn_stats");
    return 0;
}

int tdb_pointwise(const void* raw, int ld_raw, const double* stats, const float* gamma,
                  const float* beta, const float* film, int film_ld, const void* res, int ld_res,
                  void* out, int ld_out, int B, int X, int Y, int Z, int C, int G, float eps,
                  unsigned flags, int dtype, void* stream) {
    TDB_REQUIRE(raw && out, TDB_E_BADARG, "tdb_pointwise: null pointer");
    TDB_REQUIRE(!stats || (gamma && beta && G >= 1 && C % G == 0), TDB_E_BADARG, "tdb_pointwise: norm args");
    const int n = dtype == TDB_BF16 ? 8 : 4;
    TDB_REQUIRE(C % n == 0 && ld_raw % n == 0 && ld_out % n == 0 && (!res || ld_res % n == 0) &&
                    aligned16(raw) && aligned16(out) && aligned16(res),
                TDB_E_UNSUPPORTED, "tdb_pointwise: channel counts / pitches must be multiples of %d", n);
    if (G < 1) G = 1;
    Grid3 g(B, X, Y, Z);
    const int chunks = C / n;
    TDB_REQUIRE(g.vox_p < (1ll << 31) && chunks <= kThreads, TDB_E_UNSUPPORTED, "tdb_pointwise: grid too large for 32-bit indexing, or more than %d channel vectors of 16 bytes per voxel", kThreads);
    // at least two trips of four rows per thread: the per-block prologue and the launch tail stay small on the deep levels
    dim3 grid((unsigned)blocks_per_sample(ceil_div(g.vox_p * chunks, 8), B), (unsigned)B);
    const RowSplit split = make_split(g);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (size_t)2 * C * sizeof(float);
    if (dtype == TDB_BF16)
        pointwise_kernel<bf16><<<grid, kThreads, smem, s>>>((const bf16*)raw, ld_raw, stats, gamma, beta, film, film_ld,
                                                             (const bf16*)res, ld_res, (bf16*)out, ld_out, g, C, G, eps,
                                                             flags, split, chunks);
    else
        pointwise_kernel<float><<<grid, kThreads, smem, s>>>((const float*)raw, ld_raw, stats, gamma, beta, film, film_ld,
                                                              (const float*)res, ld_res, (float*)out, ld_out, g, C, G, eps,
                                                              flags, split, chunks);
    TDB_CHECK_LAUNCH("tdb_pointwise");
    return 0;
}

int tdb_trilinear(const void* in, int ld_in, int Xi, int Yi, int Zi, void* out, int ld_out, int Xo,
                  int Yo, int Zo, int B, int C, int dtype, void* stream) {
    const bool force_line = (dtype & TDB_TRILINEAR_LINE) != 0;
    dtype &= ~TDB_TRILINEAR_LINE;
    TDB_REQUIRE(in && out, TDB_E_BADARG, "tdb_trilinear: null pointer");
    const int n = dtype == TDB_BF16 ? 8 : 4;
    TDB_REQUIRE(C % n == 0 && ld_in % n == 0 && ld_out % n == 0 && aligned16(in) && aligned16(out),
                TDB_E_UNSUPPORTED, "tdb_trilinear: channel counts / pitches must be multiples of %d", n);
    Grid3 gi(B, Xi, Yi, Zi), go(B, Xo, Yo, Zo);
    const int chunks = C / n;
    TDB_REQUIRE(go.vox_p < (1ll << 31) && chunks <= kThreads, TDB_E_UNSUPPORTED, "tdb_trilinear: grid too large for 32-bit indexing, or more than %d channel vectors of 16 bytes per voxel", kThreads);
    dim3 grid((unsigned)blocks_per_sample((int64_t)go.Xp * go.Yp * chunks, B), (unsigned)B);
    const RowSplit split = make_split(go);
    auto scale_of = [](int n_in, int n_out) { return n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.0f; };
    const float sx = scale_of(Xi, Xo), sy = scale_of(Yi, Yo), sz = scale_of(Zi, Zo);
    cudaStream_t s = (cudaStream_t)stream;
    if ((int64_t)Xo * Yo * Zo < (int64_t)Xi * Yi * Zi) {
        // down-sampling: (nearly) every input voxel is read exactly once, so the gather form costs no extra traffic and
        // exposes rows x chunks parallelism (the line walker has only lines x chunks work items: 4.5 blocks per SM at level 0)
        dim3 ggrid((unsigned)blocks_per_sample(ceil_div(go.vox_p * chunks, 2), B), (unsigned)B);
        if (dtype == TDB_BF16)
            trilinear_gather_kernel<bf16><<<ggrid, kThreads, 0, s>>>((const bf16*)in, ld_in, gi, (bf16*)out, ld_out, go, C, split, chunks, sx, sy, sz);
        else
            trilinear_gather_kernel<float><<<ggrid, kThreads, 0, s>>>((const float*)in, ld_in, gi, (float*)out, ld_out, go, C, split, chunks, sx, sy, sz);
        TDB_CHECK_LAUNCH("tdb_trilinear");
        return 0;
    }
    // up-sampling with a channel-vector count that divides a warp and an x/y-blended input line per warp that fits in
    // shared memory: the two-stage line kernel (bit-identical to the walker, fewer instructions)
    static const bool no_line = std::getenv("TURBDIFF_B200_TRILINEAR_WALKER") != nullptr;
    const size_t line_smem = (size_t)go.Zp * sizeof(int4) + (size_t)TRI_WARPS * Zi * C * sizeof(float);
    // (measured at B = 8: 194x50x50 output 195 -> 177 us; 97x25x25 and below 3-10 % slower than the walker: large outputs only)
    const bool big = force_line || (int64_t)B * go.vox_p >= 1500000;
    if (!no_line && big && chunks <= 32 && 32 % chunks == 0 && line_smem <= 48 * 1024 && (int64_t)Zi * ld_in < (1ll << 31)) {
        const int64_t lines = (int64_t)go.Xp * go.Yp;
        int64_t blocks = ceil_div(lines, TRI_WARPS);
        const int64_t cap = (148 * 32) / (B < 1 ? 1 : B);  // a few waves of 8 resident blocks per SM
        if (blocks > cap) blocks = cap < 8 ? 8 : cap;
        dim3 lgrid((unsigned)blocks, (unsigned)B);
        static const bool carve = [] {  // 26 KB per 4-warp block: let eight of them share an SM
            cudaFuncSetAttribute(trilinear_up_line_kernel<bf16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(trilinear_up_line_kernel<float>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            return true;
        }();
        (void)carve;
        if (dtype == TDB_BF16)
            trilinear_up_line_kernel<bf16><<<lgrid, TRI_WARPS * 32, line_smem, s>>>((const bf16*)in, ld_in, gi, (bf16*)out, ld_out, go, C, split,
                                                                                 chunks, sx, sy, sz);
        else
            trilinear_up_line_kernel<float><<<lgrid, TRI_WARPS * 32, line_smem, s>>>((const float*)in, ld_in, gi, (float*)out, ld_out, go, C, split,
                                                                                  chunks, sx, sy, sz);
        TDB_CHECK_LAUNCH("tdb_trilinear");
        return 0;
    }
    if (dtype == TDB_BF16)
        trilinear_kernel<bf16><<<grid, kThreads, 0, s>>>((const bf16*)in, ld_in, gi, (bf16*)out, ld_out, go, C, split, chunks, sx, sy, sz);
    else
        trilinear_kernel<float><<<grid, kThreads, 0, s>>>((const float*)in, ld_in, gi, (float*)out, ld_out, go, C, split, chunks, sx, sy, sz);
    TDB_CHECK_LAUNCH("tdb_trilinear");
    return 0;
}

}  // extern "C"
