// The DDPM process around the denoiser as single-pass, HBM-bound kernels on the reference's
// NCDHW fp32 tensors: the fused ancestral-sampling update, q_sample, the masked training loss
// (+ gradient) and the bit-exact cell indexing helpers.
//
// The update arithmetic uses explicitly rounded multiplies/adds (__fmul_rn/__fadd_rn, no FMA
// contraction) in the reference's evaluation order, so given identical eps and noise the result
// is bit-identical to the reference's chain of elementwise torch kernels (ddpm.py:711-728,797-814).
#include "common.cuh"
#include "diffusion_step.cuh"

using namespace tdb;

namespace {

constexpr int kThreads = 256;

// VEC = 4: float4 path (nvox % 4 == 0 and 16-byte aligned bases); VEC = 1: scalar fallback.
template <int VEC>
__global__ void __launch_bounds__(kThreads)
ddpm_step_kernel(const float* __restrict__ x_t, const float* __restrict__ eps, const float* __restrict__ z,
                 const float* __restrict__ z_bc, const float* __restrict__ x_bcs, const uint8_t* __restrict__ mask,
                 const float* __restrict__ coef, const int32_t* __restrict__ t_ptr, float* __restrict__ x_out,
                 int64_t rows, int64_t nvox, unsigned flags, int F) {
    const int t = *t_ptr;
    const StepCoef k = load_coef(coef, t);
    const bool t0 = t == 0;
    // learned variances: eps is the (B, 2F, nvox) model output - noise prediction in the first F channels of a sample,
    // variance weights in the last F
    const bool lvar = flags & TDB_STEP_LEARNED_VAR;
    const int64_t fn = (int64_t)F * nvox;
    const bool need_z = !t0;
    const bool need_zbc = !t0 && (flags & TDB_STEP_NOISE_BCS);
    const bool need_xb = need_zbc || (flags & TDB_STEP_FINAL);
    const int64_t per_row = nvox / VEC;
    const int64_t total = rows * per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = (i % per_row) * VEC;
        const int64_t off = i * VEC;
        const int64_t eoff = lvar ? off + (off / fn) * fn : off;
        if constexpr (VEC == 4) {
            const float4 xt = *reinterpret_cast<const float4*>(x_t + off);
            const float4 e = *reinterpret_cast<const float4*>(eps + eoff);
            float4 sg = make_float4(k.sigma, k.sigma, k.sigma, k.sigma);
            if (lvar && !t0) {
                const float4 vw = *reinterpret_cast<const float4*>(eps + eoff + fn);
                sg = make_float4(learned_sigma(vw.x, k), learned_sigma(vw.y, k), learned_sigma(vw.z, k), learned_sigma(vw.w, k));
            }
            const uchar4 m = *reinterpret_cast<const uchar4*>(mask + v);
            float4 zz = make_float4(0, 0, 0, 0), zb = zz, xb = zz;
            if (need_z) zz = *reinterpret_cast<const float4*>(z + off);
            if (need_zbc) zb = *reinterpret_cast<const float4*>(z_bc + off);
            if (need_xb) xb = *reinterpret_cast<const float4*>(x_bcs + off);
            float4 o;
            o.x = step_one(xt.x, e.x, zz.x, zb.x, xb.x, m.x != 0, k, t0, flags, sg.x);
            o.y = step_one(xt.y, e.y, zz.y, zb.y, xb.y, m.y != 0, k, t0, flags, sg.y);
            o.z = step_one(xt.z, e.z, zz.z, zb.z, xb.z, m.z != 0, k, t0, flags, sg.z);
            o.w = step_one(xt.w, e.w, zz.w, zb.w, xb.w, m.w != 0, k, t0, flags, sg.w);
            *reinterpret_cast<float4*>(x_out + off) = o;
        } else {
            const float sg = (lvar && !t0) ? learned_sigma(eps[eoff + fn], k) : k.sigma;
            x_out[off] = step_one(x_t[off], eps[eoff], need_z ? z[off] : 0.f, need_zbc ? z_bc[off] : 0.f,
                                  need_xb ? x_bcs[off] : 0.f, mask[v] != 0, k, t0, flags, sg);
        }
    }
}

__global__ void __launch_bounds__(kThreads)
q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise, const int64_t* __restrict__ t,
                const float* __restrict__ coef, const uint8_t* __restrict__ mask, float* __restrict__ out, int B,
                int F, int64_t nvox, int noise_bcs) {
    const int64_t per_sample = (int64_t)F * nvox;
    const int64_t total = (int64_t)B * per_sample;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i / per_sample);
        const StepCoef k = load_coef(coef, (int)t[b]);
        float v = __fadd_rn(__fmul_rn(k.sa, x0[i]), __fmul_rn(k.s1m, noise[i]));
        if (!noise_bcs && mask[i % nvox] == 0) v = x0[i];
        out[i] = v;
    }
}

__global__ void __launch_bounds__(kThreads)
masked_loss_kernel(const float* __restrict__ eps, const float* __restrict__ noise, const uint8_t* __restrict__ mask,
                   double* __restrict__ loss_acc, float* __restrict__ grad, int64_t total, int64_t nvox,
                   double inv_count, int l1) {
    __shared__ double warp_part[kThreads / 32];
    double local = 0.0;
    const float gscale = (float)((l1 ? 1.0 : 2.0) * inv_count);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const bool inside = mask[i % nvox] != 0;
        const float d = eps[i] - noise[i];
        float g = 0.0f;
        if (inside) {
            if (l1) {
                local += (double)fabsf(d);
                g = d > 0.f ? gscale : (d < 0.f ? -gscale : 0.f);
            } else {
                local += (double)d * (double)d;
                g = gscale * d;
            }
        }
        if (grad) grad[i] = g;
    }
    local = warp_sum(local);
    if (threadIdx.x % 32 == 0) warp_part[threadIdx.x / 32] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < kThreads / 32 ? warp_part[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(loss_acc, v * inv_count);
    }
}

__global__ void __launch_bounds__(kThreads)
where_cells_kernel(const float* __restrict__ a, const float* __restrict__ other, const uint8_t* __restrict__ mask,
                   float* __restrict__ out, int64_t total, int64_t nvox) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = mask[i % nvox] ? a[i] : (other ? other[i] : 0.0f);
}

__global__ void __launch_bounds__(kThreads)
select_cells_kernel(const float* __restrict__ x, const int64_t* __restrict__ cell_idx, float* __restrict__ out,
                    int64_t rows, int64_t nvox, int64_t n_cells) {
    const int64_t total = rows * n_cells;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / n_cells, j = i % n_cells;
        out[i] = x[r * nvox + cell_idx[j]];
    }
}

__global__ void __launch_bounds__(kThreads)
scatter_cells_kernel(const float* __restrict__ samples, const int64_t* __restrict__ cell_idx, float* __restrict__ grid,
                     int B, int F, int64_t nvox, int64_t n_cells) {
    const int64_t total = (int64_t)B * n_cells * F;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int f = (int)(i % F);
        const int64_t j = (i / F) % n_cells;
        const int64_t b = i / ((int64_t)F * n_cells);
        grid[(b * F + f) * nvox + cell_idx[j]] = samples[i];
    }
}

// the steps either side of the sampler, fused.
// grid[b][f][v] = fma(scale[f], value, shift[f]): OpenFOAMData.grid_embedding (data/ofles.py:220-240) followed by
// Normalization.normalize_grid (models/normalization.py:20-24: addcmul(-mean/std, 1/std, x)); scale = 1/std,
// shift = -mean/std as torch computed them.  value =
//   * the FIXED_VALUE boundary value of channel f where the voxel's boundary class fixes it (ofles.py:233-238;
//     written AFTER the cell samples in the reference, so it wins over a cell value),
//   * samples[b][j][f] on voxel cell_idx[j],
//   * 0 elsewhere.
// code[v]: bit 0 = voxel is a cell, bits 1..7 = boundary class (0 = none); bc_has / bc_val [class][F] say which
// channels the class fixes and to what (the host resolves overlapping boundaries in the reference's write order).
// Without boundary tables code is the plain inside mask (0 / 1).  Every grid element is written exactly once.
__global__ void __launch_bounds__(kThreads)
scatter_normalize_kernel(const float* __restrict__ samples, const int64_t* __restrict__ cell_idx, const uint8_t* __restrict__ code,
                         const uint8_t* __restrict__ bc_has, const float* __restrict__ bc_val, const float* __restrict__ scale,
                         const float* __restrict__ shift, float* __restrict__ grid, int B, int F, int64_t nvox, int64_t n_cells) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // voxels whose value does not come from a cell sample: boundary value or the normalised zero
    const int64_t total_fill = (int64_t)B * F * nvox;
    for (int64_t i = t0; i < total_fill; i += stride) {
        const int64_t v = i % nvox;
        const int f = (int)((i / nvox) % F);
        const int c = code[v];
        const int cls = c >> 1;
        const bool fixed = cls != 0 && bc_has != nullptr && bc_has[cls * F + f];
        if (fixed) grid[i] = __fmaf_rn(scale[f], bc_val[cls * F + f], shift[f]);
        else if (!(c & 1)) grid[i] = __fmaf_rn(scale[f], 0.0f, shift[f]);
    }
    // cells (disjoint from the elements written above)
    const int64_t total = (int64_t)B * n_cells * F;
    for (int64_t i = t0; i < total; i += stride) {
        const int f = (int)(i % F);
        const int64_t j = (i / F) % n_cells;
        const int64_t b = i / ((int64_t)F * n_cells);
        const int64_t v = cell_idx[j];
        const int cls = code[v] >> 1;
        if (cls != 0 && bc_has != nullptr && bc_has[cls * F + f]) continue;  // a fixed boundary value overrides the sample
        grid[(b * F + f) * nvox + v] = __fmaf_rn(scale[f], samples[i], shift[f]);
    }
}

// out[b][j][f] = fma(scale[f], x[b][f][cell_idx[j]], shift[f]): Normalization.denormalize_grid (normalization.py:26-30,
// addcmul(mean, std, x)), select_cells and the "b f c -> b c f" rearrangement of SampleStore.add_samples
// (models/metrics.py:50-57); scale = std, shift = mean.
__global__ void __launch_bounds__(kThreads)
gather_denormalize_kernel(const float* __restrict__ x, const int64_t* __restrict__ cell_idx, const float* __restrict__ scale,
                          const float* __restrict__ shift, float* __restrict__ out, int B, int F, int64_t nvox, int64_t n_cells) {
    const int64_t total = (int64_t)B * n_cells * F;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int f = (int)(i % F);
        const int64_t j = (i / F) % n_cells;
        const int64_t b = i / ((int64_t)F * n_cells);
        out[i] = __fmaf_rn(scale[f], x[(b * F + f) * nvox + cell_idx[j]], shift[f]);
    }
}

__global__ void __launch_bounds__(kThreads)
build_mask_kernel(const int64_t* __restrict__ cell_idx, uint8_t* __restrict__ mask, int64_t n_cells, int64_t nvox) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = cell_idx[i];
        if (v >= 0 && v < nvox) mask[v] = 1;
    }
}

int blocks_for(int64_t n) {
    int64_t b = ceil_div(n, kThreads);
    const int64_t cap = 148 * 8;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

int tdb_ddpm_step(const float* x_t, const float* eps, const float* z, const float* z_bc, const float* x_bcs,
                  const uint8_t* mask, const float* coef, const int32_t* t_ptr, float* x_out, int B, int F,
                  int64_t nvox, unsigned flags, void* stream) {
    TDB_REQUIRE(x_t && eps && mask && coef && t_ptr && x_out && z, TDB_E_BADARG, "tdb_ddpm_step: null pointer");
    TDB_REQUIRE(x_bcs || !(flags & (TDB_STEP_NOISE_BCS | TDB_STEP_FINAL)), TDB_E_BADARG, "tdb_ddpm_step: x_bcs required");
    TDB_REQUIRE(z_bc || !(flags & TDB_STEP_NOISE_BCS), TDB_E_BADARG, "tdb_ddpm_step: z_bc required with NOISE_BCS");
    const int64_t rows = (int64_t)B * F;
    auto al = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
    const bool vec = nvox % 4 == 0 && al(x_t) && al(eps) && al(z) && al(z_bc) && al(x_bcs) && al(x_out) &&
                     ((uintptr_t)mask & 3) == 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (vec)
        ddpm_step_kernel<4><<<blocks_for(rows * nvox / 4), kThreads, 0, s>>>(x_t, eps, z, z_bc, x_bcs, mask, coef, t_ptr,
                                                                              x_out, rows, nvox, flags, F);
    else
        ddpm_step_kernel<1><<<blocks_for(rows * nvox), kThreads, 0, s>>>(x_t, eps, z, z_bc, x_bcs, mask, coef, t_ptr, x_out,
                                                                          rows, nvox, flags, F);
    TDB_CHECK_LAUNCH("tdb_ddpm_step");
    return 0;
}

int tdb_q_sample(const float* x0, const float* noise, const int64_t* t, const float* coef, const uint8_t* mask,
                 float* out, int B, int F, int64_t nvox, int noise_bcs, void* stream) {
    TDB_REQUIRE(x0 && noise && t && coef && out && (noise_bcs || mask), TDB_E_BADARG, "tdb_q_sample: null pointer");
    q_sample_kernel<<<blocks_for((int64_t)B * F * nvox), kThreads, 0, (cudaStream_t)stream>>>(x0, noise, t, coef, mask, out,
                                                                                               B, F, nvox, noise_bcs);
    TDB_CHECK_LAUNCH("tdb_q_sample");
    return 0;
}

int tdb_masked_loss(const float* eps, const float* noise, const uint8_t* mask, double* loss_acc, float* grad, int B,
                    int F, int64_t nvox, int64_t n_inside, int l1, void* stream) {
    TDB_REQUIRE(eps && noise && mask && loss_acc && n_inside > 0, TDB_E_BADARG, "tdb_masked_loss: bad argument");
    const int64_t total = (int64_t)B * F * nvox;
    const double inv_count = 1.0 / ((double)B * F * (double)n_inside);
    masked_loss_kernel<<<blocks_for(total), kThreads, 0, (cudaStream_t)stream>>>(eps, noise, mask, loss_acc, grad, total,
                                                                                  nvox, inv_count, l1);
    TDB_CHECK_LAUNCH("tdb_masked_loss");
    return 0;
}

int tdb_where_cells(const float* a, const float* other, const uint8_t* mask, float* out, int64_t rows, int64_t nvox,
                    void* stream) {
    TDB_REQUIRE(a && mask && out, TDB_E_BADARG, "tdb_where_cells: null pointer");
    if (rows * nvox == 0) return 0;
    where_cells_kernel<<<blocks_for(rows * nvox), kThreads, 0, (cudaStream_t)stream>>>(a, other, mask, out, rows * nvox, nvox);
    TDB_CHECK_LAUNCH("tdb_where_cells");
    return 0;
}

int tdb_select_cells(const float* x, const int64_t* cell_idx, float* out, int64_t rows, int64_t nvox, int64_t n_cells,
                     void* stream) {
    if (rows * n_cells == 0) return 0;
    TDB_REQUIRE(x && cell_idx && out, TDB_E_BADARG, "tdb_select_cells: null pointer");
    select_cells_kernel<<<blocks_for(rows * n_cells), kThreads, 0, (cudaStream_t)stream>>>(x, cell_idx, out, rows, nvox, n_cells);
    TDB_CHECK_LAUNCH("tdb_select_cells");
    return 0;
}

int tdb_scatter_cells(const float* samples, const int64_t* cell_idx, float* grid, int B, int F, int64_t nvox,
                      int64_t n_cells, void* stream) {
    if ((int64_t)B * F * n_cells == 0) return 0;
    TDB_REQUIRE(samples && cell_idx && grid, TDB_E_BADARG, "tdb_scatter_cells: null pointer");
    scatter_cells_kernel<<<blocks_for((int64_t)B * F * n_cells), kThreads, 0, (cudaStream_t)stream>>>(samples, cell_idx, grid,
                                                                                                       B, F, nvox, n_cells);
    TDB_CHECK_LAUNCH("tdb_scatter_cells");
    return 0;
}

int tdb_scatter_normalize(const float* samples, const int64_t* cell_idx, const uint8_t* code, const uint8_t* bc_has, const float* bc_val,
                          const float* scale, const float* shift, float* grid, int B, int F, int64_t nvox, int64_t n_cells, void* stream) {
    if ((int64_t)B * F * nvox == 0) return 0;
    TDB_REQUIRE((cell_idx || n_cells == 0) && code && scale && shift && grid && (samples || n_cells == 0), TDB_E_BADARG,
                "tdb_scatter_normalize: null pointer");
    TDB_REQUIRE((bc_has == nullptr) == (bc_val == nullptr), TDB_E_BADARG, "tdb_scatter_normalize: bc_has and bc_val come together");
    scatter_normalize_kernel<<<blocks_for((int64_t)B * F * nvox), kThreads, 0, (cudaStream_t)stream>>>(samples, cell_idx, code, bc_has, bc_val,
                                                                                                      scale, shift, grid, B, F, nvox, n_cells);
    TDB_CHECK_LAUNCH("tdb_scatter_normalize");
    return 0;
}

int tdb_gather_denormalize(const float* x, const int64_t* cell_idx, const float* scale, const float* shift, float* out, int B, int F,
                           int64_t nvox, int64_t n_cells, void* stream) {
    if ((int64_t)B * F * n_cells == 0) return 0;
    TDB_REQUIRE(x && cell_idx && scale && shift && out, TDB_E_BADARG, "tdb_gather_denormalize: null pointer");
    gather_denormalize_kernel<<<blocks_for((int64_t)B * F * n_cells), kThreads, 0, (cudaStream_t)stream>>>(x, cell_idx, scale, shift, out, B,
                                                                                                          F, nvox, n_cells);
    TDB_CHECK_LAUNCH("tdb_gather_denormalize");
    return 0;
}

int tdb_build_mask(const int64_t* cell_idx, uint8_t* mask, int64_t n_cells, int64_t nvox, void* stream) {
    if (n_cells == 0) return 0;
    TDB_REQUIRE(cell_idx && mask, TDB_E_BADARG, "tdb_build_mask: null pointer");
    build_mask_kernel<<<blocks_for(n_cells), kThreads, 0, (cudaStream_t)stream>>>(cell_idx, mask, n_cells, nvox);
    TDB_CHECK_LAUNCH("tdb_build_mask");
    return 0;
}

}  // extern "C"
