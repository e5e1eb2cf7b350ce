// The per-element arithmetic of one ancestral-sampling update (shared by tdb_ddpm_step and the fused step tail).
// Explicitly rounded multiplies / adds (__fmul_rn / __fadd_rn, no FMA contraction) in the reference's evaluation order:
// given identical eps and noise the result is bit-identical to the reference's chain of elementwise torch kernels
// (ddpm.py:711-728, 797-814).
#pragma once

#include "common.cuh"

namespace tdb {

struct StepCoef {
    float recip, recipm1, c1, c2, sigma, sa, s1m, plv;  // learned variances: sigma slot = log beta_t, plv = posterior log-variance
};

__device__ __forceinline__ StepCoef load_coef(const float* __restrict__ coef, int t) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(coef + (int64_t)t * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(coef + (int64_t)t * 8 + 4));
    return StepCoef{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
}

// learned variances (ddpm.py:732-741, 804-805): std = exp(lerp(log beta_t, posterior log-variance, sigmoid(v)) / 2) per voxel;
// torch.lerp's two-sided formula (weight < 0.5: start + w*(end - start), else end - (end - start)*(1 - w))
__device__ __forceinline__ float learned_sigma(float vw, const StepCoef& k) {
    const float w = 1.0f / (1.0f + expf(-vw));
    const float d = k.plv - k.sigma;
    const float lv = w < 0.5f ? fmaf(w, d, k.sigma) : fmaf(-d, 1.0f - w, k.plv);
    return expf(0.5f * lv);
}

__device__ __forceinline__ float step_one(float xt, float e, float z, float zbc, float xb, bool inside,
                                          const StepCoef& k, bool t0, unsigned flags, float sigma) {
    float x0 = __fsub_rn(__fmul_rn(k.recip, xt), __fmul_rn(k.recipm1, e));
    if (!(flags & TDB_STEP_NOISE_BCS) && !inside) x0 = xt;
    if (flags & TDB_STEP_CLIP) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
    float x = __fadd_rn(__fmul_rn(k.c1, x0), __fmul_rn(k.c2, xt));
    if (!t0) {
        if (flags & TDB_STEP_NOISE_BCS) {
            x = __fadd_rn(x, __fmul_rn(sigma, z));
            if (!inside) x = __fadd_rn(__fmul_rn(k.sa, xb), __fmul_rn(k.s1m, zbc));
        } else {
            x = __fadd_rn(x, __fmul_rn(sigma, inside ? z : 0.0f));
        }
    }
    if ((flags & TDB_STEP_FINAL) && !inside) x = xb;
    return x;
}

}  // namespace tdb
