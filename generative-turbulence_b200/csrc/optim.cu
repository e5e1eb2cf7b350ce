// Fused optimiser step for the training path: gradient-norm clipping
// (Lightning's gradient_clip_val = 0.1, norm; reference config/shapes_experiment.yaml:50-51) and
// torch.optim.RAdam (reference turbdiff/models/diffusion.py:216) over ALL parameter tensors in two launches:
//   tdb_grad_sqnorm   sum of squares of every gradient element (double accumulation)
//   tdb_radam_step    clip coefficient from that sum, then the RAdam update of params / exp_avg / exp_avg_sq
// Tensors are addressed through device-side pointer tables; work is split in fixed-size chunks (one block each).
#include "common.cuh"

using namespace tdb;

namespace {

constexpr int kOptThreads = 256;

struct ChunkRef {
    const float* g;
    float* p;
    float* m;
    float* v;
    int n;  // elements of this chunk
};

__device__ __forceinline__ ChunkRef chunk_of(const int64_t* pp, const int64_t* gp, const int64_t* mp, const int64_t* vp,
                                             const int64_t* numel, const int* chunk_tensor, const int64_t* chunk_off, int chunk) {
    const int t = chunk_tensor[blockIdx.x];
    const int64_t off = chunk_off[blockIdx.x];
    ChunkRef r;
    r.g = reinterpret_cast<const float*>(gp[t]) + off;
    r.p = pp ? reinterpret_cast<float*>(pp[t]) + off : nullptr;
    r.m = mp ? reinterpret_cast<float*>(mp[t]) + off : nullptr;
    r.v = vp ? reinterpret_cast<float*>(vp[t]) + off : nullptr;
    const int64_t left = numel[t] - off;
    r.n = (int)(left < chunk ? left : chunk);
    return r;
}

__global__ void __launch_bounds__(kOptThreads)
grad_sqnorm_kernel(const int64_t* __restrict__ gp, const int64_t* __restrict__ numel, const int* __restrict__ chunk_tensor,
                   const int64_t* __restrict__ chunk_off, int chunk, double* __restrict__ out) {
    const ChunkRef c = chunk_of(nullptr, gp, nullptr, nullptr, numel, chunk_tensor, chunk_off, chunk);
    float acc = 0.0f;  // <= chunk / 256 terms per thread
    for (int i = threadIdx.x; i < c.n; i += kOptThreads) acc = fmaf(c.g[i], c.g[i], acc);
    __shared__ double part[kOptThreads / 32];
    double s = warp_sum((double)acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x / 32] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int i = 0; i < kOptThreads / 32; ++i) tot += part[i];
        atomicAdd(out, tot);
    }
}

// torch.optim.RAdam (single-tensor formulation), with the clip coefficient applied to the gradient on the fly:
//   g      = grad * min(1, max_norm / (||grad|| + 1e-6))          (torch.nn.utils.clip_grad_norm_)
//   g     += weight_decay * p
//   m      = m + (g - m) * (1 - beta1)                            (lerp)
//   v      = v * beta2 + (1 - beta2) * g * g
//   p     -= step_size * m / (sqrt(v) + eps)      if rectified    (step_size = lr * rect * sqrt(bc2) / bc1)
//   p     -= step_size * m                        otherwise       (step_size = lr / bc1)
__global__ void __launch_bounds__(kOptThreads)
radam_step_kernel(const int64_t* __restrict__ pp, const int64_t* __restrict__ gp, const int64_t* __restrict__ mp,
                  const int64_t* __restrict__ vp, const int64_t* __restrict__ numel, const int* __restrict__ chunk_tensor,
                  const int64_t* __restrict__ chunk_off, int chunk, const double* __restrict__ sqnorm, float max_norm, float step_size,
                  float beta1, float beta2, float w1, float w2, float eps, float weight_decay, int rectified) {
    const ChunkRef c = chunk_of(pp, gp, mp, vp, numel, chunk_tensor, chunk_off, chunk);
    float coef = 1.0f;
    if (sqnorm) {
        const float total = (float)sqrt(*sqnorm);
        coef = fminf(max_norm / (total + 1e-6f), 1.0f);
    }
    for (int i = threadIdx.x; i < c.n; i += kOptThreads) {
        float g = c.g[i] * coef;
        const float p = c.p[i];
        if (weight_decay != 0.0f) g = fmaf(weight_decay, p, g);
        float m = c.m[i], v = c.v[i];
        m = fmaf(g - m, w1, m);
        v = fmaf(w2 * g, g, v * beta2);
        c.m[i] = m;
        c.v[i] = v;
        const float upd = rectified ? m / (sqrtf(v) + eps) : m;
        c.p[i] = fmaf(-step_size, upd, p);
    }
}

}  // namespace

extern "C" {

int tdb_grad_sqnorm(const int64_t* grad_ptrs, const int64_t* numel, const int* chunk_tensor, const int64_t* chunk_off, int n_chunks,
                    int chunk, double* out, void* stream) {
    TDB_REQUIRE(grad_ptrs && numel && chunk_tensor && chunk_off && out, TDB_E_BADARG, "tdb_grad_sqnorm: null pointer");
    TDB_REQUIRE(chunk >= kOptThreads && n_chunks >= 0, TDB_E_BADARG, "tdb_grad_sqnorm: bad chunking");
    if (n_chunks == 0) return 0;
    grad_sqnorm_kernel<<<n_chunks, kOptThreads, 0, (cudaStream_t)stream>>>(grad_ptrs, numel, chunk_tensor, chunk_off, chunk, out);
    TDB_CHECK_LAUNCH("tdb_grad_sqnorm");
    return 0;
}

int tdb_radam_step(const int64_t* param_ptrs, const int64_t* grad_ptrs, const int64_t* exp_avg_ptrs, const int64_t* exp_avg_sq_ptrs,
                   const int64_t* numel, const int* chunk_tensor, const int64_t* chunk_off, int n_chunks, int chunk, const double* sqnorm,
                   float max_norm, float step_size, double beta1, double beta2, float eps, float weight_decay, int rectified, void* stream) {
    TDB_REQUIRE(param_ptrs && grad_ptrs && exp_avg_ptrs && exp_avg_sq_ptrs && numel && chunk_tensor && chunk_off, TDB_E_BADARG,
                "tdb_radam_step: null pointer");
    TDB_REQUIRE(chunk >= kOptThreads && n_chunks >= 0, TDB_E_BADARG, "tdb_radam_step: bad chunking");
    if (n_chunks == 0) return 0;
    radam_step_kernel<<<n_chunks, kOptThreads, 0, (cudaStream_t)stream>>>(param_ptrs, grad_ptrs, exp_avg_ptrs, exp_avg_sq_ptrs, numel,
                                                                         chunk_tensor, chunk_off, chunk, sqnorm, max_norm, step_size, (float)beta1,
                                                                         (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), eps,
                                                                         weight_decay, rectified);
    TDB_CHECK_LAUNCH("tdb_radam_step");
    return 0;
}

}  // extern "C"
