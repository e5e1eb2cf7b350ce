// Timestep conditioning in one launch: Nyquist sine embedding -> process_c MLP -> every
// ResnetBlock's FiLM projection (ddpm.py:147-148, 447-452, 184/191).  grid = (row blocks, B): every
// CTA recomputes the tiny MLP (8k MACs) and then projects 256 FiLM rows with coalesced reads of the
// transposed projection matrix.
// The embedding argument is a fused multiply-add like the reference's torch.addcmul: the
// arguments reach ~480 rad where one fp32 ulp is 3e-5, so the rounding order is visible.
#include "common.cuh"

using namespace tdb;

namespace {

__global__ void __launch_bounds__(256)
time_film_kernel(const int64_t* __restrict__ t, const float* __restrict__ emb_scale,
                 const float* __restrict__ emb_bias, const float* __restrict__ w1, const float* __restrict__ b1,
                 const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ film_wt,
                 const float* __restrict__ film_b, float* __restrict__ c_out, float* __restrict__ film,
                 int dim, int film_rows) {
    extern __shared__ float sm[];
    float* emb = sm;            // [dim]
    float* h1 = emb + dim;      // [4*dim]
    float* c = h1 + 4 * dim;    // [dim]
    const int b = blockIdx.y;
    const float tf = (float)t[b];
    for (int i = threadIdx.x; i < dim; i += blockDim.x) emb[i] = sinf(fmaf(emb_scale[i], tf, emb_bias[i]));
    __syncthreads();
    for (int r = threadIdx.x; r < 4 * dim; r += blockDim.x) {
        float acc = b1[r];
        for (int k = 0; k < dim; ++k) acc = fmaf(w1[r * dim + k], emb[k], acc);
        h1[r] = silu_f(acc);
    }
    __syncthreads();
    for (int r = threadIdx.x; r < dim; r += blockDim.x) {
        float acc = b2[r];
        for (int k = 0; k < 4 * dim; ++k) acc = fmaf(w2[r * 4 * dim + k], h1[k], acc);
        c[r] = silu_f(acc);
        if (blockIdx.x == 0) c_out[(int64_t)b * dim + r] = c[r];
    }
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < film_rows) {
        float acc = film_b[r];
        for (int k = 0; k < dim; ++k) acc = fmaf(film_wt[(int64_t)k * film_rows + r], c[k], acc);
        film[(int64_t)b * film_rows + r] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Backward of the timestep conditioning (autograd of ddpm.py:447-452, 184/191 in the reference): two launches instead
// of ~20 small torch kernels with cuBLAS matmuls (which also kept the backward CUDA graph from being captured in the
// default capture mode).
//   A (grid over FiLM rows): g_film_w[r][k] = sum_b d_film[b][r] c[b][k],  g_film_b[r] = sum_b d_film[b][r],
//                            dc[b][k] += sum_r d_film[b][r] film_w[r][k]           (block partials, fp32 atomics)
//   B (one block):           the 32 -> 128 -> 32 MLP with SiLU after both linears, differentiated by hand
__device__ __forceinline__ float dsilu_exact(float z) {
    const float s = 1.0f / (1.0f + expf(-z));
    return s * (1.0f + z * (1.0f - s));
}

__global__ void __launch_bounds__(256)
time_film_bwd_rows_kernel(const float* __restrict__ d_film, const float* __restrict__ c, const float* __restrict__ film_wt,
                          float* __restrict__ g_film_w, float* __restrict__ g_film_b, float* __restrict__ dc, int B, int dim,
                          int film_rows) {
    extern __shared__ float sm[];
    float* sc = sm;              // [B][dim]
    float* sdc = sc + B * dim;   // [B][dim] block partial of dc
    for (int i = threadIdx.x; i < B * dim; i += blockDim.x) {
        sc[i] = c[i];
        sdc[i] = 0.0f;
    }
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = r < film_rows;
    float gb = 0.0f;
    constexpr int KU = 8;  // film_wt loads in flight (the k loop used to be a chain of dependent-latency loads)
    for (int k0 = 0; k0 < dim; k0 += KU) {
        float w[KU], gw[KU];
#pragma unroll
        for (int u = 0; u < KU; ++u) {
            w[u] = (on && k0 + u < dim) ? film_wt[(int64_t)(k0 + u) * film_rows + r] : 0.0f;
            gw[u] = 0.0f;
        }
        for (int b = 0; b < B; ++b) {
            const float d = on ? d_film[(int64_t)b * film_rows + r] : 0.0f;  // (L1-resident after the first trip)
            if (k0 == 0) gb += d;
#pragma unroll
            for (int u = 0; u < KU; ++u) {
                if (k0 + u >= dim) break;
                gw[u] = fmaf(d, sc[b * dim + k0 + u], gw[u]);
                const float part = warp_sum(d * w[u]);
                if ((threadIdx.x & 31) == 0) atomicAdd(&sdc[b * dim + k0 + u], part);
            }
        }
#pragma unroll
        for (int u = 0; u < KU; ++u)
            if (on && k0 + u < dim) g_film_w[(int64_t)r * dim + k0 + u] = gw[u];
    }
    if (on) g_film_b[r] = gb;
    __syncthreads();
    for (int i = threadIdx.x; i < B * dim; i += blockDim.x) atomicAdd(&dc[i], sdc[i]);
}

__global__ void __launch_bounds__(256)
time_film_bwd_mlp_kernel(const int64_t* __restrict__ t, const float* __restrict__ emb_scale, const float* __restrict__ emb_bias,
                         const float* w1, const float* __restrict__ b1, const float* w2,
                         const float* __restrict__ b2, const float* __restrict__ dc, float* __restrict__ g_w1,
                         float* __restrict__ g_b1, float* __restrict__ g_w2, float* __restrict__ g_b2, int B, int dim) {
    extern __shared__ float sm[];
    const int H = 4 * dim;
    float* emb = sm;             // [B][dim]
    float* z1 = emb + B * dim;   // [B][H]
    float* h1 = z1 + B * H;      // [B][H]
    float* dz2 = h1 + B * H;     // [B][dim]
    float* dz1 = dz2 + B * dim;  // [B][H]
    float* sw1 = dz1 + B * H;    // [H][dim]   the two weight matrices, staged with coalesced loads: the loops below walk them
    float* sw2 = sw1 + H * (dim + 1);  // [dim][H + 1]   along k, which from global memory is one latency per element
    const int p1 = dim + 1, p2 = H + 1;       // padded pitches: threads walk different rows at the same k
    for (int i = threadIdx.x; i < H * dim; i += blockDim.x) {
        sw1[(i / dim) * p1 + i % dim] = w1[i];
        sw2[(i / H) * p2 + i % H] = w2[i];
    }
    w1 = sw1;
    w2 = sw2;
    for (int i = threadIdx.x; i < B * dim; i += blockDim.x) {
        const int b = i / dim, k = i % dim;
        emb[i] = sinf(fmaf(emb_scale[k], (float)t[b], emb_bias[k]));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < B * H; i += blockDim.x) {
        const int b = i / H, r = i % H;
        float acc = b1[r];
        for (int k = 0; k < dim; ++k) acc = fmaf(w1[r * p1 + k], emb[b * dim + k], acc);
        z1[i] = acc;
        h1[i] = silu_f(acc);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < B * dim; i += blockDim.x) {
        const int b = i / dim, r = i % dim;
        float acc = b2[r];
        for (int k = 0; k < H; ++k) acc = fmaf(w2[r * p2 + k], h1[b * H + k], acc);
        dz2[i] = dc[i] * dsilu_exact(acc);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < B * H; i += blockDim.x) {
        const int b = i / H, j = i % H;
        float acc = 0.0f;
        for (int r = 0; r < dim; ++r) acc = fmaf(dz2[b * dim + r], w2[r * p2 + j], acc);
        dz1[i] = acc * dsilu_exact(z1[i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < dim * H; i += blockDim.x) {  // g_w2 [dim][H]
        const int r = i / H, j = i % H;
        float acc = 0.0f;
        for (int b = 0; b < B; ++b) acc = fmaf(dz2[b * dim + r], h1[b * H + j], acc);
        g_w2[i] = acc;
    }
    for (int i = threadIdx.x; i < H * dim; i += blockDim.x) {  // g_w1 [H][dim]
        const int j = i / dim, k = i % dim;
        float acc = 0.0f;
        for (int b = 0; b < B; ++b) acc = fmaf(dz1[b * H + j], emb[b * dim + k], acc);
        g_w1[i] = acc;
    }
    for (int r = threadIdx.x; r < dim; r += blockDim.x) {
        float acc = 0.0f;
        for (int b = 0; b < B; ++b) acc += dz2[b * dim + r];
        g_b2[r] = acc;
    }
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        float acc = 0.0f;
        for (int b = 0; b < B; ++b) acc += dz1[b * H + j];
        g_b1[j] = acc;
    }
}

}  // namespace

extern "C" int tdb_time_film_bwd(const int64_t* t, const float* emb_scale, const float* emb_bias, const float* w1, const float* b1,
                                 const float* w2, const float* b2, const float* film_wt, const float* c, const float* d_film,
                                 float* g_film_w, float* g_film_b, float* g_w1, float* g_b1, float* g_w2, float* g_b2,
                                 float* dc_scratch, int B, int dim, int film_rows, void* stream) {
    TDB_REQUIRE(t && emb_scale && emb_bias && w1 && b1 && w2 && b2 && film_wt && c && d_film && g_film_w && g_film_b && g_w1 && g_b1 &&
                    g_w2 && g_b2 && dc_scratch,
                TDB_E_BADARG, "tdb_time_film_bwd: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem_a = (size_t)2 * B * dim * sizeof(float);
    const size_t smem_b = ((size_t)B * dim * 14 + (size_t)8 * dim * dim + 5 * dim) * sizeof(float);  // emb + dz2: 2*dim, z1 + h1 + dz1: 12*dim per sample; w1, w2 (padded)
    TDB_REQUIRE(smem_a <= 48 * 1024 && smem_b <= 200 * 1024, TDB_E_UNSUPPORTED, "tdb_time_film_bwd: batch %d x dim %d does not fit", B, dim);
    cudaError_t e = cudaMemsetAsync(dc_scratch, 0, (size_t)B * dim * sizeof(float), s);
    TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_time_film_bwd: cudaMemsetAsync: %s", cudaGetErrorString(e));
    time_film_bwd_rows_kernel<<<(unsigned)ceil_div(film_rows > 0 ? film_rows : 1, 256), 256, smem_a, s>>>(d_film, c, film_wt, g_film_w, g_film_b,
                                                                                                      dc_scratch, B, dim, film_rows);
    TDB_CHECK_LAUNCH("tdb_time_film_bwd (rows)");
    if (smem_b > 40 * 1024) {
        e = cudaFuncSetAttribute(time_film_bwd_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b);
        TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_time_film_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    time_film_bwd_mlp_kernel<<<1, 256, smem_b, s>>>(t, emb_scale, emb_bias, w1, b1, w2, b2, dc_scratch, g_w1, g_b1, g_w2, g_b2, B, dim);
    TDB_CHECK_LAUNCH("tdb_time_film_bwd (mlp)");
    return 0;
}

extern "C" int tdb_time_film(const int64_t* t, const float* emb_scale, const float* emb_bias, const float* w1,
                             const float* b1, const float* w2, const float* b2, const float* film_wt,
                             const float* film_b, float* c, float* film, int B, int dim, int film_rows,
                             void* stream) {
    TDB_REQUIRE(t && emb_scale && emb_bias && w1 && b1 && w2 && b2 && film_wt && film_b && c && film, TDB_E_BADARG,
                "tdb_time_film: null pointer");
    const size_t smem = (size_t)6 * dim * sizeof(float);
    dim3 grid((unsigned)ceil_div(film_rows > 0 ? film_rows : 1, 256), (unsigned)B);
    time_film_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(t, emb_scale, emb_bias, w1, b1, w2, b2, film_wt, film_b, c,
                                                                film, dim, film_rows);
    TDB_CHECK_LAUNCH("tdb_time_film");
    return 0;
}
