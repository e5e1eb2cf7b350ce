// Bottleneck voxel self-attention (ddpm.py:295-308): one CTA per (sample, head); the whole
// sequence (108 voxels at the shapes config) lives in shared memory, so this is the
// single-tile case of a flash-style kernel: scores, softmax and the PV product never touch
// global memory.  6 MFLOP per sample - latency-bound, CUDA-core FMA with fp32 accumulation.
#include "common.cuh"

using namespace tdb;
using bf16 = __nv_bfloat16;

namespace {

constexpr int DH = 32;      // head dim == warp size: lane d owns output feature d
constexpr int WARPS = 8;
constexpr int QB = 16;     // queries per CTA: grid = (B*heads, ceil(S/QB)) so that even 108 tokens fill ~100 SMs

template <typename T>
__global__ void __launch_bounds__(WARPS * 32)
attention_kernel(const T* __restrict__ qkv, int ld_qkv, T* __restrict__ out, int ld_out, Grid3 g, int heads, int S) {
    extern __shared__ float sm[];
    float* sq = sm;                      // [S][DH+1]
    float* sk = sq + (size_t)S * (DH + 1);
    float* sv = sk + (size_t)S * (DH + 1);
    float* sp = sv + (size_t)S * (DH + 1);  // [WARPS][S] probabilities
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int hid = heads * DH;
    const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const float scale = rsqrtf((float)DH);

    for (int i = threadIdx.x; i < S * DH; i += blockDim.x) {
        const int s = i / DH, d = i % DH;
        const int z = s % g.Z, y = (s / g.Z) % g.Y, x = s / (g.Z * g.Y);
        const T* row = qkv + g.row(b, x, y, z) * ld_qkv + h * DH + d;
        sq[s * (DH + 1) + d] = (float)row[0] * scale;
        sk[s * (DH + 1) + d] = (float)row[hid];
        sv[s * (DH + 1) + d] = (float)row[2 * hid];
    }
    __syncthreads();

    float* p = sp + (size_t)warp * S;
    const int q_end = min(S, (int)(blockIdx.y + 1) * QB);
    for (int i = blockIdx.y * QB + warp; i < q_end; i += WARPS) {
        const float* qi = sq + i * (DH + 1);
        float mx = -INFINITY;
        for (int j = lane; j < S; j += 32) {
            const float* kj = sk + j * (DH + 1);
            float acc = 0.0f;
#pragma unroll
            for (int d = 0; d < DH; ++d) acc = fmaf(qi[d], kj[d], acc);
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
        __syncwarp();
        float o = 0.0f;
        for (int j = 0; j < S; ++j) o = fmaf(p[j], sv[j * (DH + 1) + lane], o);
        o /= sum;
        const int z = i % g.Z, y = (i / g.Z) % g.Y, x = i / (g.Z * g.Y);
        out[g.row(b, x, y, z) * ld_out + h * DH + lane] = (T)o;
        __syncwarp();
    }
}

}  // namespace

extern "C" int tdb_attention(const void* qkv, int ld_qkv, void* out, int ld_out, int B, int X, int Y, int Z,
                             int heads, int dh, int dtype, void* stream) {
    TDB_REQUIRE(qkv && out, TDB_E_BADARG, "tdb_attention: null pointer");
    TDB_REQUIRE(dh == DH, TDB_E_UNSUPPORTED, "tdb_attention: dim_head must be 32 (got %d)", dh);
    const int S = X * Y * Z;
    const size_t smem = ((size_t)3 * S * (DH + 1) + (size_t)WARPS * S) * sizeof(float);
    TDB_REQUIRE(smem <= 220 * 1024, TDB_E_UNSUPPORTED, "tdb_attention: sequence of %d voxels exceeds shared memory", S);
    Grid3 g(B, X, Y, Z);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e;
    if (dtype == TDB_BF16) {
        e = cudaFuncSetAttribute(attention_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            attention_kernel<bf16><<<dim3((unsigned)(B * heads), (unsigned)((S + QB - 1) / QB)), WARPS * 32, smem, s>>>((const bf16*)qkv, ld_qkv, (bf16*)out, ld_out, g, heads, S);
    } else {
        e = cudaFuncSetAttribute(attention_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)
            attention_kernel<float><<<dim3((unsigned)(B * heads), (unsigned)((S + QB - 1) / QB)), WARPS * 32, smem, s>>>((const float*)qkv, ld_qkv, (float*)out, ld_out, g, heads, S);
    }
    TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    TDB_CHECK_LAUNCH("tdb_attention");
    return 0;
}
