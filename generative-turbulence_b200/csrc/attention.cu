// Bottleneck voxel self-attention (ddpm.py:295-308): one CTA per (sample, head); the whole
// sequence (108 voxels at the shapes config) lives in shared memory, so this is the
// single-tile case of a flash-style kernel: scores, softmax and the PV product never touch
// global memory.  bf16 path: both products on the tcgen05 tensor cores with TMEM accumulators
// (attention_tc_kernel); fp32 parity path and S > 128: CUDA-core FMA with fp32 accumulation.
#include "common.cuh"
#include "ptx.cuh"

#include <cstdlib>

using namespace tdb;
using bf16 = __nv_bfloat16;

namespace {

// ------------------------------------------------------------------------------------------------------------------
// tcgen05 path (bf16 storage, S <= 128, dim_head = 32): the single-tile case of a flash-style kernel on the tensor cores.
// One CTA per (sample, head), 128 threads; thread t owns query / key / voxel t = TMEM lane t.
//   S = Q K^T   tcgen05.mma M=128 N=128 K=32, operands K-major in 128B-swizzled shared memory, accumulator in TMEM
//   softmax     each thread reads ITS row of S from TMEM (tcgen05.ld), scales, masks keys >= S, exponentiates in fp32 and
//               writes the unnormalised probabilities as the bf16 A operand of the second product
//   O = P V     tcgen05.mma M=128 N=32 K=128 against V^T (d-major rows, keys contiguous); 1/sum applied on the way out
// Scores, probabilities and the PV product never touch global memory (reference: F.scaled_dot_product_attention in
// attention.py:12-15 called from Attention.forward ddpm.py:295-308).
constexpr int TC_S = 128;                 // padded sequence length = UMMA M
constexpr uint32_t TC_ROW = 128;          // bytes per shared-memory operand row = swizzle span
constexpr uint32_t TC_TILE = TC_S * TC_ROW;             // [128 rows][64 bf16]: 16 KB
constexpr uint32_t TC_VT_TILE = 32 * TC_ROW;            // [32 rows (d)][64 keys]: 4 KB
constexpr uint32_t TC_SMEM = 2 * TC_TILE + 2 * TC_VT_TILE + 2 * TC_TILE;  // Q, K, V^T (2 key halves), P (2 key halves)

// byte offset of 16-byte chunk `c16` of row `r` inside a 1024B-aligned K-major SWIZZLE_128B tile
__device__ __forceinline__ uint32_t swz128(int r, int c16) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4));
}

__global__ void __launch_bounds__(128, 1)
attention_tc_kernel(const bf16* __restrict__ qkv, int ld_qkv, bf16* __restrict__ out, int ld_out, Grid3 g, int heads, int S) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = ptx::smem_u32(smem);
    uint8_t* sQ = smem;
    uint8_t* sK = smem + TC_TILE;
    uint8_t* sVt = smem + 2 * TC_TILE;
    uint8_t* sP = smem + 2 * TC_TILE + 2 * TC_VT_TILE;
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;

    const int t = threadIdx.x, warp = t / 32;
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int hid = heads * 32;

    // zero the operand tiles: padded keys must contribute exactly 0 to P V (0 * garbage could be NaN)
    for (uint32_t i = t; i < TC_SMEM / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (t == 0) {
        ptx::mbar_init(ptx::smem_u32(&bars[0]), 1);
        ptx::mbar_init(ptx::smem_u32(&bars[1]), 1);
        ptx::fence_barrier_init();
    }
    if (warp == 0) {
        ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 256);  // S: 128 fp32 columns, O: 32 (power of two >= 160)
        ptx::tmem_relinquish();
    }
    __syncthreads();

    // ---- stage q, k (row = voxel) and v^T (row = feature) of this head ----
    if (t < S) {
        const int z = t % g.Z, y = (t / g.Z) % g.Y, x = t / (g.Z * g.Y);
        const bf16* row = qkv + g.row(b, x, y, z) * ld_qkv + h * 32;
        uint4 q[4], k[4], v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            q[c] = *reinterpret_cast<const uint4*>(row + 8 * c);
            k[c] = *reinterpret_cast<const uint4*>(row + hid + 8 * c);
            v[c] = *reinterpret_cast<const uint4*>(row + 2 * hid + 8 * c);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            *reinterpret_cast<uint4*>(sQ + swz128(t, c)) = q[c];
            *reinterpret_cast<uint4*>(sK + swz128(t, c)) = k[c];
        }
        // V^T: element (d, key t) -> tile t / 64, row d, column t % 64
        uint8_t* vt = sVt + (t >> 6) * TC_VT_TILE;
        const int col = t & 63;
        const bf16* vv = reinterpret_cast<const bf16*>(v);
#pragma unroll
        for (int d = 0; d < 32; ++d) *reinterpret_cast<bf16*>(vt + swz128(d, col >> 3) + (col & 7) * 2) = vv[d];
    }
    ptx::fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async proxy
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t bar_s = ptx::smem_u32(&bars[0]), bar_o = ptx::smem_u32(&bars[1]);
    const uint64_t desc0 = ptx::umma_smem_desc(0, TC_ROW);

    if (t == 0) {
        const uint32_t idesc = ptx::umma_idesc_bf16(128, 128);
        const uint64_t a = desc0 | (uint64_t)(((sbase) & 0x3FFFFu) >> 4);
        const uint64_t bq = desc0 | (uint64_t)(((sbase + TC_TILE) & 0x3FFFFu) >> 4);
#pragma unroll
        for (int k = 0; k < 2; ++k) ptx::umma_f16(tmem, a + (uint64_t)(2 * k), bq + (uint64_t)(2 * k), idesc, (uint32_t)(k != 0));
        ptx::umma_commit(bar_s);
    }
    ptx::mbar_wait(bar_s, 0);
    ptx::tc_fence_after();

    // ---- softmax of row t (fp32), probabilities -> bf16 A operand ----
    const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
    const float scale = rsqrtf(32.0f);
    float sc[TC_S];
#pragma unroll
    for (int c = 0; c < TC_S; c += 16) {
        uint32_t r[16];
        ptx::tmem_ld_x16(t_row + (uint32_t)c, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) sc[c + j] = (c + j < S) ? __uint_as_float(r[j]) * scale : -INFINITY;
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < TC_S; ++j) mx = fmaxf(mx, sc[j]);
    float sum = 0.0f;
#pragma unroll
    for (int c = 0; c < TC_S; c += 8) {
        uint4 pk;
        __nv_bfloat162* hp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            const float e0 = __expf(sc[c + j] - mx), e1 = __expf(sc[c + j + 1] - mx);
            const __nv_bfloat162 pr = __floats2bfloat162_rn(e0, e1);
            // the normaliser is the sum of the ROUNDED probabilities: what the second product actually adds up
            sum += __bfloat162float(pr.x) + __bfloat162float(pr.y);
            hp[j >> 1] = pr;
        }
        *reinterpret_cast<uint4*>(sP + (c >> 6) * TC_TILE + swz128(t, (c & 63) >> 3)) = pk;
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();

    if (t == 0) {
        const uint32_t idesc = ptx::umma_idesc_bf16(128, 32);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint64_t a = desc0 | (uint64_t)(((sbase + 2 * TC_TILE + 2 * TC_VT_TILE + (uint32_t)(k >> 2) * TC_TILE) & 0x3FFFFu) >> 4);
            const uint64_t bv = desc0 | (uint64_t)(((sbase + 2 * TC_TILE + (uint32_t)(k >> 2) * TC_VT_TILE) & 0x3FFFFu) >> 4);
            ptx::umma_f16(tmem + 128u, a + (uint64_t)(2 * (k & 3)), bv + (uint64_t)(2 * (k & 3)), idesc, (uint32_t)(k != 0));
        }
        ptx::umma_commit(bar_o);
    }
    ptx::mbar_wait(bar_o, 0);
    ptx::tc_fence_after();

    {
        uint32_t r0[16], r1[16];
        ptx::tmem_ld_x16(t_row + 128u, r0);
        ptx::tmem_ld_x16(t_row + 144u, r1);
        ptx::tmem_ld_wait();
        if (t < S) {
            const float inv = 1.0f / sum;
            const int z = t % g.Z, y = (t / g.Z) % g.Y, x = t / (g.Z * g.Y);
            bf16* orow = out + g.row(b, x, y, z) * ld_out + h * 32;
            uint4 o[4];
            __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(o);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                ho[j] = __floats2bfloat162_rn(__uint_as_float(r0[2 * j]) * inv, __uint_as_float(r0[2 * j + 1]) * inv);
                ho[8 + j] = __floats2bfloat162_rn(__uint_as_float(r1[2 * j]) * inv, __uint_as_float(r1[2 * j + 1]) * inv);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(orow + 8 * c) = o[c];
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, 256);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// CUDA-core path: fp32 parity path, and sequences longer than one tensor-core tile.

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
    static const bool force_simt = std::getenv("TURBDIFF_B200_ATTN_SIMT") != nullptr;
    if (dtype == TDB_BF16 && S <= TC_S && !force_simt && ld_qkv % 8 == 0 && ld_out % 8 == 0 && ((uintptr_t)qkv & 15) == 0 &&
        ((uintptr_t)out & 15) == 0) {
        const int tc_smem = (int)TC_SMEM + 1024;
        e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem);
        TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_attention: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attention_tc_kernel<<<(unsigned)(B * heads), 128, tc_smem, s>>>((const bf16*)qkv, ld_qkv, (bf16*)out, ld_out, g, heads, S);
        TDB_CHECK_LAUNCH("tdb_attention (tcgen05)");
        return 0;
    }
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
