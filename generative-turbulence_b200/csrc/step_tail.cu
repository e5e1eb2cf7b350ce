// Fused tail of one ancestral-sampling step (sampling only): everything between the last convolution of the denoiser and
// the first convolution of the NEXT step in one HBM-bound pass over the level-0 grid.
//
//   act   = SiLU(GroupNorm(raw2)) + res          decode.0's second Block + residual     (ddpm.py:168-177, 197)
//   eps   = W_dec act + b_dec                    decode.1, 1x1x1 conv dim -> F (2F)       (ddpm.py:459, 505)
//   x'    = posterior update(x_t, eps, z, z')    tdb_ddpm_step's arithmetic, bit for bit   (ddpm.py:711-728, 797-814)
//   xin0  = W_enc x' + b_enc                     encode_x of the next step, written with its replicate halo (ddpm.py:495)
//
// Unfused, these are four launches (pointwise, decode_output, ddpm_step, encode_input) that write and re-read dec_out
// (2 x 271 MB at B = 8), eps and x_t; fused, act and eps never leave registers: 1.12 GB instead of 2.3 GB per step.
// Every intermediate is rounded exactly where the unfused kernels round it (act to the storage type, eps / x' fp32), so the
// two paths agree bit for bit given the same GroupNorm moments.
//
// One thread per haloed row (its DIM channels in registers); halo rows recompute their clamped source voxel (all loads hit
// L1/L2) and only write the encoded row; x' goes to a SECOND state buffer because halo rows still read the old state.
#include "common.cuh"
#include "diffusion_step.cuh"

#include <cstdlib>

using namespace tdb;
using bf16 = __nv_bfloat16;

namespace {

constexpr int kThreads = 128;
constexpr int FMAX = 8, DMAX = 64;
// decode.1 / encode_x weights of the running chain at FIXED offsets (immediate constant-bank operands): w_dec [Fo][dim] at 0,
// b_dec at OFF_BDEC, w_enc [dim][4] at OFF_WENC, b_enc at OFF_BENC; copied here device-to-device (stream ordered) by every
// tdb_step_tail call.  Constant-bank operands feed the FMAs directly: the 1x1x1 convolutions
// cost no load instructions (as shared-memory tables they were ~380 LDS per row and the kernel was LSU bound).
constexpr int OFF_BDEC = FMAX * DMAX, OFF_WENC = OFF_BDEC + FMAX, OFF_BENC = OFF_WENC + DMAX * 4;
__constant__ float c_tail[OFF_BENC + DMAX];

struct TailArgs {
    const void* raw; int ld_raw;
    const double* stats; const float* gamma; const float* beta;
    const void* res; int ld_res;
    const float* w_dec; const float* b_dec; int Fo;
    const float* x_in; const float* z; const float* z_bc; const float* x_bcs; const uint8_t* mask;
    const float* coef; const int32_t* t_ptr; float* x_out; float* eps_out; int F; unsigned flags;
    const float* w_enc; const float* b_enc; void* xin0; int ld_xin0;
    Grid3 g; int G; float eps_gn;
    FastDiv by_z, by_y;
};

template <typename T>
__device__ __forceinline__ float round_to(float v) {
    if constexpr (sizeof(T) == 2) return __bfloat162float(__float2bfloat16_rn(v));
    else return v;
}

template <typename T>
__device__ __forceinline__ float act_silu(float v) {
    if constexpr (sizeof(T) == 2) return silu_tanh(v);  // same function as tdb_pointwise's bf16 path
    else return silu_f(v);
}

template <typename T, int DIM, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
step_tail_kernel(const TailArgs A) {
    constexpr int N = Vec<T>::N, NV = DIM / N;
    __shared__ __align__(16) float s_ca[DIM], s_co[DIM];
    const Grid3& g = A.g;
    const int b = blockIdx.y;
    const int F = A.F, Fo = A.Fo;
    {
        const int cpg = DIM / A.G;
        const double inv_n = 1.0 / ((double)cpg * g.X * g.Y * g.Z);
        for (int c = threadIdx.x; c < DIM; c += kThreads) {
            const int gi = c / cpg;
            const double mean = A.stats[((int64_t)b * A.G + gi) * 2] * inv_n;
            const double var = fma(-mean, mean, A.stats[((int64_t)b * A.G + gi) * 2 + 1] * inv_n);
            const float mean_f = (float)mean;
            const float rstd = 1.0f / sqrtf(fmaxf((float)var, 0.0f) + A.eps_gn);
            const float a = rstd * A.gamma[c];
            s_ca[c] = a;
            s_co[c] = A.beta[c] - mean_f * a;
        }
    }
    const float* c_wdec = c_tail;               // [Fo][DIM]
    const float* c_bdec = c_tail + OFF_BDEC;    // [Fo]
    const float* c_wenc = c_tail + OFF_WENC;    // [DIM][4]
    const float* c_benc = c_tail + OFF_BENC;    // [DIM]
    __syncthreads();

    const int t = *A.t_ptr;
    const StepCoef k = load_coef(A.coef, t);
    const bool t0 = t == 0;
    const bool lvar = A.flags & TDB_STEP_LEARNED_VAR;
    const bool need_z = !t0;
    const bool need_zbc = !t0 && (A.flags & TDB_STEP_NOISE_BCS);
    const bool need_xb = need_zbc || (A.flags & TDB_STEP_FINAL);
    const int64_t nvox = (int64_t)g.X * g.Y * g.Z;
    const int64_t base = (int64_t)b * g.vox_p;
    const T* raw = static_cast<const T*>(A.raw);
    const T* res = static_cast<const T*>(A.res);
    T* xin0 = static_cast<T*>(A.xin0);

    for (uint32_t r = blockIdx.x * kThreads + threadIdx.x; r < (uint32_t)g.vox_p; r += gridDim.x * kThreads) {
        uint32_t q, zz, xx, yy;
        A.by_z.divmod(r, q, zz);
        A.by_y.divmod(q, xx, yy);
        const int xp = (int)xx, yp = (int)yy, zp = (int)zz;
        const int xs = clampi(xp, 1, g.X), ys = clampi(yp, 1, g.Y), zs = clampi(zp, 1, g.Z);
        const bool own = xs == xp && ys == yp && zs == zp;  // interior row: this thread also owns the voxel's state update
        const int64_t src = base + ((int64_t)xs * g.Yp + ys) * g.Zp + zs;
        const int64_t v = ((int64_t)(xs - 1) * g.Y + (ys - 1)) * g.Z + (zs - 1);

        uint4 rv[NV], sv[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            rv[j] = Vec<T>::load_raw(raw + src * A.ld_raw + j * N);
            sv[j] = Vec<T>::load_raw(res + src * A.ld_res + j * N);
        }
        float xt[4], zn[4], zb[4], xb[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            const int64_t o = ((int64_t)b * F + f) * nvox + v;
            const bool on = f < F;
            xt[f] = on ? A.x_in[o] : 0.0f;
            zn[f] = (on && need_z) ? A.z[o] : 0.0f;
            zb[f] = (on && need_zbc) ? A.z_bc[o] : 0.0f;
            xb[f] = (on && need_xb) ? A.x_bcs[o] : 0.0f;
        }
        const bool inside = A.mask[v] != 0;

        // decode.0 block 2: GroupNorm apply + SiLU + residual, rounded to the storage type like tdb_pointwise's output
        float act[DIM];
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            float a[N], rr[N];
            Vec<T>::unpack(rv[j], a);
            Vec<T>::unpack(sv[j], rr);
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int c = j * N + i;
                act[c] = round_to<T>(act_silu<T>(fmaf(s_ca[c], a[i], s_co[c])) + rr[i]);
            }
        }
        // decode.1 (same accumulation order as decode_output_kernel)
        float eps[FMAX];
#pragma unroll
        for (int f = 0; f < FMAX; ++f) {
            float acc = f < Fo ? c_bdec[f] : 0.0f;
            if (f < Fo) {
#pragma unroll
                for (int c = 0; c < DIM; ++c) acc = fmaf(c_wdec[f * DIM + c], act[c], acc);
            }
            eps[f] = acc;
        }
        // posterior update
        float xn[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) {
            float sg = k.sigma;
            if (lvar && !t0 && f < F) {
                float vw = 0.0f;  // eps[F + f] without a runtime-indexed (local-memory) array access
#pragma unroll
                for (int j = 0; j < FMAX; ++j)
                    if (j == F + f) vw = eps[j];
                sg = learned_sigma(vw, k);
            }
            xn[f] = f < F ? step_one(xt[f], eps[f], zn[f], zb[f], xb[f], inside, k, t0, A.flags, sg) : 0.0f;
        }
        if (own) {
#pragma unroll
            for (int f = 0; f < 4; ++f)
                if (f < F) A.x_out[((int64_t)b * F + f) * nvox + v] = xn[f];
            if (A.eps_out) {
#pragma unroll
                for (int f = 0; f < FMAX; ++f)
                    if (f < Fo) A.eps_out[((int64_t)b * Fo + f) * nvox + v] = eps[f];
            }
        }
        // encode_x of the next step (same accumulation order as encode_input_kernel), halo rows included
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            float o[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const int c = j * N + i;
                float acc = c_benc[c];
#pragma unroll
                for (int f = 0; f < 4; ++f)
                    if (f < F) acc = fmaf(c_wenc[c * 4 + f], xn[f], acc);
                o[i] = acc;
            }
            Vec<T>::store(xin0 + (base + r) * A.ld_xin0 + j * N, o);
        }
    }
}

template <typename T, int DIM>
int launch(const TailArgs& A, int B, cudaStream_t s) {
    int64_t blocks = ceil_div(A.g.vox_p, kThreads);
    const int64_t cap = (148 * 32) / (B < 1 ? 1 : B);
    if (blocks > cap) blocks = cap < 8 ? 8 : cap;
    // registers per thread trade against resident warps (the kernel is issue / latency bound: ~800 instructions per row)
    static const int variant = std::getenv("TDB_TAIL_MINB") ? std::atoi(std::getenv("TDB_TAIL_MINB")) : 4;
    const dim3 grid((unsigned)blocks, (unsigned)B);
    if (variant >= 8) step_tail_kernel<T, DIM, 8><<<grid, kThreads, 0, s>>>(A);
    else if (variant >= 6) step_tail_kernel<T, DIM, 6><<<grid, kThreads, 0, s>>>(A);
    else step_tail_kernel<T, DIM, 4><<<grid, kThreads, 0, s>>>(A);
    return 0;
}

}  // namespace

extern "C" int tdb_step_tail(const void* raw, int ld_raw, const double* stats, const float* gamma, const float* beta, const void* res,
                             int ld_res, const float* w_dec, const float* b_dec, int Fo, const float* x_in, const float* z,
                             const float* z_bc, const float* x_bcs, const uint8_t* mask, const float* coef, const int32_t* t_ptr,
                             float* x_out, float* eps_out, int F, unsigned flags, const float* w_enc, const float* b_enc, void* xin0,
                             int ld_xin0, int B, int X, int Y, int Z, int dim, int G, float eps_gn, int dtype, void* stream) {
    TDB_REQUIRE(raw && stats && gamma && beta && res && w_dec && b_dec && x_in && mask && coef && t_ptr && x_out && w_enc && b_enc && xin0,
                TDB_E_BADARG, "tdb_step_tail: null pointer");
    TDB_REQUIRE(x_in != x_out, TDB_E_BADARG, "tdb_step_tail: the state is double-buffered (x_out must differ from x_in)");
    TDB_REQUIRE(F >= 1 && F <= 4 && Fo >= F && Fo <= 8 && (!(flags & TDB_STEP_LEARNED_VAR) || Fo == 2 * F), TDB_E_UNSUPPORTED,
                "tdb_step_tail: F <= 4 state features, Fo <= 8 model outputs (F=%d Fo=%d)", F, Fo);
    TDB_REQUIRE(x_bcs || !(flags & (TDB_STEP_NOISE_BCS | TDB_STEP_FINAL)), TDB_E_BADARG, "tdb_step_tail: x_bcs required");
    TDB_REQUIRE(G >= 1 && dim % G == 0, TDB_E_BADARG, "tdb_step_tail: dim=%d G=%d", dim, G);
    const int n = dtype == TDB_BF16 ? 8 : 4;
    TDB_REQUIRE(ld_raw % n == 0 && ld_res % n == 0 && ld_xin0 % n == 0 && ((uintptr_t)raw & 15) == 0 && ((uintptr_t)res & 15) == 0 &&
                    ((uintptr_t)xin0 & 15) == 0,
                TDB_E_UNSUPPORTED, "tdb_step_tail: pitches must be multiples of %d elements and bases 16-byte aligned", n);
    TailArgs A;
    A.raw = raw; A.ld_raw = ld_raw; A.stats = stats; A.gamma = gamma; A.beta = beta; A.res = res; A.ld_res = ld_res;
    A.w_dec = w_dec; A.b_dec = b_dec; A.Fo = Fo; A.x_in = x_in; A.z = z; A.z_bc = z_bc; A.x_bcs = x_bcs; A.mask = mask;
    A.coef = coef; A.t_ptr = t_ptr; A.x_out = x_out; A.eps_out = eps_out; A.F = F; A.flags = flags;
    A.w_enc = w_enc; A.b_enc = b_enc; A.xin0 = xin0; A.ld_xin0 = ld_xin0;
    A.g = Grid3(B, X, Y, Z); A.G = G; A.eps_gn = eps_gn;
    A.by_z = FastDiv((uint32_t)A.g.Zp); A.by_y = FastDiv((uint32_t)A.g.Yp);
    TDB_REQUIRE(A.g.vox_p < (1ll << 31), TDB_E_UNSUPPORTED, "tdb_step_tail: grid too large for 32-bit indexing");
    cudaStream_t s = (cudaStream_t)stream;
    {
        // [w_dec | b_dec | w_enc | b_enc] -> constant bank, device to device on the launch stream (graph capturable)
        auto put = [&](const float* p, size_t n, size_t off) {
            return cudaMemcpyToSymbolAsync(c_tail, p, n * sizeof(float), off * sizeof(float), cudaMemcpyDeviceToDevice, s);
        };
        cudaError_t e = put(w_dec, (size_t)Fo * dim, 0);
        if (e == cudaSuccess) e = put(b_dec, (size_t)Fo, OFF_BDEC);
        if (e == cudaSuccess) e = put(b_enc, (size_t)dim, OFF_BENC);
        if (F == 4) {
            if (e == cudaSuccess) e = put(w_enc, (size_t)dim * 4, OFF_WENC);
        } else {  // rows of F < 4 weights into the fixed pitch of 4
            for (int c = 0; c < dim && e == cudaSuccess; ++c) e = put(w_enc + (size_t)c * F, (size_t)F, OFF_WENC + (size_t)c * 4);
        }
        TDB_REQUIRE(e == cudaSuccess, (int)e, "tdb_step_tail: cudaMemcpyToSymbolAsync: %s", cudaGetErrorString(e));
    }
#define TDB_TAIL(D)                                                   \
    case D:                                                           \
        if (dtype == TDB_BF16) launch<bf16, D>(A, B, s);              \
        else launch<float, D>(A, B, s);                               \
        break;
    switch (dim) {
        TDB_TAIL(8)
        TDB_TAIL(16)
        TDB_TAIL(32)
        TDB_TAIL(64)
        default:
            TDB_REQUIRE(false, TDB_E_UNSUPPORTED, "tdb_step_tail: dim must be 8, 16, 32 or 64 (got %d)", dim);
    }
#undef TDB_TAIL
    TDB_CHECK_LAUNCH("tdb_step_tail");
    return 0;
}
