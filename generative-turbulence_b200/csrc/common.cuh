// Shared helpers for libturbdiff_b200 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/turbdiff_b200.h"

namespace tdb {

// ---- error reporting / launch accounting ---------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch();

#define TDB_REQUIRE(cond, code, ...)  \
    do {                              \
        if (!(cond)) {                \
            tdb::set_error(__VA_ARGS__); \
            return (code);            \
        }                             \
    } while (0)

// Checks the launch that was just issued; returns its cudaError_t (0 = ok) from the caller.
#define TDB_CHECK_LAUNCH(name)                                                   \
    do {                                                                         \
        tdb::count_launch();                                                     \
        cudaError_t e_ = cudaGetLastError();                                     \
        if (e_ != cudaSuccess) {                                                 \
            tdb::set_error("%s: launch failed: %s", name, cudaGetErrorString(e_)); \
            return (int)e_;                                                      \
        }                                                                        \
    } while (0)

// ---- halo-grid geometry ---------------------------------------------------------------------
struct Grid3 {
    int B, X, Y, Z;     // unhaloed dims
    int Xp, Yp, Zp;     // with halo
    int64_t vox_p;      // Xp*Yp*Zp
    int64_t rows;       // B*vox_p
    __host__ __device__ Grid3() {}
    __host__ __device__ Grid3(int B_, int X_, int Y_, int Z_)
        : B(B_), X(X_), Y(Y_), Z(Z_), Xp(X_ + 2), Yp(Y_ + 2), Zp(Z_ + 2) {
        vox_p = (int64_t)Xp * Yp * Zp;
        rows = (int64_t)B * vox_p;
    }
    // row of interior voxel (x,y,z) of sample b
    __host__ __device__ int64_t row(int b, int x, int y, int z) const {
        return (((int64_t)b * Xp + (x + 1)) * Yp + (y + 1)) * Zp + (z + 1);
    }
};

// ---- division by a launch-time constant (32-bit, dividend < 2^31): one mul.hi + shift ---------
struct FastDiv {
    uint32_t d, m, s;
    FastDiv() : d(1), m(0), s(0) {}
    explicit FastDiv(uint32_t div) : d(div), m(0), s(0) {
        if (div > 1) {
            int l = 0;
            while ((1u << l) < div) ++l;
            const int p = 31 + l;
            m = (uint32_t)((((uint64_t)1 << p) + div - 1) / div);
            s = (uint32_t)(p - 32);
        }
    }
    __device__ __forceinline__ uint32_t div(uint32_t n) const { return d == 1 ? n : (__umulhi(n, m) >> s); }
    __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
        q = div(n);
        r = n - q * d;
    }
};

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---- 16-byte channel vectors -------------------------------------------------------------------
template <typename T>
struct Vec;

template <>
struct Vec<float> {
    static constexpr int N = 4;
    // raw 16-byte register image: lets a kernel issue several independent loads before it converts / uses any of them
    __device__ static uint4 load_raw(const float* p) { return *reinterpret_cast<const uint4*>(p); }
    __device__ static void unpack(const uint4& t, float (&v)[4]) {
        v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
    }
    __device__ static void load(const float* p, float (&v)[4]) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ static void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};

template <>
struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ static uint4 load_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
    __device__ static void unpack(const uint4& t, float (&v)[8]) {
        // bf16 -> fp32 is a 16-bit shift: two integer ops per pair instead of the cvt path
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xFFFF0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xFFFF0000u);
        v[4] = __uint_as_float(t.z << 16); v[5] = __uint_as_float(t.z & 0xFFFF0000u);
        v[6] = __uint_as_float(t.w << 16); v[7] = __uint_as_float(t.w & 0xFFFF0000u);
    }
    __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
        uint4 t = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __bfloat1622float2(h[i]);
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
    __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint4 t;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = t;
    }
};

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }
// x * sigmoid(x) = h + h * tanh(h), h = x / 2: ONE special-function op (tanh.approx.f32, abs. error 2^-11) instead of
// ex2 + rcp.  Only for bf16-stored results (output rounding 2^-9); the fp32 parity path keeps silu_f.
__device__ __forceinline__ float silu_tanh(float v) {
    const float h = 0.5f * v;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace tdb
