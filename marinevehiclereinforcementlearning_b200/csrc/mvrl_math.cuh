// Small device-side math layer shared by every kernel of libmvrl.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace mvrl {

// ---- precision-generic libm wrappers -------------------------------------
template <typename T> struct Real;
template <> struct Real<float> {
    static __device__ __forceinline__ void sincos(float x, float* s, float* c) { sincosf(x, s, c); }
    static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float exp(float x) { return expf(x); }
    static __device__ __forceinline__ float fmod(float a, float b) { return fmodf(a, b); }
    static __device__ __forceinline__ float floor(float a) { return floorf(a); }
    static __device__ __forceinline__ float nan() { return __int_as_float(0x7fc00000); }
};
template <> struct Real<double> {
    static __device__ __forceinline__ void sincos(double x, double* s, double* c) { ::sincos(x, s, c); }
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
    static __device__ __forceinline__ double fmod(double a, double b) { return ::fmod(a, b); }
    static __device__ __forceinline__ double floor(double a) { return ::floor(a); }
    static __device__ __forceinline__ double nan() { return __longlong_as_double(0x7ff8000000000000LL); }
};

template <typename T> __device__ __forceinline__ T tabs(T x) { return x < T(0) ? -x : x; }
template <> __device__ __forceinline__ float tabs<float>(float x) { return fabsf(x); }
template <> __device__ __forceinline__ double tabs<double>(double x) { return fabs(x); }
__device__ __forceinline__ float fmaf_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fmaf_t(double a, double b, double c) { return fma(a, b, c); }
template <typename T> __device__ __forceinline__ T tmin(T a, T b) { return a < b ? a : b; }
template <typename T> __device__ __forceinline__ T tmax(T a, T b) { return a > b ? a : b; }
template <typename T> __device__ __forceinline__ T clampt(T x, T lo, T hi) { return tmax(lo, tmin(hi, x)); }
// numpy.sign: -1, 0, +1
template <typename T> __device__ __forceinline__ T sgn(T x) { return x > T(0) ? T(1) : (x < T(0) ? T(-1) : T(0)); }
template <typename T> __device__ __forceinline__ bool finite_t(T x);
template <> __device__ __forceinline__ bool finite_t<double>(double x) { return isfinite(x); }
template <> __device__ __forceinline__ bool finite_t<float>(float x) { return isfinite(x); }

#define MVRL_TWO_PI 6.283185307179586476925286766559

// Python's float `%` (== numpy.mod): result takes the divisor's sign, and an
// exact zero remainder is +0 for a positive divisor.  The reference wraps
// angles with it (dynamicsModel_BlueROV2_Heavy_6DoF.py:560, resources.py:92-93).
template <typename T> __device__ __forceinline__ T pymod_pos(T a, T b) {  // b > 0
    T r;
    if (tabs(a) < b) r = a;  // fmod is the identity here; skips the slow path
    else r = Real<T>::fmod(a, b);
    if (r != T(0)) { if (r < T(0)) r += b; }
    else r = T(0);
    return r;
}

// resources.angleError (resources.py:75-95)
template <typename T> __device__ __forceinline__ T angle_error(T psi_d, T psi) {
    const T tp = T(MVRL_TWO_PI);
    T a = pymod_pos(psi_d - psi, tp);
    T b = pymod_pos(psi - psi_d, tp);
    return a < b ? a : -b;
}

#ifndef MVRL_POSE_COMP
#define MVRL_POSE_COMP 1
#endif
// y += inc with a Kahan carry (compensated summation across RK4 sub-steps)
template <typename T> __device__ __forceinline__ void rk4_pose_update(T& y, T& carry, T inc) {
    const T t = inc - carry;
    const T s = y + t;
    carry = (s - y) - t;
    y = s;
}

// ---- Philox4x32-10 ---------------------------------------------------------
struct Philox {
    static __device__ __forceinline__ uint4 run(uint4 c, uint2 k) {
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
            c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
            k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
        }
        return c;
    }
    // block `blk` of the draw stream of (env, episode)
    static __device__ __forceinline__ uint4 draw(uint64_t seed, uint64_t env, uint32_t episode, uint32_t stream, uint32_t blk) {
        return run(make_uint4((uint32_t)env, (uint32_t)(env >> 32), episode, stream * 65536u + blk),
                   make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    }
};
// 24-bit uniform in [0,1): exact in fp32 and fp64
template <typename T> __device__ __forceinline__ T u01(uint32_t w) { return T(w >> 8) * T(1.0 / 16777216.0); }

// ---- warp helpers ----------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v, unsigned mask) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v, unsigned mask) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(mask, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v, unsigned mask) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(mask, v, o));
    return v;
}
__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) > v) {
        unsigned long long prev = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
        if (prev == old) break;
        old = prev;
    }
}
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
    unsigned long long* a = (unsigned long long*)addr;
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) < v) {
        unsigned long long prev = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
        if (prev == old) break;
        old = prev;
    }
}

}  // namespace mvrl
