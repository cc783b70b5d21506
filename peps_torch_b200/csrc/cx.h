// Minimal real/complex scalar helpers shared by the dense kernels.
#pragma once
#include <cuda_runtime.h>

namespace ctmb {

template <bool CPLX> struct Sc;
template <> struct Sc<false> {
    using T = double;
    __device__ __forceinline__ static T zero() { return 0.0; }
    __device__ __forceinline__ static T one() { return 1.0; }
    __device__ __forceinline__ static T make(double re, double) { return re; }
    __device__ __forceinline__ static double re(T a) { return a; }
    __device__ __forceinline__ static double im(T) { return 0.0; }
    __device__ __forceinline__ static T conj(T a) { return a; }
    __device__ __forceinline__ static T add(T a, T b) { return a + b; }
    __device__ __forceinline__ static T sub(T a, T b) { return a - b; }
    __device__ __forceinline__ static T mul(T a, T b) { return a * b; }
    __device__ __forceinline__ static T scale(T a, double s) { return a * s; }
    __device__ __forceinline__ static T fma(T a, T b, T c) { return ::fma(a, b, c); }  // a*b+c
    __device__ __forceinline__ static double abs2(T a) { return a * a; }
    __device__ __forceinline__ static double abs(T a) { return fabs(a); }
    __device__ __forceinline__ static T shfl_xor(T a, int o) { return __shfl_xor_sync(0xffffffffu, a, o); }
    __device__ __forceinline__ static T div(T a, T b) { return a / b; }
    __device__ __forceinline__ static T shfl_xor_m(unsigned m, T a, int o) { return __shfl_xor_sync(m, a, o); }
};
template <> struct Sc<true> {
    using T = double2;
    __device__ __forceinline__ static T zero() { return make_double2(0.0, 0.0); }
    __device__ __forceinline__ static T one() { return make_double2(1.0, 0.0); }
    __device__ __forceinline__ static T make(double re, double im) { return make_double2(re, im); }
    __device__ __forceinline__ static double re(T a) { return a.x; }
    __device__ __forceinline__ static double im(T a) { return a.y; }
    __device__ __forceinline__ static T conj(T a) { return make_double2(a.x, -a.y); }
    __device__ __forceinline__ static T add(T a, T b) { return make_double2(a.x + b.x, a.y + b.y); }
    __device__ __forceinline__ static T sub(T a, T b) { return make_double2(a.x - b.x, a.y - b.y); }
    __device__ __forceinline__ static T mul(T a, T b) {
        return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
    }
    __device__ __forceinline__ static T scale(T a, double s) { return make_double2(a.x * s, a.y * s); }
    __device__ __forceinline__ static T fma(T a, T b, T c) {
        return make_double2(::fma(a.x, b.x, ::fma(-a.y, b.y, c.x)), ::fma(a.x, b.y, ::fma(a.y, b.x, c.y)));
    }
    __device__ __forceinline__ static double abs2(T a) { return a.x * a.x + a.y * a.y; }
    __device__ __forceinline__ static double abs(T a) { return hypot(a.x, a.y); }
    __device__ __forceinline__ static T shfl_xor(T a, int o) {
        return make_double2(__shfl_xor_sync(0xffffffffu, a.x, o), __shfl_xor_sync(0xffffffffu, a.y, o));
    }
    __device__ __forceinline__ static T shfl_xor_m(unsigned m, T a, int o) {
        return make_double2(__shfl_xor_sync(m, a.x, o), __shfl_xor_sync(m, a.y, o));
    }
    __device__ __forceinline__ static T div(T a, T b) {
        double d = b.x * b.x + b.y * b.y;
        return make_double2((a.x * b.x + a.y * b.y) / d, (a.y * b.x - a.x * b.y) / d);
    }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <bool CPLX>
__device__ __forceinline__ typename Sc<CPLX>::T warp_sum_t(typename Sc<CPLX>::T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = Sc<CPLX>::add(v, Sc<CPLX>::shfl_xor(v, o));
    return v;
}

}  // namespace ctmb
