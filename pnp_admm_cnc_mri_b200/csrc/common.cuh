// common.cuh — shared types and the pointwise math of the ADMM loop.
//
// Everything here is HOST+DEVICE so the same code is exercised by the CPU emulation harness
// (tests/host_emu/) and by the sm_100a kernels.
#pragma once

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define PNP_HD __host__ __device__ __forceinline__
#define PNP_D __device__ __forceinline__
#else
#define PNP_HD inline
#define PNP_D inline
#endif

namespace pnp {

// Interleaved complex, layout-compatible with float2 / double2 / numpy complex64/128.
template <typename T>
struct alignas(2 * sizeof(T)) cx {
    T re, im;
};

template <typename T> PNP_HD cx<T> mk(T re, T im) { cx<T> r; r.re = re; r.im = im; return r; }
template <typename T> PNP_HD cx<T> operator+(cx<T> a, cx<T> b) { return mk<T>(a.re + b.re, a.im + b.im); }
template <typename T> PNP_HD cx<T> operator-(cx<T> a, cx<T> b) { return mk<T>(a.re - b.re, a.im - b.im); }
template <typename T> PNP_HD cx<T> operator*(cx<T> a, cx<T> b) {
    return mk<T>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
template <typename T> PNP_HD cx<T> conj(cx<T> a) { return mk<T>(a.re, -a.im); }
// multiply by conj(b)
template <typename T> PNP_HD cx<T> mulc(cx<T> a, cx<T> b) {
    return mk<T>(a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im);
}
// a * (-i)  and  a * (+i)
template <typename T> PNP_HD cx<T> mul_mi(cx<T> a) { return mk<T>(a.im, -a.re); }
template <typename T> PNP_HD cx<T> mul_pi(cx<T> a) { return mk<T>(-a.im, a.re); }
// twiddle multiply: forward uses w, inverse uses conj(w)
template <bool INV, typename T> PNP_HD cx<T> twmul(cx<T> a, cx<T> w) { return INV ? mulc(a, w) : a * w; }
// multiply by -i (forward) / +i (inverse)
template <bool INV, typename T> PNP_HD cx<T> rot90(cx<T> a) { return INV ? mul_pi(a) : mul_mi(a); }

PNP_HD float  pabs(float v)  { return fabsf(v); }
PNP_HD double pabs(double v) { return fabs(v); }
PNP_HD float  pmax(float a, float b)   { return fmaxf(a, b); }
PNP_HD double pmax(double a, double b) { return fmax(a, b); }
PNP_HD float  pmin(float a, float b)   { return fminf(a, b); }
PNP_HD double pmin(double a, double b) { return fmin(a, b); }
PNP_HD float  pcopysign(float a, float b)   { return copysignf(a, b); }
PNP_HD double pcopysign(double a, double b) { return copysign(a, b); }
PNP_HD float  psqrt(float a)  { return sqrtf(a); }
PNP_HD double psqrt(double a) { return sqrt(a); }

// a1  soft(x, c) = fmax(|x| - c, 0) * sign(x), sign(0) = 0            (reference S1:18-19)
template <typename T> PNP_HD T soft(T x, T c) {
    T m = pmax(pabs(x) - c, T(0));
    return x == T(0) ? T(0) : pcopysign(m, x);
}

// Scalars of one ADMM run, rounded once from the double-precision host values.
template <typename T>
struct ProxParams {
    int prox;      // PNPADMM_PROX_L1 / _CNC / PROX_NONE (x-update only)
    T thr_l1;      // reo * lambda1                     S1:123
    T inv_b;       // 1 / b                             S4:127
    T one_m_alpha; // 1 - alpha                         S4:128
    T alpha;       // alpha
    T coef;        // alpha * reo * lambda1 * b         S4:128
    T thr_cnc;     // alpha * reo * lambda1             S4:129
};

enum { PROX_L1 = 0, PROX_CNC = 1, PROX_NONE = 2 };

// a4/a5 z-update + a6 dual update for one pixel.  x >= 0 is the fresh x-update.
template <typename T>
PNP_HD void prox_dual(const ProxParams<T>& p, T x, T& z, T& w) {
    T zn;
    if (p.prox == PROX_L1) {
        zn = soft(x + w, p.thr_l1);                                         // S1:123
    } else {
        T s = soft(z, p.inv_b);                                             // S4:127
        T t = p.one_m_alpha * z + p.alpha * (x + w) + p.coef * (z - s);     // S4:128
        zn = soft(t, p.thr_cnc);                                            // S4:129
    }
    w = w + x - zn;                                                         // S1:126
    z = zn;
}

template <typename T> PNP_HD T clamp01(T v) { return pmin(pmax(v, T(0)), T(1)); }

}  // namespace pnp
