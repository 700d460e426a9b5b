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

// state loads that must not hit a stale L1 line (another SM may have rewritten the plane)
#if defined(__CUDA_ARCH__)
PNP_D float ld_state(const float* p) { return __ldcg(p); }
#else
inline float ld_state(const float* p) { return *p; }
#endif

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
// General form (any c); the hot loops use the clamp form below when c >= 0.
template <typename T> PNP_HD T soft(T x, T c) {
    T m = pmax(pabs(x) - c, T(0));
    return x == T(0) ? T(0) : pcopysign(m, x);
}
// For c >= 0:  soft(x, c) == x - clamp(x, -c, c)  bit for bit (|x| <= c gives x - x = 0, otherwise the
// same single rounding of |x| - c with the sign restored) — 3 instructions instead of ~7.
template <typename T> PNP_HD T clampc(T x, T c) { return pmin(pmax(x, -c), c); }

// Scalars of one ADMM run, rounded once from the double-precision host values.
template <typename T>
struct ProxParams {
    int prox;      // PNPADMM_PROX_L1 / _CNC / PROX_NONE (x-update only)
    int general;   // 1 if some threshold is negative: use the general soft() form
    T thr_l1;      // reo * lambda1                     S1:123
    T inv_b;       // 1 / b                             S4:127
    T one_m_alpha; // 1 - alpha                         S4:128
    T alpha;       // alpha
    T coef;        // alpha * reo * lambda1 * b         S4:128
    T thr_cnc;     // alpha * reo * lambda1             S4:129
};

enum { PROX_L1 = 0, PROX_CNC = 1, PROX_NONE = 2 };

enum { PM_L1 = 0, PM_CNC = 1, PM_GENERAL = 3 };

// a4/a5 z-update + a6 dual update for one pixel, x >= 0 the fresh x-update.  MODE is a compile-time
// PM_* so the unrolled pixel loops of the kernels carry no run-time branch.
template <int MODE, typename T>
PNP_HD void prox_dual_m(const ProxParams<T>& p, T x, T& z, T& w) {
    T zn;
    const T xw = x + w;
    if (MODE == PM_GENERAL) {
        if (p.prox == PROX_L1) {
            zn = soft(xw, p.thr_l1);                                            // S1:123
        } else {
            T s = soft(z, p.inv_b);                                             // S4:127
            T t = p.one_m_alpha * z + p.alpha * xw + p.coef * (z - s);          // S4:128
            zn = soft(t, p.thr_cnc);                                            // S4:129
        }
    } else if (MODE == PM_L1) {
        zn = xw - clampc(xw, p.thr_l1);                                         // S1:123
    } else {
        T q = clampc(z, p.inv_b);                                               // z - soft(z, 1/b) up to rounding
        if (sizeof(T) == 8) { T s = z - q; q = z - s; }                         // fp64 build: the reference's two roundings
        T t = p.one_m_alpha * z + p.alpha * xw + p.coef * q;                    // S4:127-128
        zn = t - clampc(t, p.thr_cnc);                                          // S4:129
    }
    w = xw - zn;                                                                // S1:126  (w + x) - z
    z = zn;
}

template <typename T> PNP_HD int prox_mode(const ProxParams<T>& p) { return p.general ? PM_GENERAL : p.prox; }

// run-time dispatch (streaming kernels)
template <typename T>
PNP_HD void prox_dual(const ProxParams<T>& p, T x, T& z, T& w) {
    if (p.general) prox_dual_m<PM_GENERAL>(p, x, z, w);
    else if (p.prox == PROX_L1) prox_dual_m<PM_L1>(p, x, z, w);
    else prox_dual_m<PM_CNC>(p, x, z, w);
}

// blend_coef note.  The data-consistency residual uses cf[code] with code = m[k] + m[-k] in {0, 1, 2} and cf = {0, g / 2 N^2, g / N^2}
// (pnpadmm.cu: both rounded from the same double, so cf[2] == 2 cf[1] bit for bit).  The fp32 kernels therefore compute the coefficient as
// (float)code * cf[1]: one conversion and one multiply (blend_coef below).  The obvious `code == 0 ? 0 : (code == 1 ? cf1 : cf2)` compiles to divergent
// branches inside the unrolled 16-point loops (BSSY / BSYNC per point); replacing it is bit-identical and made K3 7 % (N = 256) to 23 %
// (N = 1024), K1 2-3 % and the columns pass at N = 1024 4 % faster (profiles/r2_measured_runs.md).
// MUL = true: (float)code * cf1 (one conversion, one multiply).  MUL = false: the select chain, kept for the columns pass at N = 256, the one
// kernel that measured slower with the product (7.77 -> 8.08 ms at B = 2048; an exact magic-number conversion without I2FP measured the same,
// so it is not the conversion pipe).
template <bool MUL> PNP_HD float blend_coef(uint32_t code, float cf1) {
    if (MUL) return (float)code * cf1;
    return code == 0 ? 0.f : (code == 1 ? cf1 : cf1 + cf1);
}

template <typename T> PNP_HD T clamp01(T v) { return pmin(pmax(v, T(0)), T(1)); }

}  // namespace pnp
