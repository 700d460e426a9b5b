// rowsep_core.cuh — K3 at N = 512 / 1024 (and 256 for A/B): per-thread pieces of the row-separable solve on the K2 line FFT
// (stream2_core.cuh: T = N / 16 threads per line, thread t owns the points n = t + T m, m < 16, before and after a transform).
//
// Same mathematics as rowsep256.cuh (derivation in cluster256_core.cuh): for a mask made of full k-space lines, m[kr][kc] = m(kc),
//     ifft2_u(G - cf .* fft2 V) = rowIFFT_u( G' - N cf(kc) .* rowFFT(V) ),   G' = colIFFT_u(G) = N cf(kc) A0 + (1 + i hb) NcS'
//     T1 = rowIFFT_u(N ms(kc) A0 + (1 + i hb) nH'),  T2 = rowIFFT_u(N ma(kc) A0 + (1 + i hb) nA'),  A0 = rowFFT(image rows)
// so a line (one image row of a packed plane) is solved on its own for all iterations.  Per line in shared memory: the padded
// FFT line, z of both images, the row of G'; in registers: the 16 points and the dual w of the thread's 16 pixels x 2 images.
// HOST+DEVICE: tests/host_emu/s2_emu.cpp runs the same code against the oracle.
#pragma once

#include "stream2_core.cuh"

namespace pnp {
namespace k3 {

using k1::cf32;

struct LineState {
    cf32 a[16];     // working points n = t + T m
    float w[32];    // dual: w[2m] image a, w[2m+1] image b
};

// codes / m[k] bits of thread class t (kc = t + T m), from mask row 0
template <int N>
PNP_HD void line_words(const uint8_t* mask, int t, uint32_t* codes, uint32_t* here) {
    constexpr int T = N / 16;
    uint32_t cw = 0, hw = 0;
    for (int m = 0; m < 16; ++m) {
        const int kc = t + T * m, mkc = (N - kc) & (N - 1);
        const uint32_t m1 = mask[kc] ? 1u : 0u, m2 = mask[mkc] ? 1u : 0u;
        cw |= (m1 + m2) << (2 * m);
        hw |= m1 << m;
    }
    *codes = cw; *here = hw;
}
PNP_HD bool column_is_constant(const uint8_t* mask, int N, int kc) {
    const bool m0 = mask[kc] != 0;
    for (int kr = 1; kr < N; ++kr)
        if ((mask[(size_t)kr * N + kc] != 0) != m0) return false;
    return true;
}
// row-major noise-term planes [3][N][N]: NcS, nH, nA (input of the column inverse transform)
PNP_HD void noise_terms(const uint8_t* mask, const cf32* noise, int N, float g_over_n2, size_t bin, cf32* planes) {
    const int kr = (int)(bin / N), kc = (int)(bin % N);
    const size_t nn = (size_t)N * N;
    const size_t mbin = (size_t)((N - kr) & (N - 1)) * N + ((N - kc) & (N - 1));
    const float m1 = mask[bin] ? 1.f : 0.f, m2 = mask[mbin] ? 1.f : 0.f;
    const cf32 n1 = noise[bin], n2 = noise[mbin];
    planes[bin] = mk<float>(g_over_n2 * 0.5f * (m1 * n1.re + m2 * n2.re), g_over_n2 * 0.5f * (m1 * n1.im - m2 * n2.im));
    planes[nn + bin] = mk<float>(0.5f * (n1.re + n2.re), 0.5f * (n1.im - n2.im));
    planes[2 * nn + bin] = mk<float>(0.5f * (n1.re - n2.re), 0.5f * (n1.im + n2.im));
}

template <int N>
PNP_HD void load_image(LineState& s, int t, const float* fa, const float* fb, const uint8_t* ua, const uint8_t* ub) {
    constexpr int T = N / 16;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int n = t + T * m;
        const float a = fa ? fa[n] : k1::unit_from_u8(ua[n]);
        const float b = fb ? fb[n] : (ub ? k1::unit_from_u8(ub[n]) : 0.f);
        s.a[m] = mk<float>(a, b);
    }
}

// after A0 = rowFFT(image rows): keep A0 (in the z row), write G' row, leave the ms branch in s.a.  ncf = N cf, `nscale` = N.
template <int N>
PNP_HD void acquire_ms(LineState& s, int t, cf32* zs, cf32* gp, const cf32* NcSp, const cf32* nHp, uint32_t codes, float ncf1,
                       float ncf2, float hb) {
    constexpr int T = N / 16;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int n = t + T * m;
        const cf32 F = s.a[m];
        zs[n] = F;
        const uint32_t code = (codes >> (2 * m)) & 3u;
        const float cf = blend_coef<true>(code, ncf1);       // code in {0, 1, 2}, ncf2 == 2 ncf1 exactly: see blend_coef note in common.cuh
        const cf32 nc = NcSp[n], nh = nHp[n];
        gp[n] = mk<float>(cf * F.re + (nc.re - hb * nc.im), cf * F.im + (nc.im + hb * nc.re));
        const float ms = (0.5f * (float)N) * (float)code;
        s.a[m] = mk<float>(ms * F.re + (nh.re - hb * nh.im), ms * F.im + (nh.im + hb * nh.re));
    }
}
template <int N>
PNP_HD void acquire_ma(LineState& s, int t, const cf32* zs, const cf32* nAp, uint32_t codes, uint32_t here, float hb) {
    constexpr int T = N / 16;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int n = t + T * m;
        const cf32 F = zs[n];
        const uint32_t code = (codes >> (2 * m)) & 3u;
        const float ma = (float)N * ((float)((here >> m) & 1u) - 0.5f * (float)code);
        const cf32 na = nAp[n];
        s.a[m] = mk<float>(ma * F.re + (na.re - hb * na.im), ma * F.im + (na.im + hb * na.re));
    }
}
PNP_HD void stash_t1(LineState& s) {
#pragma unroll
    for (int m = 0; m < 16; ++m) { s.w[2 * m] = s.a[m].re; s.w[2 * m + 1] = s.a[m].im; }
}
template <int N>
PNP_HD void zero_fill(LineState& s, int t, cf32* zs, float inv_n2, bool has_b) {
    constexpr int T = N / 16;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const float t1r = s.w[2 * m], t1i = s.w[2 * m + 1];
        const float xa = psqrt(t1r * t1r + s.a[m].im * s.a[m].im) * inv_n2;
        const float xb = has_b ? psqrt(t1i * t1i + s.a[m].re * s.a[m].re) * inv_n2 : 0.f;
        const cf32 z = mk<float>(xa, xb);
        zs[t + T * m] = z;
        s.w[2 * m] = 0.f; s.w[2 * m + 1] = 0.f;
        s.a[m] = z;
    }
}
template <int N>
PNP_HD void blend(LineState& s, int t, const cf32* gp, uint32_t codes, float ncf1, float ncf2) {
    constexpr int T = N / 16;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const cf32 gg = gp[t + T * m];
        const uint32_t code = (codes >> (2 * m)) & 3u;
        const float cf = blend_coef<true>(code, ncf1);       // code in {0, 1, 2}, ncf2 == 2 ncf1 exactly: see blend_coef note in common.cuh
        (void)ncf2;
        s.a[m] = mk<float>(gg.re - cf * s.a[m].re, gg.im - cf * s.a[m].im);
    }
}
// s.a = r (residual correction): x = |v + r|, prox, dual, next a = z - w; `last`: x, z, w of the row go to global memory
template <int MODE, int N>
PNP_HD void prox_row(LineState& s, int t, cf32* zs, const ProxParams<float>& p, bool has_b, bool last, float* xa_o, float* za_o,
                     float* wa_o, float* xb_o, float* zb_o, float* wb_o) {
    constexpr int T = N / 16;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const int n = t + T * m;
        const cf32 zz = zs[n];
        float za = zz.re, zb = zz.im, wa = s.w[2 * m], wb = s.w[2 * m + 1];
        const float xa = pabs((za - wa) + s.a[m].re);
        const float xb = has_b ? pabs((zb - wb) + s.a[m].im) : 0.f;
        prox_dual_m<MODE>(p, xa, za, wa);
        if (has_b) prox_dual_m<MODE>(p, xb, zb, wb);
        s.w[2 * m] = wa; s.w[2 * m + 1] = wb;
        if (last) {
            xa_o[n] = xa; za_o[n] = za; wa_o[n] = wa;
            if (has_b) { xb_o[n] = xb; zb_o[n] = zb; wb_o[n] = wb; }
        } else {
            zs[n] = mk<float>(za, zb);
            s.a[m] = mk<float>(za - wa, zb - wb);
        }
    }
}

}  // namespace k3
}  // namespace pnp
