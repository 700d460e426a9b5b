// s2_emu.cpp — CPU check of the K2 streaming-kernel FFT stages (stream2_core.cuh is host+device).
// Test infrastructure only.  Runs the per-thread stage code for all T threads of a line, with the
// kernels' barriers replaced by loop boundaries, against a naive double-precision DFT.
#include <cmath>
#include <complex>
#include <cstdlib>
#include <vector>

#include "../../pnp_admm_cnc_mri_b200/csrc/stream2_core.cuh"
#include "../../pnp_admm_cnc_mri_b200/csrc/rowsep_core.cuh"

using namespace pnp;
using namespace pnp::s2;

namespace {

template <bool INV, int N, class Line, class MakeLine>
double run(MakeLine make, unsigned seed) {
    typedef Plan<N> PL;
    constexpr int T = PL::T;
    std::vector<cf32> tw256(256), tw4096(4096);
    const double PI = 3.14159265358979323846;
    for (int i = 0; i < 4096; ++i) tw4096[i] = mk<float>((float)cos(-2 * PI * i / 4096), (float)sin(-2 * PI * i / 4096));
    for (int i = 0; i < 256; ++i) tw256[i] = tw4096[((i >> 4) * (i & 15) * 16) & 4095];
    std::vector<std::complex<double>> x(N), X(N);
    srand(seed);
    for (int n = 0; n < N; ++n) x[n] = std::complex<double>(rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5);
    for (int k = 0; k < N; ++k) {
        std::complex<double> s = 0;
        for (int n = 0; n < N; ++n) s += x[n] * std::polar(1.0, (INV ? 2 : -2) * PI * (double)((long)k * n % N) / N);
        X[k] = s;
    }
    std::vector<cf32> regs(T * 16);
    for (int t = 0; t < T; ++t)
        for (int m = 0; m < 16; ++m) regs[t * 16 + m] = mk<float>((float)x[t + T * m].real(), (float)x[t + T * m].imag());
    auto R = [&](int t) -> cf32(&)[16] { return *reinterpret_cast<cf32(*)[16]>(&regs[t * 16]); };
    for (int t = 0; t < T; ++t) stage1_store<INV, N>(R(t), t, make());
    for (int t = 0; t < T; ++t) stage2_load<INV, N>(R(t), t, make(), tw256.data());
    if (PL::R3 > 1) {
        for (int t = 0; t < T; ++t) stage2_store<N>(R(t), t, make());
        for (int t = 0; t < T; ++t) stage3<INV, N>(R(t), t, make(), Tw3Master{tw4096.data(), 4096 / N});
    }
    double num = 0, den = 0;
    for (int t = 0; t < T; ++t)
        for (int m = 0; m < 16; ++m) {
            const std::complex<double> got(regs[t * 16 + m].re, regs[t * 16 + m].im);
            num += std::norm(got - X[t + T * m]);
            den += std::norm(X[t + T * m]);
        }
    return std::sqrt(num / den);
}

template <bool INV, int N>
double run_both(int layout, unsigned seed) {
    if (layout == 0) {
        std::vector<cf32> buf(Plan<N>::kRowPitch + 8);
        return run<INV, N, RowLine>([&]() { RowLine l; l.line = buf.data(); return l; }, seed);
    }
    constexpr int C = 8;
    std::vector<cf32> buf((size_t)N * C);
    return run<INV, N, ColLine<C>>([&]() { ColLine<C> l; l.col = buf.data() + 3; return l; }, seed);
}

}  // namespace

extern "C" double s2_fft_check(int N, int inv, int layout, unsigned seed) {
    switch (N) {
        case 256: return inv ? run_both<true, 256>(layout, seed) : run_both<false, 256>(layout, seed);
        case 512: return inv ? run_both<true, 512>(layout, seed) : run_both<false, 512>(layout, seed);
        case 1024: return inv ? run_both<true, 1024>(layout, seed) : run_both<false, 1024>(layout, seed);
    }
    return -1.0;
}

extern "C" unsigned s2_pack_codes(const uint8_t* mcode, int N, int t, int kc) { return pack_codes_n(mcode, N, t, kc); }


// ---------------------------------------------------------------------------------------------------------------
// K3 at N = 512 / 1024 (rowsepN.cuh): the row-separable solve, one line at a time, same per-thread code as the kernel.
// ---------------------------------------------------------------------------------------------------------------
namespace {
template <int N>
int k3n_run(const uint8_t* img8, const uint8_t* mask, const float* planes_, float reo, float* x, float* z, float* w, int B, int iters,
            const ProxParams<float>& pp) {
    using namespace pnp::k3;
    constexpr int T = N / 16;
    const size_t nn = (size_t)N * N;
    for (int kc = 0; kc < N; ++kc)
        if (!column_is_constant(mask, N, kc)) return 1;
    std::vector<cf32> tw256(256), tw4096(4096);
    const double PI = 3.14159265358979323846;
    for (int i = 0; i < 4096; ++i) tw4096[i] = mk<float>((float)cos(-2 * PI * i / 4096), (float)sin(-2 * PI * i / 4096));
    for (int i = 0; i < 256; ++i) tw256[i] = tw4096[((i >> 4) * (i & 15) * 16) & 4095];
    const Tw3Master tw3{tw4096.data(), 4096 / N};
    const cf32* planes = reinterpret_cast<const cf32*>(planes_);
    uint32_t codes[T], here[T];
    for (int t = 0; t < T; ++t) line_words<N>(mask, t, codes + t, here + t);
    const double g = 1.0 / (1.0 + 1.0 / 2.0 / reo);
    const float ncf1 = (float)(0.5 * g / N), ncf2 = (float)(g / N), inv_n2 = 1.0f / ((float)N * (float)N);
    const int mode = prox_mode(pp);
    std::vector<cf32> buf(Plan<N>::kRowPitch + 8), zs(N), gp(N);
    std::vector<LineState> st(T);
    RowLine ln; ln.line = buf.data();
    auto fft = [&](bool inv) {
        for (int t = 0; t < T; ++t) { if (inv) stage1_store<true, N>(st[t].a, t, ln); else stage1_store<false, N>(st[t].a, t, ln); }
        for (int t = 0; t < T; ++t) { if (inv) stage2_load<true, N>(st[t].a, t, ln, tw256.data()); else stage2_load<false, N>(st[t].a, t, ln, tw256.data()); }
        if (Plan<N>::R3 > 1) {
            for (int t = 0; t < T; ++t) stage2_store<N>(st[t].a, t, ln);
            for (int t = 0; t < T; ++t) { if (inv) stage3<true, N>(st[t].a, t, ln, tw3); else stage3<false, N>(st[t].a, t, ln, tw3); }
        }
    };
    const int P = (B + 1) / 2;
    for (int plane = 0; plane < P; ++plane)
        for (int row = 0; row < N; ++row) {
            const int ia = 2 * plane;
            const bool has_b = ia + 1 < B;
            const float hb = has_b ? 1.f : 0.f;
            const size_t ga = (size_t)ia * nn + (size_t)row * N, gb = ga + nn, gr = (size_t)row * N;
            for (int t = 0; t < T; ++t) load_image<N>(st[t], t, nullptr, nullptr, img8 + ga, has_b ? img8 + gb : nullptr);
            fft(false);
            for (int t = 0; t < T; ++t) acquire_ms<N>(st[t], t, zs.data(), gp.data(), planes + gr, planes + nn + gr, codes[t], ncf1, ncf2, hb);
            fft(true);
            for (int t = 0; t < T; ++t) { stash_t1(st[t]); acquire_ma<N>(st[t], t, zs.data(), planes + 2 * nn + gr, codes[t], here[t], hb); }
            fft(true);
            for (int t = 0; t < T; ++t) zero_fill<N>(st[t], t, zs.data(), inv_n2, has_b);
            for (int it = 0; it < iters; ++it) {
                fft(false);
                for (int t = 0; t < T; ++t) blend<N>(st[t], t, gp.data(), codes[t], ncf1, ncf2);
                fft(true);
                const bool last = it == iters - 1;
                for (int t = 0; t < T; ++t) {
                    if (mode == PM_CNC) prox_row<PM_CNC, N>(st[t], t, zs.data(), pp, has_b, last, x + ga, z + ga, w + ga, x + gb, z + gb, w + gb);
                    else prox_row<PM_L1, N>(st[t], t, zs.data(), pp, has_b, last, x + ga, z + ga, w + ga, x + gb, z + gb, w + gb);
                }
            }
        }
    return 0;
}
}  // namespace

extern "C" void k3n_noise_terms(const uint8_t* mask, const float* noise, int N, float g_over_n2, float* planes) {
    for (size_t bin = 0; bin < (size_t)N * N; ++bin)
        pnp::k3::noise_terms(mask, reinterpret_cast<const cf32*>(noise), N, g_over_n2, bin, reinterpret_cast<cf32*>(planes));
}
extern "C" int k3n_emulate(int N, const uint8_t* img8, const uint8_t* mask, const float* planes, float reo, float* x, float* z, float* w,
                           int B, int iters, int prox, float thr_l1, float inv_b, float one_m_alpha, float alpha, float coef, float thr_cnc) {
    ProxParams<float> pp;
    pp.prox = prox; pp.general = 0; pp.thr_l1 = thr_l1; pp.inv_b = inv_b; pp.one_m_alpha = one_m_alpha;
    pp.alpha = alpha; pp.coef = coef; pp.thr_cnc = thr_cnc;
    if (N == 256) return k3n_run<256>(img8, mask, planes, reo, x, z, w, B, iters, pp);
    if (N == 512) return k3n_run<512>(img8, mask, planes, reo, x, z, w, B, iters, pp);
    if (N == 1024) return k3n_run<1024>(img8, mask, planes, reo, x, z, w, B, iters, pp);
    return -1;
}
