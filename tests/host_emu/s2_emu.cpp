// s2_emu.cpp — CPU check of the K2 streaming-kernel FFT stages (stream2_core.cuh is host+device).
// Test infrastructure only.  Runs the per-thread stage code for all T threads of a line, with the
// kernels' barriers replaced by loop boundaries, against a naive double-precision DFT.
#include <cmath>
#include <complex>
#include <cstdlib>
#include <vector>

#include "../../pnp_admm_cnc_mri_b200/csrc/stream2_core.cuh"

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
