// stream2_core.cuh — K2: per-thread FFT stages of the streaming kernels (N = 256, 512, 1024; fp32).
//
// A line of N complex points is transformed by T = N / 16 threads holding 16 points each, as a
// Stockham autosort FFT with radices (16, 16[, N / 256]) done in place in a shared-memory line:
//
//   stage 1 (Ns = 1,   R = 16): v[i] = x[t + T i]                      -> fft16 -> y[16 t + i]
//   stage 2 (Ns = 16,  R = 16): v[i] = y[t + T i] W_256^((t & 15) i)   -> fft16 -> u[16 (t & ~15) + (t & 15) + 16 i]
//   stage 3 (Ns = 256, R = N / 256 in {2, 4}), 16 / R butterflies per thread, j = t + T u:
//                               v[i] = u[j + 256 i] W_N^(j i)          -> dftR  -> X[j + 256 i]
//
// Thread t both starts and ends with the points n = t + T m (m < 16) of its line in registers, so a
// pass that chains transforms (IFFT -> pointwise -> FFT) only touches shared memory for the exchanges
// between stages, and consecutive threads own consecutive points (coalesced global accesses).
// For N = 256 stage 2 already ends in that layout; stage 3 needs no exchange of its own because its
// butterfly u combines the registers m = u + (16 / R) i.
//
// Layout policies give the address of logical point idx of "my line":
//   RowLine : contiguous line, padded by one element per 16 (stage-1 stores have stride 16);
//             raw() is the unpadded landing layout of the bulk copy
//   ColLine : column c of a [N][C] tile (lanes run along c: every access is conflict-free)
//
// All functions are HOST+DEVICE: tests/host_emu/s2_emu.cpp runs them against a naive DFT.
#pragma once

#include "cluster256_core.cuh"   // cf32, fft16, dft4

namespace pnp {
namespace s2 {

using k1::cf32;

// read-only twiddle load (non-coherent path on the device)
#if defined(__CUDA_ARCH__)
PNP_D cf32 ld_tw(const cf32* p) { const float2 v = __ldg(reinterpret_cast<const float2*>(p)); return mk<float>(v.x, v.y); }
#else
inline cf32 ld_tw(const cf32* p) { return *p; }
#endif

template <int N>
struct Plan {
    static_assert(N == 256 || N == 512 || N == 1024, "K2 register-FFT kernels: N in {256, 512, 1024}");
    static constexpr int T = N / 16;           // threads per line
    static constexpr int R3 = N / 256;         // radix of stage 3 (1: no stage 3)
    static constexpr int NB3 = 16 / R3;        // stage-3 butterflies per thread
    static constexpr int kRowPitch = N + N / 16;
};

struct RowLine {
    cf32* line;
    PNP_HD cf32& at(int idx) const { return line[idx + (idx >> 4)]; }
    PNP_HD cf32& raw(int idx) const { return line[idx]; }
};

// With C = 8 a row is 64 B = half of the banks, and the stage-1 stores of a warp (rows 16 t + i for four
// consecutive t) would all land on the same half: rows are stored at row ^ ((row >> 4) & 1), which alternates
// the half with t and keeps every group of four consecutive rows contiguous (all other accesses).
template <int C> PNP_HD int col_phys_row(int row) { return C == 8 ? (row ^ ((row >> 4) & 1)) : (C == 4 ? (row ^ ((row >> 4) & 3)) : row); }

template <int C>
struct ColLine {
    cf32* col;   // tile + c
    PNP_HD cf32& at(int idx) const { return col[col_phys_row<C>(idx) * C]; }
    PNP_HD cf32& raw(int idx) const { return col[col_phys_row<C>(idx) * C]; }
};

// stage 1: registers (n = t + T i) -> fft16 -> shared
template <bool INV, int N, class Line>
PNP_HD void stage1_store(const cf32 (&a)[16], int t, const Line& ln) {
    cf32 b[16];
    k1::fft16<INV>(a, b);
#pragma unroll
    for (int i = 0; i < 16; ++i) ln.at(16 * t + i) = b[i];
}

// stage 2: shared -> twiddle -> fft16 -> registers.  TW256[i * 16 + k] = W_256^(i k).
template <bool INV, int N, class Line>
PNP_HD void stage2_load(cf32 (&a)[16], int t, const Line& ln, const cf32* TW256) {
    constexpr int T = Plan<N>::T;
    const int k = t & 15;
    cf32 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = ln.at(t + T * i);
#pragma unroll
    for (int i = 1; i < 16; ++i) v[i] = twmul<INV>(v[i], TW256[i * 16 + k]);
    k1::fft16<INV>(v, a);
}

// N > 256 only: stage-2 results back to shared
template <int N, class Line>
PNP_HD void stage2_store(const cf32 (&a)[16], int t, const Line& ln) {
    const int base = 16 * (t & ~15) + (t & 15);
#pragma unroll
    for (int i = 0; i < 16; ++i) ln.at(base + 16 * i) = a[i];
}

// stage-3 twiddle sources: W_N^(i j), j < 256, i < R3
struct Tw3Master {          // the master table W_4096^m in global memory (stride 4096 / N), read-only path
    const cf32* m; int stride;
    PNP_HD cf32 get(int i, int j) const { return ld_tw(m + i * j * stride); }
};
struct Tw3Table {           // a [R3 - 1][256] table in shared memory: tab[(i - 1) * 256 + j]
    const cf32* tab;
    PNP_HD cf32 get(int i, int j) const { return tab[(i - 1) * 256 + j]; }
};

// N > 256 only: stage 3.
template <bool INV, int N, class Line, class Tw3>
PNP_HD void stage3(cf32 (&a)[16], int t, const Line& ln, const Tw3& tw) {
    constexpr int T = Plan<N>::T, R3 = Plan<N>::R3, NB3 = Plan<N>::NB3;
#pragma unroll
    for (int m = 0; m < 16; ++m) a[m] = ln.at(t + T * m);
    if (R3 == 2) {
#pragma unroll
        for (int u = 0; u < NB3; ++u) {
            const int j = t + T * u;
            const cf32 v1 = twmul<INV>(a[u + NB3], tw.get(1, j));
            const cf32 v0 = a[u];
            a[u] = v0 + v1;
            a[u + NB3] = v0 - v1;
        }
    } else if (R3 == 4) {
#pragma unroll
        for (int u = 0; u < NB3; ++u) {
            const int j = t + T * u;
            const cf32 v1 = twmul<INV>(a[u + NB3], tw.get(1, j));
            const cf32 v2 = twmul<INV>(a[u + 2 * NB3], tw.get(2, j));
            const cf32 v3 = twmul<INV>(a[u + 3 * NB3], tw.get(3, j));
            k1::dft4<INV>(a[u], v1, v2, v3, a[u], a[u + NB3], a[u + 2 * NB3], a[u + 3 * NB3]);
        }
    }
}

// packed mask codes for the column pass: word (t, kc) holds mcode[(t + T m) * N + kc] in bits 2m, 2m+1
PNP_HD uint32_t pack_codes_n(const uint8_t* mcode, int N, int t, int kc) {
    const int T = N / 16;
    uint32_t v = 0;
    for (int m = 0; m < 16; ++m) v |= (uint32_t)(mcode[(size_t)(t + T * m) * N + kc] & 3u) << (2 * m);
    return v;
}

}  // namespace s2
}  // namespace pnp
