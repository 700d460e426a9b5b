// rowsepN.cuh — K3 at N = 512 / 1024: whole ADMM solve under a row-separable mask, one launch, rows resident on chip.
// (N = 256 has its own kernel on K1's row-phase code, rowsep256.cuh; this one also instantiates 256 for A/B runs.)
//
// A CTA of 256 threads owns L = 256 / T consecutive rows of one packed plane (T = N / 16 threads per row: 8 rows at N = 512,
// 4 at N = 1024) for the whole solve.  Shared memory per row: the padded FFT line, z (both images) and the row of
// G' = colIFFT(G) — 12.3 KB at N = 512, 24.6 KB at N = 1024, ~100 KB per CTA, two CTAs per SM; registers: 16 points + the dual.
// The general path streams the state through HBM twice per iteration (K2: 73 N^2 bytes per plane and iteration); here an
// iteration touches no global memory at all: HBM traffic of a solve = image in, x / z / w out, three noise-term rows.
#pragma once

#include "rowsep_core.cuh"
#include "stream2.cuh"

namespace pnp {
namespace k3 {

struct RowSepNParams {
    int B, P, iters;
    const float* img; const uint8_t* img8;
    float* x; float* z; float* w;
    const cf32* planes;        // [3][N][N] row-major: colIFFT_u of NcS, nH, nA
    const uint32_t* rcodes;    // [T]
    const uint32_t* rhere;     // [T]
    const int* sep;            // blocks that saw a mask bin differ from k-space row 0 (0 = separable)
    float ncf1, ncf2;          // N g / (2 N^2), N g / N^2
    ProxParams<float> prox;
};

template <int N> struct RsGeo {
    static constexpr int T = N / 16;
    static constexpr int L = 256 / T;
    static constexpr int kPitch = s2::Plan<N>::kRowPitch;
    static constexpr int kOffZ = L * kPitch * 8;
    static constexpr int kOffG = kOffZ + L * N * 8;
    static constexpr int kOffTW = kOffG + L * N * 8;
    static constexpr int kSmemBytes = kOffTW + 256 * 8;
};

__global__ void prepare_rowsepN_kernel(const uint8_t* __restrict__ mask, const cf32* __restrict__ noise, int N, float g_over_n2,
                                       cf32* __restrict__ planes, uint32_t* __restrict__ rcodes, uint32_t* __restrict__ rhere,
                                       int* __restrict__ bad) {
    const size_t bin = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int differs = 0;
    if (bin < (size_t)N * N) {
        noise_terms(mask, noise, N, g_over_n2, bin, planes);
        differs = ((mask[bin] != 0) != (mask[bin % N] != 0)) ? 1 : 0;      // same column of k-space row 0
    }
    if (__syncthreads_or(differs) && threadIdx.x == 0) atomicAdd(bad, 1);  // `bad` zeroed by the host before the launch
    if (blockIdx.x == 0 && (int)threadIdx.x < N / 16) {
        if (N == 512) line_words<512>(mask, threadIdx.x, rcodes + threadIdx.x, rhere + threadIdx.x);
        else if (N == 1024) line_words<1024>(mask, threadIdx.x, rcodes + threadIdx.x, rhere + threadIdx.x);
        else line_words<256>(mask, threadIdx.x, rcodes + threadIdx.x, rhere + threadIdx.x);
    }
}

template <int N>
__global__ void __launch_bounds__(256, 2) rowsepN_kernel(const RowSepNParams p) {
    typedef RsGeo<N> G;
    constexpr int T = G::T, L = G::L;
    extern __shared__ __align__(128) unsigned char smem3[];
    const size_t nn = (size_t)N * N;
    if (*p.sep != 0) {   // the mask is not made of full k-space lines: fail loudly
        const float qnan = __int_as_float(0x7fc00000);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)p.B * nn; i += (size_t)gridDim.x * blockDim.x) {
            p.x[i] = qnan; p.z[i] = qnan; p.w[i] = qnan;
        }
        return;
    }
    cf32* TW = reinterpret_cast<cf32*>(smem3 + G::kOffTW);
    TW[threadIdx.x] = mk<float>(s2::g_tw256[threadIdx.x].x, s2::g_tw256[threadIdx.x].y);
    __syncthreads();
    const int tid = threadIdx.x, line = tid / T, t = tid % T;
    s2::RowLine ln;
    ln.line = reinterpret_cast<cf32*>(smem3) + line * G::kPitch;
    cf32* zs = reinterpret_cast<cf32*>(smem3 + G::kOffZ) + line * N;
    cf32* gp = reinterpret_cast<cf32*>(smem3 + G::kOffG) + line * N;
    const s2::Tw3Master tw3{reinterpret_cast<const cf32*>(g_tw_f32), kTwMax / N};
    const uint32_t codes = p.rcodes[t], here = p.rhere[t];
    const int mode = prox_mode(p.prox);
    const float inv_n2 = 1.0f / ((float)N * (float)N);
    LineState s;
    constexpr int tiles_per_plane = N / L;
    for (int task = blockIdx.x; task < p.P * tiles_per_plane; task += gridDim.x) {
        const int plane = task / tiles_per_plane, row = (task - plane * tiles_per_plane) * L + line;
        const int ia = 2 * plane;
        const bool has_b = ia + 1 < p.B;
        const float hb = has_b ? 1.f : 0.f;
        const size_t ga = (size_t)ia * nn + (size_t)row * N, gb = ga + nn, gr = (size_t)row * N;
        s2::line_sync<T, false>();   // the previous task's last shared-memory reads of this line are done
        load_image<N>(s, t, p.img ? p.img + ga : nullptr, (p.img && has_b) ? p.img + gb : nullptr,
                      p.img8 ? p.img8 + ga : nullptr, (p.img8 && has_b) ? p.img8 + gb : nullptr);
        s2::fft_regs<false, N, false>(s.a, t, ln, TW, tw3);
        acquire_ms<N>(s, t, zs, gp, p.planes + gr, p.planes + nn + gr, codes, p.ncf1, p.ncf2, hb);
        s2::line_sync<T, false>();
        s2::fft_regs<true, N, false>(s.a, t, ln, TW, tw3);
        stash_t1(s);
        acquire_ma<N>(s, t, zs, p.planes + 2 * nn + gr, codes, here, hb);
        s2::line_sync<T, false>();
        s2::fft_regs<true, N, false>(s.a, t, ln, TW, tw3);
        zero_fill<N>(s, t, zs, inv_n2, has_b);
        for (int it = 0; it < p.iters; ++it) {
            s2::line_sync<T, false>();
            s2::fft_regs<false, N, false>(s.a, t, ln, TW, tw3);
            blend<N>(s, t, gp, codes, p.ncf1, p.ncf2);
            s2::line_sync<T, false>();
            s2::fft_regs<true, N, false>(s.a, t, ln, TW, tw3);
            const bool last = it == p.iters - 1;
            float* xa = p.x + ga; float* za = p.z + ga; float* wa = p.w + ga;
            float* xb = p.x + gb; float* zb = p.z + gb; float* wb = p.w + gb;
            if (mode == PM_CNC) prox_row<PM_CNC, N>(s, t, zs, p.prox, has_b, last, xa, za, wa, xb, zb, wb);
            else if (mode == PM_L1) prox_row<PM_L1, N>(s, t, zs, p.prox, has_b, last, xa, za, wa, xb, zb, wb);
            else prox_row<PM_GENERAL, N>(s, t, zs, p.prox, has_b, last, xa, za, wa, xb, zb, wb);
        }
    }
}

}  // namespace k3
}  // namespace pnp
