// Probe: one 256-point line FFT pass both ways, lines resident on chip, no global traffic in the loop:
//   A  the production core (stream2_core.cuh): 16 points per thread, 16 threads per line, radix 16 x 16, ONE exchange per transform,
//      (56 registers for the transform alone; 128 in K1 / K3, where the dual and the prox state live in registers too -> 16 warps per SM)
//   B  8 points per thread, 32 threads (one warp) per line, radix 8 x 8 x 4, TWO exchanges per transform, <= 64 registers -> 32 warps per SM
// Each iteration: forward FFT -> scale -> inverse FFT (what a K3 iteration does around its pointwise work).  Prints line transforms per
// second, cycles per line transform and SM, and the round-trip error of both variants.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pnp_admm_cnc_mri_b200/csrc -o tools/probes/fft256_line tools/probes/fft256_line.cu
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "stream2.cuh"

using namespace pnp;
using k1::cf32;

__device__ __forceinline__ cf32 cmul_f(cf32 a, cf32 w) { return a * w; }

template <bool INV> __device__ __forceinline__ void fft8(const cf32 (&x)[8], cf32 (&X)[8]) {
    cf32 e[4], o[4];
    k1::dft4<INV>(x[0], x[2], x[4], x[6], e[0], e[1], e[2], e[3]);
    k1::dft4<INV>(x[1], x[3], x[5], x[7], o[0], o[1], o[2], o[3]);
    const float c = 0.70710678118654752f;
    o[1] = twmul<INV>(o[1], mk<float>(c, -c));
    o[2] = rot90<INV>(o[2]);
    o[3] = twmul<INV>(o[3], mk<float>(-c, -c));
#pragma unroll
    for (int k = 0; k < 4; ++k) { X[k] = e[k] + o[k]; X[k + 4] = e[k] - o[k]; }
}

__device__ __forceinline__ int pad8(int i) { return i + (i >> 3); }

// B: thread t of the warp holds a[m] = x[t + 32 m]
template <bool INV> __device__ __forceinline__ void fft256_b(cf32 (&a)[8], int t, cf32* ln, const cf32* tw64, const cf32* tw256) {
    cf32 b[8];
    fft8<INV>(a, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) ln[pad8(8 * t + i)] = b[i];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = ln[pad8(t + 32 * i)];
    const int k = t & 7;
#pragma unroll
    for (int i = 1; i < 8; ++i) a[i] = twmul<INV>(a[i], tw64[i * 8 + k]);
    fft8<INV>(a, b);
    __syncwarp();
    const int base = (t >> 3) * 64 + k;
#pragma unroll
    for (int i = 0; i < 8; ++i) ln[pad8(base + 8 * i)] = b[i];
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 8; ++m) a[m] = ln[pad8(t + 32 * m)];
    __syncwarp();
#pragma unroll
    for (int u = 0; u < 2; ++u) {                       // radix 4 over registers m = u + 2 i, j = t + 32 u < 64
        const int j = t + 32 * u;
        const cf32 v1 = twmul<INV>(a[u + 2], tw256[j]);
        const cf32 v2 = twmul<INV>(a[u + 4], tw256[64 + j]);
        const cf32 v3 = twmul<INV>(a[u + 6], tw256[128 + j]);
        k1::dft4<INV>(a[u], v1, v2, v3, a[u], a[u + 2], a[u + 4], a[u + 6]);
    }
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB) kernel_b(const float2* in, float2* out, int iters, long long* cyc) {
    __shared__ cf32 lines[8][288];
    __shared__ cf32 tw64[64], tw256[192];
    const int tid = threadIdx.x, w = tid >> 5, t = tid & 31;
    if (tid < 64) { float s, c; sincospif(-2.f * (float)((tid >> 3) * (tid & 7)) / 64.f, &s, &c); tw64[tid] = mk<float>(c, s); }
    if (tid < 192) { float s, c; sincospif(-2.f * (float)((tid / 64 + 1) * (tid % 64)) / 256.f, &s, &c); tw256[tid] = mk<float>(c, s); }
    __syncthreads();
    cf32 a[8];
    const size_t line = (size_t)blockIdx.x * 8 + w;
#pragma unroll
    for (int m = 0; m < 8; ++m) { const float2 v = in[line * 256 + t + 32 * m]; a[m] = mk<float>(v.x, v.y); }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        fft256_b<false>(a, t, lines[w], tw64, tw256);
#pragma unroll
        for (int m = 0; m < 8; ++m) { a[m].re *= (1.f / 256.f); a[m].im *= (1.f / 256.f); }
        fft256_b<true>(a, t, lines[w], tw64, tw256);
    }
    const long long t1 = clock64();
#pragma unroll
    for (int m = 0; m < 8; ++m) out[line * 256 + t + 32 * m] = make_float2(a[m].re, a[m].im);
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(256, 2) kernel_a(const float2* in, float2* out, int iters, long long* cyc) {
    constexpr int N = 256, T = 16;
    __shared__ cf32 lines[16][s2::Plan<N>::kRowPitch];
    __shared__ cf32 TW[256];
    const int tid = threadIdx.x, l = tid / T, t = tid % T;
    TW[tid] = mk<float>(s2::g_tw256[tid].x, s2::g_tw256[tid].y);
    __syncthreads();
    s2::RowLine ln; ln.line = lines[l];
    const s2::Tw3Master tw3{nullptr, 0};
    cf32 a[16];
    const size_t line = (size_t)blockIdx.x * 16 + l;
#pragma unroll
    for (int m = 0; m < 16; ++m) { const float2 v = in[line * 256 + t + T * m]; a[m] = mk<float>(v.x, v.y); }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        s2::line_sync<T, false>();
        s2::fft_regs<false, N, false>(a, t, ln, TW, tw3);
#pragma unroll
        for (int m = 0; m < 16; ++m) { a[m].re *= (1.f / 256.f); a[m].im *= (1.f / 256.f); }
        s2::line_sync<T, false>();
        s2::fft_regs<true, N, false>(a, t, ln, TW, tw3);
    }
    const long long t1 = clock64();
#pragma unroll
    for (int m = 0; m < 16; ++m) out[line * 256 + t + T * m] = make_float2(a[m].re, a[m].im);
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <class F> static void run(const char* name, F launch, int ctas_per_sm, int lines_per_cta, int sms, int iters, const float2* in, float2* out,
                                   long long* cyc, const std::vector<float2>& h_in) {
    const int grid = sms * ctas_per_sm;
    launch(grid, 2);                                                   // warm-up
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); launch(grid, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    if (cudaGetLastError() != cudaSuccess) { printf("%s: launch failed\n", name); return; }
    launch(grid, 1);
    std::vector<float2> h((size_t)grid * lines_per_cta * 256);
    cudaMemcpy(h.data(), out, h.size() * sizeof(float2), cudaMemcpyDeviceToHost);
    double err = 0, ref = 0;
    for (size_t i = 0; i < h.size(); ++i) { err += pow(h[i].x - h_in[i].x, 2) + pow(h[i].y - h_in[i].y, 2); ref += pow(h_in[i].x, 2) + pow(h_in[i].y, 2); }
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ffts = 2.0 * (double)grid * lines_per_cta * iters, rate = ffts / (best * 1e-3);
    printf("%-34s %d CTAs/SM: %7.3f ms  %7.2f G line-FFT/s  %6.1f SM-cycles per line FFT at %d MHz  = %5.1f TFLOP/s nominal (5 N log2 N)  round-trip rel-L2 %.2e\n",
           name, ctas_per_sm, best, rate / 1e9, (double)khz * 1e3 * sms / rate, khz / 1000, rate * 10240.0 / 1e12, sqrt(err / ref));
}

int main() {
    int dev = 0, sms = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int iters = 400, max_lines = sms * 4 * 16;
    std::vector<float2> h_in((size_t)max_lines * 256);
    unsigned s = 12345u;
    for (auto& v : h_in) { s = s * 1664525u + 1013904223u; v.x = (s >> 8) * (1.f / 16777216.f) - 0.5f; s = s * 1664525u + 1013904223u; v.y = (s >> 8) * (1.f / 16777216.f) - 0.5f; }
    float2 *in, *out; long long* cyc;
    cudaMalloc(&in, h_in.size() * sizeof(float2)); cudaMalloc(&out, h_in.size() * sizeof(float2)); cudaMalloc(&cyc, sizeof(long long) * sms * 8);
    cudaMemcpy(in, h_in.data(), h_in.size() * sizeof(float2), cudaMemcpyHostToDevice);
    // the production twiddle table g_tw256 is filled by the library at start-up; fill it here
    std::vector<float2> tw(256);
    for (int i = 0; i < 16; ++i) for (int k = 0; k < 16; ++k) { const double a = -2.0 * M_PI * i * k / 256.0; tw[i * 16 + k] = make_float2((float)cos(a), (float)sin(a)); }
    cudaMemcpyToSymbol(s2::g_tw256, tw.data(), sizeof(float2) * 256);
    run("A 16 pts/thread (production core)", [&](int g, int it) { kernel_a<<<g, 256>>>(in, out, it, cyc); }, 2, 16, sms, iters, in, out, cyc, h_in);
    run("A 16 pts/thread, 1 CTA/SM", [&](int g, int it) { kernel_a<<<g, 256>>>(in, out, it, cyc); }, 1, 16, sms, iters, in, out, cyc, h_in);
    run("A 16 pts/thread, 3 CTAs/SM", [&](int g, int it) { kernel_a<<<g, 256>>>(in, out, it, cyc); }, 3, 16, sms, iters, in, out, cyc, h_in);
    run("A 16 pts/thread, 4 CTAs/SM", [&](int g, int it) { kernel_a<<<g, 256>>>(in, out, it, cyc); }, 4, 16, sms, iters, in, out, cyc, h_in);
    run("B 8 pts/thread (radix 8x8x4)", [&](int g, int it) { kernel_b<4><<<g, 256>>>(in, out, it, cyc); }, 4, 8, sms, iters, in, out, cyc, h_in);
    run("B 8 pts/thread, 3 CTAs/SM", [&](int g, int it) { kernel_b<3><<<g, 256>>>(in, out, it, cyc); }, 3, 8, sms, iters, in, out, cyc, h_in);
    run("B 8 pts/thread, 2 CTAs/SM", [&](int g, int it) { kernel_b<2><<<g, 256>>>(in, out, it, cyc); }, 2, 8, sms, iters, in, out, cyc, h_in);
    return 0;
}
