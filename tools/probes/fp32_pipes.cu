// Probe: per-SM throughput of FADD, 3-register FFMA, FADD2, FFMA2 and mixes (warp-instructions / clk / SM).
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, const float* in, long long* cyc) {
    float b = in[0], c = in[1];
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = in[2 + i] + threadIdx.x;
    unsigned long long pb, pc;
    { float2 t = make_float2(b, c); pb = *reinterpret_cast<unsigned long long*>(&t); t = make_float2(c, b); pc = *reinterpret_cast<unsigned long long*>(&t); }
    unsigned long long pa[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 t = make_float2(a[2 * i], a[2 * i + 1]); pa[i] = *reinterpret_cast<unsigned long long*>(&t); }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
            if (MODE == 2 && i < 8) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(pa[i]) : "l"(pb));
            if (MODE == 3 && i < 8) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa[i]) : "l"(pb), "l"(pc));
            if (MODE == 4) { if (i & 1) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b)); else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c)); }
            if (MODE == 5) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (MODE == 6) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 15])); }   // 2 distinct regs, varying
            if (MODE == 7) { asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(a[(i + 5) & 15]), "f"(a[(i + 9) & 15])); }  // 3 distinct varying regs
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) { float2 t = *reinterpret_cast<float2*>(&pa[i]); s += t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int per_iter, int flops_per_instr_lane, float* out, float* in, long long* cyc) {
    k<MODE><<<148, 512>>>(out, in, cyc);
    cudaDeviceSynchronize();
    k<MODE><<<148, 512>>>(out, in, cyc);
    cudaDeviceSynchronize();
    double c = 0; for (int i = 0; i < 148; ++i) c += cyc[i]; c /= 148;
    double winstr = 16.0 * per_iter * ITERS;   // warp-instructions per SM (16 warps)
    printf("%-34s %7.3f warp-instr/clk/SM  -> %6.1f flop/clk/SM  (%s)\n", name, winstr / c, winstr / c * 32 * flops_per_instr_lane, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    float *out, *in; long long* cyc;
    cudaMallocManaged(&out, 148 * 512 * 4); cudaMallocManaged(&in, 64 * 4); cudaMallocManaged(&cyc, 148 * 8);
    for (int i = 0; i < 64; ++i) in[i] = 1.0f + 1e-3f * i;
    run<0>("FADD  (reg + fixed reg)", 16, 1, out, in, cyc);
    run<6>("FADD  (2 varying regs)", 16, 1, out, in, cyc);
    run<5>("FMUL  (reg * fixed reg)", 16, 1, out, in, cyc);
    run<1>("FFMA  (reg, 2 fixed regs)", 16, 2, out, in, cyc);
    run<7>("FFMA  (3 varying regs)", 16, 2, out, in, cyc);
    run<4>("FADD/FFMA alternating", 16, 1, out, in, cyc);
    run<2>("FADD2 (packed)", 8, 2, out, in, cyc);
    run<3>("FFMA2 (packed)", 8, 4, out, in, cyc);
    return 0;
}
