// Probe: DSMEM all-to-all ("transpose") throughput inside a 16-CTA cluster (256 threads, 2 CTAs per SM, the K1 geometry).
// Every CTA sends 2 KB to each of its 16 peers per round (32 KB per CTA and round, like one K1 transpose) and waits for its
// own 32 KB (mbarrier complete_tx).  Modes:
//   0  st.async 8 B per thread and store (what K1 does, SASS STAS.64): 16 stores per thread and round
//   1  st.async 16 B per thread and store (v4): 8 stores per thread and round
//   2  cp.async.bulk shared::cta -> shared::cluster, one 2 KB copy per peer (TMA engine), issued by lane 0 of each warp (2 each)
//   3  as 2, but one thread issues all 16 copies
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_bw dsmem_bw.cu && ./dsmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int CL = 16, THREADS = 256, TILE = 32768;
constexpr int OFF_SRC = 0, OFF_DST = TILE, OFF_BAR = 3 * TILE, SMEM = 3 * TILE + 64;   // dst double-buffered

__device__ __forceinline__ uint32_t mapa(uint32_t a, int r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t par) {
    long long t0 = clock64();
    while (!try_wait(bar, par)) if (clock64() - t0 > 2000000000ll) __trap();
}

__global__ void __launch_bounds__(THREADS, 2) k(int mode, int rounds, unsigned long long* cycles) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(smem);
    uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar[2] = {s0 + OFF_BAR, s0 + OFF_BAR + 8};
    for (int i = tid; i < TILE / 4; i += THREADS) reinterpret_cast<float*>(smem + OFF_SRC)[i] = (float)i;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar[b]) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int b = 0; b < 2; ++b) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar[b]), "r"(TILE) : "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        const int b = r & 1;
        const uint32_t dst_off = OFF_DST + b * TILE + rank * 2048;     // my 2 KB slot in every peer's tile
        if (mode == 0) {
            // thread = (row = tid / 16, t = tid % 16): element t of row `row` for peer j, 8 bytes
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t base = mapa(s0, j);
                const float2 v = reinterpret_cast<const float2*>(smem + OFF_SRC)[j * 256 + tid];
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(base + dst_off + tid * 8),
                             "f"(v.x), "f"(v.y), "r"(base + OFF_BAR + 8 * b)
                             : "memory");
            }
        } else if (mode == 1) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = 2 * jj + (tid >> 7);                     // half of the threads per peer, 16 B each
                const uint32_t base = mapa(s0, j);
                const float4 v = reinterpret_cast<const float4*>(smem + OFF_SRC)[j * 128 + (tid & 127)];
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(base + dst_off + (tid & 127) * 16),
                             "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(base + OFF_BAR + 8 * b)
                             : "memory");
            }
        } else if (mode == 2) {
            if (lane == 0) {
                for (int j = 2 * warp; j < 2 * warp + 2; ++j) {
                    const uint32_t base = mapa(s0, j);
                    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + dst_off),
                                 "r"(s0 + OFF_SRC + j * 2048), "r"(2048), "r"(base + OFF_BAR + 8 * b)
                                 : "memory");
                }
            }
        } else {
            if (tid == 0) {
                for (int j = 0; j < 16; ++j) {
                    const uint32_t base = mapa(s0, j);
                    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + dst_off),
                                 "r"(s0 + OFF_SRC + j * 2048), "r"(2048), "r"(base + OFF_BAR + 8 * b)
                                 : "memory");
                }
            }
        }
        wait(bar[b], (r >> 1) & 1);
        __syncthreads();
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar[b]), "r"(TILE) : "memory");
    }
    const long long t1 = clock64();
    if (tid == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int main() {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(CL * 148);
    int ncl = 0; cudaOccupancyMaxActiveClusters(&ncl, k, &cfg);
    printf("max active clusters: %d\n", ncl);
    unsigned long long* cyc; cudaMallocManaged(&cyc, sizeof(unsigned long long) * CL * 64);
    const int rounds = 2000;
    for (int clusters : {1, ncl}) {
        for (int mode = 0; mode < 4; ++mode) {
            cfg.gridDim = dim3(CL * clusters);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaLaunchKernelEx(&cfg, k, mode, 10, cyc); cudaDeviceSynchronize();
            cudaEventRecord(e0);
            cudaError_t e = cudaLaunchKernelEx(&cfg, k, mode, rounds, cyc);
            cudaEventRecord(e1);
            cudaError_t e2 = cudaDeviceSynchronize();
            if (e != cudaSuccess || e2 != cudaSuccess) { printf("mode %d: %s / %s\n", mode, cudaGetErrorString(e), cudaGetErrorString(e2)); return 1; }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double clk = (double)cyc[0] / rounds;
            // per SM: 2 CTAs x 32 KB sent (and as much received) per round
            printf("clusters %2d mode %d: %.3f ms, %.0f cycles per round, %.1f B/clk/SM sent (2 CTAs x 32 KB), %.2f us per round\n", clusters, mode,
                   ms, clk, 2.0 * TILE / clk, ms * 1e3 / rounds);
        }
    }
    return 0;
}
