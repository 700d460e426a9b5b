// Probe: how many clusters of a given shape can be co-resident (cudaOccupancyMaxActiveClusters),
// and does a launch with that many clusters actually run concurrently (clock-overlap check).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512) k512(int* out, long long* t0, long long* t1) {
    extern __shared__ unsigned char sm[];
    if (threadIdx.x == 0) { t0[blockIdx.x] = clock64(); unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid)); out[blockIdx.x] = smid; }
    long long s = clock64(); while (clock64() - s < 2000000) {}
    sm[threadIdx.x] = 1;
    if (threadIdx.x == 0) t1[blockIdx.x] = clock64();
}
__global__ void __launch_bounds__(256, 2) k256(int* out, long long* t0, long long* t1) {
    extern __shared__ unsigned char sm[];
    if (threadIdx.x == 0) { t0[blockIdx.x] = clock64(); unsigned smid; asm("mov.u32 %0, %%smid;" : "=r"(smid)); out[blockIdx.x] = smid; }
    long long s = clock64(); while (clock64() - s < 2000000) {}
    sm[threadIdx.x] = 1;
    if (threadIdx.x == 0) t1[blockIdx.x] = clock64();
}
template <typename K> void probe(K kern, const char* name, int threads, int csize, int smem) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (csize > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize * 148); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    printf("%s threads=%d cluster=%2d smem=%3d KB -> max active clusters %d (%s) = %d CTAs\n", name, threads, csize, smem / 1024, n,
           cudaGetErrorString(e), n * csize);
    cudaGetLastError();
    if (n > 0) {
        int nb = n * csize; int* out; long long *t0, *t1;
        cudaMallocManaged(&out, nb * 4); cudaMallocManaged(&t0, nb * 8); cudaMallocManaged(&t1, nb * 8);
        cfg.gridDim = dim3(nb);
        e = cudaLaunchKernelEx(&cfg, kern, out, t0, t1);
        cudaError_t e2 = cudaDeviceSynchronize();
        if (e != cudaSuccess || e2 != cudaSuccess) { printf("   launch: %s / %s\n", cudaGetErrorString(e), cudaGetErrorString(e2)); cudaGetLastError(); return; }
        long long first_end = t1[0], last_start = t0[0]; int cnt[256] = {0}; int mx = 0, used = 0;
        for (int i = 0; i < nb; ++i) { if (t1[i] < first_end) first_end = t1[i]; if (t0[i] > last_start) last_start = t0[i]; cnt[out[i]]++; }
        for (int i = 0; i < 256; ++i) { if (cnt[i]) used++; if (cnt[i] > mx) mx = cnt[i]; }
        printf("   launched %d CTAs: all concurrent=%s, SMs used=%d, max CTAs on one SM=%d\n", nb, last_start < first_end ? "yes" : "NO", used, mx);
        cudaFree(out); cudaFree(t0); cudaFree(t1);
    }
}
int main() {
    probe(k512, "k512", 512, 8, 197 * 1024);
    probe(k512, "k512", 512, 4, 197 * 1024);
    probe(k512, "k512", 512, 2, 197 * 1024);
    probe(k512, "k512", 512, 16, 197 * 1024);
    probe(k256, "k256", 256, 16, 99 * 1024);
    probe(k256, "k256", 256, 8, 99 * 1024);
    probe(k256, "k256", 256, 16, 110 * 1024);
    probe(k256, "k256", 256, 4, 99 * 1024);
    return 0;
}
