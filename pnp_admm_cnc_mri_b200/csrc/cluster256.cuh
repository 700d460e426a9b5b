// cluster256.cuh — K1 kernel: whole ADMM solve of 256x256 packed planes inside an 8-CTA cluster.
// Phase code lives in cluster256_core.cuh (shared with the CPU emulator); this file holds the
// sm_100a-only parts: cluster barriers, DSMEM stores and the persistent loop.
#pragma once

#include "cluster256_core.cuh"
#include "streaming.cuh"   // g_tw_f32

namespace pnp {
namespace k1 {

PNP_D uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
PNP_D void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
PNP_D void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
PNP_D void cluster_sync_all() { cluster_arrive(); cluster_wait(); }

// DSMEM store policy: shared::cluster address of (rank, byte offset) via mapa.
struct RemoteDsmem {
    uint32_t base[kCluster];   // shared::cluster base address of every CTA's dynamic smem
    PNP_D void init(const unsigned char* smem) {
        const uint32_t local = (uint32_t)__cvta_generic_to_shared(smem);
#pragma unroll
        for (int r = 0; r < kCluster; ++r)
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(base[r]) : "r"(local), "r"(r));
    }
    PNP_D void st(int rank, int off, cf32 v) const {
        asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(base[rank] + (uint32_t)off), "f"(v.re), "f"(v.im)
                     : "memory");
    }
};

struct ClusterParams {
    int B;                 // images
    int P;                 // packed planes
    int solo;              // one image per plane
    int iters;
    const float* z_in; const float* w_in;      // [B][256][256]
    float* x; float* z; float* w; float* xpw;  // outputs (xpw may be null)
    const cf32* G;         // [P][256][256]
    const uint8_t* mcode;  // [256][256] or [P][256][256]
    int mcode_batched;
    const float* cf;       // [3] device: blend coefficients (written by prepare)
    ProxParams<float> prox;
};

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 1)
cluster256_kernel(const ClusterParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    Ctx c;
    c.rank = (int)cluster_ctarank();
    c.tid = threadIdx.x;
    c.smem = smem;
    const int cluster_id = blockIdx.x / kCluster;
    const int nclusters = gridDim.x / kCluster;

    if (threadIdx.x < 256)
        fill_tw(reinterpret_cast<cf32*>(smem + kOffTW), reinterpret_cast<const cf32*>(g_tw_f32), threadIdx.x);
    RemoteDsmem R;
    R.init(smem);
    __syncthreads();
    cluster_sync_all();   // every CTA of the cluster is resident before any DSMEM traffic

    const float cf0 = p.cf[0], cf1 = p.cf[1], cf2 = p.cf[2];
    ThreadState s;
    const size_t nn = (size_t)kN * kN;
    for (int plane = cluster_id; plane < p.P; plane += nclusters) {
        const int ia = p.solo ? plane : 2 * plane;
        const bool has_b = !p.solo && (2 * plane + 1 < p.B);
        PlaneIO io;
        io.z_in_a = p.z_in + ia * nn; io.w_in_a = p.w_in + ia * nn;
        io.z_in_b = has_b ? io.z_in_a + nn : nullptr; io.w_in_b = has_b ? io.w_in_a + nn : nullptr;
        io.x_a = p.x + ia * nn; io.z_a = p.z ? p.z + ia * nn : nullptr; io.w_a = p.w ? p.w + ia * nn : nullptr;
        io.xpw_a = p.xpw ? p.xpw + ia * nn : nullptr;
        io.x_b = io.x_a + nn; io.z_b = io.z_a ? io.z_a + nn : nullptr; io.w_b = io.w_a ? io.w_a + nn : nullptr;
        io.xpw_b = io.xpw_a ? io.xpw_a + nn : nullptr;
        const cf32* G = p.G + plane * nn;
        const uint8_t* mcode = p.mcode + (p.mcode_batched ? plane * nn : 0);

        // prologue: first forward row FFT of z - w
        row_load_state(c, s, io);
        row_step1_write<false>(c, s);
        __syncwarp();
        row_read_step2<false>(c, s);
        row_store_remote(c, s, R);
        cluster_sync_all();

        for (int it = 0; it < p.iters; ++it) {
            // ---- column phase: col FFT -> blend -> col IFFT -> transpose back (DSMEM)
            col_load(c, s);
            __syncthreads();
            col_step1_write<false>(c, s);
            __syncthreads();
            col_read_step2<false>(c, s);
            col_blend(c, s, G, mcode, cf0, cf1, cf2);
            __syncthreads();
            col_step1_write<true>(c, s);
            __syncthreads();
            col_read_step2<true>(c, s);
            col_store_remote(c, s, R);
            cluster_sync_all();

            // ---- row phase: row IFFT -> |.| -> prox -> dual -> row FFT -> transpose (DSMEM)
            const bool last = (it == p.iters - 1);
            row_load(c, s);
            __syncwarp();
            row_step1_write<true>(c, s);
            __syncwarp();
            row_read_step2<true>(c, s);
            row_prox(c, s, p.prox, has_b, last, io);
            if (!last) {
                __syncwarp();
                row_step1_write<false>(c, s);
                __syncwarp();
                row_read_step2<false>(c, s);
                row_store_remote(c, s, R);
                cluster_sync_all();
            }
        }
    }
    cluster_sync_all();   // no CTA exits while a peer could still address its shared memory
}

}  // namespace k1
}  // namespace pnp
