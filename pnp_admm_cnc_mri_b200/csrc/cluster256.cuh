// cluster256.cuh — K1 kernel: whole ADMM solve of 256x256 packed planes inside an 8-CTA cluster.
// Phase code lives in cluster256_core.cuh (shared with the CPU emulator); this file holds the
// sm_100a-only parts: the DSMEM transposes, their synchronisation and the persistent loop.
//
// Synchronisation (no cluster barrier and no memory fence inside the loop):
//   * every transpose element is an asynchronous remote store (st.async) that completes 8 bytes of
//     transaction count on an mbarrier in the DESTINATION CTA: FULL1 guards tile B1, FULL2 tile B2.
//     A CTA starts a phase when its own barrier has seen all 64 KB of the tile.
//   * FREE1 / FREE2 are arrival-count barriers: each CTA tells all eight peers when its B1 / B2 may
//     be overwritten again; a producer waits for all eight before it stores (normally long done).
//   mbarrier waits / remote arrives use the default (.cta-scoped acquire / release) forms, as CUTLASS
//   does for cluster pipelines: the tiles live in shared memory, which no cache shadows, and the
//   cluster-scoped forms cost an L1 invalidate (CCTL.IVALL) per wait and MEMBAR + ERRBAR per arrive.
//   * the data term G of the next blend is bulk-copied (cp.async.bulk, GFULL) from L2 into the idle
//     B1 tile at the end of the row phase, so the blend reads it from shared memory.
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
PNP_D void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
PNP_D uint32_t mapa(uint32_t local, int rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
PNP_D void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
PNP_D void mbar_arm_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
PNP_D void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint
// expires) instead of spinning through issue slots the other CTA of the SM could use.
PNP_D bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
    return ok != 0;
}
// Bounded wait (~2 s of SM clock): a protocol bug must surface as a launch failure, never as a hung GPU.
PNP_D void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if ((++spins & 63u) == 0 && clock64() - t0 > 4000000000ll) __trap();
    }
}
// Poll `spin` times before going to sleep: a sleeping warp frees its issue slots for the other CTA of the SM but pays the
// hardware's wake-up latency on the critical path of the iteration (experiment knob PNPADMM_K1_SPIN).
PNP_D void mbar_wait_spin(uint32_t bar, uint32_t parity, int spin) {
    for (int i = 0; i < spin; ++i) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
    }
    mbar_wait(bar, parity);
}
PNP_D int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
PNP_D void st_release_gpu(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
PNP_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// DSMEM store policy: asynchronous remote store that signals the destination's mbarrier.
template <int CL>
struct RemoteAsync {
    typedef Geo<CL> G;
    uint32_t smem0;   // shared::cta address of this CTA's dynamic smem; peers' come from mapa per store (no register table)
    PNP_D void init(const unsigned char* smem) { smem0 = (uint32_t)__cvta_generic_to_shared(smem); }
    PNP_D void st(int rank, int off, cf32 v, int bar) const {
        const uint32_t b = mapa(smem0, rank);
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(b + (uint32_t)off),
                     "f"(v.re), "f"(v.im), "r"(b + (uint32_t)(G::kOffBar + 8 * bar))
                     : "memory");
    }
};

// Transposes with rotated destinations: CTA `rank` starts its 16 remote stores at peer 4 (rank % 4) instead of peer 0, so the
// 16 CTAs of a cluster do not all address the same destination at the same time (experiment knob PNPADMM_K1_ROT).
template <int OFF, int CL, class Remote>
PNP_D void row_store_remote_off(const Ctx<CL>& c, const ThreadState& s, const Remote& R) {
    typedef Geo<CL> G;
    const int t = c.rt();
    const int grow = G::kRows * c.rank + c.row();
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
        constexpr int dummy = 0; (void)dummy;
        const int j = (jj + OFF) & 15;
        const int off = G::kOffB2 + (grow * G::kRows + G::local(t, j)) * 8;
        R.st(G::dest(j), off, s.a[j], BAR_FULL2);
    }
}
template <int OFF, int CL, class Remote>
PNP_D void col_store_remote_off(const Ctx<CL>& c, const ThreadState& s, const Remote& R) {
    typedef Geo<CL> G;
    const int t = c.ct();
    const int kc = G::kRows * c.rank + c.cc();
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
        const int j = (jj + OFF) & 15;
        const int off = G::kOffB1 + (G::local(t, j) * kN + kc) * 8;
        R.st(G::dest(j), off, s.a[j], BAR_FULL1);
    }
}
template <int CL, class Remote>
PNP_D void row_store_remote_rot(const Ctx<CL>& c, const ThreadState& s, const Remote& R) {
    switch (c.rank & 3) {
        case 0: row_store_remote_off<0>(c, s, R); break;
        case 1: row_store_remote_off<4>(c, s, R); break;
        case 2: row_store_remote_off<8>(c, s, R); break;
        default: row_store_remote_off<12>(c, s, R); break;
    }
}
template <int CL, class Remote>
PNP_D void col_store_remote_rot(const Ctx<CL>& c, const ThreadState& s, const Remote& R) {
    switch (c.rank & 3) {
        case 0: col_store_remote_off<0>(c, s, R); break;
        case 1: col_store_remote_off<4>(c, s, R); break;
        case 2: col_store_remote_off<8>(c, s, R); break;
        default: col_store_remote_off<12>(c, s, R); break;
    }
}

struct ClusterParams {
    int B;                 // images
    int P;                 // packed planes
    int solo;              // one image per plane
    int iters;
    const float* z_in; const float* w_in;      // [B][256][256]
    float* x; float* z; float* w; float* xpw;  // outputs (xpw may be null)
    const cf32* G;         // [P][256][256]
    const uint32_t* mpack; // [16][256] or [P][16][256] packed mask codes (pack_mcode_k1_kernel)
    int mcode_batched;
    const float* cf;       // [3] device: residual coefficients (written by prepare)
    ProxParams<float> prox;
    // Chunked schedule: when P is not a multiple of the resident clusters, each plane's `iters` are cut
    // into n_chunks pieces of `chunk` iterations; task = chunk * P + plane.  Clusters claim tasks in
    // increasing order from an atomic queue, so the task a chunk depends on (same plane, previous chunk)
    // was always claimed earlier by a cluster that is running: no deadlock however few clusters are
    // resident (e.g. when another kernel shares the GPU).  A plane's z, w state is handed from one cluster
    // to the next through global memory (L2), guarded by progress[plane][rank] (chunks completed,
    // release/acquire at gpu scope).
    int chunk, n_chunks;
    int* progress;         // [P][16], zeroed before the launch
    int* queue;            // next unclaimed task, zeroed before the launch
    int dbg;               // timing experiments only (results invalid): 1 = no transposes, 2 = no G staging
    // Fused prologue (cluster256_core.cuh): the first chunk of a plane starts from the IMAGES instead of (z_in, w_in):
    // acquisition, zero-filled start and the data term G (written to `Gw` = G, this cluster's own tiles) inside the launch.
    int fused;
    const float* img; const uint8_t* img8;   // [B][256][256], one of the two (uint8 gray levels are divided by 255 on load)
    cf32* Gw;              // writable alias of G
    const cf32* tiles;     // [3][256][256] NcS, nH, nA in tile order (prepare_shared_kernel)
    const uint32_t* mhere; // [16][256] m[k] bits of each column-phase thread's 16 bins
    float cf1v, cf2v;      // residual coefficients by value
    int no_memset;         // the hand-off counters + task queue were zeroed by the preparation launch
    int rot;               // rotate the transpose destinations per rank (PNPADMM_K1_ROT, default 0)
    int spin;              // polls of a tile barrier before the waiting warp goes to sleep (PNPADMM_K1_SPIN, default 0)
};

// mcode [N][N] bytes -> packed words (one word per column-phase thread and iteration)
__global__ void pack_mcode_k1_kernel(const uint8_t* __restrict__ mcode, uint32_t* __restrict__ mpack, int planes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= planes * 16 * kN) return;
    const int plane = i / (16 * kN), r = i - plane * 16 * kN;
    mpack[i] = pack_codes(mcode + (size_t)plane * kN * kN, r / kN, r % kN);
}

// (mask, noise, reo) -> noise-term tiles + packed codes + m[k] bits for the fused prologue: 4096 threads, one per word.
__global__ void prepare_shared_kernel(const uint8_t* __restrict__ mask, const cf32* __restrict__ noise, int R, float g_over_n2,
                                      cf32* __restrict__ tiles, uint32_t* __restrict__ mpack, uint32_t* __restrict__ mhere,
                                      uint8_t* __restrict__ mcode, int* __restrict__ zero_ints, int n_zero,
                                      float* __restrict__ cf_out, float c1, float c2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    for (int k = i; k < n_zero; k += gridDim.x * blockDim.x) zero_ints[k] = 0;   // hand-off counters + task queue of the launch that follows
    if (i == 0) { cf_out[0] = 0.f; cf_out[1] = c1; cf_out[2] = c2; }             // coefficient table of the streaming kernels
    if (i >= 16 * kN) return;
    prepare_shared_word(mask, noise, R, g_over_n2, i / kN, i % kN, tiles, mpack, mhere, mcode);
}

// Cluster shape comes from the launch attribute (cudaLaunchAttributeClusterDimension = CL).
template <int CL>
__global__ void __launch_bounds__(Geo<CL>::kThreads, Geo<CL>::kCtasPerSm) cluster256_kernel(const ClusterParams p) {
    typedef Geo<CL> G;
    constexpr int kCluster = CL, kTileBytes = G::kTileBytes, kRows = G::kRows;
    extern __shared__ __align__(128) unsigned char smem[];
    Ctx<CL> c;
    c.rank = (int)cluster_ctarank();
    c.tid = threadIdx.x;
    c.smem = smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t bar0 = smem0 + G::kOffBar;
    const uint32_t bFull1 = bar0 + 8 * BAR_FULL1, bFull2 = bar0 + 8 * BAR_FULL2, bG = bar0 + 8 * BAR_GFULL;
    const uint32_t bFree1 = bar0 + 8 * BAR_FREE1, bFree2 = bar0 + 8 * BAR_FREE2;

    if (threadIdx.x < 256)
        fill_tw(reinterpret_cast<cf32*>(smem + G::kOffTW), reinterpret_cast<const cf32*>(g_tw_f32), threadIdx.x);
    if (threadIdx.x == 0) {
        mbar_init(bFull1, 1); mbar_init(bFull2, 1); mbar_init(bG, 1);
        mbar_init(bFree1, kCluster);                        // one arrival per CTA
        mbar_init(bFree2, kCluster * G::kWarps);      // one arrival per warp of every CTA
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    RemoteAsync<CL> R;
    R.init(smem);
    __syncthreads();
    cluster_sync_all();   // all CTAs resident, all barriers initialised, before any DSMEM traffic

#ifdef PNPADMM_K1_EXPERIMENTS
    const int dbg = p.dbg;     // timing experiments only (results invalid)
#else
    constexpr int dbg = 0;
#endif
    const float cf1 = p.fused ? p.cf1v : p.cf[1], cf2 = p.fused ? p.cf2v : p.cf[2];
    uint32_t nFull1 = 0, nFull2 = 0, nG = 0, nFree1 = 0, nFree2 = 0;   // completed uses (phase parity)
    ThreadState s;
    const size_t nn = (size_t)kN * kN;

    const int mode = (p.prox.prox == PROX_NONE) ? PROX_NONE : prox_mode(p.prox);
    const int ntasks = p.P * p.n_chunks;
    if ((p.dbg >> 8) && ((blockIdx.x / kCluster) & 1)) {   // experiment: stagger odd clusters by (dbg >> 8) us
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)(p.dbg >> 8) * 1965) {}
    }
    const uint32_t sTask = bar0 + 48;                     // task slot of this CTA (rank 0's is the cluster's)
    const uint32_t sTask0 = mapa(sTask, 0);
    for (;;) {
        if (c.rank == 0 && threadIdx.x == 0) {
            const int t = atomicAdd(p.queue, 1);
            asm volatile("st.shared.s32 [%0], %1;" ::"r"(sTask), "r"(t) : "memory");
        }
        cluster_sync_all();
        int task;
        asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(task) : "r"(sTask0) : "memory");
        cluster_sync_all();                               // everyone has the task before rank 0 claims the next
        if (task >= ntasks) break;
        const int plane = task % p.P, chunk = task / p.P;
        const int it0 = chunk * p.chunk;
        const int it1 = (it0 + p.chunk < p.iters) ? it0 + p.chunk : p.iters;
        const bool final_chunk = (it1 == p.iters);
        const int ia = p.solo ? plane : 2 * plane;
        const bool has_b = !p.solo && (2 * plane + 1 < p.B);
        PlaneIO io;
        // first chunk starts from the caller's z, w; later chunks from the state the previous cluster left
        const float* zsrc = (chunk == 0) ? p.z_in : p.z;
        const float* wsrc = (chunk == 0) ? p.w_in : p.w;
        io.z_in_a = zsrc + ia * nn; io.w_in_a = wsrc + ia * nn;
        io.z_in_b = has_b ? io.z_in_a + nn : nullptr; io.w_in_b = has_b ? io.w_in_a + nn : nullptr;
        io.x_a = p.x + ia * nn; io.z_a = p.z ? p.z + ia * nn : nullptr; io.w_a = p.w ? p.w + ia * nn : nullptr;
        io.xpw_a = p.xpw ? p.xpw + ia * nn : nullptr;
        io.x_b = io.x_a + nn; io.z_b = io.z_a ? io.z_a + nn : nullptr; io.w_b = io.w_a ? io.w_a + nn : nullptr;
        io.xpw_b = io.xpw_a ? io.xpw_a + nn : nullptr;
        if (chunk > 0) {   // rows of this rank were written by the same rank of another cluster
            if (threadIdx.x == 0) {
                uint32_t spins = 0;
                while (ld_acquire_gpu(p.progress + plane * 16 + c.rank) < chunk)
                    if (++spins > (1u << 26)) __trap();
            }
            __syncthreads();
        }
        // G arrives in tile order (prepare_kernel): this CTA's tile is one contiguous block of the plane
        const unsigned char* Gtile = reinterpret_cast<const unsigned char*>(p.G + plane * nn) + (size_t)c.rank * kTileBytes;
        const uint32_t* mpack = p.mpack + (p.mcode_batched ? (size_t)plane * 16 * kN : 0);

        // Stage G for the coming blend into this warp's 4 KB slice of B1 (two image rows = 512 / R staged G
        // rows), which only this warp used as FFT scratch: one 4 KB bulk copy per warp.
        auto prefetch_g = [&]() {
            if (dbg & 2) return;
            __syncwarp();
            fence_proxy_async();   // generic-proxy accesses of the slice are ordered before the async writes
            if (lane == 0) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem0 + (uint32_t)(G::kOffB1 + 4096 * warp)),
                             "l"(Gtile + 4096 * warp), "r"(4096), "r"(bG)
                             : "memory");
            }
        };
        // wait until every peer's B2 may be overwritten (signal of the previous column phase)
        auto wait_free2 = [&]() {
            if (dbg & 1) return;
            if (nFree2 > 0) mbar_wait(bFree2, (nFree2 - 1) & 1);
            ++nFree2;
        };

        if (p.fused && chunk == 0) {
            // ---- fused prologue: acquisition, zero-filled start and G from the images (see cluster256_core.cuh)
            const float hb = has_b ? 1.f : 0.f;
            cf32* Gw = p.Gw + plane * nn + (size_t)c.rank * (kTileBytes / 8);
            const cf32* tl = p.tiles + (size_t)c.rank * (kTileBytes / 8);
            cf32* Fsave = c.Zs();                          // idle until z exists
            // R0: packed image rows -> forward row FFT -> transpose
            row_load_image(c, s, p.img ? p.img + ia * nn : nullptr, (p.img && has_b) ? p.img + (ia + 1) * nn : nullptr,
                           p.img8 ? p.img8 + ia * nn : nullptr, (p.img8 && has_b) ? p.img8 + (ia + 1) * nn : nullptr);
            row_step1_write<false>(c, s);
            __syncwarp();
            row_read_step2<false>(c, s);
            wait_free2();
            row_store_remote(c, s, R);
            // C0: F = column FFT; save F, write G, inverse column FFT of the ms branch -> transpose
            const int wi = c.ct() * kN + kRows * c.rank + c.cc();
            const uint32_t codes0 = mpack[wi], here0 = p.mhere[wi];
            if (threadIdx.x == 0) mbar_arm_tx(bFull2, kTileBytes);
            mbar_wait(bFull2, nFull2 & 1); ++nFull2;
            col_load(c, s);
            __syncthreads();
            col_step1_write<false>(c, s);
            __syncthreads();
            col_read_step2<false>(c, s);
            col_acquire_ms(c, s, Fsave, Gw, tl, tl + nn, codes0, cf1, cf2, hb);
            asm volatile("fence.proxy.async;" ::: "memory");   // G (generic stores) is read back by bulk copies (async proxy)
            __syncthreads();
            if (threadIdx.x < kCluster) mbar_arrive_remote(mapa(bFree1, threadIdx.x));   // R0 is over: this CTA's B1 is free
            col_step1_write<true>(c, s);
            __syncthreads();
            col_read_step2<true>(c, s);
            mbar_wait(bFree1, nFree1 & 1); ++nFree1;       // (no FREE2 signal: B2 stays this CTA's scratch through C1)
            col_store_remote(c, s, R);
            // R1: T1 = inverse row FFT, kept in the dual's registers
            if (threadIdx.x == 0) mbar_arm_tx(bFull1, kTileBytes);
            mbar_wait(bFull1, nFull1 & 1); ++nFull1;
            row_load(c, s);
            __syncwarp();
            row_step1_write<true>(c, s);
            __syncwarp();
            row_read_step2<true>(c, s);
            row_stash_t1(s);
            // C1: inverse column FFT of the ma branch -> transpose
            __syncthreads();                               // every warp has left R1: B1 is free again
            if (threadIdx.x < kCluster) mbar_arrive_remote(mapa(bFree1, threadIdx.x));
            col_acquire_ma(c, s, Fsave, tl + 2 * nn, codes0, here0, hb);
            col_step1_write<true>(c, s);
            __syncthreads();
            col_read_step2<true>(c, s);
            fence_proxy_async();
            __syncwarp();
            if (lane < kCluster) mbar_arrive_remote(mapa(bFree2, lane));   // this warp no longer reads B2
            mbar_wait(bFree1, nFree1 & 1); ++nFree1;
            col_store_remote(c, s, R);
            // R2: T2 = inverse row FFT; z = x0 = |ifft2(y)|, w = 0; then the loop's first forward row FFT
            if (threadIdx.x == 0) mbar_arm_tx(bFull1, kTileBytes);
            mbar_wait(bFull1, nFull1 & 1); ++nFull1;
            row_load(c, s);
            __syncwarp();
            row_step1_write<true>(c, s);
            __syncwarp();
            row_read_step2<true>(c, s);
            __syncthreads();                               // F (in Zs) was last read in C1 by every warp before this point
            row_zero_fill(c, s, 1.0f / (float)(kN * kN), has_b);
        } else {
            // ---- prologue: state from the caller (or from the cluster that ran the previous chunk)
            row_load_state(c, s, io);
        }
        // first forward row FFT of z - w
        __syncwarp();
        row_step1_write<false>(c, s);
        __syncwarp();
        row_read_step2<false>(c, s);
        prefetch_g();
        wait_free2();
        if (!(dbg & 1)) { if (p.rot) row_store_remote_rot(c, s, R); else row_store_remote(c, s, R); }

        for (int it = it0; it < it1; ++it) {
            // ---- column phase: col FFT -> residual blend -> col IFFT -> transpose back (DSMEM)
            const uint32_t codes = mpack[c.ct() * kN + kRows * c.rank + c.cc()];
            // A tile barrier has one pending arrival: thread 0's expect_tx, made right before the wait, so a phase
            // cannot complete early (bytes landing first only drive the tx-count negative) and no arrival is left
            // without a wait when the kernel exits (compute-sanitizer synccheck).
            if (!(dbg & 1)) {
                if (threadIdx.x == 0) mbar_arm_tx(bFull2, kTileBytes);
                mbar_wait_spin(bFull2, nFull2 & 1, p.spin); ++nFull2;
            }
            col_load(c, s);
            __syncthreads();
            col_step1_write<false>(c, s);
            __syncthreads();
            col_read_step2<false>(c, s);
            if (!(dbg & 2)) {
                if (threadIdx.x == 0) mbar_arm_tx(bG, kTileBytes);
                mbar_wait(bG, nG & 1); ++nG;
            }
            col_blend(c, s, c.B1(), codes, cf1, cf2);
            fence_proxy_async();
            __syncthreads();                               // everyone is done with G (B1) and the scratch reads
            if (threadIdx.x < kCluster && !(dbg & 1)) mbar_arrive_remote(mapa(bFree1, threadIdx.x));
            col_step1_write<true>(c, s);
            __syncthreads();
            col_read_step2<true>(c, s);
            fence_proxy_async();
            __syncwarp();
            if (!(dbg & 1)) {
                if (lane < kCluster) mbar_arrive_remote(mapa(bFree2, lane));   // this warp no longer reads B2
                mbar_wait(bFree1, nFree1 & 1); ++nFree1;
                if (p.rot) col_store_remote_rot(c, s, R); else col_store_remote(c, s, R);
            }

            // ---- row phase: row IFFT -> |v + r| -> prox -> dual -> row FFT -> transpose (DSMEM)
            const bool last = (it == it1 - 1);
            if (!(dbg & 1)) {
                if (threadIdx.x == 0) mbar_arm_tx(bFull1, kTileBytes);
                mbar_wait_spin(bFull1, nFull1 & 1, p.spin); ++nFull1;
            }
            row_load(c, s);
            __syncwarp();
            row_step1_write<true>(c, s);
            __syncwarp();
            row_read_step2<true>(c, s);
            row_prox_dispatch(mode, c, s, p.prox, has_b, last, final_chunk, io);
            if (!last) {
                __syncwarp();
                row_step1_write<false>(c, s);
                __syncwarp();
                row_read_step2<false>(c, s);
                prefetch_g();
                wait_free2();
                if (!(dbg & 1)) { if (p.rot) row_store_remote_rot(c, s, R); else row_store_remote(c, s, R); }
            }
        }
        if (!final_chunk) {   // publish this rank's rows of z, w for the cluster that continues the plane
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                st_release_gpu(p.progress + plane * 16 + c.rank, chunk + 1);
            }
        }
    }
    cluster_sync_all();   // no CTA exits while a peer could still signal one of its barriers
}


// =============================================================================================
// Blocked-tile variant of K1 (16-CTA clusters): the two transposes of an iteration leave each CTA as 16 bulk copies of
// 2 KB (cp.async.bulk shared::cta -> shared::cluster, completion counted in bytes on the destination's FULL barrier)
// issued by one warp, instead of 16 st.async per thread through the LSU.  The LSU / MIO queue then carries only the
// local exchanges, so the second CTA of the SM keeps computing while this one's tile is on the SM-to-SM network.
//   * staging: the outgoing tile is first written to this CTA's own B1 (row phase) / B2 (column phase) as
//     [dest][256 entries] (cluster256_core.cuh), contiguous 8-byte stores, then fence.proxy.async + CTA barrier.
//   * a bulk copy out of shared memory has no completion the SENDER could wait on, so buffer reuse is ordered by two
//     arrival barriers that say "all 16 CTAs of the cluster have RECEIVED their whole tile":  RCV2 (each CTA arrives
//     at every peer after its own FULL2 wait) means every row->column copy has landed, hence every B1 staging buffer
//     has been read and every CTA has left its row phase: B1 tiles may be overwritten by the column->row copies.
//     RCV1 likewise (after FULL1) frees the B2 tiles for the next row->column copies.
//   * the data term G of the blend is read from L2 (tile-ordered copy, 256 contiguous bytes per warp and j); B1 cannot
//     stage it any more (it holds the outgoing tile until RCV2).
// =============================================================================================
enum { BK_FULL1 = 0, BK_FULL2 = 1, BK_RCV1 = 2, BK_RCV2 = 3 };

__global__ void __launch_bounds__(Geo<16>::kThreads, 2) cluster256_bk_kernel(const ClusterParams p) {
    typedef Geo<16> G;
    constexpr int kCluster = 16, kTileBytes = G::kTileBytes;
    extern __shared__ __align__(128) unsigned char smem[];
    Ctx<16> c;
    c.rank = (int)cluster_ctarank();
    c.tid = threadIdx.x;
    c.smem = smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t bar0 = smem0 + G::kOffBar;
    const uint32_t bFull1 = bar0 + 8 * BK_FULL1, bFull2 = bar0 + 8 * BK_FULL2;
    const uint32_t bRcv1 = bar0 + 8 * BK_RCV1, bRcv2 = bar0 + 8 * BK_RCV2;

    fill_tw(reinterpret_cast<cf32*>(smem + G::kOffTW), reinterpret_cast<const cf32*>(g_tw_f32), threadIdx.x);
    if (threadIdx.x == 0) {
        mbar_init(bFull1, 1); mbar_init(bFull2, 1);
        mbar_init(bRcv1, kCluster); mbar_init(bRcv2, kCluster);      // one arrival per CTA of the cluster
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();   // all CTAs resident, all barriers initialised, before any DSMEM traffic

    const float cf1 = p.cf[1], cf2 = p.cf[2];
    uint32_t nFull1 = 0, nFull2 = 0;   // completed waits = phases of FULL1 / FULL2 and of RCV1 / RCV2
    ThreadState s;
    const size_t nn = (size_t)kN * kN;
    const int mode = (p.prox.prox == PROX_NONE) ? PROX_NONE : prox_mode(p.prox);
    const int ntasks = p.P * p.n_chunks;
    const uint32_t sTask = bar0 + 48;
    const uint32_t sTask0 = mapa(sTask, 0);

    // warp 0 sends the staged tile: lane j copies block j (2 KB) into peer j, at this CTA's block, and signals `bar` there
    auto send_tile = [&](int off_src, int off_dst, int bar) {
        if (warp == 0 && lane < kCluster) {
            const int dst = (lane + c.rank) & 15;          // every CTA starts with a different peer: no hot destination
            const uint32_t peer = mapa(smem0, dst);
            asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             peer + (uint32_t)(off_dst + 2048 * c.rank)),
                         "r"(smem0 + (uint32_t)(off_src + 2048 * dst)), "r"(2048), "r"(peer + (uint32_t)(G::kOffBar + 8 * bar))
                         : "memory");
        }
    };
    auto announce = [&](uint32_t bar_local) {      // "this CTA has received its tile" -> every peer
        if (warp == 0 && lane < kCluster) mbar_arrive_remote(mapa(bar_local, lane));
    };

    for (;;) {
        if (c.rank == 0 && threadIdx.x == 0) {
            const int t = atomicAdd(p.queue, 1);
            asm volatile("st.shared.s32 [%0], %1;" ::"r"(sTask), "r"(t) : "memory");
        }
        cluster_sync_all();
        int task;
        asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(task) : "r"(sTask0) : "memory");
        cluster_sync_all();
        if (task >= ntasks) break;
        const int plane = task % p.P, chunk = task / p.P;
        const int it0 = chunk * p.chunk;
        const int it1 = (it0 + p.chunk < p.iters) ? it0 + p.chunk : p.iters;
        const bool final_chunk = (it1 == p.iters);
        const int ia = p.solo ? plane : 2 * plane;
        const bool has_b = !p.solo && (2 * plane + 1 < p.B);
        PlaneIO io;
        const float* zsrc = (chunk == 0) ? p.z_in : p.z;
        const float* wsrc = (chunk == 0) ? p.w_in : p.w;
        io.z_in_a = zsrc + ia * nn; io.w_in_a = wsrc + ia * nn;
        io.z_in_b = has_b ? io.z_in_a + nn : nullptr; io.w_in_b = has_b ? io.w_in_a + nn : nullptr;
        io.x_a = p.x + ia * nn; io.z_a = p.z ? p.z + ia * nn : nullptr; io.w_a = p.w ? p.w + ia * nn : nullptr;
        io.xpw_a = p.xpw ? p.xpw + ia * nn : nullptr;
        io.x_b = io.x_a + nn; io.z_b = io.z_a ? io.z_a + nn : nullptr; io.w_b = io.w_a ? io.w_a + nn : nullptr;
        io.xpw_b = io.xpw_a ? io.xpw_a + nn : nullptr;
        if (chunk > 0) {
            if (threadIdx.x == 0) {
                uint32_t spins = 0;
                while (ld_acquire_gpu(p.progress + plane * 16 + c.rank) < chunk)
                    if (++spins > (1u << 26)) __trap();
            }
            __syncthreads();
        }
        const cf32* Gtile = p.G + plane * nn + (size_t)c.rank * (kTileBytes / 8);
        const uint32_t* mpack = p.mpack + (p.mcode_batched ? (size_t)plane * 16 * kN : 0);

        // row -> column transpose of the tile in registers: stage in B1, then 16 bulk copies into the peers' B2
        auto send_rows = [&]() {
            __syncwarp();                  // the half-warps are done reading the scratch (= their staging slots)
            row_stage_bk(c, s);
            fence_proxy_async();           // generic-proxy stores -> visible to the bulk copies (async proxy)
            __syncthreads();
            if (warp == 0) {
                if (nFull1 > 0) mbar_wait(bRcv1, (nFull1 - 1) & 1);   // every B2 tile has been consumed and re-read
                send_tile(G::kOffB1, G::kOffB2, BK_FULL2);
            }
        };

        // ---- prologue: first forward row FFT of z - w
        row_load_state(c, s, io);
        row_step1_write_bk<false>(c, s);
        __syncwarp();
        row_read_step2_bk<false>(c, s);
        send_rows();

        for (int it = it0; it < it1; ++it) {
            // ---- column phase: col FFT -> residual blend -> col IFFT -> transpose back
            const uint32_t codes = mpack[c.ct() * kN + 16 * c.rank + c.cc()];
            if (threadIdx.x == 0) mbar_arm_tx(bFull2, kTileBytes);   // the one pending arrival, right before the wait
            mbar_wait_spin(bFull2, nFull2 & 1, p.spin); ++nFull2;
            announce(bRcv2);
            col_load(c, s);
            __syncthreads();
            col_step1_write<false>(c, s);
            __syncthreads();
            col_read_step2<false>(c, s);
            col_blend_g(c, s, Gtile, codes, cf1, cf2);
            __syncthreads();                               // scratch reads done
            col_step1_write<true>(c, s);
            __syncthreads();
            col_read_step2<true>(c, s);
            __syncthreads();                               // scratch reads done: B2 becomes the staging buffer
            col_stage_bk(c, s);
            fence_proxy_async();
            __syncthreads();
            if (warp == 0) {
                mbar_wait(bRcv2, (nFull2 - 1) & 1);         // every CTA has left its row phase, every B1 staging was read
                send_tile(G::kOffB2, G::kOffB1, BK_FULL1);
            }

            // ---- row phase: row IFFT -> |v + r| -> prox -> dual -> row FFT -> transpose
            const bool last = (it == it1 - 1);
            if (threadIdx.x == 0) mbar_arm_tx(bFull1, kTileBytes);
            mbar_wait_spin(bFull1, nFull1 & 1, p.spin); ++nFull1;
            announce(bRcv1);
            row_load_bk(c, s);
            __syncwarp();
            row_step1_write_bk<true>(c, s);
            __syncwarp();
            row_read_step2_bk<true>(c, s);
            row_prox_dispatch(mode, c, s, p.prox, has_b, last, final_chunk, io);
            if (!last) {
                __syncwarp();
                row_step1_write_bk<false>(c, s);
                __syncwarp();
                row_read_step2_bk<false>(c, s);
                send_rows();
            }
        }
        if (!final_chunk) {
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence();
                st_release_gpu(p.progress + plane * 16 + c.rank, chunk + 1);
            }
        }
    }
    cluster_sync_all();
}

}  // namespace k1
}  // namespace pnp
