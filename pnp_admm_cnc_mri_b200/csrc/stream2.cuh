// stream2.cuh — K2: TMA-staged streaming kernels for N = 256, 512, 1024 (fp32), two passes per iteration.
//
//   rows pass : row-IFFT of iteration k -> x = |v + r| -> prox -> dual -> row-FFT of iteration k+1
//   cols pass : col-FFT -> data-consistency residual blend -> col-IFFT
//
// (algorithm: streaming.cuh header; FFT stages: stream2_core.cuh).  Differences from the generic
// kernels of streaming.cuh: the 16-point butterflies run in registers (radix 16 x 16 [x 2 | x 4]) so a
// 1-D transform touches shared memory once or twice instead of log4 N times; tiles are staged
// asynchronously; every global access is a full 64/128-byte segment.
//
// rows2_kernel (one tile of 128 / T rows per CTA, 4 CTAs per SM): thread 0 issues 1-D bulk copies
//   (cp.async.bulk, SASS UBLKCP) of the K rows and the z / w rows of both images into shared memory,
//   completion on an mbarrier; results leave from registers as coalesced stores.  The block scheduler
//   overlaps the load of one CTA with the math of the three others.
// cols2_kernel (persistent, one CTA per SM for N >= 512): a tile is all N rows x C adjacent columns
//   (C * 8 = 64 or 128 contiguous bytes per row).  K tiles are double-buffered with 16-byte
//   asynchronous copies (cp.async.cg, SASS LDGSTS): tile i+1 and the data term G of tile i stream in
//   while tile i is transformed.  Lanes run along the columns, so shared-memory accesses are
//   conflict-free without padding and each global store instruction covers whole row segments.
#pragma once

#include <cuda.h>              // CUtensorMap (types only; the encoder is fetched through the runtime)

#include "cluster256.cuh"     // mbarrier / bulk-copy helpers
#include "stream2_core.cuh"
#include "streaming.cuh"      // StreamParams, modes, twiddle master table

namespace pnp {
namespace s2 {

// programmatic dependent launch (griddepcontrol): see rows2_kernel
PNP_D void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
PNP_D void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

constexpr int kRowsThreads = 128;
template <int V> struct IntC { static constexpr int value = V; };      // compile-time int tag for generic lambdas

template <int N> struct ColsGeo {
#ifndef PNP_COLS512_C
#define PNP_COLS512_C 8
#endif
#ifndef PNP_COLS1024_C
#define PNP_COLS1024_C 8
#endif
    static constexpr int C = (N == 1024) ? PNP_COLS1024_C : (N == 512 ? PNP_COLS512_C : 16);   // columns per tile
    static constexpr int kThreads = C * Plan<N>::T;          // 256 (N = 256) or 512
    static constexpr int kCtasPerSm = (N == 256 || kThreads == 256) ? 2 : 1;
    static constexpr int kTileElems = N * C;
    static constexpr int kTw3Bytes = (Plan<N>::R3 - 1) * 256 * 8;                  // stage-3 twiddle table
    static constexpr int kSmemBytes = 3 * kTileElems * 8 + 256 * 8 + kTw3Bytes;   // 2 K slots + G slot + TW256 + TW3
};
template <int N> struct RowsGeo {
    static constexpr int T = Plan<N>::T;
    static constexpr int L = kRowsThreads / T;               // rows per CTA: 8 / 4 / 2
    static constexpr int kPitch = Plan<N>::kRowPitch;
    static constexpr int kOffZW = L * kPitch * 8;
    static constexpr int kOffTW = kOffZW + 4 * L * N * 4;
    static constexpr int kOffBar = kOffTW + 256 * 8;
    static constexpr int kSmemBytes = kOffBar + 16;   // two mbarriers: K rows + twiddles, z / w (or img) rows
};

PNP_D void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
PNP_D void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
PNP_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPEND> PNP_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory"); }

// rows pass: a line's T threads sit in one warp (T <= 32) or two; columns pass: a column's threads are
// spread over every warp of the CTA (lanes run along the columns).
template <int T, bool COLS> PNP_D void line_sync() {
    if (!COLS && T <= 32) __syncwarp(); else __syncthreads();
}

// TW256[i * 16 + k] = W_256^(i k), contiguous in global memory so that it is staged with the tile copies
__device__ __align__(128) float2 g_tw256[256];

// one 1-D transform of my line, registers -> registers (n = t + T m layout on both sides)
template <bool INV, int N, bool COLS, class Line, class Tw3>
PNP_D void fft_regs(cf32 (&a)[16], int t, const Line& ln, const cf32* TW, const Tw3& tw3) {
    constexpr int T = Plan<N>::T;
    stage1_store<INV, N>(a, t, ln);
    line_sync<T, COLS>();
    stage2_load<INV, N>(a, t, ln, TW);
    if (Plan<N>::R3 > 1) {
        line_sync<T, COLS>();
        stage2_store<N>(a, t, ln);
        line_sync<T, COLS>();
        stage3<INV, N>(a, t, ln, tw3);
    }
}

// ----------------------------------------------------------------------------------------------
// Rows pass.  grid = tiles = planes * N / L; block = 128 threads = L rows x T threads.
// ----------------------------------------------------------------------------------------------
template <int N, int MODE>
__global__ void __launch_bounds__(kRowsThreads, 4) rows2_kernel(const StreamParams<float> p) {
    typedef RowsGeo<N> G;
    constexpr int T = G::T, L = G::L;
    extern __shared__ __align__(128) unsigned char smem2[];
    cf32* Ks = reinterpret_cast<cf32*>(smem2);
    float* zw = reinterpret_cast<float*>(smem2 + G::kOffZW);      // [za | wa | zb | wb], each [L][N]
    cf32* TW = reinterpret_cast<cf32*>(smem2 + G::kOffTW);
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(smem2);
    const uint32_t bar = smem0 + G::kOffBar;

    const int tid = threadIdx.x, line = tid / T, t = tid % T;
    constexpr int tiles_per_plane = N / L;
    const int plane = blockIdx.x / tiles_per_plane;
    const int r0 = (blockIdx.x - plane * tiles_per_plane) * L;
    const size_t nn = (size_t)N * N;
    const size_t tile_off = (size_t)plane * nn + (size_t)r0 * N;     // element offset of the tile in a [planes][N][N] array
    const int ia = p.solo ? plane : 2 * plane;
    const bool has_b = !p.solo && (2 * plane + 1 < p.B);
    const size_t offa = (size_t)ia * nn + (size_t)r0 * N;
    const size_t offb = offa + nn;
    constexpr uint32_t kRowBytes = N * 8, kRealBytes = L * N * 4;

    constexpr bool kLoadK = (MODE == RM_INV_PROX_FWD || MODE == RM_INV_X || MODE == RM_INV_ABS);
    constexpr bool kLoadZW = (MODE == RM_INV_PROX_FWD || MODE == RM_INV_X || MODE == RM_FWD_ZW);
    const uint32_t bar2 = bar + 8;          // z / w / img rows: needed only after the inverse transform
    if (tid == 0) {
        k1::mbar_init(bar, 1);
        k1::mbar_init(bar2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        uint32_t bytes = 256 * 8, bytes2 = 0;
        if (kLoadK) bytes += L * kRowBytes;
        if (kLoadZW) bytes2 += (has_b ? 4 : 2) * kRealBytes;
        if (MODE == RM_FWD_IMG) bytes2 += kRealBytes;
        k1::mbar_arm_tx(bar, bytes);
        if (bytes2) k1::mbar_arm_tx(bar2, bytes2);
    }
    __syncthreads();                        // barriers initialised and armed before any copy or wait
    // Programmatic dependent launch: everything above needs nothing from the kernel before this one in the stream; its output
    // is complete and visible after the wait (a no-op for an ordinary launch).  The dependents may be scheduled right away: they
    // wait at the same point.
    pdl_wait();
    pdl_launch_dependents();
    // Issuing a bulk copy costs its thread ~140 cycles (measured on K5), so the copies are issued by one lane of
    // each of the four warps instead of one after the other by thread 0 (the K rows, needed first, by warps 0 and 3).
    if (tid == 0) {
        bulk_g2s(smem0 + G::kOffTW, g_tw256, 256 * 8, bar);
        if (kLoadK) {
            const cf32* src = (MODE == RM_INV_ABS ? p.cin : p.K) + tile_off;
#pragma unroll
            for (int l = 0; l < (L >= 4 ? L / 2 : L); ++l) bulk_g2s(smem0 + l * G::kPitch * 8, src + (size_t)l * N, kRowBytes, bar);
        }
    } else if (tid == 96) {
        if (kLoadK && L >= 4) {
            const cf32* src = (MODE == RM_INV_ABS ? p.cin : p.K) + tile_off;
#pragma unroll
            for (int l = L / 2; l < L; ++l) bulk_g2s(smem0 + l * G::kPitch * 8, src + (size_t)l * N, kRowBytes, bar);
        }
    } else if (tid == 32) {
        if (kLoadZW) {
            bulk_g2s(smem0 + G::kOffZW, p.z + offa, kRealBytes, bar2);
            bulk_g2s(smem0 + G::kOffZW + kRealBytes, p.w + offa, kRealBytes, bar2);
        }
        if (MODE == RM_FWD_IMG) bulk_g2s(smem0 + G::kOffZW, p.img + tile_off, kRealBytes, bar2);
    } else if (tid == 64) {
        if (kLoadZW && has_b) {
            bulk_g2s(smem0 + G::kOffZW + 2 * kRealBytes, p.z + offb, kRealBytes, bar2);
            bulk_g2s(smem0 + G::kOffZW + 3 * kRealBytes, p.w + offb, kRealBytes, bar2);
        }
    }
    k1::mbar_wait(bar, 0);

    RowLine ln;
    ln.line = Ks + line * G::kPitch;
    const Tw3Master tw3{reinterpret_cast<const cf32*>(g_tw_f32), kTwMax / N};
    const float* za_s = zw + line * N;
    const float* wa_s = za_s + L * N;
    const float* zb_s = wa_s + L * N;
    const float* wb_s = zb_s + L * N;
    const size_t ga = offa + (size_t)line * N, gb = offb + (size_t)line * N;   // my row in the real planes
    const size_t gk = tile_off + (size_t)line * N;                             // my row in the complex planes
    cf32 a[16];

    if (kLoadK) {   // inverse transform of the landed row
#pragma unroll
        for (int m = 0; m < 16; ++m) a[m] = ln.raw(t + T * m);
        line_sync<T, false>();                     // landing layout fully read before the padded layout is written
        fft_regs<true, N, false>(a, t, ln, TW, tw3);
    }
    if (kLoadZW || MODE == RM_FWD_IMG) k1::mbar_wait(bar2, 0);

    if (MODE == RM_INV_ABS) {               // zero-filled reconstruction |ifft2(y)|          (S1:100)
#pragma unroll
        for (int m = 0; m < 16; ++m) p.x[gk + t + T * m] = psqrt(a[m].re * a[m].re + a[m].im * a[m].im) * p.scale;
        return;
    }
    if (MODE == RM_INV_X) {                 // x-update only: x and x + w for the denoiser    (S3:259-276)
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int n = t + T * m;
            const float wa = wa_s[n];
            const float xa = pabs((za_s[n] - wa) + a[m].re);
            p.x[ga + n] = xa;
            if (p.xpw) p.xpw[ga + n] = xa + wa;
            if (has_b) {
                const float wb = wb_s[n];
                const float xb = pabs((zb_s[n] - wb) + a[m].im);
                p.x[gb + n] = xb;
                if (p.xpw) p.xpw[gb + n] = xb + wb;
            }
        }
        return;
    }
    if (MODE == RM_INV_PROX_FWD) {
        const int pm = prox_mode(p.prox);
        // The common case (both images of the plane present, not the last iteration, L1 or CNC in clamp form) runs a loop without a single
        // run-time condition: with the mode / has_b / last tests inside it, every pixel of the unrolled loop becomes its own basic blocks and
        // the 32 independent prox chains cannot be interleaved.  Everything else takes the general loop below.
        if (has_b && !p.last && pm != PM_GENERAL) {
            auto fast = [&](auto mode_c) {
                constexpr int PM = decltype(mode_c)::value;
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const int n = t + T * m;
                    float za = za_s[n], wa = wa_s[n], zb = zb_s[n], wb = wb_s[n];
                    const float xa = pabs((za - wa) + a[m].re), xb = pabs((zb - wb) + a[m].im);
                    prox_dual_m<PM>(p.prox, xa, za, wa);
                    prox_dual_m<PM>(p.prox, xb, zb, wb);
                    p.z[ga + n] = za; p.w[ga + n] = wa;
                    p.z[gb + n] = zb; p.w[gb + n] = wb;
                    a[m] = mk<float>(za - wa, zb - wb);
                }
            };
            if (pm == PM_CNC) fast(IntC<PM_CNC>{}); else fast(IntC<PM_L1>{});
        } else
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int n = t + T * m;
            float za = za_s[n], wa = wa_s[n];
            const float xa = pabs((za - wa) + a[m].re);
            if (pm == PM_CNC) prox_dual_m<PM_CNC>(p.prox, xa, za, wa);
            else if (pm == PM_L1) prox_dual_m<PM_L1>(p.prox, xa, za, wa);
            else prox_dual_m<PM_GENERAL>(p.prox, xa, za, wa);
            p.z[ga + n] = za; p.w[ga + n] = wa;
            if (p.last) p.x[ga + n] = xa;
            float zb = 0.f, wb = 0.f;
            if (has_b) {
                zb = zb_s[n]; wb = wb_s[n];
                const float xb = pabs((zb - wb) + a[m].im);
                if (pm == PM_CNC) prox_dual_m<PM_CNC>(p.prox, xb, zb, wb);
                else if (pm == PM_L1) prox_dual_m<PM_L1>(p.prox, xb, zb, wb);
                else prox_dual_m<PM_GENERAL>(p.prox, xb, zb, wb);
                p.z[gb + n] = zb; p.w[gb + n] = wb;
                if (p.last) p.x[gb + n] = xb;
            }
            a[m] = mk<float>(za - wa, zb - wb);
        }
        if (p.last) return;
        line_sync<T, false>();                     // last reads of the inverse transform done before stage-1 stores
    }
    if (MODE == RM_FWD_ZW) {
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int n = t + T * m;
            a[m] = mk<float>(za_s[n] - wa_s[n], has_b ? zb_s[n] - wb_s[n] : 0.f);
        }
    }
    if (MODE == RM_FWD_IMG) {
#pragma unroll
        for (int m = 0; m < 16; ++m) a[m] = mk<float>(za_s[t + T * m], 0.f);
    }
    fft_regs<false, N, false>(a, t, ln, TW, tw3);
    cf32* out = (MODE == RM_FWD_IMG ? p.cout : p.K) + gk;
#pragma unroll
    for (int m = 0; m < 16; ++m) out[t + T * m] = a[m];
}

// ----------------------------------------------------------------------------------------------
// Columns pass.  Persistent: CTA b handles tiles b, b + gridDim.x, ...; tile = (plane, C columns).
// thread = (t = tid / C, c = tid % C) holds rows t + T m of column c0 + c.
// ----------------------------------------------------------------------------------------------
template <int N, int MODE>
__global__ void __launch_bounds__(ColsGeo<N>::kThreads, ColsGeo<N>::kCtasPerSm) cols2_kernel(const StreamParams<float> p,
                                                                                             const uint32_t* __restrict__ mpack) {
    typedef ColsGeo<N> G;
    constexpr int T = Plan<N>::T, C = G::C, NT = G::kThreads, TE = G::kTileElems;
    extern __shared__ __align__(128) unsigned char smem2[];
    cf32* Kslot = reinterpret_cast<cf32*>(smem2);                  // [2][N][C]
    cf32* Gs = Kslot + 2 * TE;                                     // [N][C]
    cf32* TW = Gs + TE;
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(smem2);

    const int tid = threadIdx.x, t = tid / C, c = tid % C;
    constexpr int tiles_per_plane = N / C;
    const int ntiles = p.P * tiles_per_plane;       // P = planes (packed) or images (per-image modes)
    const size_t nn = (size_t)N * N;
    const cf32* in = (MODE == CM_FWD_BLEND_INV) ? p.K : p.cin;
    cf32* out = (MODE == CM_FWD_BLEND_INV) ? p.K : p.cout;
    const cf32* aux = (MODE == CM_FWD_BLEND_INV) ? p.G : p.noise;   // second streamed operand
    const bool aux_batched = (MODE == CM_FWD_BLEND_INV) ? true : (p.noise_batched != 0);
    pdl_wait();
    pdl_launch_dependents();
    const float cf1 = (MODE == CM_FWD_BLEND_INV) ? p.cf[1] : 0.f;      // cf[2] == 2 cf[1]: common.cuh, blend_coef note

    // tile copy: N rows x (C * 8) bytes = N * C / 2 pieces of 16 bytes, 8 per thread
    // `permute`: land the rows in the ColLine order (K tiles); the G / noise tile is read in natural order
    auto issue_tile = [&](const cf32* base, int tile, uint32_t dst, bool batched, bool permute) {
        const int plane = tile / tiles_per_plane, c0 = (tile - plane * tiles_per_plane) * C;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(base + (batched ? (size_t)plane * nn : 0) + c0);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int q = tid + NT * e;
            const int row = q / (C / 2), part = q % (C / 2);
            const int prow = permute ? col_phys_row<C>(row) : row;
            cp_async16(dst + (uint32_t)(prow * C * 8 + part * 16), src + (size_t)row * N * 8 + part * 16);
        }
    };

    if (tid < 128) cp_async16(smem0 + 3 * TE * 8 + tid * 16, reinterpret_cast<const unsigned char*>(g_tw256) + tid * 16);
    // stage-3 twiddles W_N^(i j) (N > 256) into shared memory once per CTA (visible after the first tile barrier)
    cf32* TW3 = TW + 256;
    for (int q = tid; q < (Plan<N>::R3 - 1) * 256; q += NT)
        TW3[q] = ld_tw(reinterpret_cast<const cf32*>(g_tw_f32) + ((q >> 8) + 1) * (q & 255) * (kTwMax / N));
    const Tw3Table tw3{TW3};
    int tile = blockIdx.x;
    if (tile < ntiles) issue_tile(in, tile, smem0, true, true);
    cp_async_commit();
    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        const int plane = tile / tiles_per_plane, c0 = (tile - plane * tiles_per_plane) * C;
        if (MODE != CM_INV) issue_tile(aux, tile, smem0 + 2 * TE * 8, aux_batched, false);
        cp_async_commit();
        const int next = tile + gridDim.x;
        if (next < ntiles) issue_tile(in, next, smem0 + (uint32_t)((s ^ 1) * TE * 8), true, true);
        cp_async_commit();
        uint32_t codes = 0;
        if (MODE == CM_FWD_BLEND_INV) codes = mpack[(p.mcode_batched ? (size_t)plane * (nn / 16) : 0) + (size_t)t * N + c0 + c];

        cp_async_wait<2>();
        __syncthreads();                     // K tile landed (every thread's pieces)
        ColLine<C> ln;
        ln.col = Kslot + s * TE + c;
        cf32 a[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) a[m] = ln.raw(t + T * m);
        __syncthreads();                     // tile fully read before it is reused as exchange scratch
        if (MODE == CM_INV) {
            fft_regs<true, N, true>(a, t, ln, TW, tw3);
        } else {
            fft_regs<false, N, true>(a, t, ln, TW, tw3);
            cp_async_wait<1>();
            __syncthreads();                 // G / noise tile landed; forward-transform scratch reads done
            const cf32* g = Gs + c;
            if (MODE == CM_FWD_BLEND_INV) {
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const cf32 gg = g[(t + T * m) * C];
                    const uint32_t code = (codes >> (2 * m)) & 3u;
                    const float cf = blend_coef<(N != 256)>(code, cf1);        // code in {0, 1, 2}, cf2 == 2 cf1 exactly: see blend_coef note in common.cuh
                    a[m] = mk<float>(gg.re - cf * a[m].re, gg.im - cf * a[m].im);
                }
                fft_regs<true, N, true>(a, t, ln, TW, tw3);
            } else {   // CM_FWD_ACQ: y = fft2(img) * mask + noises                         (S1:99)
                const uint8_t* mk8 = p.mask + (p.mask_batched ? (size_t)plane * nn : 0) + c0 + c;
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const cf32 nz = g[(t + T * m) * C];
                    const float mm = mk8[(size_t)(t + T * m) * N] ? 1.f : 0.f;
                    a[m] = mk<float>(a[m].re * mm + nz.re, a[m].im * mm + nz.im);
                }
            }
        }
        cf32* o = out + (size_t)plane * nn + c0 + c;
#pragma unroll
        for (int m = 0; m < 16; ++m) o[(size_t)(t + T * m) * N] = a[m];
        __syncthreads();                     // all shared-memory reads of this tile done: slots reusable
    }
    cp_async_wait<0>();
}

// ----------------------------------------------------------------------------------------------
// Columns pass of the iteration (col FFT -> residual blend -> col IFFT) with the K and G tiles loaded by 2-D TMA
// (cp.async.bulk.tensor.2d, boxes of C columns x 256 rows from the row-major planes [planes N][N]): the strided 8 C-byte row
// pieces of a tile never go through the LSU.  Same tiling, persistent loop, FFT stages and stores as cols2_kernel.
// ----------------------------------------------------------------------------------------------
PNP_D void tma_load_2d(uint32_t dst, const void* tmap, int x, int y, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(tmap), "r"(x), "r"(y), "r"(bar)
                 : "memory");
}

PNP_D void tma_store_2d(const void* tmap, int x, int y, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap), "r"(x), "r"(y), "r"(src)
                 : "memory");
}

// TMA_STORE: the results also leave through the tile slot and 2-D TMA stores instead of 8-byte stores of the LSU.
template <int N, bool TMA_STORE>
__global__ void __launch_bounds__(ColsGeo<N>::kThreads, ColsGeo<N>::kCtasPerSm)
cols2_tma_kernel(const StreamParams<float> p, const uint32_t* __restrict__ mpack, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmG) {
    typedef ColsGeo<N> G;
    constexpr int T = Plan<N>::T, C = G::C, NT = G::kThreads, TE = G::kTileElems;
    constexpr int kBoxRows = 256, kBoxes = N / kBoxRows;
    extern __shared__ __align__(128) unsigned char smem2[];
    cf32* Kslot = reinterpret_cast<cf32*>(smem2);                  // [2][N][C]
    cf32* Gs = Kslot + 2 * TE;                                     // [N][C]
    cf32* TW = Gs + TE;
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(smem2);
    const uint32_t bar0 = smem0 + G::kSmemBytes;                   // fullK[0], fullK[1], fullG (after the tables)

    const int tid = threadIdx.x, t = tid / C, c = tid % C;
    constexpr int tiles_per_plane = N / C;
    const int ntiles = p.P * tiles_per_plane;
    const size_t nn = (size_t)N * N;

    if (tid == 0) {
        k1::mbar_init(bar0, 1); k1::mbar_init(bar0 + 8, 1); k1::mbar_init(bar0 + 16, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int q = tid; q < 256; q += NT) TW[q] = mk<float>(g_tw256[q].x, g_tw256[q].y);
    cf32* TW3 = TW + 256;
    for (int q = tid; q < (Plan<N>::R3 - 1) * 256; q += NT)
        TW3[q] = ld_tw(reinterpret_cast<const cf32*>(g_tw_f32) + ((q >> 8) + 1) * (q & 255) * (kTwMax / N));
    const Tw3Table tw3{TW3};
    __syncthreads();
    pdl_wait();                    // tables and barriers are set up; the K plane of the previous pass is complete from here on
    pdl_launch_dependents();
    const float cf1 = p.cf[1];                 // cf[2] == 2 cf[1]: common.cuh, blend_coef note

    // one thread: expect the tile's bytes, then one box per 256 rows
    auto issue = [&](const CUtensorMap* tm, int tile, uint32_t dst, uint32_t bar) {
        const int plane = tile / tiles_per_plane, c0 = (tile - plane * tiles_per_plane) * C;
        k1::mbar_arm_tx(bar, TE * 8);
#pragma unroll
        for (int q = 0; q < kBoxes; ++q) tma_load_2d(dst + q * kBoxRows * C * 8, tm, c0, plane * N + q * kBoxRows, bar);
    };

    int tile = blockIdx.x;
    if (tid == 0 && tile < ntiles) issue(&tmK, tile, smem0, bar0);
    for (int it = 0; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it & 1;
        const int plane = tile / tiles_per_plane, c0 = (tile - plane * tiles_per_plane) * C;
        const int next = tile + gridDim.x;
        if (tid == 0) {
            issue(&tmG, tile, smem0 + 2 * TE * 8, bar0 + 16);
            if (next < ntiles) {
                // slot s ^ 1 still feeds the TMA store of the previous tile: wait until it has been read
                if (TMA_STORE) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                issue(&tmK, next, smem0 + (uint32_t)((s ^ 1) * TE * 8), bar0 + 8 * (s ^ 1));
            }
        }
        const uint32_t codes = mpack[(p.mcode_batched ? (size_t)plane * (nn / 16) : 0) + (size_t)t * N + c0 + c];

        k1::mbar_wait(bar0 + 8 * s, (it >> 1) & 1);          // K tile landed (natural row order)
        ColLine<C> ln;
        ln.col = Kslot + s * TE + c;
        cf32 a[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) a[m] = ln.col[(t + T * m) * C];
        __syncthreads();                     // tile fully read before it is reused as exchange scratch
        fft_regs<false, N, true>(a, t, ln, TW, tw3);
        k1::mbar_wait(bar0 + 16, it & 1);                    // G tile landed
        __syncthreads();                     // forward-transform scratch reads done
        const cf32* g = Gs + c;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const cf32 gg = g[(t + T * m) * C];
            const uint32_t code = (codes >> (2 * m)) & 3u;
            const float cf = blend_coef<(N != 256)>(code, cf1);        // code in {0, 1, 2}, cf2 == 2 cf1 exactly: see blend_coef note in common.cuh
            a[m] = mk<float>(gg.re - cf * a[m].re, gg.im - cf * a[m].im);
        }
        fft_regs<true, N, true>(a, t, ln, TW, tw3);
        if (TMA_STORE) {
            __syncthreads();                 // last scratch reads of the inverse transform done
#pragma unroll
            for (int m = 0; m < 16; ++m) ln.col[(t + T * m) * C] = a[m];          // natural row order = the box layout
            k1::fence_proxy_async();         // staged results (and my earlier slot accesses) visible to the async proxy
            __syncthreads();
            if (tid == 0) {
#pragma unroll
                for (int q = 0; q < kBoxes; ++q)
                    tma_store_2d(&tmK, c0, plane * N + q * kBoxRows, smem0 + (uint32_t)(s * TE * 8) + q * kBoxRows * C * 8);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else {
            cf32* o = p.K + (size_t)plane * nn + c0 + c;
#pragma unroll
            for (int m = 0; m < 16; ++m) o[(size_t)(t + T * m) * N] = a[m];
            k1::fence_proxy_async();         // my generic accesses of the slots precede the next TMA writes into them
            __syncthreads();                 // all shared-memory reads of this tile done: slots reusable
        }
    }
    if (TMA_STORE && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores done before the CTA's smem goes away
}

// mcode [N][N] bytes -> packed words [T][N] (one word per column-pass thread and tile)
__global__ void pack_mcode_kernel(const uint8_t* __restrict__ mcode, uint32_t* __restrict__ mpack, int planes, int N) {
    const int T = N / 16;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t per = (size_t)T * N;
    if (i >= per * planes) return;
    const int plane = (int)(i / per);
    const int r = (int)(i - (size_t)plane * per);
    mpack[i] = pack_codes_n(mcode + (size_t)plane * N * N, N, r / N, r % N);
}

}  // namespace s2
}  // namespace pnp
