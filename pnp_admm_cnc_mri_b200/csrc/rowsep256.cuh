// rowsep256.cuh — K3: whole ADMM solve for 256x256 images under a ROW-SEPARABLE sampling mask (full k-space lines,
// m[kr][kc] = m(kc): Cartesian undersampling, CS_MRI/Q_Cartesian30), fp32, from the images, in one launch.
//
// The blend coefficient of the x-update (S1:117-118) then depends on kc only and commutes with the column transforms, so
// every image row is an independent 1-D problem for all iterations (derivation in cluster256_core.cuh): no column FFTs in
// the loop, no transposes, no thread-block cluster, half the arithmetic of the general kernel.  A CTA of 256 threads owns
// 16 consecutive rows of one packed plane (two real images) for the whole solve: z rows, the exchange scratch and the rows
// of G' = colIFFT(G) live in shared memory (98 KB, two CTAs per SM), the 16 points and the dual w of a thread in registers;
// half-warps never synchronise with each other.  Per iteration and row: row FFT -> a = G' - N cf(kc) a -> row IFFT ->
// x = |v + r| -> prox (L1 | CNC) -> dual.  HBM traffic of a solve: the image in, x / z / w out, three noise-term rows.
//
// The caller asserts separability (kernel = PNPADMM_KERNEL_ROWSEP); prepare_rowsep_kernel verifies it on the device and a
// mask that violates it makes the solve write NaN into x, z and w instead of a wrong reconstruction.
#pragma once

#include "cluster256_core.cuh"
#include "streaming.cuh"   // g_tw_f32

namespace pnp {
namespace k3 {

using k1::cf32;
using k1::kN;

struct RowSepParams {
    int B, P, iters;
    const float* img; const uint8_t* img8;   // [B][256][256], one of the two
    float* x; float* z; float* w;            // [B][256][256] out
    const cf32* planes;        // [3][256][256] row-major: colIFFT of NcS, nH, nA
    const uint32_t* rcodes;    // [16] packed codes of kc = t + 16 j
    const uint32_t* rhere;     // [16] m[kc] bits
    const int* sep;            // blocks of the preparation launch that saw a mask bin differ from k-space row 0 (0 = separable)
    float ncf1, ncf2;          // N * g / (2 N^2), N * g / N^2
    ProxParams<float> prox;
};

// every thread: the noise terms of one bin, and whether its mask bin equals the bin of k-space row 0 in the same column (`bad`
// counts the blocks that saw a difference; zeroed by the host before the launch); block 0 also writes the 2 x 16 words
__global__ void prepare_rowsep_kernel(const uint8_t* __restrict__ mask, const cf32* __restrict__ noise, float g_over_n2,
                                      cf32* __restrict__ planes, uint32_t* __restrict__ rcodes, uint32_t* __restrict__ rhere,
                                      int* __restrict__ bad) {
    const int bin = blockIdx.x * blockDim.x + threadIdx.x;
    int differs = 0;
    if (bin < kN * kN) {
        k1::rsep_noise_terms(mask, noise, g_over_n2, bin, planes);
        differs = ((mask[bin] != 0) != (mask[bin % kN] != 0)) ? 1 : 0;
    }
    if (__syncthreads_or(differs) && threadIdx.x == 0) atomicAdd(bad, 1);
    if (blockIdx.x == 0 && threadIdx.x < 16) k1::rsep_words(mask, threadIdx.x, rcodes + threadIdx.x, rhere + threadIdx.x);
}

__global__ void __launch_bounds__(256, 2) rowsep256_kernel(const RowSepParams p) {
    typedef k1::Geo<16> G;
    extern __shared__ __align__(128) unsigned char smem[];
    k1::Ctx<16> c;
    c.tid = threadIdx.x;
    c.smem = smem;
    c.rank = 0;
    const size_t nn = (size_t)kN * kN;
    if (*p.sep != 0) {   // the mask is not made of full k-space lines: fail loudly
        const float qnan = __int_as_float(0x7fc00000);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)p.B * nn; i += (size_t)gridDim.x * blockDim.x) {
            p.x[i] = qnan; p.z[i] = qnan; p.w[i] = qnan;
        }
        return;
    }
    k1::fill_tw(reinterpret_cast<cf32*>(smem + G::kOffTW), reinterpret_cast<const cf32*>(g_tw_f32), threadIdx.x);
    __syncthreads();
    const uint32_t codes = p.rcodes[c.rt()], here = p.rhere[c.rt()];
    const int mode = prox_mode(p.prox);
    k1::ThreadState s;
    for (int task = blockIdx.x; task < p.P * 16; task += gridDim.x) {
        const int plane = task >> 4;
        c.rank = task & 15;
        const int ia = 2 * plane;
        const bool has_b = ia + 1 < p.B;
        const float hb = has_b ? 1.f : 0.f;
        k1::PlaneIO io;
        io.z_in_a = io.w_in_a = io.z_in_b = io.w_in_b = nullptr;
        io.x_a = p.x + ia * nn; io.z_a = p.z + ia * nn; io.w_a = p.w + ia * nn; io.xpw_a = nullptr;
        io.x_b = io.x_a + nn; io.z_b = io.z_a + nn; io.w_b = io.w_a + nn; io.xpw_b = nullptr;
        __syncwarp();   // the previous task's last reads of this half-warp's shared-memory rows are done
        // ---- prologue: A0 = row FFT of the image rows; G' row; zero-filled start from the ms / ma branches
        k1::row_load_image(c, s, p.img ? p.img + ia * nn : nullptr, (p.img && has_b) ? p.img + (ia + 1) * nn : nullptr,
                           p.img8 ? p.img8 + ia * nn : nullptr, (p.img8 && has_b) ? p.img8 + (ia + 1) * nn : nullptr);
        k1::row_step1_write<false>(c, s);
        __syncwarp();
        k1::row_read_step2<false>(c, s);
        k1::rsep_acquire_ms(c, s, p.planes, p.planes + nn, codes, p.ncf1, p.ncf2, hb);
        __syncwarp();
        k1::row_step1_write<true>(c, s);
        __syncwarp();
        k1::row_read_step2<true>(c, s);
        k1::row_stash_t1(s);
        k1::rsep_acquire_ma(c, s, p.planes + 2 * nn, codes, here, hb);
        __syncwarp();
        k1::row_step1_write<true>(c, s);
        __syncwarp();
        k1::row_read_step2<true>(c, s);
        k1::row_zero_fill(c, s, 1.0f / (float)(kN * kN), has_b);
        // ---- iterations, all of them on this half-warp's own row
        for (int it = 0; it < p.iters; ++it) {
            __syncwarp();
            k1::row_step1_write<false>(c, s);
            __syncwarp();
            k1::row_read_step2<false>(c, s);
            k1::rsep_blend(c, s, codes, p.ncf1, p.ncf2);
            __syncwarp();
            k1::row_step1_write<true>(c, s);
            __syncwarp();
            k1::row_read_step2<true>(c, s);
            k1::row_prox_dispatch(mode, c, s, p.prox, has_b, it == p.iters - 1, true, io);
        }
    }
}

}  // namespace k3
}  // namespace pnp
