// metrics.cuh — per-image PSNR / SSIM / RE of a batch of reconstructions on the device, so that a
// batched run never copies images back just to score them.
//
// Follows utils/utils_image.py of the reference, evaluated in double precision like NumPy:
//   calculate_psnr :543-556   20 log10(255 / sqrt(mean((E - H)^2)))
//   calculate_ssim :570-615   11x11 Gaussian (sigma 1.5) window, "valid" region, C1 = 6.5025, C2 = 58.5225
//   calculate_re   :622-636   ||H - E||_2 / ||H||_2
// with E = 255 x (S1:133, S4:139; float image) or E = uint8(round(255 clip(x, 0, 1))) (S6:315, S6:531)
// and H the uint8 ground truth.  border = 0 as in every reference call site.
//
// One CTA scores a 32 x 32 block of SSIM outputs (42 x 42 input pixels) with a separable filter in
// shared memory and also owns the 32 x 32 image pixels at the same origin for the PSNR / RE sums;
// per-image sums are accumulated with double atomics and turned into the three scores by
// metrics_finalize_kernel.
#pragma once

#include "common.cuh"

namespace pnp {

constexpr int kMetTile = 32, kMetHalo = 10, kMetIn = kMetTile + kMetHalo;   // 42

struct MetricsAcc { double se, hh, ssim_sum, pad; };   // sum (E-H)^2, sum H^2, sum of the SSIM map
constexpr int kMetSmemBytes = (2 * kMetIn * (kMetIn + 1) + 5 * kMetIn * (kMetTile + 1) + 12 + 24) * 8;

template <typename T>
__global__ void __launch_bounds__(256) metrics_tile_kernel(const T* __restrict__ x, const uint8_t* __restrict__ ref,
                                                           int N, int quantize, MetricsAcc* __restrict__ acc) {
    extern __shared__ __align__(16) unsigned char met_smem[];
    typedef double RowIn[kMetIn + 1];
    typedef double RowH[kMetTile + 1];
    RowIn* sa = reinterpret_cast<RowIn*>(met_smem);                     // [42][43]
    RowIn* sb = sa + kMetIn;                                            // [42][43]
    RowH (*h)[kMetIn] = reinterpret_cast<RowH(*)[kMetIn]>(sb + kMetIn); // [5][42][33]
    double* gk = reinterpret_cast<double*>(h + 5);                      // [11]
    double (*red)[8] = reinterpret_cast<double(*)[8]>(gk + 12);         // [3][8]
    const int img = blockIdx.z, i0 = blockIdx.y * kMetTile, j0 = blockIdx.x * kMetTile;
    const T* xi = x + (size_t)img * N * N;
    const uint8_t* ri = ref + (size_t)img * N * N;
    const int tid = threadIdx.x;
    if (tid < 11) {   // cv2.getGaussianKernel(11, 1.5): exp(-(i-5)^2 / (2 sigma^2)) normalised
        double s = 0.0;
        for (int k = 0; k < 11; ++k) s += exp(-((k - 5.0) * (k - 5.0)) / 4.5);
        gk[tid] = exp(-((tid - 5.0) * (tid - 5.0)) / 4.5) / s;
    }
    double se = 0.0, hh = 0.0;
    for (int q = tid; q < kMetIn * kMetIn; q += 256) {
        const int r = q / kMetIn, c = q - r * kMetIn;
        const int gi = i0 + r, gj = j0 + c;
        double a = 0.0, b = 0.0;
        if (gi < N && gj < N) {
            const T v = xi[(size_t)gi * N + gj];
            if (quantize) a = (double)rintf(fminf(fmaxf((float)v, 0.f), 1.f) * 255.f);   // util.single2uint
            else a = 255.0 * (double)v;                                                    // img_E = x * 255
            b = (double)ri[(size_t)gi * N + gj];
            if (r < kMetTile && c < kMetTile) { se += (a - b) * (a - b); hh += b * b; }
        }
        sa[r][c] = a; sb[r][c] = b;
    }
    __syncthreads();
    // horizontal pass: 42 rows x 32 columns x 5 maps
    for (int q = tid; q < kMetIn * kMetTile; q += 256) {
        const int r = q / kMetTile, c = q - r * kMetTile;
        double m1 = 0, m2 = 0, s11 = 0, s22 = 0, s12 = 0;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const double a = sa[r][c + k], b = sb[r][c + k], g = gk[k];
            m1 += g * a; m2 += g * b; s11 += g * a * a; s22 += g * b * b; s12 += g * a * b;
        }
        h[0][r][c] = m1; h[1][r][c] = m2; h[2][r][c] = s11; h[3][r][c] = s22; h[4][r][c] = s12;
    }
    __syncthreads();
    const double C1 = (0.01 * 255) * (0.01 * 255), C2 = (0.03 * 255) * (0.03 * 255);
    const int V = N - kMetHalo;   // valid outputs per dimension
    double ss = 0.0;
    for (int q = tid; q < kMetTile * kMetTile; q += 256) {
        const int r = q / kMetTile, c = q - r * kMetTile;
        if (i0 + r < V && j0 + c < V) {
            double m1 = 0, m2 = 0, s11 = 0, s22 = 0, s12 = 0;
#pragma unroll
            for (int k = 0; k < 11; ++k) {
                const double g = gk[k];
                m1 += g * h[0][r + k][c]; m2 += g * h[1][r + k][c];
                s11 += g * h[2][r + k][c]; s22 += g * h[3][r + k][c]; s12 += g * h[4][r + k][c];
            }
            const double m11 = m1 * m1, m22 = m2 * m2, m12 = m1 * m2;
            ss += ((2 * m12 + C1) * (2 * (s12 - m12) + C2)) / ((m11 + m22 + C1) * ((s11 - m11) + (s22 - m22) + C2));
        }
    }
    // block reduction -> one atomic per quantity
    for (int o = 16; o > 0; o >>= 1) {
        se += __shfl_down_sync(0xffffffffu, se, o);
        hh += __shfl_down_sync(0xffffffffu, hh, o);
        ss += __shfl_down_sync(0xffffffffu, ss, o);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = se; red[1][tid >> 5] = hh; red[2][tid >> 5] = ss; }
    __syncthreads();
    if (tid < 3) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[tid][w];
        double* dst = tid == 0 ? &acc[img].se : (tid == 1 ? &acc[img].hh : &acc[img].ssim_sum);
        atomicAdd(dst, s);
    }
}

// acc[B] -> out[B][3] = (psnr, ssim, re)
__global__ void metrics_finalize_kernel(const MetricsAcc* __restrict__ acc, double* __restrict__ out, int B, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    const double n2 = (double)N * N, v = (double)(N - kMetHalo);
    const double mse = acc[i].se / n2;
    out[3 * i + 0] = mse == 0.0 ? INFINITY : 20.0 * log10(255.0 / sqrt(mse));
    out[3 * i + 1] = acc[i].ssim_sum / (v * v);
    out[3 * i + 2] = sqrt(acc[i].se) / sqrt(acc[i].hh);
}

}  // namespace pnp
