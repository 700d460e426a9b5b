// streaming.cuh — K2: generic (any power-of-two N, float or double) two-pass-per-iteration
// ADMM kernels that stream the state through L2 / HBM.
//
//   rows pass : row-IFFT of iteration k -> |Re|,|Im| -> prox -> dual -> row-FFT of iteration k+1
//   cols pass : col-FFT -> data-consistency blend -> col-IFFT
//
// Two real images are packed into one complex plane (a + i b).  Because the mask blend of the
// reference (S1:117-118) is followed by Re(ifft2(.)) (S1:119), it can be replaced by the blend
// with the Hermitian-symmetrised data term, which is linear with REAL coefficients and therefore
// acts on the packed plane directly.  With v = z - w (packed V = v_a + i v_b), C = fft2(V),
// g = 1/(1+La2), ms[k] = (m[k]+m[-k])/2 and Ys[k] = (m[k] y[k] + m[-k] conj(y[-k]))/2:
//     Re ifft2(blend(fft2 v)) = v + ifft2_unnormalised(R),   R = G - cf .* C
//     G = g/N^2 * (Ys_a + i Ys_b),   cf = g * ms / N^2  in {0, g/2N^2, g/N^2}
//     x_a = |v_a + Re r|,  x_b = |v_b + Im r|,  r = ifft2_unnormalised(R)
// This RESIDUAL form is algebraically the reference's x-update, but the identity part of the
// blend passes through exactly and FFT round-off only enters scaled by g (~0.03-0.1), which
// buys ~20x lower fp32 error than transforming the full signal both ways (1e-14 vs the reference
// restatement in fp64; fp32 figures in DESIGN.md).
//
// FFT: shared-memory Stockham autosort, radix-4 stages (+ one radix-2 stage when log2 N is odd),
// twiddles from a table computed in double precision.
#pragma once

#include "common.cuh"

namespace pnp {

constexpr int kTwMax = 4096;   // master twiddle table: W_4096^k, serves every N <= 4096

__device__ float2  g_tw_f32[kTwMax];
__device__ double2 g_tw_f64[kTwMax];

template <typename T> struct TwTable;
template <> struct TwTable<float>  { static PNP_D const cx<float>*  get() { return reinterpret_cast<const cx<float>*>(g_tw_f32); } };
template <> struct TwTable<double> { static PNP_D const cx<double>* get() { return reinterpret_cast<const cx<double>*>(g_tw_f64); } };

// ----------------------------------------------------------------------------------------------
// Batched in-smem Stockham FFT over `lines` lines of length N (line pitch `pitch` elements).
// All threads of the CTA participate.  `src` must be fully written and synchronised on entry;
// the returned buffer is synchronised on exit.  tw = W_N^k table (N entries) in shared memory.
// ----------------------------------------------------------------------------------------------
template <bool INV, typename T>
__device__ cx<T>* fft_lines(cx<T>* src, cx<T>* dst, const cx<T>* tw, int lines, int N, int log2N, int pitch) {
    int Ns = 1, logNs = 0;
    if (log2N & 1) {   // one radix-2 stage (Ns = 1: no twiddles)
        const int half = N >> 1, lh = log2N - 1;
        for (int g = threadIdx.x; g < lines * half; g += blockDim.x) {
            const int line = g >> lh, j = g & (half - 1);
            const cx<T>* s = src + line * pitch;
            cx<T>* d = dst + line * pitch;
            cx<T> v0 = s[j], v1 = s[j + half];
            d[2 * j] = v0 + v1;
            d[2 * j + 1] = v0 - v1;
        }
        __syncthreads();
        cx<T>* t = src; src = dst; dst = t;
        Ns = 2; logNs = 1;
    }
    const int q = N >> 2, lq = log2N - 2;
    for (; Ns < N; Ns <<= 2, logNs += 2) {
        const int twstride = N >> (logNs + 2);   // N / (4 Ns)
        for (int g = threadIdx.x; g < lines * q; g += blockDim.x) {
            const int line = g >> lq, j = g & (q - 1);
            const cx<T>* s = src + line * pitch;
            cx<T>* d = dst + line * pitch;
            const int k = j & (Ns - 1);
            cx<T> v0 = s[j], v1 = s[j + q], v2 = s[j + 2 * q], v3 = s[j + 3 * q];
            if (Ns > 1) {
                const int ti = k * twstride;
                v1 = twmul<INV>(v1, tw[ti]);
                v2 = twmul<INV>(v2, tw[2 * ti]);
                v3 = twmul<INV>(v3, tw[3 * ti]);
            }
            cx<T> a0 = v0 + v2, a1 = v0 - v2, a2 = v1 + v3, a3 = rot90<INV>(v1 - v3);
            const int j0 = ((j - k) << 2) + k;
            d[j0] = a0 + a2;
            d[j0 + Ns] = a1 + a3;
            d[j0 + 2 * Ns] = a0 - a2;
            d[j0 + 3 * Ns] = a1 - a3;
        }
        __syncthreads();
        cx<T>* t = src; src = dst; dst = t;
    }
    return src;
}

template <typename T>
__device__ void load_tw(cx<T>* tw_s, int N) {
    const cx<T>* m = TwTable<T>::get();
    const int stride = kTwMax / N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) tw_s[i] = m[i * stride];
}

// ----------------------------------------------------------------------------------------------
template <typename T>
struct StreamParams {
    int N, log2N;
    int B;            // images
    int P;            // packed planes (pairs); solo => P == B
    int solo;         // 1: one image per plane (per-image masks), imaginary slot unused
    int lines;        // lines (rows or columns) per CTA tile
    cx<T>* K;         // [P][N][N] packed work plane
    const cx<T>* G;   // [P][N][N] data term (prepare)
    const uint8_t* mcode;   // [N][N] or [P][N][N]: m[k] + m[-k] in {0,1,2}
    int mcode_batched;
    const T* cf;      // [3] device: g * mcode / (2 N^2) (written by prepare)
    T* x; T* z; T* w; T* xpw;   // [B][N][N] planes (any may be null depending on mode)
    const T* img;     // acquire input
    const cx<T>* cin; // per-image complex input  [B][N][N]
    cx<T>* cout;      // per-image complex output [B][N][N]
    const uint8_t* mask; int mask_batched;
    const cx<T>* noise; int noise_batched;
    T scale;          // 1/N^2 for the zero-filled inverse
    int round_f32;    // acquire: round each FFT pass to float32 (NumPy >= 2 semantics for a float32 image)
    int last;         // rows-prox pass: last iteration (emit x, no forward FFT)
    ProxParams<T> prox;
};

enum RowsMode { RM_FWD_ZW = 0, RM_FWD_IMG = 1, RM_INV_X = 2, RM_INV_ABS = 3, RM_INV_PROX_FWD = 4 };
enum ColsMode { CM_FWD_ACQ = 0, CM_INV = 1, CM_FWD_BLEND_INV = 2 };

extern __shared__ __align__(16) unsigned char pnp_smem_raw[];

// ----------------------------------------------------------------------------------------------
// Rows pass.  grid = (N / lines, planes); a tile is `lines` consecutive rows (contiguous memory).
// ----------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(256) rows_kernel(const StreamParams<T> p) {
    const int N = p.N, lines = p.lines;
    cx<T>* buf0 = reinterpret_cast<cx<T>*>(pnp_smem_raw);
    cx<T>* buf1 = buf0 + lines * N;
    cx<T>* tw = buf1 + lines * N;
    load_tw<T>(tw, N);

    const int plane = blockIdx.y;
    const int r0 = blockIdx.x * lines;
    const size_t plane_off = (size_t)plane * N * N + (size_t)r0 * N;   // element offset of the tile
    const int cnt = lines * N;
    // image indices of a packed plane
    const int ia = p.solo ? plane : 2 * plane;
    const bool has_b = !p.solo && (2 * plane + 1 < p.B);
    const size_t offa = (size_t)ia * N * N + (size_t)r0 * N;
    const size_t offb = offa + (size_t)N * N;

    if (MODE == RM_FWD_ZW) {
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            T re = p.z[offa + i] - p.w[offa + i];
            T im = has_b ? (p.z[offb + i] - p.w[offb + i]) : T(0);
            buf0[i] = mk<T>(re, im);
        }
        __syncthreads();
        cx<T>* res = fft_lines<false, T>(buf0, buf1, tw, lines, N, p.log2N, N);
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) p.K[plane_off + i] = res[i];
    } else if (MODE == RM_FWD_IMG) {
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) buf0[i] = mk<T>(p.img[plane_off + i], T(0));
        __syncthreads();
        cx<T>* res = fft_lines<false, T>(buf0, buf1, tw, lines, N, p.log2N, N);
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            cx<T> v = res[i];
            if (p.round_f32) v = mk<T>((T)(float)v.re, (T)(float)v.im);
            p.cout[plane_off + i] = v;
        }
    } else if (MODE == RM_INV_ABS) {
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) buf0[i] = p.cin[plane_off + i];
        __syncthreads();
        cx<T>* res = fft_lines<true, T>(buf0, buf1, tw, lines, N, p.log2N, N);
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            cx<T> c = res[i];
            p.x[plane_off + i] = psqrt(c.re * c.re + c.im * c.im) * p.scale;
        }
    } else if (MODE == RM_INV_X) {
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) buf0[i] = p.K[plane_off + i];
        __syncthreads();
        cx<T>* res = fft_lines<true, T>(buf0, buf1, tw, lines, N, p.log2N, N);
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            cx<T> c = res[i];
            const T wa = p.w[offa + i];
            T xa = pabs((p.z[offa + i] - wa) + c.re);
            p.x[offa + i] = xa;
            if (p.xpw) p.xpw[offa + i] = xa + wa;
            if (has_b) {
                const T wb = p.w[offb + i];
                T xb = pabs((p.z[offb + i] - wb) + c.im);
                p.x[offb + i] = xb;
                if (p.xpw) p.xpw[offb + i] = xb + wb;
            }
        }
    } else {   // RM_INV_PROX_FWD
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) buf0[i] = p.K[plane_off + i];
        __syncthreads();
        cx<T>* res = fft_lines<true, T>(buf0, buf1, tw, lines, N, p.log2N, N);
        cx<T>* other = (res == buf0) ? buf1 : buf0;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            cx<T> c = res[i];
            T za = p.z[offa + i], wa = p.w[offa + i];
            T xa = pabs((za - wa) + c.re);
            prox_dual(p.prox, xa, za, wa);
            p.z[offa + i] = za; p.w[offa + i] = wa;
            if (p.last) p.x[offa + i] = xa;
            T zb = T(0), wb = T(0);
            if (has_b) {
                zb = p.z[offb + i]; wb = p.w[offb + i];
                T xb = pabs((zb - wb) + c.im);
                prox_dual(p.prox, xb, zb, wb);
                p.z[offb + i] = zb; p.w[offb + i] = wb;
                if (p.last) p.x[offb + i] = xb;
            }
            res[i] = mk<T>(za - wa, zb - wb);
        }
        if (!p.last) {
            __syncthreads();
            cx<T>* f = fft_lines<false, T>(res, other, tw, lines, N, p.log2N, N);
            for (int i = threadIdx.x; i < cnt; i += blockDim.x) p.K[plane_off + i] = f[i];
        }
    }
}

// ----------------------------------------------------------------------------------------------
// Columns pass.  grid = (N / lines, planes); a tile is `lines` adjacent columns, all N rows.
// smem holds the tile transposed: line = column, pitch N + 4 (bank spreading).
// ----------------------------------------------------------------------------------------------
template <typename T, int MODE>
__global__ void __launch_bounds__(256) cols_kernel(const StreamParams<T> p) {
    const int N = p.N, lines = p.lines, pitch = N + 4;
    cx<T>* buf0 = reinterpret_cast<cx<T>*>(pnp_smem_raw);
    cx<T>* buf1 = buf0 + lines * pitch;
    cx<T>* tw = buf1 + lines * pitch;
    load_tw<T>(tw, N);

    const int plane = blockIdx.y;
    const int c0 = blockIdx.x * lines;
    const size_t base = (size_t)plane * N * N + c0;
    const int cnt = lines * N;
    int ll = 0; while ((1 << ll) < lines) ++ll;   // log2(lines)

    const cx<T>* in = (MODE == CM_FWD_BLEND_INV) ? p.K : p.cin;
    cx<T>* out = (MODE == CM_FWD_BLEND_INV) ? p.K : p.cout;

    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const int r = i >> ll, c = i & (lines - 1);
        buf0[c * pitch + r] = in[base + (size_t)r * N + c];
    }
    __syncthreads();
    cx<T>* res;
    if (MODE == CM_INV) {
        res = fft_lines<true, T>(buf0, buf1, tw, lines, N, p.log2N, pitch);
    } else {
        res = fft_lines<false, T>(buf0, buf1, tw, lines, N, p.log2N, pitch);
    }
    if (MODE == CM_FWD_ACQ) {
        // y = fft2(img) * mask + noises      (reference S1:99)
        const uint8_t* m = p.mask + (p.mask_batched ? (size_t)plane * N * N : 0) + c0;
        const cx<T>* nz = p.noise + (p.noise_batched ? (size_t)plane * N * N : 0) + c0;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            const int r = i >> ll, c = i & (lines - 1);
            const size_t g = (size_t)r * N + c;
            cx<T> v = res[c * pitch + r];
            if (p.round_f32) v = mk<T>((T)(float)v.re, (T)(float)v.im);
            const T mm = m[g] ? T(1) : T(0);
            cx<T> n = nz[g];
            out[base + g] = mk<T>(v.re * mm + n.re, v.im * mm + n.im);
        }
    } else if (MODE == CM_INV) {
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            const int r = i >> ll, c = i & (lines - 1);
            out[base + (size_t)r * N + c] = res[c * pitch + r];
        }
    } else {   // CM_FWD_BLEND_INV
        const uint8_t* mc = p.mcode + (p.mcode_batched ? (size_t)plane * N * N : 0) + c0;
        const cx<T>* G = p.G + base;
        const T cf0 = p.cf[0], cf1 = p.cf[1], cf2 = p.cf[2];
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            const int r = i >> ll, c = i & (lines - 1);
            const size_t g = (size_t)r * N + c;
            cx<T> v = res[c * pitch + r];
            const int code = mc[g];
            const T cf = code == 0 ? cf0 : (code == 1 ? cf1 : cf2);
            cx<T> gg = G[g];
            res[c * pitch + r] = mk<T>(gg.re - cf * v.re, gg.im - cf * v.im);
        }
        __syncthreads();
        cx<T>* other = (res == buf0) ? buf1 : buf0;
        cx<T>* f = fft_lines<true, T>(res, other, tw, lines, N, p.log2N, pitch);
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
            const int r = i >> ll, c = i & (lines - 1);
            out[base + (size_t)r * N + c] = f[c * pitch + r];
        }
    }
}

// ----------------------------------------------------------------------------------------------
// prepare: G = g/N^2 (Ys_a + i Ys_b), mcode = m[k] + m[-k] = 2 ms      (see file header)
// one thread per bin of a packed plane.
// ----------------------------------------------------------------------------------------------
template <typename T>
__global__ void prepare_kernel(const cx<T>* __restrict__ y, const uint8_t* __restrict__ mask,
                               cx<T>* __restrict__ G, uint8_t* __restrict__ mcode,
                               int B, int P, int N, int solo, int mask_batched, T g_over_n2,
                               cx<T>* __restrict__ Gt, int tileR) {
    const size_t nn = (size_t)N * N;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nn * P) return;
    const int plane = (int)(idx / nn);
    const size_t bin = idx - (size_t)plane * nn;
    const int r = (int)(bin / N), c = (int)(bin - (size_t)r * N);
    const size_t mbin = (size_t)((N - r) & (N - 1)) * N + ((N - c) & (N - 1));   // -k mod N
    const int ia = solo ? plane : 2 * plane;
    const bool has_b = !solo && (2 * plane + 1 < B);
    const uint8_t* m = mask + (mask_batched ? (size_t)ia * nn : 0);
    const T m1 = m[bin] ? T(1) : T(0), m2 = m[mbin] ? T(1) : T(0);
    const cx<T>* ya = y + (size_t)ia * nn;
    cx<T> a1 = ya[bin], a2 = ya[mbin];
    // Ys_a = (m1*y[k] + m2*conj(y[-k])) / 2
    T sar = T(0.5) * (m1 * a1.re + m2 * a2.re), sai = T(0.5) * (m1 * a1.im - m2 * a2.im);
    T sbr = T(0), sbi = T(0);
    if (has_b) {
        const cx<T>* yb = ya + nn;
        cx<T> b1 = yb[bin], b2 = yb[mbin];
        sbr = T(0.5) * (m1 * b1.re + m2 * b2.re);
        sbi = T(0.5) * (m1 * b1.im - m2 * b2.im);
    }
    // Ys_a + i Ys_b = (sar - sbi) + i (sai + sbr)
    const cx<T> gv = mk<T>(g_over_n2 * (sar - sbi), g_over_n2 * (sai + sbr));
    G[idx] = gv;
    // N = 256 fp32: second copy in the cluster kernel's tile order [plane][rank][kr][c] (cluster256_core.cuh)
    if (Gt) Gt[(size_t)plane * nn + (size_t)(c / tileR) * ((size_t)N * tileR) + (size_t)r * tileR + (c % tileR)] = gv;
    if (!mcode) return;                      // codes already written (fused prologue's preparation launch)
    if (!mask_batched) {
        if (plane == 0) mcode[bin] = (uint8_t)((m[bin] ? 1 : 0) + (m[mbin] ? 1 : 0));
    } else {
        mcode[idx] = (uint8_t)((m[bin] ? 1 : 0) + (m[mbin] ? 1 : 0));
    }
}

// ----------------------------------------------------------------------------------------------
// pointwise PnP pieces
// ----------------------------------------------------------------------------------------------
template <typename T>
__global__ void soft_kernel(const T* __restrict__ x, T* __restrict__ out, T c, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = soft(x[i], c);
}

template <typename T>
__global__ void cnc_combine_kernel(const T* __restrict__ z, const T* __restrict__ x, const T* __restrict__ w,
                                   const T* __restrict__ s, T* __restrict__ t, T one_m_alpha, T alpha, T coef, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        T zz = z[i];
        t[i] = one_m_alpha * zz + alpha * (x[i] + w[i]) + coef * (zz - s[i]);   // S6:301
    }
}

template <typename T>
__global__ void dual_update_kernel(T* __restrict__ x, T* __restrict__ z, T* __restrict__ w, int clamp, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        T xx = x[i], zz = z[i];
        T ww = w[i] + xx - zz;                                                  // S3:293
        if (clamp) { x[i] = clamp01(xx); z[i] = clamp01(zz); ww = clamp01(ww); }  // S3:294-296
        w[i] = ww;
    }
}

template <typename T>
__global__ void u8_to_unit_kernel(const uint8_t* __restrict__ in, T* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (T)((float)in[i] / 255.0f);       // utils_image.uint2single: np.float32(img / 255.)
}

template <typename T>
__global__ void copy_zero_kernel(const T* __restrict__ src, T* __restrict__ dst, T* __restrict__ zero, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (dst) dst[i] = src[i];
        zero[i] = T(0);
    }
}

}  // namespace pnp
