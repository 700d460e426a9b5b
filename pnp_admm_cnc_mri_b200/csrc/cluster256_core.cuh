// cluster256_core.cuh — K1: per-thread phases of the cluster-resident 256x256 ADMM solve.
//
// One thread-block cluster of CL CTAs (CL = 8 or 16) owns one packed plane (two real images
// a + i b) for the whole solve.  With R = 256 / CL, CTA `rank` owns image rows [R rank, R rank + R)
// in the ROW phases and frequency columns [R rank, R rank + R) in the COLUMN phases; the two
// all-to-all transposes per iteration are remote shared-memory stores (DSMEM).  16 R threads x 16
// points = the CTA's whole R x 256 tile lives in registers while it is transformed, so the tile
// buffers double as the exchange scratch.
//   CL = 8 : R = 32, 512 threads, 194 KB smem, one CTA per SM   (15 clusters co-resident on B200)
//   CL = 16: R = 16, 256 threads,  98 KB smem, two CTAs per SM  (14 clusters co-resident): the two
//            CTAs of an SM belong to different planes, so one computes while the other transposes.
//
//   shared memory per CTA (fp32):
//     Zs [R][256] float2   z of image a / b interleaved
//     B1 [R][256] float2   row-layout tile    (written remotely by the column phase)
//     B2 [256][R] float2   column-layout tile (written remotely by the row phase)
//     TW [16][16] float2   W_256^(k1*n2)                                           2 KB
//   B1 is idle during the column phase until the peers' transposes land, so the data term G of
//   the coming blend is staged there (bulk async copy issued at the end of the row phase).
//   registers per thread: 16 complex points + the dual w of its 16 pixels x 2 images (32 floats,
//   resident for the whole solve).
//
// 256-point FFT = 16 x 16 Cooley-Tukey: thread t of a 16-thread group holds x[t + 16 j];
//   step 1: 16-point DFT over j  -> b[k1];  b[k1] *= W_256^(t k1)
//   exchange (16x16 transpose inside the group through shared memory)
//   step 2: 16-point DFT over n2 -> X[t + 16 k2]          (same "t + 16 j" layout as the input)
// Row phases: group = half-warp (exchange guarded by __syncwarp, XOR-swizzled slots).
// Column phases: thread = (t = tid / R, c = tid % R), lanes run along columns (exchange guarded by
// __syncthreads, conflict-free because lanes are always contiguous), so every remote store is a
// contiguous 8 R byte segment.
//
// All functions are HOST+DEVICE: tests/host_emu/ runs the same code for 8 x 512 emulated threads.
#pragma once

#include "common.cuh"

namespace pnp {
namespace k1 {

typedef cx<float> cf32;

constexpr int kN = 256;

// mbarrier slots
enum { BAR_FULL1 = 0, BAR_FULL2 = 1, BAR_GFULL = 2, BAR_FREE1 = 3, BAR_FREE2 = 4 };

// Geometry of a cluster of CL CTAs.
template <int CL>
struct Geo {
    static constexpr int kCluster = CL;
    static constexpr int kRows = kN / CL;              // rows (columns) per CTA
    static constexpr int kThreads = 16 * kRows;
    static constexpr int kWarps = kThreads / 32;
    static constexpr int kTileBytes = kRows * kN * 8;  // one tile = one transpose = one G stage
    static constexpr int kOffZs = 0;
    static constexpr int kOffB1 = kTileBytes;
    static constexpr int kOffB2 = 2 * kTileBytes;
    static constexpr int kOffTW = 3 * kTileBytes;
    static constexpr int kOffBar = kOffTW + 256 * 8;   // 5 mbarriers (device kernel only)
    static constexpr int kSmemBytes = kOffBar + 64;
    static constexpr int kCtasPerSm = (CL == 16) ? 2 : 1;
    // element k = t + 16 j (t < 16) of a 256-line lives in CTA k / kRows at local index k % kRows
    static PNP_HD int dest(int j) { return kRows == 32 ? (j >> 1) : j; }
    static PNP_HD int local(int t, int j) { return kRows == 32 ? t + 16 * (j & 1) : t; }
};

struct ThreadState {
    cf32 a[16];      // working points
    float w[32];     // dual variable: w[2j] image a, w[2j+1] image b, pixel column t + 16 j
};

template <int CL>
struct Ctx {
    typedef Geo<CL> G;
    int rank;                 // CTA rank in the cluster
    int tid;                  // thread index in the CTA
    unsigned char* smem;      // this CTA's dynamic shared memory
    PNP_HD cf32* Zs() const { return reinterpret_cast<cf32*>(smem + G::kOffZs); }
    PNP_HD cf32* B1() const { return reinterpret_cast<cf32*>(smem + G::kOffB1); }
    PNP_HD cf32* B2() const { return reinterpret_cast<cf32*>(smem + G::kOffB2); }
    PNP_HD const cf32* TW() const { return reinterpret_cast<const cf32*>(smem + G::kOffTW); }
    // row-phase mapping: half-warp = one row
    PNP_HD int row() const { return (tid >> 5) * 2 + ((tid >> 4) & 1); }
    PNP_HD int rt() const { return tid & 15; }
    // column-phase mapping: t = tid / R, column = tid % R
    PNP_HD int ct() const { return tid / G::kRows; }
    PNP_HD int cc() const { return tid % G::kRows; }
};

// ---------------------------------------------------------------------------------------------
// 16-point DFT, natural order in / natural order out, radix 4 x 4, fully unrolled in registers.
// ---------------------------------------------------------------------------------------------
#define PNP_C8  0.92387953251128673848f   /* cos(pi/8) */
#define PNP_S8  0.38268343236508978178f   /* sin(pi/8) */
#define PNP_R2  0.70710678118654752440f   /* sqrt(1/2) */

template <bool INV> PNP_HD cf32 mul_w16_1(cf32 a) { return twmul<INV>(a, mk<float>(PNP_C8, -PNP_S8)); }
template <bool INV> PNP_HD cf32 mul_w16_3(cf32 a) { return twmul<INV>(a, mk<float>(PNP_S8, -PNP_C8)); }
template <bool INV> PNP_HD cf32 mul_w16_9(cf32 a) { return twmul<INV>(a, mk<float>(-PNP_C8, PNP_S8)); }
// W16^2 = r(1 - i), W16^6 = -r(1 + i)   (conjugated for the inverse)
template <bool INV> PNP_HD cf32 mul_w16_2(cf32 a) {
    return INV ? mk<float>(PNP_R2 * (a.re - a.im), PNP_R2 * (a.re + a.im))
               : mk<float>(PNP_R2 * (a.re + a.im), PNP_R2 * (a.im - a.re));
}
template <bool INV> PNP_HD cf32 mul_w16_6(cf32 a) {
    return INV ? mk<float>(-PNP_R2 * (a.re + a.im), PNP_R2 * (a.re - a.im))
               : mk<float>(PNP_R2 * (a.im - a.re), -PNP_R2 * (a.re + a.im));
}

template <bool INV>
PNP_HD void dft4(cf32 v0, cf32 v1, cf32 v2, cf32 v3, cf32& y0, cf32& y1, cf32& y2, cf32& y3) {
    cf32 a0 = v0 + v2, a1 = v0 - v2, a2 = v1 + v3, a3 = rot90<INV>(v1 - v3);
    y0 = a0 + a2; y1 = a1 + a3; y2 = a0 - a2; y3 = a1 - a3;
}

template <bool INV>
PNP_HD void fft16(const cf32 (&x)[16], cf32 (&X)[16]) {
    cf32 t[16];   // t[4*k1 + n2]
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2)
        dft4<INV>(x[n2], x[4 + n2], x[8 + n2], x[12 + n2], t[n2], t[4 + n2], t[8 + n2], t[12 + n2]);
    // twiddles W16^(n2*k1)
    t[5] = mul_w16_1<INV>(t[5]);   t[6] = mul_w16_2<INV>(t[6]);    t[7] = mul_w16_3<INV>(t[7]);
    t[9] = mul_w16_2<INV>(t[9]);   t[10] = rot90<INV>(t[10]);      t[11] = mul_w16_6<INV>(t[11]);
    t[13] = mul_w16_3<INV>(t[13]); t[14] = mul_w16_6<INV>(t[14]);  t[15] = mul_w16_9<INV>(t[15]);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
        dft4<INV>(t[4 * k1], t[4 * k1 + 1], t[4 * k1 + 2], t[4 * k1 + 3], X[k1], X[k1 + 4], X[k1 + 8], X[k1 + 12]);
}

// step 1 of the 256-point transform for group-thread t: 16-pt DFT over j, then W_256^(t*k1)
template <bool INV>
PNP_HD void fft256_step1(cf32 (&a)[16], int t, const cf32* TW) {
    cf32 b[16];
    fft16<INV>(a, b);
    a[0] = b[0];
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) a[k1] = twmul<INV>(b[k1], TW[k1 * 16 + t]);
}

template <bool INV>
PNP_HD void fft256_step2(cf32 (&c)[16], cf32 (&out)[16]) { fft16<INV>(c, out); }

// ---------------------------------------------------------------------------------------------
// ROW phases (thread = (row, t); holds columns n = t + 16 j of image row R*rank + row)
// ---------------------------------------------------------------------------------------------
struct PlaneIO {            // global-memory planes of the two images of a packed plane
    const float* z_in_a; const float* w_in_a; const float* z_in_b; const float* w_in_b;   // b may be null
    float* x_a; float* z_a; float* w_a; float* xpw_a;
    float* x_b; float* z_b; float* w_b; float* xpw_b;
};

// prologue: global z, w -> Zs / registers; a = z - w
template <int CL>
PNP_HD void row_load_state(const Ctx<CL>& c, ThreadState& s, const PlaneIO& io) {
    const int row = c.row(), t = c.rt();
    const int g0 = (Geo<CL>::kRows * c.rank + row) * kN;
    cf32* Zs = c.Zs() + row * kN;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int n = t + 16 * j;
        const float za = ld_state(io.z_in_a + g0 + n), wa = ld_state(io.w_in_a + g0 + n);
        float zb = 0.f, wb = 0.f;
        if (io.z_in_b) { zb = ld_state(io.z_in_b + g0 + n); wb = ld_state(io.w_in_b + g0 + n); }
        Zs[n] = mk<float>(za, zb);
        s.w[2 * j] = wa; s.w[2 * j + 1] = wb;
        s.a[j] = mk<float>(za - wa, zb - wb);
    }
}

template <int CL>
PNP_HD void row_load(const Ctx<CL>& c, ThreadState& s) {
    const cf32* src = c.B1() + c.row() * kN + c.rt();
#pragma unroll
    for (int j = 0; j < 16; ++j) s.a[j] = src[16 * j];
}

// step 1 + scatter into the exchange scratch (= this row's slot of B1), XOR swizzle
template <bool INV, int CL>
PNP_HD void row_step1_write(const Ctx<CL>& c, ThreadState& s) {
    const int t = c.rt();
    fft256_step1<INV>(s.a, t, c.TW());
    cf32* sc = c.B1() + c.row() * kN;
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) sc[k1 * 16 + (t ^ k1)] = s.a[k1];
}

template <bool INV, int CL>
PNP_HD void row_read_step2(const Ctx<CL>& c, ThreadState& s) {
    const int t = c.rt();
    const cf32* sc = c.B1() + c.row() * kN + t * 16;
    cf32 v[16];
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = sc[n2 ^ t];
    fft256_step2<INV>(v, s.a);
}

// after the inverse row FFT (a = r, the residual correction): x = |v + r|; prox; dual; next a = z - w.
// MODE: PM_L1 / PM_CNC / PM_GENERAL, or PROX_NONE (x-update only: emit x and x + w, state untouched).
// `last`: end of this cluster's run on the plane -> z, w go to global memory (x too if want_x).
template <int MODE, int CL>
PNP_HD void row_prox(const Ctx<CL>& c, ThreadState& s, const ProxParams<float>& p, bool has_b, bool last, bool want_x,
                     const PlaneIO& io) {
    const int row = c.row(), t = c.rt();
    const int g0 = (Geo<CL>::kRows * c.rank + row) * kN;
    cf32* Zs = c.Zs() + row * kN;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int n = t + 16 * j;
        cf32 zz = Zs[n];
        float za = zz.re, zb = zz.im, wa = s.w[2 * j], wb = s.w[2 * j + 1];
        // residual form: x = |v + r|, v = z - w the very value that was transformed
        const float xa = pabs((za - wa) + s.a[j].re);
        const float xb = has_b ? pabs((zb - wb) + s.a[j].im) : 0.f;
        if (MODE == PROX_NONE) {
            io.x_a[g0 + n] = xa;
            if (io.xpw_a) io.xpw_a[g0 + n] = xa + wa;
            if (has_b) {
                io.x_b[g0 + n] = xb;
                if (io.xpw_b) io.xpw_b[g0 + n] = xb + wb;
            }
        } else {
            prox_dual_m<MODE>(p, xa, za, wa);
            if (has_b) prox_dual_m<MODE>(p, xb, zb, wb);
            s.w[2 * j] = wa; s.w[2 * j + 1] = wb;
            if (last) {
                io.z_a[g0 + n] = za; io.w_a[g0 + n] = wa;
                if (want_x) io.x_a[g0 + n] = xa;
                if (has_b) {
                    io.z_b[g0 + n] = zb; io.w_b[g0 + n] = wb;
                    if (want_x) io.x_b[g0 + n] = xb;
                }
            } else {
                Zs[n] = mk<float>(za, zb);
                s.a[j] = mk<float>(za - wa, zb - wb);
            }
        }
    }
}

template <int CL>
PNP_HD void row_prox_dispatch(int mode, const Ctx<CL>& c, ThreadState& s, const ProxParams<float>& p, bool has_b, bool last,
                              bool want_x, const PlaneIO& io) {
    switch (mode) {
        case PM_L1: row_prox<PM_L1>(c, s, p, has_b, last, want_x, io); break;
        case PM_CNC: row_prox<PM_CNC>(c, s, p, has_b, last, want_x, io); break;
        case PROX_NONE: row_prox<PROX_NONE>(c, s, p, has_b, last, want_x, io); break;
        default: row_prox<PM_GENERAL>(c, s, p, has_b, last, want_x, io); break;
    }
}

// forward row FFT output X[row][k = t + 16 j] -> CTA (k / R), B2[R*rank + row][k % R]
template <int CL, class Remote>
PNP_HD void row_store_remote(const Ctx<CL>& c, const ThreadState& s, const Remote& R) {
    typedef Geo<CL> G;
    const int t = c.rt();
    const int grow = G::kRows * c.rank + c.row();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int off = G::kOffB2 + (grow * G::kRows + G::local(t, j)) * 8;
        R.st(G::dest(j), off, s.a[j], BAR_FULL2);
    }
}

// ---------------------------------------------------------------------------------------------
// COLUMN phases (thread = (t, c); holds rows r = t + 16 j of column R*rank + c)
// ---------------------------------------------------------------------------------------------
template <int CL>
PNP_HD void col_load(const Ctx<CL>& c, ThreadState& s) {
    constexpr int R = Geo<CL>::kRows;
    const cf32* src = c.B2() + c.ct() * R + c.cc();
#pragma unroll
    for (int j = 0; j < 16; ++j) s.a[j] = src[16 * j * R];
}

template <bool INV, int CL>
PNP_HD void col_step1_write(const Ctx<CL>& c, ThreadState& s) {
    constexpr int R = Geo<CL>::kRows;
    const int t = c.ct();
    fft256_step1<INV>(s.a, t, c.TW());
    cf32* sc = c.B2() + t * R + c.cc();
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) sc[k1 * 16 * R] = s.a[k1];
}

template <bool INV, int CL>
PNP_HD void col_read_step2(const Ctx<CL>& c, ThreadState& s) {
    constexpr int R = Geo<CL>::kRows;
    const cf32* sc = c.B2() + (c.ct() * 16) * R + c.cc();
    cf32 v[16];
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = sc[n2 * R];
    fft256_step2<INV>(v, s.a);
}

// data-consistency residual on the packed spectrum: a = G - cf[code] * a      (see streaming.cuh)
// Gs: this CTA's G tile staged in shared memory as [kr = 256][c = R] (the B1 buffer);
// codes: 2 bits per j (mcode of bin (t + 16 j, kc)), packed by pack_mcode_k1.
template <int CL>
PNP_HD void col_blend(const Ctx<CL>& c, ThreadState& s, const cf32* Gs, uint32_t codes, float cf1, float cf2) {
    constexpr int R = Geo<CL>::kRows;
    const cf32* g = Gs + c.ct() * R + c.cc();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const cf32 gg = g[16 * j * R];
        const uint32_t code = (codes >> (2 * j)) & 3u;
        const float cf = blend_coef<true>(code, cf1);        // code in {0, 1, 2}, cf2 == 2 cf1 exactly: see blend_coef note in common.cuh
        s.a[j] = mk<float>(gg.re - cf * s.a[j].re, gg.im - cf * s.a[j].im);
    }
}

// global G[kr][kc] rows -> staged tile [kr][R]: byte offsets of row kr for CTA `rank`
template <int CL> PNP_HD int g_stage_src_off(int rank, int kr) { return (kr * kN + Geo<CL>::kRows * rank) * 8; }
template <int CL> PNP_HD int g_stage_dst_off(int kr) { return Geo<CL>::kOffB1 + kr * Geo<CL>::kRows * 8; }
// K1-tiled copy of G written by prepare: [plane][rank][kr][c], kc = R rank + c, so the tile a CTA stages is
// one contiguous kTileBytes block and a warp's slice of it one contiguous 4 KB piece (element offset in the plane).
PNP_HD int g_tiled_elem(int R, int kr, int kc) { return (kc / R) * (kN * R) + kr * R + (kc % R); }

// packed mask codes: word (t, kc) holds mcode[(t + 16 j) * 256 + kc] in bits 2j, 2j+1
PNP_HD uint32_t pack_codes(const uint8_t* mcode, int t, int kc) {
    uint32_t v = 0;
    for (int j = 0; j < 16; ++j) v |= (uint32_t)(mcode[(t + 16 * j) * kN + kc] & 3u) << (2 * j);
    return v;
}

// inverse column FFT output at image row r = t + 16 j, column kc -> CTA (r / R), B1[r % R][kc]
template <int CL, class Remote>
PNP_HD void col_store_remote(const Ctx<CL>& c, const ThreadState& s, const Remote& R) {
    typedef Geo<CL> G;
    const int t = c.ct();
    const int kc = G::kRows * c.rank + c.cc();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int off = G::kOffB1 + (G::local(t, j) * kN + kc) * 8;
        R.st(G::dest(j), off, s.a[j], BAR_FULL1);
    }
}

// ---------------------------------------------------------------------------------------------
// Blocked-tile variant (CL = 16 only): both transposes leave a CTA as 16 contiguous 2 KB blocks, one per peer, so they
// can be sent with the TMA engine (cp.async.bulk shared::cta -> shared::cluster) instead of 4096 8-byte remote stores
// through the LSU.  Every tile is addressed as [peer 16][256 entries]:
//   B2  [src rank s][row r][col c]   = element (image row 16 s + r, column-of-this-CTA c)   (same bytes as [256][16])
//   B1  [src rank s][row r][col c]   = element (row-of-this-CTA r, frequency column 16 s + c)
//   outgoing staging (in B1 at the end of a row phase, in B2 at the end of a column phase): [dest j][tid]
// With tid = 16 row + t (row phases) = 16 t + c (column phases) every access below is `base + 256 j + tid`: a warp
// touches 256 contiguous bytes.  A warp's row-FFT scratch is its own 16 input chunks (256 B in each peer block).
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// Fused prologue: acquisition (S1:99), zero-filled start (S1:100-105) and the data term of the blend, inside the solve.
// With F = fft2(a + i b) of the two real images of a plane, m' the mirrored mask, ms = (m + m') / 2, ma = (m - m') / 2
// and the noise split into its Hermitian / anti-Hermitian parts nH = (n[k] + conj n[-k]) / 2, nA = (n[k] - conj n[-k]) / 2:
//     G  = cf .* F + (1 + i) NcS,  NcS = g / N^2 * (m n[k] + m' conj n[-k]) / 2        (the term prepare_kernel builds from y)
//     T1 = ifft2_unnormalised(ms .* F + (1 + i) nH) = (Re ifft2 y_a) + i (Re ifft2 y_b)
//     T2 = ifft2_unnormalised(ma .* F + (1 + i) nA) = -(Im ifft2 y_b) + i (Im ifft2 y_a)
//     x0_a = |Re T1 + i Im T2| / N^2,   x0_b = |Im T1 - i Re T2| / N^2                  (= |ifft2(y)|, y never materialised)
// because a real image times a symmetric (antisymmetric) real mask has a real (imaginary) inverse transform.  NcS, nH, nA
// depend on (mask, noise, reo) only and come from prepare_shared_kernel in this kernel's tile order.  A plane with one
// image (odd batch) uses (1 + 0 i) instead of (1 + i): its imaginary slot stays exactly zero.
// Phases: R0 image rows -> row FFT -> transpose;  C0 col FFT = F, save F (in the idle Zs tile), write G, col IFFT of the
// ms branch -> transpose;  R1 row IFFT = T1, stashed in the dual's registers;  C1 reload F, col IFFT of the ma branch ->
// transpose;  R2 row IFFT = T2, z = x0, w = 0, then the first forward row FFT of the loop.
// ---------------------------------------------------------------------------------------------
PNP_HD float unit_from_u8(uint8_t v) { return (float)v / 255.0f; }   // utils_image.uint2single: np.float32(img / 255.)

template <int CL>
PNP_HD void row_load_image(const Ctx<CL>& c, ThreadState& s, const float* fa, const float* fb, const uint8_t* ua, const uint8_t* ub) {
    const int row = c.row(), t = c.rt();
    const int g0 = (Geo<CL>::kRows * c.rank + row) * kN;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int n = g0 + t + 16 * j;
        const float a = fa ? fa[n] : unit_from_u8(ua[n]);
        const float b = fb ? fb[n] : (ub ? unit_from_u8(ub[n]) : 0.f);
        s.a[j] = mk<float>(a, b);
    }
}

// tile-order index of this thread's j-th point in a column phase (G / noise-term tiles, F save area)
template <int CL> PNP_HD int col_tile_index(const Ctx<CL>& c, int j) { return c.ct() * Geo<CL>::kRows + c.cc() + 16 * j * Geo<CL>::kRows; }

// C0: s.a holds F.  hb = 1 if the plane has a second image, else 0.
template <int CL>
PNP_HD void col_acquire_ms(const Ctx<CL>& c, ThreadState& s, cf32* Fsave, cf32* Gtile, const cf32* NcS, const cf32* nH,
                           uint32_t codes, float cf1, float cf2, float hb) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int i = col_tile_index(c, j);
        const cf32 F = s.a[j];
        Fsave[i] = F;
        const uint32_t code = (codes >> (2 * j)) & 3u;
        const float cf = blend_coef<true>(code, cf1);        // code in {0, 1, 2}, cf2 == 2 cf1 exactly: see blend_coef note in common.cuh
        const cf32 nc = NcS[i], nh = nH[i];
        Gtile[i] = mk<float>(cf * F.re + (nc.re - hb * nc.im), cf * F.im + (nc.im + hb * nc.re));
        const float ms = 0.5f * (float)code;
        s.a[j] = mk<float>(ms * F.re + (nh.re - hb * nh.im), ms * F.im + (nh.im + hb * nh.re));
    }
}

// C1: reload F; ma = m[k] - (m[k] + m[-k]) / 2;  `here` bit j = m[k] of the thread's j-th bin
template <int CL>
PNP_HD void col_acquire_ma(const Ctx<CL>& c, ThreadState& s, const cf32* Fsave, const cf32* nA, uint32_t codes, uint32_t here, float hb) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int i = col_tile_index(c, j);
        const cf32 F = Fsave[i];
        const uint32_t code = (codes >> (2 * j)) & 3u;
        const float ma = (float)((here >> j) & 1u) - 0.5f * (float)code;
        const cf32 na = nA[i];
        s.a[j] = mk<float>(ma * F.re + (na.re - hb * na.im), ma * F.im + (na.im + hb * na.re));
    }
}

// R1: T1 -> the dual's registers (w is zero until the loop starts)
PNP_HD void row_stash_t1(ThreadState& s) {
#pragma unroll
    for (int j = 0; j < 16; ++j) { s.w[2 * j] = s.a[j].re; s.w[2 * j + 1] = s.a[j].im; }
}

// R2: s.a = T2.  z = x0 = |ifft2(y)|, w = 0, a = z - w
template <int CL>
PNP_HD void row_zero_fill(const Ctx<CL>& c, ThreadState& s, float inv_n2, bool has_b) {
    cf32* Zs = c.Zs() + c.row() * kN;
    const int t = c.rt();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float t1r = s.w[2 * j], t1i = s.w[2 * j + 1];
        const float xa = psqrt(t1r * t1r + s.a[j].im * s.a[j].im) * inv_n2;
        const float xb = has_b ? psqrt(t1i * t1i + s.a[j].re * s.a[j].re) * inv_n2 : 0.f;
        const cf32 z = mk<float>(xa, xb);
        Zs[t + 16 * j] = z;
        s.w[2 * j] = 0.f; s.w[2 * j + 1] = 0.f;
        s.a[j] = z;
    }
}

// (mask, noise, reo) -> the three noise-term tiles, packed codes and the m[k] bits, one thread per (t, kc) word.
// tiles: [3][rank][kr][c] (NcS, nH, nA);  mpack / mhere: [16][256] words.
PNP_HD void prepare_shared_word(const uint8_t* mask, const cf32* noise, int R, float g_over_n2, int t, int kc, cf32* tiles,
                                uint32_t* mpack, uint32_t* mhere, uint8_t* mcode) {
    uint32_t codes = 0, here = 0;
    const int mkc = (kN - kc) & (kN - 1);
    for (int j = 0; j < 16; ++j) {
        const int kr = t + 16 * j, mkr = (kN - kr) & (kN - 1);
        const int bin = kr * kN + kc, mbin = mkr * kN + mkc;
        const float m1 = mask[bin] ? 1.f : 0.f, m2 = mask[mbin] ? 1.f : 0.f;
        const cf32 n1 = noise[bin], n2 = noise[mbin];
        const uint32_t code = (uint32_t)(m1 + m2);
        codes |= code << (2 * j);
        here |= (uint32_t)m1 << j;
        if (mcode) mcode[bin] = (uint8_t)code;
        const int i = g_tiled_elem(R, kr, kc);
        tiles[i] = mk<float>(g_over_n2 * 0.5f * (m1 * n1.re + m2 * n2.re), g_over_n2 * 0.5f * (m1 * n1.im - m2 * n2.im));
        tiles[kN * kN + i] = mk<float>(0.5f * (n1.re + n2.re), 0.5f * (n1.im - n2.im));
        tiles[2 * kN * kN + i] = mk<float>(0.5f * (n1.re - n2.re), 0.5f * (n1.im + n2.im));
    }
    mpack[t * kN + kc] = codes;
    mhere[t * kN + kc] = here;
}

// ---------------------------------------------------------------------------------------------
// Row-separable masks (K3, rowsep256.cuh): when the sampling pattern consists of full k-space lines, m[kr][kc] = m(kc)
// (CS_MRI/Q_Cartesian30: 76 full columns), the blend coefficient depends on kc only and therefore commutes with the
// column transforms:   ifft2(G - cf .* fft2 V) = rowIFFT( colIFFT(G) - N cf(kc) .* rowFFT(V) ).
// Every image row then is an independent 1-D problem for the whole solve: no column FFTs, no transposes, no cluster.  The
// same holds for the fused prologue (ms, ma depend on kc only and colIFFT(colFFT(.)) = N):
//     G'  = colIFFT(G)  = N cf(kc) A0 + (1 + i hb) NcS',    A0 = rowFFT(image row pair),   X' = colIFFT_unnormalised(X)
//     T1  = rowIFFT(N ms(kc) A0 + (1 + i hb) nH'),  T2 = rowIFFT(N ma(kc) A0 + (1 + i hb) nA'),  x0 as in the fused prologue / N^2
// A CTA of 256 threads owns 16 consecutive rows of a plane and reuses the row-phase code above with the Geo<16> shared
// memory layout: Zs = z rows, B1 = exchange scratch, B2 = the rows of G' ([16][256], resident for the whole solve).
// Thread (row, t) handles kc = t + 16 j: its 16 codes / m[k] bits are one word each (rcodes[t], rhere[t]).
// ---------------------------------------------------------------------------------------------
// prologue after A0 = rowFFT(image rows): keep A0 (in Zs), write G' (B2), leave the ms branch in s.a
PNP_HD void rsep_acquire_ms(const Ctx<16>& c, ThreadState& s, const cf32* NcSp, const cf32* nHp, uint32_t codes, float ncf1, float ncf2,
                            float hb) {
    const int row = c.row(), t = c.rt();
    cf32* A0 = c.Zs() + row * kN;
    cf32* Gp = c.B2() + row * kN;
    const int g0 = (16 * c.rank + row) * kN;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int n = t + 16 * j;
        const cf32 F = s.a[j];
        A0[n] = F;
        const uint32_t code = (codes >> (2 * j)) & 3u;
        const float cf = blend_coef<true>(code, ncf1);       // code in {0, 1, 2}, ncf2 == 2 ncf1 exactly: see blend_coef note in common.cuh
        const cf32 nc = NcSp[g0 + n], nh = nHp[g0 + n];
        Gp[n] = mk<float>(cf * F.re + (nc.re - hb * nc.im), cf * F.im + (nc.im + hb * nc.re));
        const float ms = (0.5f * (float)kN) * (float)code;
        s.a[j] = mk<float>(ms * F.re + (nh.re - hb * nh.im), ms * F.im + (nh.im + hb * nh.re));
    }
}

PNP_HD void rsep_acquire_ma(const Ctx<16>& c, ThreadState& s, const cf32* nAp, uint32_t codes, uint32_t here, float hb) {
    const int row = c.row(), t = c.rt();
    const cf32* A0 = c.Zs() + row * kN;
    const int g0 = (16 * c.rank + row) * kN;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int n = t + 16 * j;
        const cf32 F = A0[n];
        const uint32_t code = (codes >> (2 * j)) & 3u;
        const float ma = (float)kN * ((float)((here >> j) & 1u) - 0.5f * (float)code);
        const cf32 na = nAp[g0 + n];
        s.a[j] = mk<float>(ma * F.re + (na.re - hb * na.im), ma * F.im + (na.im + hb * na.re));
    }
}

// the residual blend of a row: a = G' - N cf(kc) a
PNP_HD void rsep_blend(const Ctx<16>& c, ThreadState& s, uint32_t codes, float ncf1, float ncf2) {
    const cf32* Gp = c.B2() + c.row() * kN + c.rt();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const cf32 gg = Gp[16 * j];
        const uint32_t code = (codes >> (2 * j)) & 3u;
        const float cf = blend_coef<true>(code, ncf1);       // code in {0, 1, 2}, ncf2 == 2 ncf1 exactly: see blend_coef note in common.cuh
        s.a[j] = mk<float>(gg.re - cf * s.a[j].re, gg.im - cf * s.a[j].im);
    }
}

// Preparation for K3, thread kc of 256: is the mask the same on every k-space row?  codes / m[k] words of the 16 row-phase
// thread classes, and the noise-term planes in ROW-major order [3][256][256] (input of the column inverse transform).
PNP_HD bool rsep_column_is_constant(const uint8_t* mask, int kc) {
    const bool m0 = mask[kc] != 0;
    for (int kr = 1; kr < kN; ++kr)
        if ((mask[kr * kN + kc] != 0) != m0) return false;
    return true;
}
PNP_HD void rsep_words(const uint8_t* mask, int t, uint32_t* codes, uint32_t* here) {
    uint32_t cw = 0, hw = 0;
    for (int j = 0; j < 16; ++j) {
        const int kc = t + 16 * j, mkc = (kN - kc) & (kN - 1);
        const uint32_t m1 = mask[kc] ? 1u : 0u, m2 = mask[mkc] ? 1u : 0u;     // row 0 speaks for every row (checked separately)
        cw |= (m1 + m2) << (2 * j);
        hw |= m1 << j;
    }
    *codes = cw; *here = hw;
}
PNP_HD void rsep_noise_terms(const uint8_t* mask, const cf32* noise, float g_over_n2, int bin, cf32* planes) {
    const int kr = bin / kN, kc = bin % kN;
    const int mbin = ((kN - kr) & (kN - 1)) * kN + ((kN - kc) & (kN - 1));
    const float m1 = mask[bin] ? 1.f : 0.f, m2 = mask[mbin] ? 1.f : 0.f;
    const cf32 n1 = noise[bin], n2 = noise[mbin];
    planes[bin] = mk<float>(g_over_n2 * 0.5f * (m1 * n1.re + m2 * n2.re), g_over_n2 * 0.5f * (m1 * n1.im - m2 * n2.im));
    planes[kN * kN + bin] = mk<float>(0.5f * (n1.re + n2.re), 0.5f * (n1.im - n2.im));
    planes[2 * kN * kN + bin] = mk<float>(0.5f * (n1.re - n2.re), 0.5f * (n1.im + n2.im));
}

#if defined(__CUDA_ARCH__)
PNP_D cf32 ld_g(const cf32* p) { const float2 v = __ldg(reinterpret_cast<const float2*>(p)); return mk<float>(v.x, v.y); }
#else
inline cf32 ld_g(const cf32* p) { return *p; }
#endif

PNP_HD void row_load_bk(const Ctx<16>& c, ThreadState& s) {
    const cf32* src = c.B1() + c.tid;
#pragma unroll
    for (int j = 0; j < 16; ++j) s.a[j] = src[256 * j];
}

// scratch slot of row-local index k (0..255) for the half-warp of `tid`: chunk k / 16, entry (tid & ~15) + k % 16
template <bool INV>
PNP_HD void row_step1_write_bk(const Ctx<16>& c, ThreadState& s) {
    const int t = c.rt();
    fft256_step1<INV>(s.a, t, c.TW());
    cf32* sc = c.B1() + (c.tid & ~15);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) sc[k1 * 256 + (t ^ k1)] = s.a[k1];
}

template <bool INV>
PNP_HD void row_read_step2_bk(const Ctx<16>& c, ThreadState& s) {
    const int t = c.rt();
    const cf32* sc = c.B1() + t * 256 + (c.tid & ~15);
    cf32 v[16];
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = sc[n2 ^ t];
    fft256_step2<INV>(v, s.a);
}

// forward row FFT output X[row][k = t + 16 j] -> staging block of peer j (lands in its B2 at block `rank`)
PNP_HD void row_stage_bk(const Ctx<16>& c, const ThreadState& s) {
    cf32* st = c.B1() + c.tid;
#pragma unroll
    for (int j = 0; j < 16; ++j) st[256 * j] = s.a[j];
}

// inverse column FFT output at image row t + 16 j -> staging block of peer j (lands in its B1 at block `rank`)
PNP_HD void col_stage_bk(const Ctx<16>& c, const ThreadState& s) {
    cf32* st = c.B2() + c.tid;
#pragma unroll
    for (int j = 0; j < 16; ++j) st[256 * j] = s.a[j];
}

// the blend with the data term read straight from the tile-ordered global copy Gt (this CTA's 32 KB block)
PNP_HD void col_blend_g(const Ctx<16>& c, ThreadState& s, const cf32* Gtile, uint32_t codes, float cf1, float cf2) {
    const cf32* g = Gtile + c.tid;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const cf32 gg = ld_g(g + 256 * j);
        const uint32_t code = (codes >> (2 * j)) & 3u;
        const float cf = blend_coef<true>(code, cf1);        // code in {0, 1, 2}, cf2 == 2 cf1 exactly: see blend_coef note in common.cuh
        s.a[j] = mk<float>(gg.re - cf * s.a[j].re, gg.im - cf * s.a[j].im);
    }
}

// TW[k1*16 + n2] = W_256^(k1*n2) from the master table W_4096^m
PNP_HD void fill_tw(cf32* TW, const cf32* master4096, int i) {
    const int k1 = i >> 4, n2 = i & 15;
    TW[i] = master4096[(k1 * n2 * 16) & 4095];
}

}  // namespace k1
}  // namespace pnp
