// CPU emulation of the K1 cluster kernel (TEST INFRASTRUCTURE).
// Runs the SAME per-thread phase code as cluster256.cuh (cluster256_core.cuh is host+device) for
// 8 emulated CTAs x 512 threads, with every barrier turned into a phase boundary.  It validates the
// FFT decomposition, swizzles, thread<->pixel mappings and DSMEM offsets without a GPU; it cannot
// detect races (compute-sanitizer on the GPU box does that).
#include <cmath>
#include <cstring>
#include <vector>

#include "../../pnp_admm_cnc_mri_b200/csrc/cluster256_core.cuh"

using namespace pnp;
using namespace pnp::k1;

namespace {
struct RemoteHost {
    unsigned char** all;
    void st(int rank, int off, cf32 v, int /*bar*/) const { *reinterpret_cast<cf32*>(all[rank] + off) = v; }
};
}  // namespace

extern "C" int k1_emulate(const float* z_in, const float* w_in, float* x, float* z, float* w, float* xpw,
                          const float* G_, const uint8_t* mcode_, int mcode_batched, float /*cf0*/, float cf1, float cf2,
                          int B, int P, int solo, int iters, int prox, float thr_l1, float inv_b, float one_m_alpha,
                          float alpha, float coef, float thr_cnc) {
    std::vector<cf32> master(4096);
    for (int i = 0; i < 4096; ++i) {
        const double a = -2.0 * M_PI * i / 4096.0;
        master[i] = mk<float>((float)std::cos(a), (float)std::sin(a));
    }
    std::vector<std::vector<unsigned char>> smem(kCluster, std::vector<unsigned char>(kSmemBytes, 0));
    unsigned char* all[kCluster];
    for (int r = 0; r < kCluster; ++r) all[r] = smem[r].data();
    RemoteHost R{all};
    std::vector<ThreadState> st((size_t)kCluster * kThreads);
    ProxParams<float> pp;
    pp.prox = prox; pp.general = 0; pp.thr_l1 = thr_l1; pp.inv_b = inv_b; pp.one_m_alpha = one_m_alpha;
    pp.alpha = alpha; pp.coef = coef; pp.thr_cnc = thr_cnc;

    for (int r = 0; r < kCluster; ++r)
        for (int i = 0; i < 256; ++i) fill_tw(reinterpret_cast<cf32*>(all[r] + kOffTW), master.data(), i);

#define FOR_ALL(body)                                                   \
    for (int r = 0; r < kCluster; ++r)                                  \
        for (int t = 0; t < kThreads; ++t) {                            \
            Ctx c; c.rank = r; c.tid = t; c.smem = all[r];              \
            ThreadState& s = st[(size_t)r * kThreads + t];              \
            body;                                                       \
        }

    const size_t nn = (size_t)kN * kN;
    for (int plane = 0; plane < P; ++plane) {
        const int ia = solo ? plane : 2 * plane;
        const bool has_b = !solo && (2 * plane + 1 < B);
        PlaneIO io;
        io.z_in_a = z_in + ia * nn; io.w_in_a = w_in + ia * nn;
        io.z_in_b = has_b ? io.z_in_a + nn : nullptr; io.w_in_b = has_b ? io.w_in_a + nn : nullptr;
        io.x_a = x + ia * nn; io.z_a = z ? z + ia * nn : nullptr; io.w_a = w ? w + ia * nn : nullptr;
        io.xpw_a = xpw ? xpw + ia * nn : nullptr;
        io.x_b = io.x_a + nn; io.z_b = io.z_a ? io.z_a + nn : nullptr; io.w_b = io.w_a ? io.w_a + nn : nullptr;
        io.xpw_b = io.xpw_a ? io.xpw_a + nn : nullptr;
        const cf32* G = reinterpret_cast<const cf32*>(G_) + plane * nn;
        const uint8_t* mcode = mcode_ + (mcode_batched ? plane * nn : 0);

        // the bulk-async G prefetch into B1 of the device kernel, as a plain copy
        auto stage_g = [&]() {
            for (int r = 0; r < kCluster; ++r)
                for (int kr = 0; kr < kN; ++kr)
                    std::memcpy(all[r] + g_stage_dst_off(kr), reinterpret_cast<const unsigned char*>(G) + g_stage_src_off(r, kr),
                                kRows * 8);
        };
        FOR_ALL(row_load_state(c, s, io));
        FOR_ALL(row_step1_write<false>(c, s));
        FOR_ALL(row_read_step2<false>(c, s));
        stage_g();
        FOR_ALL(row_store_remote(c, s, R));
        for (int it = 0; it < iters; ++it) {
            FOR_ALL(col_load(c, s));
            FOR_ALL(col_step1_write<false>(c, s));
            FOR_ALL(col_read_step2<false>(c, s);
                    col_blend(c, s, c.B1(), pack_codes(mcode, c.ct(), kRows * c.rank + c.cc()), cf1, cf2));
            FOR_ALL(col_step1_write<true>(c, s));
            FOR_ALL(col_read_step2<true>(c, s));
            FOR_ALL(col_store_remote(c, s, R));
            const bool last = (it == iters - 1);
            FOR_ALL(row_load(c, s));
            FOR_ALL(row_step1_write<true>(c, s));
            FOR_ALL(row_read_step2<true>(c, s); row_prox_dispatch(prox == PROX_NONE ? PROX_NONE : prox_mode(pp), c, s, pp, has_b, last, true, io));
            if (!last) {
                FOR_ALL(row_step1_write<false>(c, s));
                FOR_ALL(row_read_step2<false>(c, s));
                stage_g();
                FOR_ALL(row_store_remote(c, s, R));
            }
        }
    }
    return 0;
}

// 256-point FFT of one line through the same step1 / exchange / step2 code (unit test hook)
extern "C" void k1_fft256_line(const float* in, float* out, int inverse) {
    std::vector<cf32> master(4096), TW(256);
    for (int i = 0; i < 4096; ++i) {
        const double a = -2.0 * M_PI * i / 4096.0;
        master[i] = mk<float>((float)std::cos(a), (float)std::sin(a));
    }
    for (int i = 0; i < 256; ++i) fill_tw(TW.data(), master.data(), i);
    cf32 a[16][16], sc[256];
    for (int t = 0; t < 16; ++t)
        for (int j = 0; j < 16; ++j) a[t][j] = mk<float>(in[2 * (t + 16 * j)], in[2 * (t + 16 * j) + 1]);
    for (int t = 0; t < 16; ++t) {
        if (inverse) fft256_step1<true>(a[t], t, TW.data()); else fft256_step1<false>(a[t], t, TW.data());
        for (int k1 = 0; k1 < 16; ++k1) sc[k1 * 16 + t] = a[t][k1];
    }
    for (int t = 0; t < 16; ++t) {
        cf32 v[16], o[16];
        for (int n2 = 0; n2 < 16; ++n2) v[n2] = sc[t * 16 + n2];
        if (inverse) fft256_step2<true>(v, o); else fft256_step2<false>(v, o);
        for (int k2 = 0; k2 < 16; ++k2) { out[2 * (t + 16 * k2)] = o[k2].re; out[2 * (t + 16 * k2) + 1] = o[k2].im; }
    }
}
