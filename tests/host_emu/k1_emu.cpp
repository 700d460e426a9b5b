// CPU emulation of the K1 cluster kernel (TEST INFRASTRUCTURE).
// Runs the SAME per-thread phase code as cluster256.cuh (cluster256_core.cuh is host+device) for
// CL emulated CTAs x 16*256/CL threads (CL = 8 or 16), with every barrier turned into a phase boundary.
// It validates the FFT decomposition, swizzles, thread<->pixel mappings and DSMEM offsets without a
// GPU; it cannot detect races (compute-sanitizer on the GPU box does that).
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

std::vector<cf32> master_table() {
    std::vector<cf32> m(4096);
    for (int i = 0; i < 4096; ++i) {
        const double a = -2.0 * M_PI * i / 4096.0;
        m[i] = mk<float>((float)std::cos(a), (float)std::sin(a));
    }
    return m;
}

template <int CL>
int emulate(const float* z_in, const float* w_in, float* x, float* z, float* w, float* xpw, const float* G_,
            const uint8_t* mcode_, int mcode_batched, float cf1, float cf2, int B, int P, int solo, int iters,
            const ProxParams<float>& pp) {
    typedef Geo<CL> G;
    std::vector<cf32> master = master_table();
    std::vector<std::vector<unsigned char>> smem(CL, std::vector<unsigned char>(G::kSmemBytes, 0));
    unsigned char* all[CL];
    for (int r = 0; r < CL; ++r) all[r] = smem[r].data();
    RemoteHost R{all};
    std::vector<ThreadState> st((size_t)CL * G::kThreads);
    for (int r = 0; r < CL; ++r)
        for (int i = 0; i < 256; ++i) fill_tw(reinterpret_cast<cf32*>(all[r] + G::kOffTW), master.data(), i);
    const int mode = pp.prox == PROX_NONE ? PROX_NONE : prox_mode(pp);

#define FOR_ALL(body)                                                   \
    for (int r = 0; r < CL; ++r)                                        \
        for (int t = 0; t < G::kThreads; ++t) {                         \
            Ctx<CL> c; c.rank = r; c.tid = t; c.smem = all[r];          \
            ThreadState& s = st[(size_t)r * G::kThreads + t];           \
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
        const cf32* Gp = reinterpret_cast<const cf32*>(G_) + plane * nn;
        const uint8_t* mcode = mcode_ + (mcode_batched ? plane * nn : 0);
        // the bulk-async G prefetch into B1 of the device kernel, as a plain copy
        auto stage_g = [&]() {
            for (int r = 0; r < CL; ++r)
                for (int kr = 0; kr < kN; ++kr)
                    std::memcpy(all[r] + g_stage_dst_off<CL>(kr),
                                reinterpret_cast<const unsigned char*>(Gp) + g_stage_src_off<CL>(r, kr), G::kRows * 8);
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
                    col_blend(c, s, c.B1(), pack_codes(mcode, c.ct(), G::kRows * c.rank + c.cc()), cf1, cf2));
            FOR_ALL(col_step1_write<true>(c, s));
            FOR_ALL(col_read_step2<true>(c, s));
            FOR_ALL(col_store_remote(c, s, R));
            const bool last = (it == iters - 1);
            FOR_ALL(row_load(c, s));
            FOR_ALL(row_step1_write<true>(c, s));
            FOR_ALL(row_read_step2<true>(c, s); row_prox_dispatch(mode, c, s, pp, has_b, last, true, io));
            if (!last) {
                FOR_ALL(row_step1_write<false>(c, s));
                FOR_ALL(row_read_step2<false>(c, s));
                stage_g();
                FOR_ALL(row_store_remote(c, s, R));
            }
        }
    }
#undef FOR_ALL
    return 0;
}

// Blocked-tile variant (cluster256_bk_kernel): staging + 16 bulk copies of 2 KB per CTA and transpose, G read from
// the tile-ordered global copy.
int emulate_bk(const float* z_in, const float* w_in, float* x, float* z, float* w, float* xpw, const float* G_,
               const uint8_t* mcode_, int mcode_batched, float cf1, float cf2, int B, int P, int solo, int iters,
               const ProxParams<float>& pp) {
    constexpr int CL = 16;
    typedef Geo<CL> G;
    std::vector<cf32> master = master_table();
    std::vector<std::vector<unsigned char>> smem(CL, std::vector<unsigned char>(G::kSmemBytes, 0));
    unsigned char* all[CL];
    for (int r = 0; r < CL; ++r) all[r] = smem[r].data();
    std::vector<ThreadState> st((size_t)CL * G::kThreads);
    for (int r = 0; r < CL; ++r)
        for (int i = 0; i < 256; ++i) fill_tw(reinterpret_cast<cf32*>(all[r] + G::kOffTW), master.data(), i);
    const int mode = pp.prox == PROX_NONE ? PROX_NONE : prox_mode(pp);
#define FOR_ALL(body)                                                   \
    for (int r = 0; r < CL; ++r)                                        \
        for (int t = 0; t < G::kThreads; ++t) {                         \
            Ctx<CL> c; c.rank = r; c.tid = t; c.smem = all[r];          \
            ThreadState& s = st[(size_t)r * G::kThreads + t];           \
            body;                                                       \
        }
    // send_tile: block j of CTA r's staging buffer -> block r of CTA j's destination tile
    auto send = [&](int off_src, int off_dst) {
        std::vector<std::vector<unsigned char>> snap(CL);
        for (int r = 0; r < CL; ++r) snap[r].assign(all[r] + off_src, all[r] + off_src + G::kTileBytes);
        for (int r = 0; r < CL; ++r)
            for (int j = 0; j < CL; ++j) std::memcpy(all[j] + off_dst + 2048 * r, snap[r].data() + 2048 * j, 2048);
    };
    const size_t nn = (size_t)kN * kN;
    std::vector<cf32> Gt(nn);
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
        const cf32* Gp = reinterpret_cast<const cf32*>(G_) + plane * nn;
        for (int kr = 0; kr < kN; ++kr)                 // what prepare_kernel writes for the cluster kernel
            for (int kc = 0; kc < kN; ++kc) Gt[g_tiled_elem(G::kRows, kr, kc)] = Gp[kr * kN + kc];
        const uint8_t* mcode = mcode_ + (mcode_batched ? plane * nn : 0);
        FOR_ALL(row_load_state(c, s, io));
        FOR_ALL(row_step1_write_bk<false>(c, s));
        FOR_ALL(row_read_step2_bk<false>(c, s));
        FOR_ALL(row_stage_bk(c, s));
        send(G::kOffB1, G::kOffB2);
        for (int it = 0; it < iters; ++it) {
            FOR_ALL(col_load(c, s));
            FOR_ALL(col_step1_write<false>(c, s));
            FOR_ALL(col_read_step2<false>(c, s);
                    col_blend_g(c, s, Gt.data() + (size_t)c.rank * (G::kTileBytes / 8),
                                pack_codes(mcode, c.ct(), G::kRows * c.rank + c.cc()), cf1, cf2));
            FOR_ALL(col_step1_write<true>(c, s));
            FOR_ALL(col_read_step2<true>(c, s));
            FOR_ALL(col_stage_bk(c, s));
            send(G::kOffB2, G::kOffB1);
            const bool last = (it == iters - 1);
            FOR_ALL(row_load_bk(c, s));
            FOR_ALL(row_step1_write_bk<true>(c, s));
            FOR_ALL(row_read_step2_bk<true>(c, s); row_prox_dispatch(mode, c, s, pp, has_b, last, true, io));
            if (!last) {
                FOR_ALL(row_step1_write_bk<false>(c, s));
                FOR_ALL(row_read_step2_bk<false>(c, s));
                FOR_ALL(row_stage_bk(c, s));
                send(G::kOffB1, G::kOffB2);
            }
        }
    }
#undef FOR_ALL
    return 0;
}
}  // namespace

extern "C" int k1_emulate(const float* z_in, const float* w_in, float* x, float* z, float* w, float* xpw,
                          const float* G_, const uint8_t* mcode_, int mcode_batched, float /*cf0*/, float cf1, float cf2,
                          int B, int P, int solo, int iters, int prox, float thr_l1, float inv_b, float one_m_alpha,
                          float alpha, float coef, float thr_cnc, int cluster) {
    ProxParams<float> pp;
    pp.prox = prox; pp.general = 0; pp.thr_l1 = thr_l1; pp.inv_b = inv_b; pp.one_m_alpha = one_m_alpha;
    pp.alpha = alpha; pp.coef = coef; pp.thr_cnc = thr_cnc;
    if (cluster == 116)   // 16-CTA geometry, blocked tiles + bulk-copy transposes
        return emulate_bk(z_in, w_in, x, z, w, xpw, G_, mcode_, mcode_batched, cf1, cf2, B, P, solo, iters, pp);
    if (cluster == 16)
        return emulate<16>(z_in, w_in, x, z, w, xpw, G_, mcode_, mcode_batched, cf1, cf2, B, P, solo, iters, pp);
    return emulate<8>(z_in, w_in, x, z, w, xpw, G_, mcode_, mcode_batched, cf1, cf2, B, P, solo, iters, pp);
}

// Fused prologue + loop (cluster256_kernel with ClusterParams::fused): images, mask, noise in; x, z, w out.
template <int CL>
int emulate_fused(const float* img, const uint8_t* img8, const uint8_t* mask, const float* noise_, float reo, float* x, float* z,
                  float* w, int B, int iters, const ProxParams<float>& pp) {
    typedef Geo<CL> G;
    const size_t nn = (size_t)kN * kN;
    const int P = (B + 1) / 2;
    const cf32* noise = reinterpret_cast<const cf32*>(noise_);
    std::vector<cf32> master = master_table();
    std::vector<std::vector<unsigned char>> smem(CL, std::vector<unsigned char>(G::kSmemBytes, 0));
    unsigned char* all[CL];
    for (int r = 0; r < CL; ++r) all[r] = smem[r].data();
    RemoteHost R{all};
    std::vector<ThreadState> st((size_t)CL * G::kThreads);
    for (int r = 0; r < CL; ++r)
        for (int i = 0; i < 256; ++i) fill_tw(reinterpret_cast<cf32*>(all[r] + G::kOffTW), master.data(), i);
    const int mode = prox_mode(pp);
    const double La2 = 1.0 / 2.0 / reo, g = 1.0 / (1.0 + La2);
    const float cf1 = (float)(0.5 * g / nn), cf2 = (float)(g / nn);
    std::vector<cf32> tiles(3 * nn), Gt(nn);
    std::vector<uint32_t> mpack(16 * kN), mhere(16 * kN);
    for (int i = 0; i < 16 * kN; ++i)
        prepare_shared_word(mask, noise, G::kRows, (float)(g / nn), i / kN, i % kN, tiles.data(), mpack.data(), mhere.data(), nullptr);
#define FOR_ALL(body)                                                   \
    for (int r = 0; r < CL; ++r)                                        \
        for (int t = 0; t < G::kThreads; ++t) {                         \
            Ctx<CL> c; c.rank = r; c.tid = t; c.smem = all[r];          \
            ThreadState& s = st[(size_t)r * G::kThreads + t];           \
            body;                                                       \
        }
    const int tile = G::kTileBytes / 8;
    for (int plane = 0; plane < P; ++plane) {
        const int ia = 2 * plane;
        const bool has_b = 2 * plane + 1 < B;
        const float hb = has_b ? 1.f : 0.f;
        PlaneIO io;
        io.z_in_a = io.w_in_a = io.z_in_b = io.w_in_b = nullptr;
        io.x_a = x + ia * nn; io.z_a = z + ia * nn; io.w_a = w + ia * nn; io.xpw_a = nullptr;
        io.x_b = io.x_a + nn; io.z_b = io.z_a + nn; io.w_b = io.w_a + nn; io.xpw_b = nullptr;
        auto stage_g = [&]() {   // bulk copy of this CTA's G tile (tile order: one contiguous block) into B1
            for (int r = 0; r < CL; ++r) std::memcpy(all[r] + G::kOffB1, Gt.data() + (size_t)r * tile, G::kTileBytes);
        };
#define WORD(c) ((c).ct() * kN + G::kRows * (c).rank + (c).cc())
        FOR_ALL(row_load_image(c, s, img ? img + ia * nn : nullptr, (img && has_b) ? img + (ia + 1) * nn : nullptr,
                               img8 ? img8 + ia * nn : nullptr, (img8 && has_b) ? img8 + (ia + 1) * nn : nullptr));
        FOR_ALL(row_step1_write<false>(c, s));
        FOR_ALL(row_read_step2<false>(c, s));
        FOR_ALL(row_store_remote(c, s, R));
        FOR_ALL(col_load(c, s));
        FOR_ALL(col_step1_write<false>(c, s));
        FOR_ALL(col_read_step2<false>(c, s);
                col_acquire_ms(c, s, c.Zs(), Gt.data() + (size_t)c.rank * tile, tiles.data() + (size_t)c.rank * tile,
                               tiles.data() + nn + (size_t)c.rank * tile, mpack[WORD(c)], cf1, cf2, hb));
        FOR_ALL(col_step1_write<true>(c, s));
        FOR_ALL(col_read_step2<true>(c, s));
        FOR_ALL(col_store_remote(c, s, R));
        FOR_ALL(row_load(c, s));
        FOR_ALL(row_step1_write<true>(c, s));
        FOR_ALL(row_read_step2<true>(c, s); row_stash_t1(s));
        FOR_ALL(col_acquire_ma(c, s, c.Zs(), tiles.data() + 2 * nn + (size_t)c.rank * tile, mpack[WORD(c)], mhere[WORD(c)], hb));
        FOR_ALL(col_step1_write<true>(c, s));
        FOR_ALL(col_read_step2<true>(c, s));
        FOR_ALL(col_store_remote(c, s, R));
        FOR_ALL(row_load(c, s));
        FOR_ALL(row_step1_write<true>(c, s));
        FOR_ALL(row_read_step2<true>(c, s));
        FOR_ALL(row_zero_fill(c, s, 1.0f / (float)(kN * kN), has_b));
        FOR_ALL(row_step1_write<false>(c, s));
        FOR_ALL(row_read_step2<false>(c, s));
        stage_g();
        FOR_ALL(row_store_remote(c, s, R));
        for (int it = 0; it < iters; ++it) {
            FOR_ALL(col_load(c, s));
            FOR_ALL(col_step1_write<false>(c, s));
            FOR_ALL(col_read_step2<false>(c, s); col_blend(c, s, c.B1(), mpack[WORD(c)], cf1, cf2));
            FOR_ALL(col_step1_write<true>(c, s));
            FOR_ALL(col_read_step2<true>(c, s));
            FOR_ALL(col_store_remote(c, s, R));
            const bool last = (it == iters - 1);
            FOR_ALL(row_load(c, s));
            FOR_ALL(row_step1_write<true>(c, s));
            FOR_ALL(row_read_step2<true>(c, s); row_prox_dispatch(mode, c, s, pp, has_b, last, true, io));
            if (!last) {
                FOR_ALL(row_step1_write<false>(c, s));
                FOR_ALL(row_read_step2<false>(c, s));
                stage_g();
                FOR_ALL(row_store_remote(c, s, R));
            }
        }
#undef WORD
    }
#undef FOR_ALL
    return 0;
}

extern "C" int k1_emulate_fused(const float* img, const uint8_t* img8, const uint8_t* mask, const float* noise, float reo, float* x,
                                float* z, float* w, int B, int iters, int prox, float thr_l1, float inv_b, float one_m_alpha,
                                float alpha, float coef, float thr_cnc, int cluster) {
    ProxParams<float> pp;
    pp.prox = prox; pp.general = 0; pp.thr_l1 = thr_l1; pp.inv_b = inv_b; pp.one_m_alpha = one_m_alpha;
    pp.alpha = alpha; pp.coef = coef; pp.thr_cnc = thr_cnc;
    if (cluster == 16) return emulate_fused<16>(img, img8, mask, noise, reo, x, z, w, B, iters, pp);
    return emulate_fused<8>(img, img8, mask, noise, reo, x, z, w, B, iters, pp);
}

// K3 (rowsep256.cuh): row-separable mask, every row solved on its own.  `planes` = column inverse transforms of the three
// noise-term planes (the test computes them with NumPy from rsep_noise_terms' output, see k3_noise_terms).
extern "C" void k3_noise_terms(const uint8_t* mask, const float* noise, float g_over_n2, float* planes) {
    for (int bin = 0; bin < kN * kN; ++bin)
        rsep_noise_terms(mask, reinterpret_cast<const cf32*>(noise), g_over_n2, bin, reinterpret_cast<cf32*>(planes));
}
extern "C" int k3_emulate(const float* img, const uint8_t* img8, const uint8_t* mask, const float* planes_, float reo, float* x, float* z,
                          float* w, int B, int iters, int prox, float thr_l1, float inv_b, float one_m_alpha, float alpha, float coef,
                          float thr_cnc) {
    typedef Geo<16> G;
    ProxParams<float> pp;
    pp.prox = prox; pp.general = 0; pp.thr_l1 = thr_l1; pp.inv_b = inv_b; pp.one_m_alpha = one_m_alpha;
    pp.alpha = alpha; pp.coef = coef; pp.thr_cnc = thr_cnc;
    int n_const = 0;
    for (int kc = 0; kc < kN; ++kc) n_const += rsep_column_is_constant(mask, kc) ? 1 : 0;
    if (n_const != kN) return 1;
    const size_t nn = (size_t)kN * kN;
    const cf32* planes = reinterpret_cast<const cf32*>(planes_);
    std::vector<cf32> master = master_table();
    std::vector<unsigned char> smem(G::kSmemBytes, 0);
    for (int i = 0; i < 256; ++i) fill_tw(reinterpret_cast<cf32*>(smem.data() + G::kOffTW), master.data(), i);
    uint32_t rcodes[16], rhere[16];
    for (int t = 0; t < 16; ++t) rsep_words(mask, t, rcodes + t, rhere + t);
    const double La2 = 1.0 / 2.0 / reo, g = 1.0 / (1.0 + La2);
    const float ncf1 = (float)(0.5 * g / kN), ncf2 = (float)(g / kN);
    const int mode = prox_mode(pp);
    std::vector<ThreadState> st(G::kThreads);
    const int P = (B + 1) / 2;
    for (int task = 0; task < P * 16; ++task) {
        const int plane = task >> 4, ia = 2 * plane;
        const bool has_b = ia + 1 < B;
        const float hb = has_b ? 1.f : 0.f;
        PlaneIO io;
        io.z_in_a = io.w_in_a = io.z_in_b = io.w_in_b = nullptr;
        io.x_a = x + ia * nn; io.z_a = z + ia * nn; io.w_a = w + ia * nn; io.xpw_a = nullptr;
        io.x_b = io.x_a + nn; io.z_b = io.z_a + nn; io.w_b = io.w_a + nn; io.xpw_b = nullptr;
#define FOR_T(body)                                                     \
        for (int t = 0; t < G::kThreads; ++t) {                          \
            Ctx<16> c; c.rank = task & 15; c.tid = t; c.smem = smem.data(); \
            ThreadState& s = st[t];                                      \
            body;                                                        \
        }
        FOR_T(row_load_image(c, s, img ? img + ia * nn : nullptr, (img && has_b) ? img + (ia + 1) * nn : nullptr,
                             img8 ? img8 + ia * nn : nullptr, (img8 && has_b) ? img8 + (ia + 1) * nn : nullptr));
        FOR_T(row_step1_write<false>(c, s));
        FOR_T(row_read_step2<false>(c, s); rsep_acquire_ms(c, s, planes, planes + nn, rcodes[c.rt()], ncf1, ncf2, hb));
        FOR_T(row_step1_write<true>(c, s));
        FOR_T(row_read_step2<true>(c, s); row_stash_t1(s); rsep_acquire_ma(c, s, planes + 2 * nn, rcodes[c.rt()], rhere[c.rt()], hb));
        FOR_T(row_step1_write<true>(c, s));
        FOR_T(row_read_step2<true>(c, s); row_zero_fill(c, s, 1.0f / (float)(kN * kN), has_b));
        for (int it = 0; it < iters; ++it) {
            FOR_T(row_step1_write<false>(c, s));
            FOR_T(row_read_step2<false>(c, s); rsep_blend(c, s, rcodes[c.rt()], ncf1, ncf2));
            FOR_T(row_step1_write<true>(c, s));
            FOR_T(row_read_step2<true>(c, s); row_prox_dispatch(mode, c, s, pp, has_b, it == iters - 1, true, io));
        }
#undef FOR_T
    }
    return 0;
}

// 256-point FFT of one line through the same step1 / exchange / step2 code (unit test hook)
extern "C" void k1_fft256_line(const float* in, float* out, int inverse) {
    std::vector<cf32> master = master_table(), TW(256);
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
