// dncnn_tc.cuh — K5: DnCNN / FDnCNN forward on the 5th-generation tensor cores.
//
// The denoiser of the PnP variants (reference models/network_dncnn.py:36-67, 120-141: conv3x3 1->64, ReLU,
// (nb - 2) x [conv3x3 64->64, ReLU], conv3x3 64->1; DnCNN returns x - n(x)) is 99.8 % of a PnP-ADMM-CNC
// iteration (BASELINE config 3).  The 64->64 layers are an implicit GEMM per image-row segment:
//
//     D[m = pixel x0+m of row y][n = c_out] = sum over 9 taps (ky, kx) and 64 c_in of
//         act[y + ky - 1][x0 + m + kx - 1][c_in] * w[c_out][c_in][ky][kx]            M = 128, N = 64, K = 576
//
// issued as 36 `tcgen05.mma.cta_group::1.kind::f16` (128 x 64 x 16, bf16 in, fp32 accumulate in TMEM) by one
// thread.  Activations are NHWC bf16 (128 B per pixel).  Shared memory uses the NO-SWIZZLE K-major canonical
// layout (core matrix = 8 rows x 16 B, contiguous 128 B):
//     A row buffer : [k-chunk 0..7][slot 0..130][8 c_in]   slot s <-> pixel x0 - 1 + s (1-pixel halo each side)
// so the tap shift kx is a 16-byte shift of the descriptor start address and the tap shift ky selects another
// row buffer: the im2col matrix is never materialised and every input row is loaded ONCE per strip and reused by
// the 9 taps of three output rows.  The row buffers form a 6-deep ring filled by four producer warps with 16-byte
// `cp.async` (zero-fill outside the image = the convolution's zero padding); the 72 KB of weights of the layer
//     B            : [tap 0..8][k-chunk 0..7][c_out 0..N)[8 c_in]
// stay resident for the persistent CTA's whole life.  Accumulators are double-buffered in TMEM so the epilogue
// warps (tcgen05.ld -> bias -> ReLU -> bf16 -> 128 B per pixel) overlap the next row's MMAs.
//
//   warps 0-3 : epilogue (TMEM lane quadrant = warp)          mbarriers: full[6] / empty[6]  (producer <-> MMA)
//   warp  4   : TMEM alloc, one lane issues the MMAs                      tfull[2] / tempty[2] (MMA <-> epilogue)
//   warps 5-8 : producers
//
// The first layer (c_in = 1 or 2) and nothing else runs on the CUDA cores (it is a 0.3 ms NHWC write); the last
// layer (c_out = 1) reuses the tensor-core kernel with N = 16 (rows 1..15 of B are zero) and an fp32 epilogue
// that applies the residual  out = x - n(x).
#pragma once

#include <cuda_bf16.h>

#include "cluster256.cuh"   // mbarrier helpers (k1::mbar_*), fence_proxy_async

namespace pnp {
namespace tc {

constexpr int kTileM = 128;                    // output pixels per tile
constexpr int kSlots = kTileM + 2;             // staged input pixels per row
constexpr int kPPad = 131;                     // slot pitch of a k-chunk plane, odd: the 8 chunks of a pixel hit 8 bank groups
constexpr int kChunkBytes = kPPad * 16;        // = LBO of the A descriptor
constexpr int kRowBytes = 8 * kChunkBytes;     // 16768
constexpr int kStages = 6;
constexpr int kLag = 2;                        // producer signals a row two rows after issuing it (copies stay in flight)
constexpr int kThreads = 288;
constexpr int kProducers = 128;
constexpr int kOffW = 0;
constexpr int kWBytesMax = 9 * 8 * 64 * 16;    // 73728
constexpr int kOffRing = kWBytesMax;
constexpr int kOffBar = kOffRing + kStages * kRowBytes;
constexpr int kNumBars = 2 * kStages + 4;
constexpr int kOffTmemPtr = kOffBar + 8 * kNumBars;
constexpr int kOffBias = kOffTmemPtr + 16;
constexpr int kSmemBytes = kOffBias + 64 * 4;
constexpr int kTmemCols = 128;                 // two accumulator stages of 64 columns

struct ConvParams {
    const __nv_bfloat16* in;    // [B][H][W][64]
    __nv_bfloat16* out;         // [B][H][W][64]           (N = 64 layers)
    const float* resid;         // tail: residual source, pixel (b, y, x) at resid[b * resid_bstride + y * W + x] (may be null)
    float* out_f32;             // tail: [B][H][W]
    long long resid_bstride;
    const void* w;              // packed weights of this layer (see file header)
    const float* bias;          // [64] (tail: [1])
    int B, H, W, strip, xtiles, ystrips, items;
    int relu;
    int dbg;                    // timing experiments only (results invalid): 1 = no input copies, 2 = no output stores, 4 = no MMAs
};

PNP_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
PNP_D void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
PNP_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> PNP_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
PNP_D void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
PNP_D void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
PNP_D void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
PNP_D void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, both K-major; `acc` = 0 overwrites D
PNP_D void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// shared-memory matrix descriptor, no swizzle, K-major: core matrices of 8 rows x 16 B; `lbo` = byte distance between the
// two core matrices of one K = 16 step, `sbo` = byte distance between consecutive 8-row groups  (cute/arch/mma_sm100_desc.hpp)
PNP_D uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// instruction descriptor: D fp32, A and B bf16, both K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc(int n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

#define PNP_TMEM_LD32(taddr, v, o)                                                                                   \
    asm volatile(                                                                                                    \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "  \
        "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"               \
        : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]),            \
          "=r"(v[o + 6]), "=r"(v[o + 7]), "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]),          \
          "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15]), "=r"(v[o + 16]), "=r"(v[o + 17]),      \
          "=r"(v[o + 18]), "=r"(v[o + 19]), "=r"(v[o + 20]), "=r"(v[o + 21]), "=r"(v[o + 22]), "=r"(v[o + 23]),      \
          "=r"(v[o + 24]), "=r"(v[o + 25]), "=r"(v[o + 26]), "=r"(v[o + 27]), "=r"(v[o + 28]), "=r"(v[o + 29]),      \
          "=r"(v[o + 30]), "=r"(v[o + 31])                                                                           \
        : "r"(taddr))

PNP_D void st_global_v8(void* p, const uint32_t (&o)[8]) {      // one 256-bit store (SASS STG.256)
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]),
                 "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                 : "memory");
}

// one lane of the (converged) warp
PNP_D bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, px;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

PNP_D uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// work item -> (image, x tile, y strip)
struct Item { int b, x0, y0, rows; };
PNP_D Item decode_item(const ConvParams& p, int item) {
    Item it;
    const int xt = item % p.xtiles;
    const int r = item / p.xtiles;
    const int ys = r % p.ystrips;
    it.b = r / p.ystrips;
    it.x0 = xt * kTileM;
    it.y0 = ys * p.strip;
    it.rows = (p.H - it.y0 < p.strip) ? p.H - it.y0 : p.strip;
    return it;
}

// NOUT = 64: bf16 NHWC output with bias (+ ReLU).  NOUT = 16: last layer, only c_out 0 is real; fp32 output,
// out = resid - (conv + bias) when resid != null, else conv + bias.
template <int NOUT>
__global__ void __launch_bounds__(kThreads, 1) conv64_tc_kernel(const ConvParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int kWBytes = 9 * 8 * NOUT * 16;
    constexpr uint32_t kIdesc = make_idesc(NOUT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t s0 = smem_u32(smem);
    const uint32_t ring = s0 + kOffRing, bars = s0 + kOffBar;
    auto bFull = [&](uint32_t i) { return bars + 8 * i; };
    auto bEmpty = [&](uint32_t i) { return bars + 8 * (kStages + i); };
    auto bTFull = [&](uint32_t i) { return bars + 8 * (2 * kStages + i); };
    auto bTEmpty = [&](uint32_t i) { return bars + 8 * (2 * kStages + 2 + i); };
    float* bias_s = reinterpret_cast<float*>(smem + kOffBias);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kOffTmemPtr);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { k1::mbar_init(bFull(i), kProducers); k1::mbar_init(bEmpty(i), 1); }
        for (int i = 0; i < 2; ++i) { k1::mbar_init(bTFull(i), 1); k1::mbar_init(bTEmpty(i), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < kWBytes / 16; i += kThreads)
        cp_async16(s0 + kOffW + 16 * i, reinterpret_cast<const unsigned char*>(p.w) + 16 * i, 16);
    cp_async_commit();
    if (threadIdx.x < 64) bias_s[threadIdx.x] = (threadIdx.x < (NOUT == 64 ? 64 : 1)) ? p.bias[threadIdx.x] : 0.f;
    cp_async_wait<0>();
    k1::fence_proxy_async();                   // weights (generic-proxy writes) visible to the tensor core's async proxy
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s0 + kOffTmemPtr), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 5) {
        // ------------------------------------------------------------------ producers: input rows -> ring
        const int pt = threadIdx.x - 160;
        uint32_t e = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const Item it = decode_item(p, item);
            for (int r = 0; r < it.rows + 2; ++r, ++e) {
                const uint32_t st = e % kStages;
                k1::mbar_wait(bEmpty(st), ((e / kStages) & 1u) ^ 1u);
                const int y = it.y0 - 1 + r;
                const bool yok = (y >= 0) && (y < p.H);
                const __nv_bfloat16* rowp = p.in + ((size_t)it.b * p.H + (yok ? y : 0)) * p.W * 64;
                const uint32_t dst0 = ring + st * kRowBytes;
                for (int i = pt; i < kSlots * 8; i += kProducers) {
                    const int s = i >> 3, ch = i & 7;
                    const int x = it.x0 - 1 + s;
                    const bool ok = yok && (x >= 0) && (x < p.W);
                    const __nv_bfloat16* src = ok ? rowp + (size_t)x * 64 + ch * 8 : p.in;
                    if (!(p.dbg & 1)) cp_async16(dst0 + ch * kChunkBytes + s * 16, src, ok ? 16u : 0u);
                }
                cp_async_commit();
                if (e >= (uint32_t)kLag) {
                    cp_async_wait<kLag>();
                    k1::fence_proxy_async();
                    mbar_arrive(bFull((e - kLag) % kStages));
                }
            }
        }
        cp_async_wait<0>();
        k1::fence_proxy_async();
        for (uint32_t k = (e >= (uint32_t)kLag ? e - kLag : 0u); k < e; ++k) mbar_arrive(bFull(k % kStages));
    } else if (warp == 4) {
        // ------------------------------------------------------------------ MMA issuer
        // The whole warp runs the loop (warp-uniform control and address math stay on the uniform datapath); one
        // elected lane issues the tensor-core instructions.
        uint32_t e_base = 0, waited = 0, t = 0;
        const uint64_t desc_hi_a = ((uint64_t)(kChunkBytes >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
        const uint64_t b_desc0 = make_desc(s0 + kOffW, NOUT * 16, 128);
        const bool leader = elect_one();
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const Item it = decode_item(p, item);
            for (int j = 0; j < it.rows; ++j, ++t) {
                const uint32_t as = t & 1u;
                k1::mbar_wait(bTEmpty(as), ((t >> 1) & 1u) ^ 1u);
                while (waited <= e_base + j + 2) {
                    k1::mbar_wait(bFull(waited % kStages), (waited / kStages) & 1u);
                    ++waited;
                }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 64;
                if (leader && !(p.dbg & 4)) {
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint32_t a_row = ring + ((e_base + j + dy) % kStages) * kRowBytes;
                        const uint64_t a_desc0 = desc_hi_a | (uint64_t)((a_row & 0x3FFFFu) >> 4);
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                // start-address field += byte offset / 16 (never carries out of its 14 bits: smem < 256 KB)
                                const uint64_t ad = a_desc0 + (uint64_t)((dx * 16 + ks * 2 * kChunkBytes) >> 4);
                                const uint64_t bd = b_desc0 + (uint64_t)((((dy * 3 + dx) * 8 + ks * 2) * (NOUT * 16)) >> 4);
                                tc_mma_bf16(d_tmem, ad, bd, kIdesc, (dy | dx | ks) != 0);
                            }
                        }
                    }
                }
                __syncwarp();
                if (leader) {
                    tc_commit(bEmpty((e_base + j) % kStages));           // input row j is not needed by later output rows
                    if (j == it.rows - 1) {
                        tc_commit(bEmpty((e_base + j + 1) % kStages));
                        tc_commit(bEmpty((e_base + j + 2) % kStages));
                    }
                    tc_commit(bTFull(as));
                }
                __syncwarp();
            }
            e_base += it.rows + 2;
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps
        uint32_t t = 0;
        for (int item = blockIdx.x; item < p.items; item += gridDim.x) {
            const Item it = decode_item(p, item);
            for (int j = 0; j < it.rows; ++j, ++t) {
                const uint32_t as = t & 1u;
                k1::mbar_wait(bTFull(as), (t >> 1) & 1u);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + as * 64;
                const int x = it.x0 + warp * 32 + lane, y = it.y0 + j;
                const size_t pix = ((size_t)it.b * p.H + y) * p.W + x;
                if (NOUT == 64) {
                    uint32_t v[64];
                    PNP_TMEM_LD32(taddr, v, 0);
                    PNP_TMEM_LD32(taddr + 32, v, 32);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bTEmpty(as));
                    if (x < p.W && !(p.dbg & 2)) {
                        __nv_bfloat16* dst = p.out + pix * 64;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {           // 16 channels = one 32-byte sector per store
                            float f[16];
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4) {
                                const float4 bb = reinterpret_cast<const float4*>(bias_s)[4 * q + k4];
                                f[4 * k4 + 0] = bb.x; f[4 * k4 + 1] = bb.y; f[4 * k4 + 2] = bb.z; f[4 * k4 + 3] = bb.w;
                            }
                            uint32_t o[8];
#pragma unroll
                            for (int k = 0; k < 16; ++k) {
                                f[k] += __uint_as_float(v[16 * q + k]);
                                if (p.relu) f[k] = fmaxf(f[k], 0.f);
                            }
#pragma unroll
                            for (int k = 0; k < 8; ++k) o[k] = pack_bf16x2(f[2 * k], f[2 * k + 1]);
                            st_global_v8(dst + 16 * q, o);
                        }
                    }
                } else {
                    uint32_t v0;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v0) : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bTEmpty(as));
                    if (x < p.W) {
                        const float n = __uint_as_float(v0) + bias_s[0];
                        p.out_f32[pix] = p.resid ? p.resid[(size_t)it.b * p.resid_bstride + (size_t)y * p.W + x] - n : n;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 4)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
}

// First layer: conv3x3 CIN -> 64 + bias + ReLU on the CUDA cores (K = 9 CIN is too thin for the tensor cores and the
// layer is bound by its 128 B/pixel NHWC write).  x: [B][CIN][H][W] fp32 (rounded to bf16 like a bf16 PyTorch module
// would), w: [64][CIN][3][3] fp32 (bf16-representable values), out: [B][H][W][64] bf16.
// Thread = (8 output channels, pixel lane): its 72 CIN weights live in registers for the whole kernel and it walks over
// kHeadPix pixels; the 8 threads of a pixel write its 128 bytes as one contiguous segment.
constexpr int kHeadPix = 16;     // pixels per thread
template <int CIN>
__global__ void __launch_bounds__(256) dncnn_head_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                         const float* __restrict__ w, const float* __restrict__ bias, int B,
                                                         int H, int W) {
    const int ch = threadIdx.x & 7, pl = threadIdx.x >> 3;
    float wr[CIN][8][9], br[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        br[c] = __ldg(bias + ch * 8 + c);
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
            for (int t = 0; t < 9; ++t) wr[ci][c][t] = __ldg(w + ((ch * 8 + c) * CIN + ci) * 9 + t);
    }
    const uint32_t npix = (uint32_t)B * H * W;          // < 2^31 (checked by the host)
    uint32_t pix = blockIdx.x * (32u * kHeadPix) + pl;
    if (pix >= npix) return;
    int xx = (int)(pix % (uint32_t)W);
    int yy = (int)((pix / (uint32_t)W) % (uint32_t)H);
    uint32_t b = pix / ((uint32_t)W * H);
    for (int it = 0; it < kHeadPix; ++it, pix += 32, xx += 32) {
        if (pix >= npix) return;
        while (xx >= W) { xx -= W; if (++yy == H) { yy = 0; ++b; } }
        float f[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) f[c] = br[c];
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
            const float* img = x + ((size_t)b * CIN + ci) * (size_t)H * W;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int y2 = yy + ky - 1, x2 = xx + kx - 1;
                    float v = 0.f;
                    if (y2 >= 0 && y2 < H && x2 >= 0 && x2 < W) v = __ldg(img + (size_t)y2 * W + x2);
                    v = __bfloat162float(__float2bfloat16(v));
#pragma unroll
                    for (int c = 0; c < 8; ++c) f[c] = fmaf(wr[ci][c][ky * 3 + kx], v, f[c]);
                }
        }
        uint4 o;
        o.x = pack_bf16x2(fmaxf(f[0], 0.f), fmaxf(f[1], 0.f)); o.y = pack_bf16x2(fmaxf(f[2], 0.f), fmaxf(f[3], 0.f));
        o.z = pack_bf16x2(fmaxf(f[4], 0.f), fmaxf(f[5], 0.f)); o.w = pack_bf16x2(fmaxf(f[6], 0.f), fmaxf(f[7], 0.f));
        *reinterpret_cast<uint4*>(out + (size_t)pix * 64 + ch * 8) = o;
    }
}

}  // namespace tc
}  // namespace pnp
