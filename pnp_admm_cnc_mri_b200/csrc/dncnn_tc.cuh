// dncnn_tc.cuh — K5: DnCNN / FDnCNN forward on the 5th-generation tensor cores.
//
// The denoiser of the PnP variants (reference models/network_dncnn.py:36-67, 120-141: conv3x3 1->64, ReLU,
// (nb - 2) x [conv3x3 64->64, ReLU], conv3x3 64->1; DnCNN returns x - n(x)) is 99.8 % of a PnP-ADMM-CNC
// iteration (BASELINE config 3).  The 64->64 layers are an implicit GEMM per image-row segment:
//
//     D[m = pixel x0+m of row y][n = c_out] = sum over 9 taps (ky, kx) and 64 c_in of
//         act[y + ky - 1][x0 + m + kx - 1][c_in] * w[c_out][c_in][ky][kx]            M = 128, N = 64, K = 576
//
// issued as `tcgen05.mma.cta_group::1.kind::f16` (bf16 in, fp32 accumulate in TMEM) by one elected thread: 12 instructions
// of 128 x 192 x 16 per input row, the three ky taps stacked along N (input-stationary schedule, see the kernel).
// Activations are bf16 in CHUNK-PLANAR ROWS, [B][H][8 chunks][W][8 channels]: a row of the image is 8 planes of
// W x 16 bytes.  Shared memory uses the NO-SWIZZLE K-major canonical layout (core matrix = 8 rows x 16 B = 128 B):
//     A row buffer : [k-chunk 0..7][slot 0..128 + 2 dil)[8 c_in]   slot s <-> pixel x0 - dil + s (halo of dil = 1 pixel each side;
//                    IRCNN's dilated layers: dil = 2, 3, 4, see ConvParams::dil)
// so the tap shift kx is a 16-byte shift of the descriptor start address and the tap shift ky selects another row
// buffer: the im2col matrix is never materialised, every input row is loaded ONCE per strip, and a k-chunk plane of a
// row buffer is 2080 (dil = 1) CONTIGUOUS bytes of global memory: one 1-D bulk copy (cp.async.bulk, TMA engine) per plane, no
// tensor map and no LSU instruction on the load path; zero padding comes from a zero buffer.  The row buffers form a
// 5-deep ring; the 72 KB of weights of the layer
//     B            : [kx 0..2][k-chunk 0..7][ky 0..2][c_out 0..N)[8 c_in]   (the three ky taps stacked along N)
// stay resident for the persistent CTA's whole life.  Accumulators live in an 8-block ring in TMEM (see the kernel)
// so the epilogue warps (tcgen05.ld -> ReLU -> bf16 -> staging tile -> 8 bulk stores of 2 KB) overlap the MMAs.
//
//   warps 0-7  : two epilogue groups, alternate output rows (TMEM lane quadrant = warp & 3)
//   warp  8    : TMEM alloc, one elected lane issues the MMAs
//   warps 9-10 : producers (one lane each issues the bulk copies of four k-chunk planes)
//   mbarriers  : full[5] / empty[5] (producers <-> MMA, one pair per ring stage), tfull[8] / tempty[8] (MMA <-> epilogue,
//                one pair per TMEM block)
//
// The first layer (c_in = 1 or 2) and nothing else runs on the CUDA cores (bound by its 128 B per pixel write); the last
// layer (c_out = 1) reuses the tensor-core kernel with N = 16 (rows 1..15 of B are zero) and an fp32 epilogue
// that applies the residual  out = x - n(x).
#pragma once

#include <cuda_bf16.h>

#include "cluster256.cuh"   // mbarrier helpers (k1::mbar_*), fence_proxy_async

namespace pnp {
namespace tc {

constexpr int kTileM = 128;                    // output pixels per tile
#ifndef PNP_TC_MAXDIL
#define PNP_TC_MAXDIL 4                        // A/B builds: -DPNP_TC_MAXDIL=1 gives the undilated kernel's row-buffer pitch (131 slots)
#endif
constexpr int kMaxDil = PNP_TC_MAXDIL;         // largest dilation (IRCNN: 1, 2, 3, 4, 3, 2, 1)
constexpr int kSlotsMax = kTileM + 2 * kMaxDil;    // staged input pixels per row: 128 + a halo of `dil` pixels on each side
constexpr int kPPad = kSlotsMax + 1;           // slot pitch of a k-chunk plane (137; odd, so the 8 chunk planes start in 8 different bank groups)
constexpr int kChunkBytes = kPPad * 16;        // = LBO of the A descriptor
constexpr int kRowBytes = 8 * kChunkBytes;     // 17536
constexpr int kStages = 5;
constexpr int kEpiGroups = 2;                  // epilogue groups of four warps; group g takes the output rows t = g (mod kEpiGroups)
constexpr int kMmaWarp = 4 * kEpiGroups;       // warps [0, 4 g): epilogue, then the MMA warp, then the producer warps
constexpr int kProdWarps = 2;                  // producer warps: issuing a bulk copy costs its thread ~140 cycles, so the 8 copies of a row are split
constexpr int kThreads = 32 * (kMmaWarp + 1 + kProdWarps);
constexpr int kOffW = 0;
constexpr int kWBytesMax = 9 * 8 * 64 * 16;    // 73728
constexpr int kOffRing = kWBytesMax;
constexpr int kOffBar = kOffRing + kStages * kRowBytes;
constexpr int kBlocks = 8;                     // accumulator blocks (output rows in flight) in TMEM
constexpr int kNumBars = 2 * kStages + 2 * kBlocks;
constexpr int kOffTmemPtr = kOffBar + 8 * kNumBars;
constexpr int kOffBias = kOffTmemPtr + 16;
constexpr int kOffOut = kOffBias + 64 * 4;     // staging tiles [8 chunks][128 pixels][16 B], two per epilogue group
constexpr int kSmemBytes = kOffOut + 2 * kEpiGroups * kTileM * 128;

// source of the zero padding (rows above / below the image, the pixel left / right of it)
__device__ __align__(128) unsigned char g_zero[kSlotsMax * 16];

struct ConvParams {
    const __nv_bfloat16* in;    // [B][H][8 chunks][W][8 channels]   (chunk-planar rows, see the file header)
    __nv_bfloat16* out;         // same layout                       (N = 64 layers)
    const float* resid;         // tail: residual source, pixel (b, y, x) at resid[b * resid_bstride + y * W + x] (may be null)
    float* out_f32;             // tail: [B][H][W]
    long long resid_bstride;
    const void* w;              // packed weights of this layer (see file header)
    const float* bias;          // [64] (tail: [1])
    int B, H, W, xtiles, ystrips, items;
    int relu;
    int dil;                    // dilation (= padding) of the 3x3 taps, 1..kMaxDil.  The rows y = py (mod dil) form `dil` independent
                                // row-interleaved sub-images in which the taps are ADJACENT rows, so a work item is a strip of one
                                // sub-image and the schedule is the undilated one; along x the tap shift is dil slots of the row buffer
    int kchunks;                // 8-channel chunks of the input that are real: 8 (64 channels) or 2 (a thin first layer, K = 16 per tap);
                                // the input layout then is [B][H][kchunks][W][8] and a tap costs kchunks / 2 instructions
    int cout;                   // tail (NOUT = 16): real output channels, 1 (DnCNN / FDnCNN) or 4 (FFDNet, pixel-shuffled on the way out)
    int out_H, out_W;           // tail with cout = 4: size of the pixel-shuffled output image [B][out_H][out_W] (H = ceil(out_H / 2))
    // PNPADMM_TC_DEBUG, timing experiments only, compiled in with -DPNPADMM_TC_EXPERIMENTS (PNPADMM_NVCC_EXTRA of build.py);
    // every bit but 256 makes the results invalid: 1 = no input copies, 2 = no
    // output stores, 4 = no MMAs, 8 / 16 = no tcgen05.ld / no accumulator re-init in the epilogue, 64 = no MMA <-> epilogue
    // handshake (epilogue off), 128 = no producer <-> MMA handshake (producers off), 256 = wait-time attribution (g_tc_prof)
    int dbg;
};

PNP_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
PNP_D void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
PNP_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> PNP_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
PNP_D void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
// Waits sleep in hardware (k1::mbar_wait: try_wait with a suspend-time hint, clock-bounded so that a protocol bug traps):
// the kernel runs power-capped (~1.45 GHz under load), so a spinning waiter costs clock for the warps that work.
PNP_D void mbar_wait(uint32_t bar, uint32_t parity) { k1::mbar_wait(bar, parity); }
// Optional wait-time attribution (PNPADMM_TC_DEBUG bit 256): cycles each role spent blocked on each barrier kind, summed
// over CTAs: [0] producer/empty [1] MMA/tempty [2] MMA/full [3] epilogue/tfull [4] MMA loop total [5] producer total [6] epilogue total
__device__ unsigned long long g_tc_prof[8];
PNP_D void mbar_wait_t(uint32_t bar, uint32_t parity, unsigned long long& acc, bool on) {
    if (!on) { mbar_wait(bar, parity); return; }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += (unsigned long long)(clock64() - t0);
}
PNP_D void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {      // global -> shared, counts on `bar`
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
PNP_D void bulk_store(void* dst, uint32_t src, uint32_t bytes) {                          // shared -> global, bulk async-group
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
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

#define PNP_TMEM_ST32(taddr, v, o)                                                                                   \
    asm volatile(                                                                                                    \
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "  \
        "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"                 \
        ::"r"(taddr), "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]),      \
          "r"(v[o + 6]), "r"(v[o + 7]), "r"(v[o + 8]), "r"(v[o + 9]), "r"(v[o + 10]), "r"(v[o + 11]),                \
          "r"(v[o + 12]), "r"(v[o + 13]), "r"(v[o + 14]), "r"(v[o + 15]), "r"(v[o + 16]), "r"(v[o + 17]),            \
          "r"(v[o + 18]), "r"(v[o + 19]), "r"(v[o + 20]), "r"(v[o + 21]), "r"(v[o + 22]), "r"(v[o + 23]),            \
          "r"(v[o + 24]), "r"(v[o + 25]), "r"(v[o + 26]), "r"(v[o + 27]), "r"(v[o + 28]), "r"(v[o + 29]),            \
          "r"(v[o + 30]), "r"(v[o + 31])                                                                             \
        : "memory")

PNP_D uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// work item -> (image, x tile, row sub-image py, strip of its rows): output rows y = py + dil (j0 + j), j < rows.
// A sub-image's rows are cut into `ystrips` nearly equal strips (never empty: the host checks H >= dil * ystrips).
struct Item { int b, x0, py, j0, rows; };
PNP_D Item decode_item(const ConvParams& p, int item) {
    Item it;
    const int xt = item % p.xtiles;
    int r = item / p.xtiles;
    const int ys = r % p.ystrips;
    r /= p.ystrips;
    it.py = r % p.dil;
    it.b = r / p.dil;
    it.x0 = xt * kTileM;
    const int sub = (p.H - it.py + p.dil - 1) / p.dil;           // rows of this sub-image
    it.j0 = (int)((long long)ys * sub / p.ystrips);
    it.rows = (int)((long long)(ys + 1) * sub / p.ystrips) - it.j0;
    return it;
}

// NOUT = 64: bf16 NHWC output with bias (+ ReLU).  NOUT = 16: last layer, only c_out 0 is real; fp32 output,
// out = resid - (conv + bias) when resid != null, else conv + bias.
//
// Input-stationary schedule.  Input row i of a strip contributes to the output rows j = i - dy (dy = 0, 1, 2), and
// the weights of the three dy taps are stacked along N in shared memory, so ONE instruction of N = 3 NOUT per
// (dx, k-step) feeds all three output rows from one fetch of the A slab:  12 instructions of 128 x 192 x 16 per row
// instead of 36 of 128 x 64 x 16 (shared-memory operand traffic 120 KB instead of 216 KB per row).  The accumulators
// of consecutive output rows therefore sit in consecutive TMEM column blocks, in DESCENDING row order: output row t
// (running count) lives in block (-t) mod 8, so the window of input row i is blocks [blk(j = i), blk(i - 1),
// blk(i - 2)] = b, b + 1, b + 2 (split in two instructions where it wraps).  A block is complete after its third
// input row; the epilogue drains it and re-initialises it with the BIAS (tcgen05.st), so every MMA accumulates.
template <int NOUT>
__global__ void __launch_bounds__(kThreads, 1) conv64_tc_kernel(const ConvParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int kWBytes = 9 * 8 * NOUT * 16;
    constexpr uint32_t kTmemColsK = kBlocks * NOUT;          // 512 (all of TMEM) or 128
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef PNPADMM_TC_EXPERIMENTS
    const int dbg = p.dbg;                     // timing experiments (see ConvParams::dbg); production builds compile them out
#else
    constexpr int dbg = 0;
#endif
    const bool prof = (dbg & 256) != 0;
    // Programmatic dependent launch (the layers of a forward are launched with it): the next layer may be scheduled as soon as SMs
    // free up and runs its set-up (weights, TMEM, barriers: nothing the previous layer wrote) while this one drains; it waits below,
    // before it first touches an activation buffer.  A no-op for an ordinary launch.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const uint32_t s0 = smem_u32(smem);
    const uint32_t ring = s0 + kOffRing, bars = s0 + kOffBar;
    auto bFull = [&](uint32_t i) { return bars + 8 * i; };
    auto bEmpty = [&](uint32_t i) { return bars + 8 * (kStages + i); };
    auto bTFull = [&](uint32_t i) { return bars + 8 * (2 * kStages + i); };
    auto bTEmpty = [&](uint32_t i) { return bars + 8 * (2 * kStages + kBlocks + i); };
    auto blk = [](uint32_t t) { return (kBlocks - (t & (kBlocks - 1))) & (kBlocks - 1); };
    float* bias_s = reinterpret_cast<float*>(smem + kOffBias);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kOffTmemPtr);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { k1::mbar_init(bFull(i), 1); k1::mbar_init(bEmpty(i), 1); }
        for (int i = 0; i < kBlocks; ++i) { k1::mbar_init(bTFull(i), 1); k1::mbar_init(bTEmpty(i), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < kWBytes / 16; i += kThreads)
        cp_async16(s0 + kOffW + 16 * i, reinterpret_cast<const unsigned char*>(p.w) + 16 * i, 16);
    cp_async_commit();
    if (threadIdx.x < 64) bias_s[threadIdx.x] = ((int)threadIdx.x < (NOUT == 64 ? 64 : p.cout)) ? p.bias[threadIdx.x] : 0.f;
    cp_async_wait<0>();
    k1::fence_proxy_async();                   // weights (generic-proxy writes) visible to the tensor core's async proxy
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s0 + kOffTmemPtr), "r"(kTmemColsK) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // epilogue warps: the bias is the accumulators' initial value (this thread's TMEM lane, all columns of a block)
    auto init_block = [&](uint32_t taddr) {
        if (NOUT == 64) {
#pragma unroll
            for (int h = 0; h < 4; ++h) {          // 16 columns at a time keeps the register peak low
                uint32_t br[16];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint4 b4 = reinterpret_cast<const uint4*>(bias_s)[4 * h + k];
                    br[4 * k] = b4.x; br[4 * k + 1] = b4.y; br[4 * k + 2] = b4.z; br[4 * k + 3] = b4.w;
                }
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                    ::"r"(taddr + 16 * h), "r"(br[0]), "r"(br[1]), "r"(br[2]), "r"(br[3]), "r"(br[4]), "r"(br[5]), "r"(br[6]), "r"(br[7]),
                      "r"(br[8]), "r"(br[9]), "r"(br[10]), "r"(br[11]), "r"(br[12]), "r"(br[13]), "r"(br[14]), "r"(br[15])
                    : "memory");
            }
        } else if (p.cout == 4) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(bias_s[0])),
                         "r"(__float_as_uint(bias_s[1])), "r"(__float_as_uint(bias_s[2])), "r"(__float_as_uint(bias_s[3]))
                         : "memory");
        } else {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(bias_s[0])) : "memory");
        }
    };
    if (warp < kMmaWarp) {
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        for (int b = (warp >> 2); b < kBlocks; b += kEpiGroups) init_block(lane_addr + b * NOUT);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    asm volatile("griddepcontrol.wait;" ::: "memory");     // the previous layer is complete: its output may be read, its input overwritten

    if (warp > kMmaWarp) {
        // ------------------------------------------------------------------ producer: input rows -> ring
        // One lane of each producer warp streams its half of the k-chunk planes of the input rows with 1-D bulk copies (TMA
        // engine, no LSU): in the chunk-planar layout the
        // 130 pixels a row buffer needs from one 8-channel chunk are 2080 contiguous bytes of global memory and of the
        // A operand's plane, so a row is 8 copies (+ 16-byte zero copies for the pixel left / right of the image and
        // whole zero planes for the rows above / below it: the convolution's padding).
        unsigned long long w0 = 0;
        const long long tstart = clock64();
        const int pw = warp - (kMmaWarp + 1), c0 = pw * (8 / kProdWarps);      // this warp's k-chunk planes
        if (lane == 0) {
            const unsigned char* in_b = reinterpret_cast<const unsigned char*>(p.in);
            const size_t plane = (size_t)p.W * 16, rowb = (size_t)p.kchunks * plane;
            const int c1 = c0 + 8 / kProdWarps < p.kchunks ? c0 + 8 / kProdWarps : p.kchunks;   // this warp's planes: [c0, c1)
            const uint32_t slots = kTileM + 2 * p.dil;
            const uint32_t row_tx = (uint32_t)p.kchunks * slots * 16;
            uint32_t e = 0;
            for (int item = blockIdx.x; item < ((dbg & 128) ? 0 : p.items); item += gridDim.x) {
                const Item it = decode_item(p, item);
                const int xs = it.x0 - p.dil;
                const int lo = xs < 0 ? 0 : xs;
                const int hi = it.x0 + kTileM + p.dil < p.W ? it.x0 + kTileM + p.dil : p.W;
                const uint32_t nleft = lo - xs, nvalid = hi - lo, nright = slots - nleft - nvalid;
                for (int r = 0; r < it.rows + 2; ++r, ++e) {
                    const uint32_t st = e % kStages;
                    mbar_wait_t(bEmpty(st), ((e / kStages) & 1u) ^ 1u, w0, prof);
                    const uint32_t bar = bFull(st), dst0 = ring + st * kRowBytes;
                    if (dbg & 1) { if (pw == 0) mbar_arrive(bar); continue; }
                    if (pw == 0) k1::mbar_arm_tx(bar, row_tx);       // the other producer warps' bytes may land first: fine
                    const int y = it.py + p.dil * (it.j0 - 1 + r);
                    if (y < 0 || y >= p.H) {
#pragma unroll 4
                        for (int c = c0; c < c1; ++c) bulk_load(dst0 + c * kChunkBytes, g_zero, slots * 16, bar);
                    } else {
                        const unsigned char* src = in_b + ((size_t)it.b * p.H + y) * rowb + (size_t)lo * 16;
#pragma unroll 4
                        for (int c = c0; c < c1; ++c) {
                            const uint32_t d = dst0 + c * kChunkBytes;
                            if (nleft) bulk_load(d, g_zero, nleft * 16, bar);
                            bulk_load(d + nleft * 16, src + c * plane, nvalid * 16, bar);
                            if (nright) bulk_load(d + (nleft + nvalid) * 16, g_zero, nright * 16, bar);
                        }
                    }
                }
            }
            if (prof && pw == 0) { atomicAdd(&g_tc_prof[0], w0); atomicAdd(&g_tc_prof[5], (unsigned long long)(clock64() - tstart)); }
        }
        __syncwarp();
    } else if (warp == kMmaWarp) {
        // ------------------------------------------------------------------ MMA issuer
        // The whole warp runs the loop (warp-uniform control and address math stay on the uniform datapath); one
        // elected lane issues the tensor-core instructions.
        uint32_t e = 0, t_base = 0;
        unsigned long long w1 = 0, w2 = 0;
        const long long tstart = clock64();
        const uint64_t desc_hi_a = ((uint64_t)(kChunkBytes >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
        const uint64_t b_desc0 = make_desc(s0 + kOffW, 3 * NOUT * 16, 128);
        const bool leader = elect_one();
        const int ksteps = p.kchunks >> 1;             // K = 16 per instruction = two chunks
        const int dil16 = p.dil * 16;                  // tap shift along x in bytes of the row buffer
        // The tensor pipe queues only an instruction or two, so whatever the issuing thread does between the last MMA
        // of a row and the first of the next is a bubble.  The waits for row e + 1 (pre) and the commits of row e - 1
        // (post) are therefore placed INSIDE row e's MMA stream; nothing but loop control sits between rows.
        auto pre = [&](uint32_t ee, int i, int R, uint32_t tb) {           // barriers of input row (ee, i)
            if (i < R && !(dbg & 64)) {                              // block of the output row that starts with this input row
                const uint32_t t = tb + i;
                mbar_wait_t(bTEmpty(blk(t)), ((t / kBlocks) & 1u) ^ 1u, w1, prof);
            }
            if (!(dbg & 128)) mbar_wait_t(bFull(ee % kStages), (ee / kStages) & 1u, w2, prof);
            tc_fence_after();
        };
        auto issue = [&](uint32_t ee, int i, int R, uint32_t tb, int dx0, int dx1) {   // MMAs of taps kx in [dx0, dx1)
            if (!leader || (dbg & 4)) return;
            const uint32_t a_row = ring + (ee % kStages) * kRowBytes;
            const uint64_t a_desc0 = desc_hi_a | (uint64_t)((a_row & 0x3FFFFu) >> 4);
            const int dy_hi = i < 2 ? i : 2;
            int dy = i - (R - 1) > 0 ? i - (R - 1) : 0;
            while (dy <= dy_hi) {
                const uint32_t b = blk(tb + i - dy);
                int n = dy_hi - dy + 1;
                if (n > (int)(kBlocks - b)) n = kBlocks - b;
                const uint32_t idesc = make_idesc(n * NOUT);
                const uint32_t d_tmem = tmem_base + b * NOUT;
                const uint64_t b_dy = b_desc0 + (uint64_t)((dy * NOUT * 16) >> 4);
                auto mma_ks = [&](int dx, int ks) {
                    // start-address field += byte offset / 16 (never carries out of its 14 bits: smem < 256 KB)
                    const uint64_t ad = a_desc0 + (uint64_t)((dx * dil16 + ks * 2 * kChunkBytes) >> 4);
                    const uint64_t bd = b_dy + (uint64_t)(((dx * 8 + ks * 2) * (3 * NOUT * 16)) >> 4);
                    tc_mma_bf16(d_tmem, ad, bd, idesc, 1);
                };
                for (int dx = dx0; dx < dx1; ++dx) {
                    if (ksteps == 4) {                       // 64 input channels: the hot path, fully unrolled
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) mma_ks(dx, ks);
                    } else {
                        for (int ks = 0; ks < ksteps; ++ks) mma_ks(dx, ks);
                    }
                }
                dy += n;
            }
        };
        // completion signals of input row (ee, i): its ring stage is free, and output row i - 2 has all three input rows
        auto post = [&](uint32_t ee, int i, uint32_t tb) {
            __syncwarp();
            if (leader) {
                if (!(dbg & 128)) tc_commit(bEmpty(ee % kStages));
                if (i >= 2 && !(dbg & 64)) tc_commit(bTFull(blk(tb + i - 2)));
            }
            __syncwarp();
        };
        int item = blockIdx.x;
        if (item < p.items) {
            int R = decode_item(p, item).rows, i = 0;
            bool have_prev = false;
            uint32_t p_e = 0, p_tb = 0;
            int p_i = 0;
            pre(e, i, R, t_base);
            for (;;) {
                // coordinates of the next input row
                int n_item = item, n_i = i + 1, n_R = R;
                uint32_t n_tb = t_base;
                bool has_next = true;
                if (n_i == R + 2) {
                    n_item = item + gridDim.x; n_i = 0; n_tb = t_base + R;
                    has_next = n_item < p.items;
                    if (has_next) n_R = decode_item(p, n_item).rows;
                }
                // row e: taps kx = 0 | the PREVIOUS row's commits (they then also cover these four MMAs: a few hundred
                // cycles later, but off the critical gap between rows) | kx = 1 | the NEXT row's waits | kx = 2
                issue(e, i, R, t_base, 0, 1);
                if (have_prev) post(p_e, p_i, p_tb);
                issue(e, i, R, t_base, 1, 2);
                if (has_next) pre(e + 1, n_i, n_R, n_tb);
                issue(e, i, R, t_base, 2, 3);
                have_prev = true; p_e = e; p_i = i; p_tb = t_base;
                if (!has_next) break;
                item = n_item; i = n_i; R = n_R; t_base = n_tb; ++e;
            }
            post(p_e, p_i, p_tb);
        }
        if (prof && lane == 0) { atomicAdd(&g_tc_prof[1], w1); atomicAdd(&g_tc_prof[2], w2); atomicAdd(&g_tc_prof[4], (unsigned long long)(clock64() - tstart)); }
    } else {
        // ------------------------------------------------------------------ epilogue warps
        // Groups of four warps (TMEM lane quadrant = warp & 3) take alternate output rows: a row's chain (wait ->
        // tcgen05.ld -> re-init -> pack -> stage -> barrier -> bulk stores) is ~1.7 k cycles of mostly latency, and with
        // the MMA warp's waits hidden it is what bounds the pipeline; two rows in flight halve it.
        uint32_t t = 0;
        unsigned long long w3 = 0;
        const long long tstart = clock64();
        const uint32_t grp = warp >> 2, quad = warp & 3;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
        for (int item = blockIdx.x; item < ((dbg & 64) ? 0 : p.items); item += gridDim.x) {
            const Item it = decode_item(p, item);
            for (int j = 0; j < it.rows; ++j, ++t) {
                if (kEpiGroups > 1 && (t % kEpiGroups) != grp) continue;
                const uint32_t b = blk(t);
                mbar_wait_t(bTFull(b), (t / kBlocks) & 1u, w3, prof);
                tc_fence_after();
                const uint32_t taddr = lane_addr + b * NOUT;
                const int x = it.x0 + quad * 32 + lane, y = it.py + p.dil * (it.j0 + j);
                if (NOUT == 64) {
                    uint32_t v[64];
                    if (!(dbg & 8)) {
                    PNP_TMEM_LD32(taddr, v, 0);
                    PNP_TMEM_LD32(taddr + 32, v, 32);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    }
                    if (!(dbg & 16)) {
                    init_block(taddr);                           // next use of the block starts from the bias again
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bTEmpty(b));
                    // bf16 row of this pixel -> staging tile [chunk][pixel][16 B] (the global layout of a row segment; lanes
                    // write consecutive 16-byte slots: conflict-free) -> 8 bulk copies of 2 KB by one thread.  The LSU never
                    // sees a global store.  Two staging tiles per group: the copies of the group's previous row must have READ their tile before
                    // the group's next row overwrites it; the issuing lanes check that before the one barrier of this row.
                    const uint32_t stile = s0 + kOffOut + (2 * grp + ((t / kEpiGroups) & 1u)) * (kTileM * 128);
                    const uint32_t m = quad * 32 + lane;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        uint32_t o[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            float f0 = __uint_as_float(v[8 * q + 2 * k]), f1 = __uint_as_float(v[8 * q + 2 * k + 1]);
                            if (p.relu) { f0 = fmaxf(f0, 0.f); f1 = fmaxf(f1, 0.f); }
                            o[k] = pack_bf16x2(f0, f1);
                        }
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stile + q * (kTileM * 16) + m * 16), "r"(o[0]), "r"(o[1]),
                                     "r"(o[2]), "r"(o[3])
                                     : "memory");
                    }
                    k1::fence_proxy_async();
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                    if (lane == 0) {                           // each warp's lane 0 sends two of the eight chunk planes
                        const int npx = p.W - it.x0 < kTileM ? p.W - it.x0 : kTileM;
                        unsigned char* dst = reinterpret_cast<unsigned char*>(p.out) + (((size_t)it.b * p.H + y) * 8 * p.W + it.x0) * 16;
                        if (!(dbg & 2)) {
#pragma unroll
                            for (int q = 2 * quad; q < 2 * quad + 2; ++q) bulk_store(dst + (size_t)q * p.W * 16, stile + q * (kTileM * 16), npx * 16);
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                } else if (p.cout == 4) {
                    // FFDNet tail: four output channels = the 2 x 2 sub-pixels of the full-resolution image
                    // (F.pixel_shuffle: out[2 y + dy][2 x + dx] = conv[2 dy + dx][y][x]), cropped to out_H x out_W
                    uint32_t v[4];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    init_block(taddr);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bTEmpty(b));
                    if (x < p.W && !(dbg & 2)) {
#pragma unroll
                        for (int dy = 0; dy < 2; ++dy) {
                            const int oy = 2 * y + dy, ox = 2 * x;
                            if (oy >= p.out_H) continue;
                            float* o = p.out_f32 + ((size_t)it.b * p.out_H + oy) * p.out_W + ox;
                            const float f0 = __uint_as_float(v[2 * dy]), f1 = __uint_as_float(v[2 * dy + 1]);
                            if (!(p.out_W & 1)) *reinterpret_cast<float2*>(o) = make_float2(f0, f1);      // ox + 1 < out_W, 8-byte aligned
                            else { o[0] = f0; if (ox + 1 < p.out_W) o[1] = f1; }
                        }
                    }
                } else {
                    uint32_t v0;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v0) : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    init_block(taddr);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bTEmpty(b));
                    if (x < p.W && !(dbg & 2)) {
                        const float n = __uint_as_float(v0);
                        const size_t pix = ((size_t)it.b * p.H + y) * p.W + x;
                        p.out_f32[pix] = p.resid ? p.resid[(size_t)it.b * p.resid_bstride + (size_t)y * p.W + x] - n : n;
                    }
                }
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // all staged rows have been written out
        if (prof && (threadIdx.x & 127) == 0) { atomicAdd(&g_tc_prof[3], w3); atomicAdd(&g_tc_prof[6], (unsigned long long)(clock64() - tstart)); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kMmaWarp)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemColsK) : "memory");
}

// First layer: conv3x3 CIN -> 64 + bias + ReLU on the CUDA cores (K = 9 CIN is too thin for the tensor cores and the
// layer is bound by its 128 B/pixel write).  x: [B][CIN][H][W] fp32 (rounded to bf16 like a bf16 PyTorch module would),
// w: [64][CIN][3][3] fp32 (bf16-representable values), out: [B][H][8][W][8] bf16 (chunk-planar rows).
// Block = a patch of 32 columns x kHeadRows rows; warp = 8-channel chunk, lane = column.  A thread keeps its 72 CIN
// weights in registers and slides a 3x3 window DOWN its column: three new (coalesced) loads per pixel instead of nine;
// a warp stores 32 consecutive pixels of one chunk plane (512 contiguous bytes) per row.
constexpr int kHeadRows = 64;
template <int CIN>
__global__ void __launch_bounds__(256) dncnn_head_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                         const float* __restrict__ w, const float* __restrict__ bias, int B,
                                                         int H, int W) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // the first tensor-core layer may start its set-up
    const int ch = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float wr[CIN][8][9], br[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        br[c] = __ldg(bias + ch * 8 + c);
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
            for (int t = 0; t < 9; ++t) wr[ci][c][t] = __ldg(w + ((ch * 8 + c) * CIN + ci) * 9 + t);
    }
    const int xt = (W + 31) / 32, yt = (H + kHeadRows - 1) / kHeadRows;
    const int bx = blockIdx.x % xt, by = (blockIdx.x / xt) % yt, b = blockIdx.x / (xt * yt);
    const int xx = bx * 32 + lane, y0 = by * kHeadRows;
    // the block's input patch (kHeadRows + 2 rows x 34 columns, zero outside the image, rounded to bf16) goes to shared
    // memory first: one round trip to L2 per block instead of one per row of every warp
    __shared__ float patch[CIN][kHeadRows + 2][34];
    for (int i = threadIdx.x; i < CIN * (kHeadRows + 2) * 34; i += 256) {
        const int c = i % 34, r = (i / 34) % (kHeadRows + 2), ci = i / (34 * (kHeadRows + 2));
        const int y = y0 - 1 + r, xq = bx * 32 - 1 + c;
        float v = 0.f;
        if (y >= 0 && y < H && xq >= 0 && xq < W) v = __ldg(x + (((size_t)b * CIN + ci) * H + y) * W + xq);
        patch[ci][r][c] = __bfloat162float(__float2bfloat16(v));
    }
    __syncthreads();
    if (xx >= W) return;
    auto load_row = [&](float (&v)[CIN][3], int r) {
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
            v[ci][0] = patch[ci][r][lane]; v[ci][1] = patch[ci][r][lane + 1]; v[ci][2] = patch[ci][r][lane + 2];
        }
    };
    float r0[CIN][3], r1[CIN][3], r2[CIN][3];
    load_row(r0, 0);
    load_row(r1, 1);
    const int y1 = y0 + kHeadRows < H ? y0 + kHeadRows : H;
    for (int y = y0; y < y1; ++y) {
        load_row(r2, y - y0 + 2);
        float f[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) f[c] = br[c];
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    f[c] = fmaf(wr[ci][c][k], r0[ci][k], f[c]);
                    f[c] = fmaf(wr[ci][c][3 + k], r1[ci][k], f[c]);
                    f[c] = fmaf(wr[ci][c][6 + k], r2[ci][k], f[c]);
                }
        uint4 o;
        o.x = pack_bf16x2(fmaxf(f[0], 0.f), fmaxf(f[1], 0.f)); o.y = pack_bf16x2(fmaxf(f[2], 0.f), fmaxf(f[3], 0.f));
        o.z = pack_bf16x2(fmaxf(f[4], 0.f), fmaxf(f[5], 0.f)); o.w = pack_bf16x2(fmaxf(f[6], 0.f), fmaxf(f[7], 0.f));
        *reinterpret_cast<uint4*>(out + ((((size_t)b * H + y) * 8 + ch) * W + xx) * 8) = o;      // chunk-planar rows
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
            for (int k = 0; k < 3; ++k) { r0[ci][k] = r1[ci][k]; r1[ci][k] = r2[ci][k]; }
    }
}

// FFDNet input (reference models/network_ffdnet.py:56-70: replicate-pad to even size, PixelUnShuffle(2), concatenate the
// noise-level map): x [B][H][W] fp32 -> [B][H2][2 chunks][W2][8] bf16, chunk 0 = {x[2y][2x], x[2y][2x+1], x[2y+1][2x], x[2y+1][2x+1],
// sigma, 0, 0, 0}, chunk 1 = 0 (the thin first layer reads K = 16 per tap: ConvParams::kchunks = 2).  Values are rounded to
// bf16 like the input of a bf16 PyTorch module.
__global__ void __launch_bounds__(256) ffdnet_pack_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, float sigma, int B,
                                                          int H, int W, int H2, int W2) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * H2 * W2) return;
    const int xx = (int)(i % W2), yy = (int)((i / W2) % H2), b = (int)(i / ((size_t)W2 * H2));
    const int y0 = 2 * yy, y1 = (2 * yy + 1 < H) ? 2 * yy + 1 : H - 1, x0 = 2 * xx, x1 = (2 * xx + 1 < W) ? 2 * xx + 1 : W - 1;
    const float* img = x + (size_t)b * H * W;
    uint4 o;
    o.x = pack_bf16x2(__ldg(img + (size_t)y0 * W + x0), __ldg(img + (size_t)y0 * W + x1));
    o.y = pack_bf16x2(__ldg(img + (size_t)y1 * W + x0), __ldg(img + (size_t)y1 * W + x1));
    o.z = pack_bf16x2(sigma, 0.f);
    o.w = 0u;
    uint4* row = reinterpret_cast<uint4*>(out) + ((size_t)b * H2 + yy) * 2 * W2;
    row[xx] = o;
    row[W2 + xx] = make_uint4(0u, 0u, 0u, 0u);
}

}  // namespace tc
}  // namespace pnp
