// pnpadmm.cu — C ABI (include/pnpadmm.h) of the B200-native ADMM CS-MRI path.
// Single translation unit: nvcc -gencode arch=compute_100a,code=sm_100a -shared ... pnpadmm.cu
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

#include "../../include/pnpadmm.h"
#include "cluster256.cuh"
#include "rowsep256.cuh"
#include "rowsepN.cuh"
#include "streaming.cuh"
#include "stream2.cuh"
#include "metrics.cuh"
#include "dncnn_tc.cuh"

using namespace pnp;

namespace {

thread_local char g_err[512] = "no error";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess)                                                              \
            return fail(PNPADMM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define LAUNCH_CHECK(name)                                                                  \
    do {                                                                                    \
        cudaError_t e_ = cudaGetLastError();                                                \
        if (e_ != cudaSuccess)                                                              \
            return fail(PNPADMM_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_)); \
    } while (0)

constexpr int kMaxDevices = 64;
constexpr size_t kAlign = 256;
constexpr int kMaxSmemOptin = 227 * 1024;

struct DeviceState {
    bool ready = false;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    int max_clusters_256 = 0;     // best of the two geometries below
    int max_cl8 = 0, max_cl16 = 0;  // co-resident clusters of the 8-CTA / 16-CTA variants of K1
    int k1_cluster = 8;           // geometry used by default
    cudaStream_t side = nullptr;  // hybrid schedule: the K2 share of a batch runs here, beside K1
    cudaStream_t side2 = nullptr; // second half of the K2 share
    cudaEvent_t fork_ev = nullptr, join_ev = nullptr, join_ev2 = nullptr;   // fork / join of the hybrid schedule (created once, with the side streams)
    std::mutex mu;                // guards the side stream's fork/join pairs and this device's K2 graph cache
    // timing model of the hybrid schedule, measured once per device by calibrate_hybrid() (literals = fallback, round-1 B200 fit)
    double tau1_us = 10.1;        // K1: one plane-iteration of one cluster, all clusters busy
    double k2_a_us = 11.0;        // K2 on the SMs outside the clusters, beside a running K1: per iteration, fixed part ...
    double k2_b_us = 2.7;         // ... and per plane
    double k2_pro_us = 45.0;      // K2 share of a fused reconstruct: its own acquisition / zero-fill / prepare launches
    bool calibrated = false;
};
DeviceState g_dev[kMaxDevices];
std::mutex g_mu;

size_t align_up(size_t v) { return (v + kAlign - 1) / kAlign * kAlign; }

int ilog2(int n) { int l = 0; while ((1 << l) < n) ++l; return l; }

template <typename T> constexpr int max_n() { return sizeof(T) == 4 ? 2048 : 1024; }

int check_n(int N, bool f64) {
    const int mx = f64 ? 1024 : 2048;
    if (N < 16 || N > mx || (N & (N - 1)) != 0)
        return fail(PNPADMM_ERR_BAD_SIZE, "N=%d unsupported: need a power of two in [16, %d] for %s", N, mx, f64 ? "f64" : "f32");
    return PNPADMM_OK;
}

// ------------------------------------------------------------------------------------------
// tile shapes of the streaming kernels
// ------------------------------------------------------------------------------------------
template <typename T> int rows_lines(int N) { int l = 2048 / N; if (l > N) l = N; return l < 1 ? 1 : l; }
template <typename T> int cols_lines(int N) {
    const int cap = sizeof(T) == 4 ? 8 : 4;
    const int maxpts = sizeof(T) == 4 ? 8192 : 4096;
    int l = maxpts / N;
    if (l > cap) l = cap;
    if (l < 1) l = 1;
    return l;
}
template <typename T> size_t rows_smem(int N) { return (size_t)(2 * rows_lines<T>(N) * N + N) * sizeof(cx<T>); }
template <typename T> size_t cols_smem(int N) { return (size_t)(2 * cols_lines<T>(N) * (N + 4) + N) * sizeof(cx<T>); }

template <typename T>
cudaError_t set_stream_attrs() {
    cudaError_t e;
#define SET_ATTR(k)                                                                             \
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemOptin);    \
    if (e != cudaSuccess) return e;
    SET_ATTR((rows_kernel<T, RM_FWD_ZW>))
    SET_ATTR((rows_kernel<T, RM_FWD_IMG>))
    SET_ATTR((rows_kernel<T, RM_INV_X>))
    SET_ATTR((rows_kernel<T, RM_INV_ABS>))
    SET_ATTR((rows_kernel<T, RM_INV_PROX_FWD>))
    SET_ATTR((cols_kernel<T, CM_FWD_ACQ>))
    SET_ATTR((cols_kernel<T, CM_INV>))
    SET_ATTR((cols_kernel<T, CM_FWD_BLEND_INV>))
#undef SET_ATTR
    return cudaSuccess;
}


// ------------------------------------------------------------------------------------------
// K2 register-FFT streaming kernels (stream2.cuh): fp32, N in {256, 512, 1024}
// ------------------------------------------------------------------------------------------
bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

// 2-D tensor map of `planes` row-major complex64 planes stacked as [planes N rows][N columns], boxes of C columns x 256 rows
// (the column-pass tile in four pieces at N = 1024).  The encoder comes from the driver through the runtime's entry-point
// query, so the library does not link against libcuda.  Returns false if it is unavailable (the cp.async kernel runs).
bool plane_tensor_map(CUtensorMap* tm, const void* base, int planes, int N, int C) {
    typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static Encode enc = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            (void)cudaGetLastError();
            f = nullptr;
        }
        return reinterpret_cast<Encode>(f);
    }();
    if (!enc || (((uintptr_t)base) & 15)) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)planes * N};
    const cuuint64_t strides[1] = {(cuuint64_t)N * 8};
    const cuuint32_t box[2] = {(cuuint32_t)C, 256};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T> struct S2 {
    static bool ok(int) { return false; }
    static cudaError_t set_attrs() { return cudaSuccess; }
    template <int MODE> static int rows(const StreamParams<T>&, int, cudaStream_t) { return PNPADMM_ERR_UNSUPPORTED; }
    template <int MODE> static int cols(const StreamParams<T>&, int, const uint32_t*, int, cudaStream_t) { return PNPADMM_ERR_UNSUPPORTED; }
};
template <> struct S2<float> {
    static bool ok(int N) {
        static const bool force_v1 = getenv("PNPADMM_STREAM_V1") != nullptr;   // A/B against the generic kernels
        return !force_v1 && (N == 256 || N == 512 || N == 1024);
    }
    template <int N> static cudaError_t set_attrs_n() {
        cudaError_t e;
#define SET2(k, bytes)                                                                          \
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);            \
    if (e != cudaSuccess) return e;
        SET2((s2::rows2_kernel<N, RM_FWD_ZW>), s2::RowsGeo<N>::kSmemBytes)
        SET2((s2::rows2_kernel<N, RM_FWD_IMG>), s2::RowsGeo<N>::kSmemBytes)
        SET2((s2::rows2_kernel<N, RM_INV_X>), s2::RowsGeo<N>::kSmemBytes)
        SET2((s2::rows2_kernel<N, RM_INV_ABS>), s2::RowsGeo<N>::kSmemBytes)
        SET2((s2::rows2_kernel<N, RM_INV_PROX_FWD>), s2::RowsGeo<N>::kSmemBytes)
        SET2((s2::cols2_kernel<N, CM_FWD_ACQ>), s2::ColsGeo<N>::kSmemBytes)
        SET2((s2::cols2_kernel<N, CM_INV>), s2::ColsGeo<N>::kSmemBytes)
        SET2((s2::cols2_kernel<N, CM_FWD_BLEND_INV>), s2::ColsGeo<N>::kSmemBytes)
        SET2((s2::cols2_tma_kernel<N, false>), s2::ColsGeo<N>::kSmemBytes + 32)
        SET2((s2::cols2_tma_kernel<N, true>), s2::ColsGeo<N>::kSmemBytes + 32)
#undef SET2
        return cudaSuccess;
    }
    static cudaError_t set_attrs() {
        cudaError_t e = set_attrs_n<256>(); if (e != cudaSuccess) return e;
        e = set_attrs_n<512>(); if (e != cudaSuccess) return e;
        return set_attrs_n<1024>();
    }
    // The passes of the iteration are launched with programmatic stream serialisation: a pass may be scheduled while the
    // previous one drains and waits (griddepcontrol.wait) right before it first touches the data (PNPADMM_NO_PDL=1: plain launches).
    static bool pdl_enabled() {
        static const bool on = [] { const char* e = getenv("PNPADMM_NO_PDL"); return !(e && atoi(e) != 0); }();
        return on;
    }
    template <typename... KArgs, typename... Args>
    static void launch(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, bool pdl, Args&&... args) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
        (void)cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
    }
    // `planes`: packed planes, or images for the per-image modes (FWD_IMG, INV_ABS, FWD_ACQ, INV)
    template <int N, int MODE> static void rows_n(const StreamParams<float>& p, int planes, cudaStream_t st) {
        typedef s2::RowsGeo<N> G;
        constexpr bool kIter = (MODE == RM_INV_PROX_FWD);
        launch(s2::rows2_kernel<N, MODE>, (unsigned)(planes * (N / G::L)), s2::kRowsThreads, G::kSmemBytes, st, kIter, p);
    }
    template <int MODE> static int rows(const StreamParams<float>& p, int planes, cudaStream_t st) {
        switch (p.N) {
            case 256: rows_n<256, MODE>(p, planes, st); break;
            case 512: rows_n<512, MODE>(p, planes, st); break;
            default: rows_n<1024, MODE>(p, planes, st); break;
        }
        return PNPADMM_OK;
    }
    template <int N, int MODE> static void cols_n(StreamParams<float> p, int planes, const uint32_t* mpack, int sm_count, cudaStream_t st) {
        typedef s2::ColsGeo<N> G;
        p.P = planes;
        const long ntiles = (long)planes * (N / G::C);
        const long cap = (long)sm_count * G::kCtasPerSm;
        if (MODE == CM_FWD_BLEND_INV) {      // iteration pass: K and G tiles through 2-D TMA (PNPADMM_COLS_LSU=1: cp.async variant)
            static const bool lsu = getenv("PNPADMM_COLS_LSU") != nullptr;
            CUtensorMap tmK, tmG;
            if (!lsu && plane_tensor_map(&tmK, p.K, planes, N, G::C) && plane_tensor_map(&tmG, p.G, planes, N, G::C)) {
                static const bool tma_store = getenv("PNPADMM_COLS_TMA_STORE") != nullptr;   // experiment: measured 1 % slower
                const unsigned grid = (unsigned)(ntiles < cap ? ntiles : cap);
                if (tma_store) launch(s2::cols2_tma_kernel<N, true>, grid, G::kThreads, G::kSmemBytes + 32, st, true, p, mpack, tmK, tmG);
                else launch(s2::cols2_tma_kernel<N, false>, grid, G::kThreads, G::kSmemBytes + 32, st, true, p, mpack, tmK, tmG);
                return;
            }
        }
        s2::cols2_kernel<N, MODE><<<(unsigned)(ntiles < cap ? ntiles : cap), G::kThreads, G::kSmemBytes, st>>>(p, mpack);
    }
    template <int MODE> static int cols(const StreamParams<float>& p, int planes, const uint32_t* mpack, int sm_count, cudaStream_t st) {
        switch (p.N) {
            case 256: cols_n<256, MODE>(p, planes, mpack, sm_count, st); break;
            case 512: cols_n<512, MODE>(p, planes, mpack, sm_count, st); break;
            default: cols_n<1024, MODE>(p, planes, mpack, sm_count, st); break;
        }
        return PNPADMM_OK;
    }
};

template <int CL>
int probe_clusters(int sm_count) {
    typedef k1::Geo<CL> G;
    if (cudaFuncSetAttribute(k1::cluster256_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::kSmemBytes) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    if (CL > 8 && cudaFuncSetAttribute(k1::cluster256_kernel<CL>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    if (CL == 16 && (cudaFuncSetAttribute(k1::cluster256_bk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G::kSmemBytes) != cudaSuccess ||
                     cudaFuncSetAttribute(k1::cluster256_bk_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)) {
        (void)cudaGetLastError();
        return 0;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(CL * sm_count);
    cfg.blockDim = dim3(G::kThreads);
    cfg.dynamicSmemBytes = G::kSmemBytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int ncl = 0;
    if (cudaOccupancyMaxActiveClusters(&ncl, k1::cluster256_kernel<CL>, &cfg) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return ncl;
}

void calibrate_hybrid(DeviceState* d);

int ensure_device(DeviceState** out) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) return fail(PNPADMM_ERR_CUDA, "device ordinal %d out of range", dev);
    DeviceState& d = g_dev[dev];
    std::lock_guard<std::mutex> lk(g_mu);      // first-time initialisation only does real work under it
    if (!d.ready) {
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
        d.sm_count = prop.multiProcessorCount;
        d.cc_major = prop.major;
        d.cc_minor = prop.minor;
        if (prop.major != 10)
            return fail(PNPADMM_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                        dev, prop.major, prop.minor);
        // master twiddle tables, computed in double precision
        std::vector<float2> tf(kTwMax);
        std::vector<double2> td(kTwMax);
        for (int i = 0; i < kTwMax; ++i) {
            const long double a = -2.0L * 3.14159265358979323846264338327950288L * i / kTwMax;
            td[i] = make_double2((double)cosl(a), (double)sinl(a));
            tf[i] = make_float2((float)td[i].x, (float)td[i].y);
        }
        CUDA_TRY(cudaMemcpyToSymbol(g_tw_f32, tf.data(), sizeof(float2) * kTwMax));
        CUDA_TRY(cudaMemcpyToSymbol(g_tw_f64, td.data(), sizeof(double2) * kTwMax));
        std::vector<float2> t256(256);
        for (int i = 0; i < 256; ++i) t256[i] = tf[((i >> 4) * (i & 15) * 16) & (kTwMax - 1)];   // = k1::fill_tw
        CUDA_TRY(cudaMemcpyToSymbol(s2::g_tw256, t256.data(), sizeof(float2) * 256));
        CUDA_TRY(set_stream_attrs<float>());
        CUDA_TRY(set_stream_attrs<double>());
        CUDA_TRY(S2<float>::set_attrs());
        CUDA_TRY(cudaFuncSetAttribute(k3::rowsep256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k1::Geo<16>::kSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(k3::rowsepN_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, k3::RsGeo<256>::kSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(k3::rowsepN_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, k3::RsGeo<512>::kSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(k3::rowsepN_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, k3::RsGeo<1024>::kSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(metrics_tile_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMetSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(metrics_tile_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMetSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(tc::conv64_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes));
        CUDA_TRY(cudaFuncSetAttribute(tc::conv64_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes));
        d.max_cl8 = probe_clusters<8>(d.sm_count);
        d.max_cl16 = probe_clusters<16>(d.sm_count);
        // Default geometry: 16 half-size CTAs (two planes share an SM, so one computes while the other
        // transposes; measured 3-4 % faster than 8 full-size CTAs on B200).  PNPADMM_K1_CLUSTER=8|16 overrides.
        d.k1_cluster = (d.max_cl16 > 0 && d.max_cl16 * 16 >= d.max_cl8 * 8 * 9 / 10) ? 16 : 8;
        if (const char* e = getenv("PNPADMM_K1_CLUSTER")) {
            const int want = atoi(e);
            if (want == 16 && d.max_cl16 > 0) d.k1_cluster = 16;
            if (want == 8 && d.max_cl8 > 0) d.k1_cluster = 8;
        }
        if (const char* e = getenv("PNPADMM_K1_MAXCL")) {   // experiments: override the occupancy query
            const int n = atoi(e);
            if (n > 0) { d.max_cl8 = n; d.max_cl16 = n; }
        }
        d.max_clusters_256 = d.k1_cluster == 16 ? d.max_cl16 : d.max_cl8;
        CUDA_TRY(cudaStreamCreateWithFlags(&d.side, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&d.fork_ev, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&d.join_ev, cudaEventDisableTiming));
        CUDA_TRY(cudaStreamCreateWithFlags(&d.side2, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&d.join_ev2, cudaEventDisableTiming));
        d.ready = true;
        calibrate_hybrid(&d);     // best effort: the literals stay if it cannot run
    }
    *out = &d;
    return PNPADMM_OK;
}

// ------------------------------------------------------------------------------------------
// workspace
// ------------------------------------------------------------------------------------------
template <typename T>
struct Workspace {
    T* cf;            // [3] residual coefficients g*mcode/(2N^2) written by prepare
    cx<T>* K;         // [P][N][N]
    cx<T>* G;         // [P][N][N]
    cx<T>* T1;        // [B][N][N] per-image complex scratch
    uint8_t* mcode;   // [N][N] or [P][N][N]
    uint32_t* mpack;  // N == 256 only: [16][256] or [P][16][256] packed codes for the cluster kernel
    cx<T>* Gt;        // N == 256 fp32 only: G in the cluster kernel's tile order [P][rank][256][R]
    int* progress;    // [P][8] hand-off counters of the chunked cluster schedule
    cx<T>* Y;         // [B][N][N] measurements of the reconstruct entry points (acquisition output when it is not fused away)
    cx<T>* tiles;     // N == 256 fp32 only: [6][N][N]: noise-term tiles of the fused prologue (NcS, nH, nA); K3 uses [0,3) as the
                      // row-major terms and [3,6) as their column inverse transforms
    uint32_t* mhere;  // N == 256 fp32 only: [16][256] m[k] bits; K3: rcodes [16] | rhere [16] | separability count
    int P, solo;
};

size_t ws_bytes_impl(int B, int N, size_t elt, int mask_batched) {
    const size_t nn = (size_t)N * N;
    const size_t P = mask_batched ? (size_t)B : (size_t)(B + 1) / 2;
    size_t s = kAlign;                       // header (cf table)
    s += align_up((size_t)B * nn * 2 * elt); // T1
    s += align_up(P * nn * 2 * elt);         // K
    s += align_up(P * nn * 2 * elt);         // G
    s += align_up((mask_batched ? P : 1) * nn);
    s += align_up((mask_batched ? P : 1) * nn / 4);   // packed codes (2 bits per bin, 32-bit words)
    s += align_up((P * 16 + 16) * sizeof(int));       // progress counters [P][<=16 ranks] + task queue
    if (N == 256 && elt == 4) s += align_up(P * nn * 2 * elt);   // Gt (cluster kernel)
    s += align_up((size_t)B * nn * 2 * elt);                     // Y
    if ((N == 256 || N == 512 || N == 1024) && elt == 4) s += align_up(6 * nn * 2 * elt) + align_up(16 * 256 * sizeof(uint32_t));   // fused prologue / K3
    return s;
}

// `scratch_only`: the call touches only the header + per-image scratch T1 (acquire, zero_filled),
// whose position does not depend on the pairing, so any workspace sized for (B, N) is accepted.
template <typename T>
int carve(void* ws, size_t ws_bytes, int B, int N, int mask_batched, Workspace<T>* out, bool scratch_only = false) {
    if (!ws) return fail(PNPADMM_ERR_WORKSPACE, "workspace is NULL");
    if (((uintptr_t)ws) % kAlign) return fail(PNPADMM_ERR_WORKSPACE, "workspace must be %zu-byte aligned", kAlign);
    const size_t nn = (size_t)N * N;
    const size_t need = scratch_only ? kAlign + align_up((size_t)B * nn * 2 * sizeof(T))
                                     : ws_bytes_impl(B, N, sizeof(T), mask_batched);
    if (ws_bytes < need) return fail(PNPADMM_ERR_WORKSPACE, "workspace too small: %zu < %zu bytes", ws_bytes, need);
    const size_t P = mask_batched ? (size_t)B : (size_t)(B + 1) / 2;
    unsigned char* p = (unsigned char*)ws;
    out->cf = (T*)p; p += kAlign;
    out->T1 = (cx<T>*)p; p += align_up((size_t)B * nn * 2 * sizeof(T));
    out->K = (cx<T>*)p; p += align_up(P * nn * 2 * sizeof(T));
    out->G = (cx<T>*)p; p += align_up(P * nn * 2 * sizeof(T));
    out->mcode = p; p += align_up((mask_batched ? P : 1) * nn);
    out->mpack = (uint32_t*)p; p += align_up((mask_batched ? P : 1) * nn / 4);
    out->progress = (int*)p; p += align_up((P * 16 + 16) * sizeof(int));
    out->Gt = (N == 256 && sizeof(T) == 4) ? (cx<T>*)p : nullptr;
    if (out->Gt) p += align_up(P * nn * 2 * sizeof(T));
    out->Y = nullptr; out->tiles = nullptr; out->mhere = nullptr;
    if (!scratch_only) {
        out->Y = (cx<T>*)p; p += align_up((size_t)B * nn * 2 * sizeof(T));
        if ((N == 256 || N == 512 || N == 1024) && sizeof(T) == 4) {
            out->tiles = (cx<T>*)p; p += align_up(6 * nn * 2 * sizeof(T));
            out->mhere = (uint32_t*)p;
        }
    }
    out->P = (int)P;
    out->solo = mask_batched ? 1 : 0;
    return PNPADMM_OK;
}

template <typename T>
ProxParams<T> make_prox(int prox, double lambda1, double reo, double alpha, double b) {
    ProxParams<T> p;
    p.prox = prox;
    p.thr_l1 = (T)(reo * lambda1);
    p.inv_b = (T)(1.0 / b);
    p.one_m_alpha = (T)(1.0 - alpha);
    p.alpha = (T)alpha;
    p.coef = (T)(alpha * reo * lambda1 * b);
    p.thr_cnc = (T)(alpha * reo * lambda1);
    p.general = (p.thr_l1 < T(0) || p.inv_b < T(0) || p.thr_cnc < T(0)) ? 1 : 0;
    return p;
}

int check_prox(int prox, int iters, double reo, double b) {
    if (prox != PNPADMM_PROX_L1 && prox != PNPADMM_PROX_CNC) return fail(PNPADMM_ERR_BAD_ARG, "unknown prox %d", prox);
    if (iters < 0) return fail(PNPADMM_ERR_BAD_ARG, "iters=%d < 0", iters);
    if (!(reo > 0.0)) return fail(PNPADMM_ERR_BAD_ARG, "reo=%g must be > 0", reo);
    if (prox == PNPADMM_PROX_CNC && b == 0.0) return fail(PNPADMM_ERR_BAD_ARG, "b must be non-zero for PROX_CNC");
    return PNPADMM_OK;
}

template <typename T>
StreamParams<T> base_params(const Workspace<T>& w, int B, int N) {
    StreamParams<T> p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.log2N = ilog2(N); p.B = B; p.P = w.P; p.solo = w.solo;
    p.K = w.K; p.G = w.G; p.mcode = w.mcode; p.mcode_batched = w.solo;
    p.cf = w.cf;
    p.scale = (T)(1.0 / ((double)N * N));
    return p;
}

int grid_1d(size_t n, int sm_count) {
    size_t g = (n + 255) / 256;
    const size_t cap = (size_t)sm_count * 16;
    return (int)(g < cap ? (g ? g : 1) : cap);
}

// ------------------------------------------------------------------------------------------
// implementations (templated on precision)
// ------------------------------------------------------------------------------------------
template <typename T>
int acquire_impl(const T* img, const uint8_t* mask, const T* noise, T* y, int B, int N, int mask_batched,
                 int noise_batched, int round_f32, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!img || !mask || !noise || !y || B <= 0) return fail(PNPADMM_ERR_BAD_ARG, "acquire: NULL pointer or B <= 0");
    int rc = check_n(N, sizeof(T) == 8); if (rc) return rc;
    DeviceState* d; rc = ensure_device(&d); if (rc) return rc;
    Workspace<T> w; rc = carve<T>(ws, ws_bytes, B, N, 0, &w, true); if (rc) return rc;
    StreamParams<T> p = base_params(w, B, N);
    p.img = img; p.cout = w.T1; p.round_f32 = (sizeof(T) == 8) ? round_f32 : 0;
    if (S2<T>::ok(N) && aligned16(img) && aligned16(noise) && aligned16(y)) {
        S2<T>::template rows<RM_FWD_IMG>(p, B, st);
        LAUNCH_CHECK("rows2_kernel<FWD_IMG>");
        p.cin = w.T1; p.cout = reinterpret_cast<cx<T>*>(y);
        p.mask = mask; p.mask_batched = mask_batched;
        p.noise = reinterpret_cast<const cx<T>*>(noise); p.noise_batched = noise_batched;
        S2<T>::template cols<CM_FWD_ACQ>(p, B, nullptr, d->sm_count, st);
        LAUNCH_CHECK("cols2_kernel<FWD_ACQ>");
        return PNPADMM_OK;
    }
    p.lines = rows_lines<T>(N);
    rows_kernel<T, RM_FWD_IMG><<<dim3(N / p.lines, B), 256, rows_smem<T>(N), st>>>(p);
    LAUNCH_CHECK("rows_kernel<FWD_IMG>");
    p.cin = w.T1; p.cout = reinterpret_cast<cx<T>*>(y);
    p.mask = mask; p.mask_batched = mask_batched;
    p.noise = reinterpret_cast<const cx<T>*>(noise); p.noise_batched = noise_batched;
    p.lines = cols_lines<T>(N);
    cols_kernel<T, CM_FWD_ACQ><<<dim3(N / p.lines, B), 256, cols_smem<T>(N), st>>>(p);
    LAUNCH_CHECK("cols_kernel<FWD_ACQ>");
    return PNPADMM_OK;
}

template <typename T>
int zero_filled_impl(const T* y, T* x0, int B, int N, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!y || !x0 || B <= 0) return fail(PNPADMM_ERR_BAD_ARG, "zero_filled: NULL pointer or B <= 0");
    int rc = check_n(N, sizeof(T) == 8); if (rc) return rc;
    DeviceState* d; rc = ensure_device(&d); if (rc) return rc;
    Workspace<T> w; rc = carve<T>(ws, ws_bytes, B, N, 0, &w, true); if (rc) return rc;
    StreamParams<T> p = base_params(w, B, N);
    p.cin = reinterpret_cast<const cx<T>*>(y); p.cout = w.T1;
    if (S2<T>::ok(N) && aligned16(y) && aligned16(x0)) {
        S2<T>::template cols<CM_INV>(p, B, nullptr, d->sm_count, st);
        LAUNCH_CHECK("cols2_kernel<INV>");
        p.cin = w.T1; p.x = x0;
        S2<T>::template rows<RM_INV_ABS>(p, B, st);
        LAUNCH_CHECK("rows2_kernel<INV_ABS>");
        return PNPADMM_OK;
    }
    p.lines = cols_lines<T>(N);
    cols_kernel<T, CM_INV><<<dim3(N / p.lines, B), 256, cols_smem<T>(N), st>>>(p);
    LAUNCH_CHECK("cols_kernel<INV>");
    p.cin = w.T1; p.x = x0;
    p.lines = rows_lines<T>(N);
    rows_kernel<T, RM_INV_ABS><<<dim3(N / p.lines, B), 256, rows_smem<T>(N), st>>>(p);
    LAUNCH_CHECK("rows_kernel<INV_ABS>");
    return PNPADMM_OK;
}

template <typename T>
__global__ void write_cf_kernel(T* cf, T c0, T c1, T c2) { cf[0] = c0; cf[1] = c1; cf[2] = c2; }

template <typename T>
int prepare_impl(const T* y, const uint8_t* mask, int B, int N, int mask_batched, double reo, void* ws,
                 size_t ws_bytes, cudaStream_t st) {
    if (!y || !mask || B <= 0) return fail(PNPADMM_ERR_BAD_ARG, "prepare: NULL pointer or B <= 0");
    if (!(reo > 0.0)) return fail(PNPADMM_ERR_BAD_ARG, "reo=%g must be > 0", reo);
    int rc = check_n(N, sizeof(T) == 8); if (rc) return rc;
    DeviceState* d; rc = ensure_device(&d); if (rc) return rc;
    Workspace<T> w; rc = carve<T>(ws, ws_bytes, B, N, mask_batched, &w); if (rc) return rc;
    const double La2 = 1.0 / 2.0 / reo;            // S1:117
    const double g = 1.0 / (1.0 + La2);
    const double n2 = (double)N * N;
    write_cf_kernel<T><<<1, 1, 0, st>>>(w.cf, (T)0, (T)(0.5 * g / n2), (T)(g / n2));
    LAUNCH_CHECK("write_cf_kernel");
    const size_t total = (size_t)w.P * N * N;
    const bool k1_ready = (N == k1::kN && sizeof(T) == 4 && d->max_clusters_256 > 0);
    prepare_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(reinterpret_cast<const cx<T>*>(y), mask, w.G,
                                                                      w.mcode, B, w.P, N, w.solo, mask_batched,
                                                                      (T)(g / n2), k1_ready ? w.Gt : nullptr,
                                                                      k1::kN / d->k1_cluster);
    LAUNCH_CHECK("prepare_kernel");
    if (sizeof(T) == 4 && (N == 256 || N == 512 || N == 1024)) {   // packed codes of the column passes (K1 and K2)
        const int planes = mask_batched ? w.P : 1;
        const size_t words = (size_t)planes * (N / 16) * N;
        s2::pack_mcode_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(w.mcode, w.mpack, planes, N);
        LAUNCH_CHECK("pack_mcode_kernel");
    }
    return PNPADMM_OK;
}

int pick_kernel(int kernel, int N, bool f64, const DeviceState* d, bool* use_cluster, bool allow_rowsep = false) {
    if (kernel == PNPADMM_KERNEL_ROWSEP) {
        if (!allow_rowsep || (N != 256 && N != 512 && N != 1024) || f64)
            return fail(PNPADMM_ERR_UNSUPPORTED, "the row-separable kernel serves pnpadmm_reconstruct_f32 / the host entry points at N in "
                        "{256, 512, 1024} with one mask and one noise array for the batch (N=%d, f64=%d)", N, (int)f64);
        *use_cluster = false;
        return PNPADMM_OK;
    }
    if (kernel != PNPADMM_KERNEL_AUTO && kernel != PNPADMM_KERNEL_CLUSTER && kernel != PNPADMM_KERNEL_STREAMING)
        return fail(PNPADMM_ERR_BAD_ARG, "unknown kernel selector %d", kernel);
    const bool can = (N == 256) && !f64 && d->max_clusters_256 > 0;
    if (kernel == PNPADMM_KERNEL_CLUSTER && !can)
        return fail(PNPADMM_ERR_UNSUPPORTED, "cluster kernel needs N == 256, f32 and a device that can co-schedule an 8-CTA cluster (N=%d, f64=%d, max clusters=%d)",
                    N, (int)f64, d->max_clusters_256);
    *use_cluster = (kernel == PNPADMM_KERNEL_STREAMING) ? false : can;
    return PNPADMM_OK;
}

// What a chunk boundary costs, in iteration-times: the plane's z, w go through L2, the pipeline of phases drains and refills
// (state reload + first row FFT + first transposes).  Round 2, B200: 29 planes in 10 chunks on 14 clusters take 1.236 ms
// = 21 steps x (5 + 1.0) x 9.82 us; the round-1 value 0.35 made the planner prefer such schedules to whole rounds.
constexpr double kChunkHandoff = 1.0;

// Pick the chunking of the static cluster schedule: minimise ceil(P n / ncl) * (ceil(iters / n) + hand-off).
void plan_chunks(int P, int iters, int max_clusters, int* chunk, int* n_chunks) {
    *chunk = iters; *n_chunks = 1;
    if (P <= max_clusters || P % max_clusters == 0 || iters < 2) return;
    const double handoff = kChunkHandoff;   // iterations' worth of time a chunk boundary costs
    double best = 1e30;
    for (int n = 1; n <= 10 && n <= iters; ++n) {
        const int c = (iters + n - 1) / n, nn = (iters + c - 1) / c;
        const long steps = ((long)P * nn + max_clusters - 1) / max_clusters;
        const double cost = steps * (c + (nn > 1 ? handoff : 0.0));
        if (cost < best - 1e-9) { best = cost; *chunk = c; *n_chunks = nn; }
    }
}

// 16-CTA geometry: transposes as bulk copies through the TMA engine (cluster256_bk_kernel); PNPADMM_K1_BULK=0 selects the
// st.async kernel for A/B runs.
bool k1_bulk_enabled() {
    static const bool on = [] { const char* e = getenv("PNPADMM_K1_BULK"); return e ? atoi(e) != 0 : false; }();
    return on;
}

template <int CL>
int launch_cluster_t(k1::ClusterParams& cp, int max_clusters, cudaStream_t st) {
    typedef k1::Geo<CL> G;
    plan_chunks(cp.P, cp.iters, max_clusters, &cp.chunk, &cp.n_chunks);
    if (const char* e = getenv("PNPADMM_K1_CHUNKS")) {   // experiments: force the chunk count
        const int n = atoi(e);
        if (n >= 1 && n <= cp.iters) { cp.chunk = (cp.iters + n - 1) / n; cp.n_chunks = (cp.iters + cp.chunk - 1) / cp.chunk; }
    }
    if (!cp.no_memset) CUDA_TRY(cudaMemsetAsync(cp.progress, 0, sizeof(int) * (cp.P * 16 + 16), st));   // hand-off counters + task queue
    cp.queue = cp.progress + cp.P * 16;
    const long ntasks = (long)cp.P * cp.n_chunks;
    int ncl = ntasks < max_clusters ? (int)ntasks : max_clusters;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(ncl * CL);
    cfg.blockDim = dim3(G::kThreads);
    cfg.dynamicSmemBytes = G::kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (const char* e = getenv("PNPADMM_K1_POLICY")) {   // experiments: 1 = spread, 2 = load balancing
        attr[1].id = cudaLaunchAttributeClusterSchedulingPolicyPreference;
        attr[1].val.clusterSchedulingPolicyPreference =
            atoi(e) == 2 ? cudaClusterSchedulingPolicyLoadBalancing : cudaClusterSchedulingPolicySpread;
        cfg.numAttrs = 2;
    }
    if (const char* e = getenv("PNPADMM_K1_STAGGER")) cp.dbg = atoi(e) << 8;   // experiments: start delay of odd clusters (us)
    if (CL == 16 && k1_bulk_enabled() && !cp.fused) {   // the experiment kernel has no fused prologue
        CUDA_TRY(cudaLaunchKernelEx(&cfg, k1::cluster256_bk_kernel, (const k1::ClusterParams)cp));
        return PNPADMM_OK;
    }
    CUDA_TRY(cudaLaunchKernelEx(&cfg, k1::cluster256_kernel<CL>, (const k1::ClusterParams)cp));
    return PNPADMM_OK;
}

int launch_cluster(k1::ClusterParams& cp, const DeviceState* d, cudaStream_t st) {
    return d->k1_cluster == 16 ? launch_cluster_t<16>(cp, d->max_cl16, st) : launch_cluster_t<8>(cp, d->max_cl8, st);
}

template <typename T> struct ClusterDispatch {
    static int run(const Workspace<T>&, const T*, const T*, T*, T*, T*, T*, int, int, const ProxParams<T>&,
                   const DeviceState*, cudaStream_t) {
        return fail(PNPADMM_ERR_UNSUPPORTED, "cluster kernel is f32 only");
    }
};
template <> struct ClusterDispatch<float> {
    static int run(const Workspace<float>& w, const float* z_in, const float* w_in, float* x, float* z, float* wo,
                   float* xpw, int B, int iters, const ProxParams<float>& pp, const DeviceState* d, cudaStream_t st) {
        k1::ClusterParams cp;
        memset(&cp, 0, sizeof(cp));
        cp.B = B; cp.P = w.P; cp.solo = w.solo; cp.iters = iters;
        cp.z_in = z_in; cp.w_in = w_in; cp.x = x; cp.z = z; cp.w = wo; cp.xpw = xpw;
        cp.G = reinterpret_cast<const k1::cf32*>(w.Gt); cp.mpack = w.mpack; cp.mcode_batched = w.solo;
        cp.progress = w.progress;
        cp.cf = w.cf;
        cp.prox = pp;
        const char* dbg = getenv("PNPADMM_K1_DEBUG");   // timing experiments only
        cp.dbg = dbg ? atoi(dbg) : 0;
        static const int spin = [] { const char* e = getenv("PNPADMM_K1_SPIN"); return e ? atoi(e) : 0; }();
        cp.spin = spin;
        static const int rot = [] { const char* e = getenv("PNPADMM_K1_ROT"); return e ? atoi(e) : 0; }();
        cp.rot = rot;
        return launch_cluster(cp, d, st);
    }
};

// K2 iteration loop on the planes of `w` (fp32, N in {256, 512, 1024}); `sms` caps the persistent grid of the
// columns pass (the hybrid schedule runs it on the SMs K1 leaves free).
template <typename T>
int stream2_launch_sequence(const StreamParams<T>& p0, int planes, const uint32_t* mpack, int iters, int sms, cudaStream_t st) {
    StreamParams<T> p = p0;
    S2<T>::template rows<RM_FWD_ZW>(p, planes, st);
    for (int it = 0; it < iters; ++it) {
        S2<T>::template cols<CM_FWD_BLEND_INV>(p, planes, mpack, sms, st);
        p.last = (it == iters - 1);
        S2<T>::template rows<RM_INV_PROX_FWD>(p, planes, st);
    }
    LAUNCH_CHECK("K2 iteration kernels");
    return PNPADMM_OK;
}

// The sequence is 2 * iters + 1 launches with fixed arguments: it is captured into a CUDA graph (small per-device
// cache keyed by every kernel argument) and replayed with one launch, so a caller that reconstructs batch after
// batch through the same buffers spends ~10 us of host time per call instead of ~0.4 ms.  A key is only captured
// the SECOND time it is seen (among the last 8 distinct keys; callers that allocate fresh x, z, w per call would otherwise pay a capture
// + instantiation per call and thrash the cache).  Skipped when the caller is itself capturing the stream, or with
// PNPADMM_NO_GRAPH=1.  The cache is guarded by the device's own mutex (DeviceState::mu), not the global one.
struct K2GraphKey {
    StreamParams<float> p;
    const uint32_t* mpack;
    int planes, iters, sms, pad;
};
struct K2GraphEntry {
    bool valid = false;
    K2GraphKey key;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    unsigned long long stamp = 0;
};
constexpr int kGraphSlots = 6;
constexpr int kMissRing = 8;
struct K2GraphCache {
    K2GraphEntry slot[kGraphSlots];
    K2GraphKey seen[kMissRing];      // keys launched directly so far (ring): a key found here is captured
    int n_seen = 0, next_seen = 0;
    unsigned long long clock = 0;
};
K2GraphCache g_k2graphs[kMaxDevices];

// Keys are compared byte-wise: both sides are built by memset(0) + memcpy of structs whose padding is itself
// zeroed (base_params memsets StreamParams before filling it), so padding never carries stack garbage.
void make_key(K2GraphKey* key, const StreamParams<float>& p, const uint32_t* mpack, int planes, int iters, int sms) {
    memset(key, 0, sizeof(*key));
    memcpy(&key->p, &p, sizeof(p));
    key->mpack = mpack; key->planes = planes; key->iters = iters; key->sms = sms;
}

template <typename T>
int stream2_replay(const StreamParams<T>& p, int planes, const uint32_t* mpack, int iters, int sms, cudaStream_t st, bool) {
    return stream2_launch_sequence<T>(p, planes, mpack, iters, sms, st);
}
// `locked`: the caller already holds the device mutex (hybrid schedule).
template <>
int stream2_replay<float>(const StreamParams<float>& p, int planes, const uint32_t* mpack, int iters, int sms, cudaStream_t st,
                          bool locked) {
    static const char* off = getenv("PNPADMM_NO_GRAPH");
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if ((off && atoi(off) != 0) || iters < 4 || st == nullptr || st == cudaStreamLegacy ||   // the legacy stream cannot be captured
        cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
        (void)cudaGetLastError();
        return stream2_launch_sequence<float>(p, planes, mpack, iters, sms, st);
    }
    int dev = 0; CUDA_TRY(cudaGetDevice(&dev));
    K2GraphKey key;
    make_key(&key, p, mpack, planes, iters, sms);
    std::unique_lock<std::mutex> lk(g_dev[dev].mu, std::defer_lock);
    if (!locked) lk.lock();
    K2GraphCache& c = g_k2graphs[dev];
    K2GraphEntry* slot = nullptr;
    K2GraphEntry* victim = nullptr;
    for (int i = 0; i < kGraphSlots; ++i) {
        K2GraphEntry& e = c.slot[i];
        if (e.valid && memcmp(&e.key, &key, sizeof(key)) == 0) { slot = &e; break; }
        if (!victim || (victim->valid && (!e.valid || e.stamp < victim->stamp))) victim = &e;
    }
    if (!slot) {
        bool seen = false;
        for (int i = 0; i < c.n_seen && !seen; ++i) seen = memcmp(&c.seen[i], &key, sizeof(key)) == 0;
        if (!seen) {   // first sighting: launch directly, remember it (a caller cycling through a few buffer sets still hits)
            memcpy(&c.seen[c.next_seen], &key, sizeof(key));
            c.next_seen = (c.next_seen + 1) % kMissRing;
            if (c.n_seen < kMissRing) ++c.n_seen;
            return stream2_launch_sequence<float>(p, planes, mpack, iters, sms, st);
        }
        slot = victim;
        if (slot->valid) {   // evict before the capture starts (graph destruction is not a capture-safe call)
            (void)cudaGraphExecDestroy(slot->exec);
            (void)cudaGraphDestroy(slot->graph);
            slot->valid = false;
        }
        if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            (void)cudaGetLastError();
            return stream2_launch_sequence<float>(p, planes, mpack, iters, sms, st);
        }
        const int rc = stream2_launch_sequence<float>(p, planes, mpack, iters, sms, st);
        cudaGraph_t g = nullptr;
        const cudaError_t e_end = cudaStreamEndCapture(st, &g);
        if (rc != PNPADMM_OK || e_end != cudaSuccess || !g) {
            if (g) (void)cudaGraphDestroy(g);
            (void)cudaGetLastError();
            return rc != PNPADMM_OK ? rc : fail(PNPADMM_ERR_CUDA, "capturing the K2 launch sequence failed: %s", cudaGetErrorString(e_end));
        }
        cudaGraphExec_t ex = nullptr;
        const cudaError_t e_inst = cudaGraphInstantiate(&ex, g, 0);
        if (e_inst != cudaSuccess) {
            (void)cudaGraphDestroy(g);
            return fail(PNPADMM_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e_inst));
        }
        memcpy(&slot->key, &key, sizeof(key)); slot->graph = g; slot->exec = ex; slot->valid = true;
    }
    slot->stamp = ++c.clock;
    CUDA_TRY(cudaGraphLaunch(slot->exec, st));
    return PNPADMM_OK;
}

template <typename T>
int stream2_iterate(const Workspace<T>& w, T* x, T* z, T* wv, int B, int N, const ProxParams<T>& pp, int iters, int sms,
                    cudaStream_t st, bool locked = false) {
    StreamParams<T> p = base_params(w, B, N);
    p.z = z; p.w = wv; p.x = x; p.prox = pp;
    return stream2_replay<T>(p, w.P, w.mpack, iters, sms, st, locked);
}

// Hybrid schedule for N = 256: the cluster kernel can only use the SMs that form whole 8-SM groups inside
// a GPC (112 or 120 of 148); the planes [P1, P) run on the K2 streaming kernels on a side stream and land
// on the SMs K1 leaves free.  Returns P1 (== P: no split).  Model (fitted on B200, tools/k1_bench.py sweeps in
// profiles/r1_experiments.txt): K1 takes tau1 = 10.1 us per plane-iteration and cluster; K2 on the leftover
// SMs takes 11 + 2.7 p2 us per iteration for p2 planes (one wave of `cap` planes has a 21 us latency floor).
int plan_hybrid(const DeviceState* d, int P, int iters, bool fused = false) {
    static const char* off = getenv("PNPADMM_NO_HYBRID");
    if (off && atoi(off) != 0) return P;
    const int ncl = d->max_clusters_256;
    const int left = d->sm_count - 8 * ncl;
    const int cap = left / 8;                      // planes per K2 wave: 32 row-CTAs (4 per SM) / 16 col-tiles (2 per SM) per plane
    if (ncl <= 0 || cap < 1 || iters < 4 || P <= ncl) return P;
    if (const char* e = getenv("PNPADMM_HYBRID_P2")) {   // experiments: force the K2 share
        const int p2 = atoi(e);
        return (p2 >= 0 && p2 < P) ? P - p2 : P;
    }
    // per-device constants (calibrate_hybrid); a fused reconstruct adds the prologue to both shares: ~1.5 iteration-times per plane
    // on K1 (three extra row phases, two column phases), a handful of small launches on the K2 side
    const double tau1 = d->tau1_us, handoff = kChunkHandoff;
    double best = 1e30; int best_p2 = 0;
    for (int p2 = 0; p2 <= P - ncl && p2 <= P / 3; ++p2) {
        const int p1 = P - p2;
        int chunk, nch; plan_chunks(p1, iters, ncl, &chunk, &nch);
        const long steps = ((long)p1 * nch + ncl - 1) / ncl;
        const long rounds = ((long)p1 + ncl - 1) / ncl;
        const double t1 = (steps * (chunk + (nch > 1 ? handoff : 0.0)) + (fused ? 1.5 * rounds : 0.0)) * tau1;
        const double t2 = p2 ? (double)iters * (d->k2_a_us + d->k2_b_us * p2) + (fused ? d->k2_pro_us : 0.0) : 0.0;
        const double t = t1 > t2 ? t1 : t2;
        if (t < best - 1e-9) { best = t; best_p2 = p2; }
    }
    return P - best_p2;
}

// Measure the constants of the hybrid planner on this device (first use of the library on it; PNPADMM_NO_CALIBRATE=1 keeps the
// literals): K1's time per plane-iteration with every resident cluster busy (two run lengths, the difference removes the launch
// and the prologue), and K2's time per iteration for `cap` and 2 `cap` planes on the SMs outside the clusters WHILE a long K1 run
// occupies the rest (the situation the planner models).  Zero-filled scratch problem: the kernels' timing does not depend on data.
void calibrate_hybrid(DeviceState* d) {
    static const bool off = [] { const char* e = getenv("PNPADMM_NO_CALIBRATE"); return e && atoi(e) != 0; }();
    const int ncl = d->max_clusters_256, left = d->sm_count - 8 * ncl, cap = left / 8;
    if (off || ncl <= 0 || cap < 1) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(cudaStreamPerThread, &cs) != cudaSuccess) { (void)cudaGetLastError(); return; }
    const int N = k1::kN, P = ncl + 2 * cap, B = 2 * P;
    const size_t nn = (size_t)N * N, wsb = ws_bytes_impl(B, N, 4, 0), sb = (size_t)B * nn * 4;
    unsigned char* buf = nullptr;
    cudaStream_t sa = nullptr, sbm = nullptr;
    cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
    bool ok = cudaMalloc(&buf, wsb + 3 * sb) == cudaSuccess && cudaMemset(buf, 0, wsb + 3 * sb) == cudaSuccess &&
              cudaStreamCreateWithFlags(&sa, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&sbm, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 4 && ok; ++i) ok = cudaEventCreate(&e[i]) == cudaSuccess;
    Workspace<float> w;
    ok = ok && carve<float>(buf, wsb, B, N, 0, &w) == PNPADMM_OK;
    if (ok) {
        float* x = reinterpret_cast<float*>(buf + wsb); float* z = x + (size_t)B * nn; float* wv = z + (size_t)B * nn;
        const ProxParams<float> pp = make_prox<float>(PROX_CNC, 0.5, 0.05, 0.45, 64.0);
        auto k1_run = [&](int planes, int iters, cudaStream_t st) {
            Workspace<float> w1 = w; w1.P = planes;
            return ClusterDispatch<float>::run(w1, z, wv, x, z, wv, nullptr, 2 * planes, iters, pp, d, st);
        };
        auto ms = [&](cudaEvent_t a, cudaEvent_t b) { float t = 0.f; return cudaEventElapsedTime(&t, a, b) == cudaSuccess ? (double)t : -1.0; };
        // K1: ncl planes (one per cluster), 8 and 24 iterations
        ok = k1_run(ncl, 8, sa) == PNPADMM_OK;                                         // warm-up
        double t8 = -1, t24 = -1;
        if (ok) { cudaEventRecord(e[0], sa); ok = k1_run(ncl, 8, sa) == PNPADMM_OK; cudaEventRecord(e[1], sa); }
        if (ok) { cudaEventRecord(e[2], sa); ok = k1_run(ncl, 24, sa) == PNPADMM_OK; cudaEventRecord(e[3], sa); }
        if (ok && cudaStreamSynchronize(sa) == cudaSuccess) { t8 = ms(e[0], e[1]); t24 = ms(e[2], e[3]); }
        const double tau1 = (t24 - t8) * 1e3 / 16.0;
        // K2 beside a running K1: planes [ncl, ncl + cap) and [ncl, ncl + 2 cap), 16 iterations each
        double ta = -1, tb = -1;
        if (ok && S2<float>::ok(N)) {
            StreamParams<float> p = base_params(w, B, N);
            const size_t o = (size_t)ncl * nn, ob = 2 * o;
            p.K = w.K + o; p.G = w.G + o; p.z = z + ob; p.w = wv + ob; p.x = x + ob; p.prox = pp;
            ok = k1_run(ncl, 400, sa) == PNPADMM_OK;                                   // ~4 ms of K1 to run beside
            stream2_launch_sequence<float>(p, cap, w.mpack, 4, left, sbm);              // warm-up
            cudaEventRecord(e[0], sbm); stream2_launch_sequence<float>(p, cap, w.mpack, 16, left, sbm); cudaEventRecord(e[1], sbm);
            cudaEventRecord(e[2], sbm); stream2_launch_sequence<float>(p, 2 * cap, w.mpack, 16, left, sbm); cudaEventRecord(e[3], sbm);
            if (ok && cudaStreamSynchronize(sbm) == cudaSuccess && cudaStreamSynchronize(sa) == cudaSuccess) { ta = ms(e[0], e[1]); tb = ms(e[2], e[3]); }
        }
        if (tau1 > 2.0 && tau1 < 50.0) d->tau1_us = tau1;
        if (ta > 0 && tb > ta) {
            const double a_it = ta * 1e3 / 16.0, b_it = tb * 1e3 / 16.0, per_plane = (b_it - a_it) / cap, fixed = a_it - per_plane * cap;
            if (per_plane > 0.2 && per_plane < 20.0 && fixed > 0.0 && fixed < 100.0) { d->k2_b_us = per_plane; d->k2_a_us = fixed; }
        }
        d->calibrated = true;
    }
    (void)cudaDeviceSynchronize();
    for (int i = 0; i < 4; ++i) if (e[i]) (void)cudaEventDestroy(e[i]);
    if (sa) (void)cudaStreamDestroy(sa);
    if (sbm) (void)cudaStreamDestroy(sbm);
    if (buf) (void)cudaFree(buf);
    (void)cudaGetLastError();
}

bool k2_split_enabled();

template <typename T>
int xupdate_impl(const T* z, const T* wv, T* x, T* xpw, int B, int N, int mask_batched, int kernel, void* ws,
                 size_t ws_bytes, cudaStream_t st) {
    if (!z || !wv || !x || B <= 0) return fail(PNPADMM_ERR_BAD_ARG, "xupdate: NULL pointer or B <= 0");
    int rc = check_n(N, sizeof(T) == 8); if (rc) return rc;
    DeviceState* d; rc = ensure_device(&d); if (rc) return rc;
    Workspace<T> w; rc = carve<T>(ws, ws_bytes, B, N, mask_batched, &w); if (rc) return rc;
    bool use_cluster; rc = pick_kernel(kernel, N, sizeof(T) == 8, d, &use_cluster); if (rc) return rc;
    if (use_cluster) {
        ProxParams<T> pp = make_prox<T>(PROX_NONE, 0, 1, 0, 1);
        return ClusterDispatch<T>::run(w, z, wv, x, nullptr, nullptr, xpw, B, 1, pp, d, st);
    }
    StreamParams<T> p = base_params(w, B, N);
    p.z = const_cast<T*>(z); p.w = const_cast<T*>(wv); p.x = x; p.xpw = xpw;
    if (S2<T>::ok(N) && aligned16(z) && aligned16(wv) && aligned16(x) && aligned16(xpw)) {
        S2<T>::template rows<RM_FWD_ZW>(p, w.P, st);
        S2<T>::template cols<CM_FWD_BLEND_INV>(p, w.P, w.mpack, d->sm_count, st);
        S2<T>::template rows<RM_INV_X>(p, w.P, st);
        LAUNCH_CHECK("K2 x-update kernels");
        return PNPADMM_OK;
    }
    p.lines = rows_lines<T>(N);
    rows_kernel<T, RM_FWD_ZW><<<dim3(N / p.lines, w.P), 256, rows_smem<T>(N), st>>>(p);
    LAUNCH_CHECK("rows_kernel<FWD_ZW>");
    p.lines = cols_lines<T>(N);
    cols_kernel<T, CM_FWD_BLEND_INV><<<dim3(N / p.lines, w.P), 256, cols_smem<T>(N), st>>>(p);
    LAUNCH_CHECK("cols_kernel<FWD_BLEND_INV>");
    p.lines = rows_lines<T>(N);
    rows_kernel<T, RM_INV_X><<<dim3(N / p.lines, w.P), 256, rows_smem<T>(N), st>>>(p);
    LAUNCH_CHECK("rows_kernel<INV_X>");
    return PNPADMM_OK;
}

template <typename T>
int iterate_impl(T* x, T* z, T* wv, int B, int N, int mask_batched, int prox, int iters, double lambda1, double reo,
                 double alpha, double b, int kernel, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!x || !z || !wv || B <= 0) return fail(PNPADMM_ERR_BAD_ARG, "iterate: NULL pointer or B <= 0");
    int rc = check_n(N, sizeof(T) == 8); if (rc) return rc;
    rc = check_prox(prox, iters, reo, b); if (rc) return rc;
    DeviceState* d; rc = ensure_device(&d); if (rc) return rc;
    Workspace<T> w; rc = carve<T>(ws, ws_bytes, B, N, mask_batched, &w); if (rc) return rc;
    bool use_cluster; rc = pick_kernel(kernel, N, sizeof(T) == 8, d, &use_cluster); if (rc) return rc;
    if (iters == 0) return PNPADMM_OK;
    ProxParams<T> pp = make_prox<T>(prox, lambda1, reo, alpha, b);
    if (use_cluster) {
        const int P1 = (kernel == PNPADMM_KERNEL_AUTO && S2<T>::ok(N) && aligned16(z) && aligned16(wv) && aligned16(x))
                           ? plan_hybrid(d, w.P, iters) : w.P;
        if (P1 == w.P) return ClusterDispatch<T>::run(w, z, wv, x, z, wv, nullptr, B, iters, pp, d, st);
        // K1 on planes [0, P1) in `st`; K2 on planes [P1, P) in the side stream, joined back into `st`
        const size_t nn = (size_t)N * N;
        const int B1 = w.solo ? P1 : 2 * P1;
        Workspace<T> w1 = w, w2 = w;
        w1.P = P1;
        w2.P = w.P - P1;
        w2.K = w.K + (size_t)P1 * nn; w2.G = w.G + (size_t)P1 * nn;
        if (w.solo) { w2.mcode = w.mcode + (size_t)P1 * nn; w2.mpack = w.mpack + (size_t)P1 * (nn / 16); }
        // The fork / join events live in DeviceState (no allocation per call).  A stream wait captures the record
        // that precedes it, so one pair serves every caller as long as each record + wait pair (and the side stream's
        // launch order) is not interleaved with another host thread's: the device mutex.
        std::lock_guard<std::mutex> lk(d->mu);
        CUDA_TRY(cudaEventRecord(d->fork_ev, st));
        rc = ClusterDispatch<T>::run(w1, z, wv, x, z, wv, nullptr, B1, iters, pp, d, st);
        if (rc) return rc;
        const int sms = d->sm_count - 8 * d->max_clusters_256;
        auto share = [&](int pa, int pb, cudaStream_t s2, cudaEvent_t join) -> int {   // planes [pa, pb) on the streaming kernels in s2
            const int ba = w.solo ? pa : 2 * pa, bb0 = w.solo ? pb : 2 * pb, bb = bb0 < B ? bb0 : B;
            Workspace<T> ws2 = w;
            ws2.P = pb - pa;
            ws2.K = w.K + (size_t)pa * nn; ws2.G = w.G + (size_t)pa * nn;
            if (w.solo) { ws2.mcode = w.mcode + (size_t)pa * nn; ws2.mpack = w.mpack + (size_t)pa * (nn / 16); }
            CUDA_TRY(cudaStreamWaitEvent(s2, d->fork_ev, 0));
            const int r = stream2_iterate<T>(ws2, x + (size_t)ba * nn, z + (size_t)ba * nn, wv + (size_t)ba * nn, bb - ba, N, pp, iters, sms,
                                             s2, /*locked=*/true);
            if (r) return r;
            CUDA_TRY(cudaEventRecord(join, s2));
            CUDA_TRY(cudaStreamWaitEvent(st, join, 0));
            return PNPADMM_OK;
        };
        const int P2 = w.P - P1;
        if (k2_split_enabled() && P2 >= 2) {
            const int pm = P1 + (P2 + 1) / 2;
            rc = share(P1, pm, d->side, d->join_ev);
            if (rc == PNPADMM_OK) rc = share(pm, w.P, d->side2, d->join_ev2);
        } else {
            rc = share(P1, w.P, d->side, d->join_ev);
        }
        return rc;
    }

    StreamParams<T> p = base_params(w, B, N);
    p.z = z; p.w = wv; p.x = x; p.prox = pp;
    if (S2<T>::ok(N) && aligned16(z) && aligned16(wv) && aligned16(x))
        return stream2_iterate<T>(w, x, z, wv, B, N, pp, iters, d->sm_count, st);
    const int rl = rows_lines<T>(N), cl = cols_lines<T>(N);
    const size_t rs = rows_smem<T>(N), cs = cols_smem<T>(N);
    p.lines = rl;
    rows_kernel<T, RM_FWD_ZW><<<dim3(N / rl, w.P), 256, rs, st>>>(p);
    LAUNCH_CHECK("rows_kernel<FWD_ZW>");
    for (int it = 0; it < iters; ++it) {
        p.lines = cl;
        cols_kernel<T, CM_FWD_BLEND_INV><<<dim3(N / cl, w.P), 256, cs, st>>>(p);
        p.lines = rl;
        p.last = (it == iters - 1);
        rows_kernel<T, RM_INV_PROX_FWD><<<dim3(N / rl, w.P), 256, rs, st>>>(p);
    }
    LAUNCH_CHECK("streaming iteration kernels");
    return PNPADMM_OK;
}

template <typename T>
int solve_impl(const T* y, const uint8_t* mask, T* x, T* z, T* wv, int B, int N, int mask_batched, int prox, int iters,
               double lambda1, double reo, double alpha, double b, int kernel, void* ws, size_t ws_bytes,
               cudaStream_t st) {
    if (!y || !mask || !x || !z || !wv || B <= 0) return fail(PNPADMM_ERR_BAD_ARG, "solve: NULL pointer or B <= 0");
    int rc = check_n(N, sizeof(T) == 8); if (rc) return rc;
    rc = check_prox(prox, iters, reo, b); if (rc) return rc;
    DeviceState* d; rc = ensure_device(&d); if (rc) return rc;
    // x = |ifft2(y)|; z = x; w = 0                                              (S1:100-105)
    rc = zero_filled_impl<T>(y, z, B, N, ws, ws_bytes, st); if (rc) return rc;
    const size_t n = (size_t)B * N * N;
    copy_zero_kernel<T><<<grid_1d(n, d->sm_count), 256, 0, st>>>(z, x, wv, n);
    LAUNCH_CHECK("copy_zero_kernel");
    rc = prepare_impl<T>(y, mask, B, N, mask_batched, reo, ws, ws_bytes, st); if (rc) return rc;
    return iterate_impl<T>(x, z, wv, B, N, mask_batched, prox, iters, lambda1, reo, alpha, b, kernel, ws, ws_bytes, st);
}

// ------------------------------------------------------------------------------------------
// Whole reconstruction from IMAGES (what the reference's functions do per image, S1:97-132 / S4:101-138): acquisition,
// zero-filled start, data term, iterations.  For N == 256, fp32, one mask and one noise array for the batch, all of it runs
// inside the cluster kernel (fused prologue, cluster256_core.cuh) after ONE small preparation launch; the planes the hybrid
// schedule gives to the streaming kernels go through acquisition / zero-fill / prepare on the side stream, concurrently.
// Everything else: acquisition into the workspace's Y region, then solve.  img or img8 (uint8 gray levels, / 255) is given.
// ------------------------------------------------------------------------------------------
// Experiment knob PNPADMM_K2_SPLIT=1: the K2 share of the hybrid schedule in two halves on two side streams (measured in round 2:
// within run-to-run noise of the single stream, so it stays off).
bool k2_split_enabled() {
    static const bool on = [] { const char* e = getenv("PNPADMM_K2_SPLIT"); return e ? atoi(e) != 0 : false; }();
    return on;
}

template <typename T> struct FusedDispatch {
    static bool run(int*, const Workspace<T>&, const T*, const uint8_t*, const uint8_t*, const T*, T*, T*, T*, int, int, int, int,
                    double, double, double, double, int, DeviceState*, cudaStream_t) { return false; }
};
template <> struct FusedDispatch<float> {
    // K3: row-separable mask (full k-space lines): preparation + column inverse transform of the three noise-term planes + ONE kernel
    static int run_rowsep(const Workspace<float>& w, const float* img, const uint8_t* img8, const uint8_t* mask, const float* noise,
                          float* x, float* z, float* wv, int B, int prox, int iters, double lambda1, double reo, double alpha, double b,
                          DeviceState* d, cudaStream_t st) {
        const int N = k1::kN;
        const size_t nn = (size_t)N * N;
        if (!w.tiles || w.solo) return fail(PNPADMM_ERR_UNSUPPORTED, "row-separable kernel: one mask for the batch, fp32, N == 256");
        if (iters < 1) return fail(PNPADMM_ERR_BAD_ARG, "row-separable kernel: iters=%d < 1", iters);
        const double La2 = 1.0 / 2.0 / reo, g = 1.0 / (1.0 + La2), n2 = (double)nn;
        uint32_t* rcodes = w.mhere; uint32_t* rhere = w.mhere + 16; int* sep = reinterpret_cast<int*>(w.mhere + 32);
        CUDA_TRY(cudaMemsetAsync(sep, 0, sizeof(int), st));
        k3::prepare_rowsep_kernel<<<(unsigned)(nn / 256), 256, 0, st>>>(mask, reinterpret_cast<const k1::cf32*>(noise), (float)(g / n2),
                                                                       reinterpret_cast<k1::cf32*>(w.tiles), rcodes, rhere, sep);
        LAUNCH_CHECK("prepare_rowsep_kernel");
        StreamParams<float> sp = base_params(w, 3, N);
        sp.cin = w.tiles; sp.cout = w.tiles + 3 * nn;
        S2<float>::template cols<CM_INV>(sp, 3, nullptr, d->sm_count, st);
        LAUNCH_CHECK("cols2_kernel<INV> (noise terms)");
        k3::RowSepParams rp;
        memset(&rp, 0, sizeof(rp));
        rp.B = B; rp.P = (B + 1) / 2; rp.iters = iters;
        rp.img = img; rp.img8 = img ? nullptr : img8;
        rp.x = x; rp.z = z; rp.w = wv;
        rp.planes = reinterpret_cast<const k1::cf32*>(w.tiles + 3 * nn);
        rp.rcodes = rcodes; rp.rhere = rhere; rp.sep = sep;
        rp.ncf1 = (float)(0.5 * g / N); rp.ncf2 = (float)(g / N);
        rp.prox = make_prox<float>(prox, lambda1, reo, alpha, b);
        const int tasks = rp.P * 16, cap = 2 * d->sm_count;
        k3::rowsep256_kernel<<<tasks < cap ? tasks : cap, 256, k1::Geo<16>::kSmemBytes, st>>>(rp);
        LAUNCH_CHECK("rowsep256_kernel");
        return PNPADMM_OK;
    }
    // K3 at N = 512 / 1024 (rowsepN.cuh, the K2 line FFT): same three launches
    template <int N>
    static void launch_rowsep_n(const k3::RowSepNParams& rp, int sm_count, cudaStream_t st) {
        typedef k3::RsGeo<N> G;
        const int tasks = rp.P * (N / G::L), cap = 2 * sm_count;
        k3::rowsepN_kernel<N><<<tasks < cap ? tasks : cap, 256, G::kSmemBytes, st>>>(rp);
    }
    static int run_rowsep_n(const Workspace<float>& w, const float* img, const uint8_t* img8, const uint8_t* mask, const float* noise,
                            float* x, float* z, float* wv, int B, int N, int prox, int iters, double lambda1, double reo, double alpha,
                            double b, DeviceState* d, cudaStream_t st) {
        const size_t nn = (size_t)N * N;
        if (!w.tiles || w.solo) return fail(PNPADMM_ERR_UNSUPPORTED, "row-separable kernel: one mask for the batch, fp32, N in {256, 512, 1024}");
        if (iters < 1) return fail(PNPADMM_ERR_BAD_ARG, "row-separable kernel: iters=%d < 1", iters);
        const double La2 = 1.0 / 2.0 / reo, g = 1.0 / (1.0 + La2), n2 = (double)nn;
        uint32_t* rcodes = w.mhere; uint32_t* rhere = w.mhere + 64; int* sep = reinterpret_cast<int*>(w.mhere + 128);
        CUDA_TRY(cudaMemsetAsync(sep, 0, sizeof(int), st));
        k3::prepare_rowsepN_kernel<<<(unsigned)(nn / 256), 256, 0, st>>>(mask, reinterpret_cast<const k1::cf32*>(noise), N, (float)(g / n2),
                                                                        reinterpret_cast<k1::cf32*>(w.tiles), rcodes, rhere, sep);
        LAUNCH_CHECK("prepare_rowsepN_kernel");
        StreamParams<float> sp = base_params(w, 3, N);
        sp.cin = w.tiles; sp.cout = w.tiles + 3 * nn;
        S2<float>::template cols<CM_INV>(sp, 3, nullptr, d->sm_count, st);
        LAUNCH_CHECK("cols2_kernel<INV> (noise terms)");
        k3::RowSepNParams rp;
        memset(&rp, 0, sizeof(rp));
        rp.B = B; rp.P = (B + 1) / 2; rp.iters = iters;
        rp.img = img; rp.img8 = img ? nullptr : img8;
        rp.x = x; rp.z = z; rp.w = wv;
        rp.planes = reinterpret_cast<const k1::cf32*>(w.tiles + 3 * nn);
        rp.rcodes = rcodes; rp.rhere = rhere; rp.sep = sep;
        rp.ncf1 = (float)(0.5 * g / N); rp.ncf2 = (float)(g / N);
        rp.prox = make_prox<float>(prox, lambda1, reo, alpha, b);
        if (N == 256) launch_rowsep_n<256>(rp, d->sm_count, st);
        else if (N == 512) launch_rowsep_n<512>(rp, d->sm_count, st);
        else launch_rowsep_n<1024>(rp, d->sm_count, st);
        LAUNCH_CHECK("rowsepN_kernel");
        return PNPADMM_OK;
    }
    // returns true if it handled the call (*rc holds the status)
    static bool run(int* rc, const Workspace<float>& w, const float* img, const uint8_t* img8, const uint8_t* mask, const float* noise,
                    float* x, float* z, float* wv, int B, int N, int prox, int iters, double lambda1, double reo, double alpha,
                    double b, int kernel, DeviceState* d, cudaStream_t st) {
        if (kernel == PNPADMM_KERNEL_ROWSEP) {
            // N = 256 has two implementations: rowsepN<256> on the K2 line FFT (default: measured 7-9 % faster, 24 B of spills instead of
            // 136) and rowsep256 on K1's row-phase code (PNPADMM_K3_K1CODE=1, kept for A/B)
            static const bool k3_k1code = [] { const char* e = getenv("PNPADMM_K3_K1CODE"); return e && atoi(e) != 0; }();
            *rc = (N == k1::kN && k3_k1code) ? run_rowsep(w, img, img8, mask, noise, x, z, wv, B, prox, iters, lambda1, reo, alpha, b, d, st)
                                           : run_rowsep_n(w, img, img8, mask, noise, x, z, wv, B, N, prox, iters, lambda1, reo, alpha, b, d, st);
            return true;
        }
        static const bool off = [] { const char* e = getenv("PNPADMM_NO_FUSED_PROLOGUE"); return e && atoi(e) != 0; }();
        if (off || N != k1::kN || iters < 1 || w.solo || kernel == PNPADMM_KERNEL_STREAMING || d->max_clusters_256 <= 0 || !w.tiles ||
            !aligned16(x) || !aligned16(z) || !aligned16(wv) || (img && !aligned16(img)) || !aligned16(noise))
            return false;
        const size_t nn = (size_t)N * N;
        const double La2 = 1.0 / 2.0 / reo, g = 1.0 / (1.0 + La2), n2 = (double)nn;
        const ProxParams<float> pp = make_prox<float>(prox, lambda1, reo, alpha, b);
        const int P1 = (kernel == PNPADMM_KERNEL_AUTO && S2<float>::ok(N)) ? plan_hybrid(d, w.P, iters, true) : w.P;
        const int B1 = (2 * P1 < B) ? 2 * P1 : B;
        // one launch: noise-term tiles, packed mask codes (K1 and K2 share the words), m[k] bits, mcode bytes, cf table,
        // and the zeroed hand-off counters + task queue of the cluster launch
        k1::prepare_shared_kernel<<<16, 256, 0, st>>>(mask, reinterpret_cast<const k1::cf32*>(noise), k1::kN / d->k1_cluster,
                                                      (float)(g / n2), reinterpret_cast<k1::cf32*>(w.tiles), w.mpack, w.mhere, w.mcode,
                                                      w.progress, w.P * 16 + 16, w.cf, (float)(0.5 * g / n2), (float)(g / n2));
        if (cudaGetLastError() != cudaSuccess) { *rc = fail(PNPADMM_ERR_CUDA, "launch of prepare_shared_kernel failed"); return true; }
        k1::ClusterParams cp;
        memset(&cp, 0, sizeof(cp));
        cp.B = B1; cp.P = P1; cp.solo = 0; cp.iters = iters;
        cp.x = x; cp.z = z; cp.w = wv;
        cp.G = reinterpret_cast<const k1::cf32*>(w.Gt); cp.Gw = reinterpret_cast<k1::cf32*>(w.Gt);
        cp.mpack = w.mpack; cp.mcode_batched = 0; cp.progress = w.progress; cp.cf = w.cf; cp.prox = pp;
        cp.fused = 1; cp.img = img; cp.img8 = img ? nullptr : img8;
        cp.tiles = reinterpret_cast<const k1::cf32*>(w.tiles); cp.mhere = w.mhere;
        cp.cf1v = (float)(0.5 * g / n2); cp.cf2v = (float)(g / n2);
        cp.no_memset = 1;
        if (P1 == w.P) { *rc = launch_cluster(cp, d, st); return true; }
        // hybrid: planes [P1, P) = images [B1, B) on the streaming kernels, prologue included, in the side stream
        std::lock_guard<std::mutex> lk(d->mu);
        if (cudaEventRecord(d->fork_ev, st) != cudaSuccess) {
            *rc = fail(PNPADMM_ERR_CUDA, "fork of the hybrid schedule failed: %s", cudaGetErrorString(cudaGetLastError())); return true;
        }
        *rc = launch_cluster(cp, d, st);
        if (*rc) return true;
        const int sms = d->sm_count - 8 * d->max_clusters_256;
        auto share = [&](int pa, int pb, cudaStream_t s2, cudaEvent_t join) -> int {
            const int ba = 2 * pa, bb = (2 * pb < B) ? 2 * pb : B, B2 = bb - ba, P2 = pb - pa;
            CUDA_TRY(cudaStreamWaitEvent(s2, d->fork_ev, 0));
            float* x2 = x + (size_t)ba * nn; float* z2 = z + (size_t)ba * nn; float* w2v = wv + (size_t)ba * nn;
            const float* img2 = img ? img + (size_t)ba * nn : nullptr;
            if (!img2) {   // uint8 input: the streaming kernels want float images; x2 is free until the iterations write it
                u8_to_unit_kernel<float><<<grid_1d((size_t)B2 * nn, sms), 256, 0, s2>>>(img8 + (size_t)ba * nn, x2, (size_t)B2 * nn);
                img2 = x2;
            }
            StreamParams<float> p = base_params(w, B2, N);
            p.P = P2;
            cx<float>* T1 = w.T1 + (size_t)ba * nn; cx<float>* Y2 = w.Y + (size_t)ba * nn;
            p.img = img2; p.cout = T1;
            S2<float>::template rows<RM_FWD_IMG>(p, B2, s2);
            p.cin = T1; p.cout = Y2; p.mask = mask; p.mask_batched = 0;
            p.noise = reinterpret_cast<const cx<float>*>(noise); p.noise_batched = 0;
            S2<float>::template cols<CM_FWD_ACQ>(p, B2, nullptr, sms, s2);
            p.cin = Y2; p.cout = T1;
            S2<float>::template cols<CM_INV>(p, B2, nullptr, sms, s2);
            p.cin = T1; p.x = z2;
            S2<float>::template rows<RM_INV_ABS>(p, B2, s2);
            copy_zero_kernel<float><<<grid_1d((size_t)B2 * nn, sms), 256, 0, s2>>>(z2, nullptr, w2v, (size_t)B2 * nn);
            const size_t total = (size_t)P2 * nn;
            prepare_kernel<float><<<(unsigned)((total + 255) / 256), 256, 0, s2>>>(Y2, mask, w.G + (size_t)pa * nn, nullptr, B2, P2, N, 0, 0,
                                                                                (float)(g / n2), nullptr, 0);
            LAUNCH_CHECK("streaming prologue of the hybrid share");
            Workspace<float> w2 = w;
            w2.P = P2; w2.K = w.K + (size_t)pa * nn; w2.G = w.G + (size_t)pa * nn;
            const int r = stream2_iterate<float>(w2, x2, z2, w2v, B2, N, pp, iters, sms, s2, /*locked=*/true);
            if (r) return r;
            CUDA_TRY(cudaEventRecord(join, s2));
            CUDA_TRY(cudaStreamWaitEvent(st, join, 0));
            return PNPADMM_OK;
        };
        const int P2 = w.P - P1;
        if (k2_split_enabled() && P2 >= 2) {
            const int pm = P1 + (P2 + 1) / 2;
            *rc = share(P1, pm, d->side, d->join_ev);
            if (*rc == PNPADMM_OK) *rc = share(pm, w.P, d->side2, d->join_ev2);
        } else {
            *rc = share(P1, w.P, d->side, d->join_ev);
        }
        return true;
    }
};

template <typename T>
int reconstruct_impl(const T* img, const uint8_t* img8, const uint8_t* mask, const T* noise, T* x, T* z, T* wv, int B, int N,
                     int mask_batched, int noise_batched, int prox, int iters, double lambda1, double reo, double alpha, double b,
                     int kernel, void* ws, size_t ws_bytes, cudaStream_t st) {
    if ((!img && !img8) || !mask || !noise || !x || !z || !wv || B <= 0) return fail(PNPADMM_ERR_BAD_ARG, "reconstruct: NULL pointer or B <= 0");
    int rc = check_n(N, sizeof(T) == 8); if (rc) return rc;
    rc = check_prox(prox, iters, reo, b); if (rc) return rc;
    DeviceState* d; rc = ensure_device(&d); if (rc) return rc;
    Workspace<T> w; rc = carve<T>(ws, ws_bytes, B, N, mask_batched, &w); if (rc) return rc;
    bool use_cluster; rc = pick_kernel(kernel, N, sizeof(T) == 8, d, &use_cluster, true); if (rc) return rc;
    if (kernel == PNPADMM_KERNEL_ROWSEP && (mask_batched || noise_batched))
        return fail(PNPADMM_ERR_UNSUPPORTED, "the row-separable kernel needs one mask and one noise array for the batch");
    if ((use_cluster || kernel == PNPADMM_KERNEL_ROWSEP) && !mask_batched && !noise_batched &&
        FusedDispatch<T>::run(&rc, w, img, img8, mask, noise, x, z, wv, B, N, prox, iters, lambda1, reo, alpha, b, kernel, d, st))
        return rc;
    const size_t n = (size_t)B * N * N;
    if (!img) {   // uint8 gray levels -> unit scale; x is free until the solve writes it
        u8_to_unit_kernel<T><<<grid_1d(n, d->sm_count), 256, 0, st>>>(img8, x, n);
        LAUNCH_CHECK("u8_to_unit_kernel");
        img = x;
    }
    T* y = reinterpret_cast<T*>(w.Y);
    rc = acquire_impl<T>(img, mask, noise, y, B, N, mask_batched, noise_batched, sizeof(T) == 8 ? 1 : 0, ws, ws_bytes, st); if (rc) return rc;
    return solve_impl<T>(y, mask, x, z, wv, B, N, mask_batched, prox, iters, lambda1, reo, alpha, b, kernel, ws, ws_bytes, st);
}

template <typename T>
int soft_impl(const T* x, T* out, double c, size_t n, cudaStream_t st) {
    if (!x || !out) return fail(PNPADMM_ERR_BAD_ARG, "soft: NULL pointer");
    if (n == 0) return PNPADMM_OK;
    DeviceState* d; int rc = ensure_device(&d); if (rc) return rc;
    soft_kernel<T><<<grid_1d(n, d->sm_count), 256, 0, st>>>(x, out, (T)c, n);
    LAUNCH_CHECK("soft_kernel");
    return PNPADMM_OK;
}

template <typename T>
int combine_impl(const T* z, const T* x, const T* w, const T* s, T* t, double alpha, double coef, size_t n,
                 cudaStream_t st) {
    if (!z || !x || !w || !s || !t) return fail(PNPADMM_ERR_BAD_ARG, "cnc_combine: NULL pointer");
    if (n == 0) return PNPADMM_OK;
    DeviceState* d; int rc = ensure_device(&d); if (rc) return rc;
    cnc_combine_kernel<T><<<grid_1d(n, d->sm_count), 256, 0, st>>>(z, x, w, s, t, (T)(1.0 - alpha), (T)alpha, (T)coef, n);
    LAUNCH_CHECK("cnc_combine_kernel");
    return PNPADMM_OK;
}

template <typename T>
int dual_impl(T* x, T* z, T* w, int clamp, size_t n, cudaStream_t st) {
    if (!x || !z || !w) return fail(PNPADMM_ERR_BAD_ARG, "dual_update: NULL pointer");
    if (n == 0) return PNPADMM_OK;
    DeviceState* d; int rc = ensure_device(&d); if (rc) return rc;
    dual_update_kernel<T><<<grid_1d(n, d->sm_count), 256, 0, st>>>(x, z, w, clamp, n);
    LAUNCH_CHECK("dual_update_kernel");
    return PNPADMM_OK;
}

template <typename T>
int metrics_impl(const T* x, const uint8_t* ref, int B, int N, int quantize, double* out, void* scratch,
                 size_t scratch_bytes, cudaStream_t st) {
    if (!x || !ref || !out || B <= 0) return fail(PNPADMM_ERR_BAD_ARG, "metrics: NULL pointer or B <= 0");
    if (N < 16 || N > 4096) return fail(PNPADMM_ERR_BAD_SIZE, "metrics: N=%d unsupported (16 <= N <= 4096)", N);
    if (!scratch || scratch_bytes < sizeof(MetricsAcc) * (size_t)B || ((uintptr_t)scratch & 15))
        return fail(PNPADMM_ERR_WORKSPACE, "metrics: scratch must be 16-byte aligned and hold %zu bytes", sizeof(MetricsAcc) * (size_t)B);
    DeviceState* d; int rc = ensure_device(&d); if (rc) return rc;
    MetricsAcc* acc = (MetricsAcc*)scratch;
    CUDA_TRY(cudaMemsetAsync(acc, 0, sizeof(MetricsAcc) * (size_t)B, st));
    const int tiles = (N + kMetTile - 1) / kMetTile;
    metrics_tile_kernel<T><<<dim3(tiles, tiles, B), 256, kMetSmemBytes, st>>>(x, ref, N, quantize, acc);
    LAUNCH_CHECK("metrics_tile_kernel");
    metrics_finalize_kernel<<<(B + 127) / 128, 128, 0, st>>>(acc, out, B, N);
    LAUNCH_CHECK("metrics_finalize_kernel");
    return PNPADMM_OK;
}

// img_E = saturate_cast<uint8>(255 x), rounding half to even like cv2.imwrite of the reference's float image (S1:133-138)
// and np.uint8((clip(x, 0, 1) * 255).round()) of util.single2uint (S6:315)
__global__ void unit_to_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = fminf(fmaxf(rintf(255.0f * x[i]), 0.f), 255.f);
        out[i] = (uint8_t)v;
    }
}

__global__ void __launch_bounds__(256) debug_copy_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// FP32 FMA throughput probe: 8 independent FMA chains per thread.
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters, float b, float c) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ------------------------------------------------------------------------------------------
// K5: DnCNN / FDnCNN forward on the tensor cores (dncnn_tc.cuh)
// ------------------------------------------------------------------------------------------
// Work decomposition of one conv64 launch: B x xtiles x dil x ystrips items on `sms` persistent CTAs.  A strip of R output rows costs
// R + 2 input rows (the halo), and the launch takes ceil(items / sms) strips per CTA: `ystrips` minimises waves x (R + 2).  Big
// batches end up with long strips (DnCNN, B = 256, 256^2: 2 strips of 128 rows, 7 waves); small ones are cut until every SM has
// work (B = 1: 37 strips of 7 rows instead of 4 of 64 on 8 SMs); FFDNet's half-resolution layers at B = 256 take 4 strips of 32
// rows (7 waves x 34) instead of 2 of 64 (3.46 -> 4 waves x 66).
int conv_geometry(tc::ConvParams& p, int B, int H, int W, const char* who, int sms, int dil = 1) {
    if (B <= 0 || H <= 0 || W <= 0) return fail(PNPADMM_ERR_BAD_ARG, "%s: B=%d H=%d W=%d must be positive", who, B, H, W);
    if ((long long)B * H * W >= (1ll << 31)) return fail(PNPADMM_ERR_BAD_SIZE, "%s: B*H*W = %lld pixels exceeds 2^31", who, (long long)B * H * W);
    if (dil < 1 || dil > tc::kMaxDil) return fail(PNPADMM_ERR_UNSUPPORTED, "%s: dilation %d (1..%d)", who, dil, tc::kMaxDil);
    if (H < dil) return fail(PNPADMM_ERR_BAD_SIZE, "%s: H=%d is smaller than the dilation %d", who, H, dil);
    p.B = B; p.H = H; p.W = W;
    p.kchunks = 8; p.cout = 1; p.dil = dil;
    p.xtiles = (W + tc::kTileM - 1) / tc::kTileM;
    const int sub_max = (H + dil - 1) / dil, sub_min = H / dil;      // rows of the largest / smallest row sub-image (ConvParams::dil)
    const long long per_strip = (long long)B * p.xtiles * dil;
    long long best = -1;
    for (int ys = 1; ys <= sub_min && ys <= 64; ++ys) {             // ys <= sub_min: no strip of any sub-image is empty (tc::decode_item)
        const int rows = (sub_max + ys - 1) / ys;
        if (rows < 4 && ys > 1) break;                                // below four rows the halo alone is a third of the work
        const long long waves = (per_strip * ys + sms - 1) / sms, cost = waves * (rows + 2);
        if (best < 0 || cost < best) { best = cost; p.ystrips = ys; }
    }
    if (per_strip * p.ystrips >= (1ll << 31)) return fail(PNPADMM_ERR_BAD_SIZE, "%s: too many work items", who);
    p.items = (int)(per_strip * p.ystrips);
    return PNPADMM_OK;
}

// A layer that follows another kernel of the same forward is launched with programmatic stream serialisation (see the kernel);
// PNPADMM_NO_PDL=1 gives plain launches.
template <int NOUT>
void launch_conv(const tc::ConvParams& p, int grid, cudaStream_t st, bool dependent) {
    static const bool pdl_on = [] { const char* e = getenv("PNPADMM_NO_PDL"); return !(e && atoi(e) != 0); }();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(tc::kThreads); cfg.dynamicSmemBytes = tc::kSmemBytes; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = (dependent && pdl_on) ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, tc::conv64_tc_kernel<NOUT>, p);
}

int conv64_impl(const void* in, void* out, const void* w, const float* bias, int B, int H, int W, int relu, cudaStream_t st, int dil = 1) {
    if (!in || !out || !w || !bias) return fail(PNPADMM_ERR_BAD_ARG, "conv64: NULL pointer");
    if ((((uintptr_t)in) | ((uintptr_t)out) | ((uintptr_t)w)) & 15) return fail(PNPADMM_ERR_BAD_ARG, "conv64: pointers must be 16-byte aligned");
    DeviceState* d; int rc = ensure_device(&d); if (rc) return rc;
    tc::ConvParams p{};
    rc = conv_geometry(p, B, H, W, "conv64", d->sm_count, dil); if (rc) return rc;
    p.in = static_cast<const __nv_bfloat16*>(in); p.out = static_cast<__nv_bfloat16*>(out);
    p.w = w; p.bias = bias; p.relu = relu;
    if (const char* e = getenv("PNPADMM_TC_DEBUG")) p.dbg = atoi(e);
    const int grid = p.items < d->sm_count ? p.items : d->sm_count;
    launch_conv<64>(p, grid, st, false);
    LAUNCH_CHECK("conv64_tc_kernel<64>");
    return PNPADMM_OK;
}

int dncnn_forward_impl(const float* x, float* out, int B, int cin, int H, int W, int n_mid, const float* w_head,
                       const float* b_head, const void* w_mid, const float* b_mid, const void* w_tail, const float* b_tail,
                       int residual, void* act0, void* act1, cudaStream_t st, const int* dil_mid = nullptr) {
    if (!x || !out || !w_head || !b_head || !w_tail || !b_tail || !act0 || !act1 || (n_mid > 0 && (!w_mid || !b_mid)))
        return fail(PNPADMM_ERR_BAD_ARG, "dncnn_forward: NULL pointer");
    if (cin != 1 && cin != 2) return fail(PNPADMM_ERR_UNSUPPORTED, "dncnn_forward: %d input channels (1 = DnCNN, 2 = FDnCNN)", cin);
    if (n_mid < 0) return fail(PNPADMM_ERR_BAD_ARG, "dncnn_forward: n_mid=%d", n_mid);
    if ((((uintptr_t)act0) | ((uintptr_t)act1) | ((uintptr_t)w_mid) | ((uintptr_t)w_tail)) & 15)
        return fail(PNPADMM_ERR_BAD_ARG, "dncnn_forward: activation / weight pointers must be 16-byte aligned");
    DeviceState* d; int rc = ensure_device(&d); if (rc) return rc;
    tc::ConvParams p{};
    rc = conv_geometry(p, B, H, W, "dncnn_forward", d->sm_count); if (rc) return rc;
    __nv_bfloat16* act[2] = {static_cast<__nv_bfloat16*>(act0), static_cast<__nv_bfloat16*>(act1)};
    const unsigned hgrid = (unsigned)B * ((W + 31) / 32) * ((H + tc::kHeadRows - 1) / tc::kHeadRows);
    if (cin == 1) tc::dncnn_head_kernel<1><<<hgrid, 256, 0, st>>>(x, act[0], w_head, b_head, B, H, W);
    else tc::dncnn_head_kernel<2><<<hgrid, 256, 0, st>>>(x, act[0], w_head, b_head, B, H, W);
    LAUNCH_CHECK("dncnn_head_kernel");
    const int grid = p.items < d->sm_count ? p.items : d->sm_count;
    p.relu = 1;
    if (const char* e = getenv("PNPADMM_TC_DEBUG")) p.dbg = atoi(e);     // timing experiments only
    for (int l = 0; l < n_mid; ++l) {
        if (dil_mid) {                         // IRCNN: per-layer dilation (work items are strips of row sub-images)
            rc = conv_geometry(p, B, H, W, "dncnn_forward", d->sm_count, dil_mid[l]); if (rc) return rc;
        }
        const int g = p.items < d->sm_count ? p.items : d->sm_count;
        p.in = act[l & 1]; p.out = act[(l + 1) & 1];
        p.w = static_cast<const unsigned char*>(w_mid) + (size_t)l * tc::kWBytesMax;
        p.bias = b_mid + 64 * l;
        launch_conv<64>(p, g, st, true);
    }
    LAUNCH_CHECK("conv64_tc_kernel<64>");
    if (dil_mid) { rc = conv_geometry(p, B, H, W, "dncnn_forward", d->sm_count, 1); if (rc) return rc; }
    p.in = act[n_mid & 1]; p.out = nullptr; p.out_f32 = out;
    p.w = w_tail; p.bias = b_tail; p.relu = 0;
    p.resid = residual ? x : nullptr;          // channel 0 of x
    p.resid_bstride = (long long)cin * H * W;
    launch_conv<16>(p, grid, st, true);
    LAUNCH_CHECK("conv64_tc_kernel<16>");
    return PNPADMM_OK;
}

// FFDNet forward (reference models/network_ffdnet.py:31-73, called from S3:64-66): pack (pad, pixel-unshuffle, noise-level map) ->
// thin first layer (5 real input channels: K = 16 per tap) -> n_mid x conv64 -> four-channel tail with the pixel shuffle in its
// epilogue.  All convolutions run on conv64_tc_kernel at half resolution.
int ffdnet_forward_impl(const float* x, float* out, int B, int H, int W, float sigma, int n_mid, const void* w_head, const float* b_head,
                        const void* w_mid, const float* b_mid, const void* w_tail, const float* b_tail, void* act0, void* act1,
                        cudaStream_t st) {
    if (!x || !out || !w_head || !b_head || !w_tail || !b_tail || !act0 || !act1 || (n_mid > 0 && (!w_mid || !b_mid)))
        return fail(PNPADMM_ERR_BAD_ARG, "ffdnet_forward: NULL pointer");
    if (n_mid < 0) return fail(PNPADMM_ERR_BAD_ARG, "ffdnet_forward: n_mid=%d", n_mid);
    if ((((uintptr_t)act0) | ((uintptr_t)act1) | ((uintptr_t)w_head) | ((uintptr_t)w_mid) | ((uintptr_t)w_tail)) & 15)
        return fail(PNPADMM_ERR_BAD_ARG, "ffdnet_forward: activation / weight pointers must be 16-byte aligned");
    if (((uintptr_t)out) & 7) return fail(PNPADMM_ERR_BAD_ARG, "ffdnet_forward: out must be 8-byte aligned");
    if (B <= 0 || H <= 0 || W <= 0) return fail(PNPADMM_ERR_BAD_ARG, "ffdnet_forward: B=%d H=%d W=%d must be positive", B, H, W);
    DeviceState* d; int rc = ensure_device(&d); if (rc) return rc;
    const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
    tc::ConvParams p{};
    rc = conv_geometry(p, B, H2, W2, "ffdnet_forward", d->sm_count); if (rc) return rc;
    __nv_bfloat16* act[2] = {static_cast<__nv_bfloat16*>(act0), static_cast<__nv_bfloat16*>(act1)};
    const size_t npx = (size_t)B * H2 * W2;
    tc::ffdnet_pack_kernel<<<(unsigned)((npx + 255) / 256), 256, 0, st>>>(x, act[1], sigma, B, H, W, H2, W2);
    LAUNCH_CHECK("ffdnet_pack_kernel");
    const int grid = p.items < d->sm_count ? p.items : d->sm_count;
    if (const char* e = getenv("PNPADMM_TC_DEBUG")) p.dbg = atoi(e);     // timing experiments only
    p.relu = 1;
    p.in = act[1]; p.out = act[0]; p.w = w_head; p.bias = b_head; p.kchunks = 2;
    launch_conv<64>(p, grid, st, false);               // after the pack kernel (an ordinary kernel without the early trigger)
    p.kchunks = 8;
    for (int l = 0; l < n_mid; ++l) {
        p.in = act[l & 1]; p.out = act[(l + 1) & 1];
        p.w = static_cast<const unsigned char*>(w_mid) + (size_t)l * tc::kWBytesMax;
        p.bias = b_mid + 64 * l;
        launch_conv<64>(p, grid, st, true);
    }
    LAUNCH_CHECK("conv64_tc_kernel<64>");
    p.in = act[n_mid & 1]; p.out = nullptr; p.out_f32 = out;
    p.w = w_tail; p.bias = b_tail; p.relu = 0; p.resid = nullptr;
    p.cout = 4; p.out_H = H; p.out_W = W;
    launch_conv<16>(p, grid, st, true);
    LAUNCH_CHECK("conv64_tc_kernel<16>");
    return PNPADMM_OK;
}

}  // namespace

// ==========================================================================================
// extern "C" surface
// ==========================================================================================
extern "C" {

int pnpadmm_abi_version(void) { return PNPADMM_ABI_VERSION; }
const char* pnpadmm_last_error_string(void) { return g_err; }

int pnpadmm_device_info(int* sm_count, int* max_clusters_256, int* cc_major, int* cc_minor) {
    DeviceState* d; int rc = ensure_device(&d); if (rc) return rc;
    if (sm_count) *sm_count = d->sm_count;
    if (max_clusters_256) *max_clusters_256 = d->max_clusters_256;
    if (cc_major) *cc_major = d->cc_major;
    if (cc_minor) *cc_minor = d->cc_minor;
    return PNPADMM_OK;
}

// Kernel launches of the calls below, kept next to the code that makes them (bench.py reports them as gpu_launches):
//   acquire: rows<FWD_IMG> + cols<FWD_ACQ>;  zero_filled: cols<INV> + rows<INV_ABS>;  solve adds copy_zero, write_cf,
//   prepare, pack_mcode (fp32, N in {256, 512, 1024}), then the iteration kernels.
int pnpadmm_plan_info(int B, int N, int mask_batched, int iters, int kernel, int* planes_cluster, int* planes_streaming,
                      int* chunks, int* launches_acquire, int* launches_solve, int* launches_reconstruct) {
    if (B <= 0 || iters < 0) return fail(PNPADMM_ERR_BAD_ARG, "plan_info: B <= 0 or iters < 0");
    int rc = check_n(N, false); if (rc) return rc;
    DeviceState* d; rc = ensure_device(&d); if (rc) return rc;
    bool use_cluster; rc = pick_kernel(kernel, N, false, d, &use_cluster); if (rc) return rc;
    const int P = mask_batched ? B : (B + 1) / 2;
    int p1 = 0, p2 = P, nch = 1;
    if (use_cluster) {
        p1 = (kernel == PNPADMM_KERNEL_AUTO && S2<float>::ok(N)) ? plan_hybrid(d, P, iters, !mask_batched && iters >= 1) : P;
        p2 = P - p1;
        int chunk; plan_chunks(p1, iters, d->max_clusters_256, &chunk, &nch);
    }
    const bool s2 = S2<float>::ok(N);
    int it_launches = 0;
    if (iters > 0) it_launches = (p1 > 0 ? 1 : 0) + (p2 > 0 ? 1 + 2 * iters : 0);
    if (planes_cluster) *planes_cluster = p1;
    if (planes_streaming) *planes_streaming = p2;
    if (chunks) *chunks = nch;
    if (launches_acquire) *launches_acquire = 2;
    const int solve = 2 /* zero-fill */ + 1 /* copy_zero */ + 2 /* write_cf, prepare */ + (s2 ? 1 : 0) /* pack */ + it_launches;
    if (launches_solve) *launches_solve = solve;
    if (launches_reconstruct) {
        // fused prologue (N == 256, fp32, shared mask and noise): prepare_shared + the cluster kernel; the streaming share of
        // the hybrid schedule adds rows<FWD_IMG>, cols<FWD_ACQ>, cols<INV>, rows<INV_ABS>, copy_zero, prepare + its iterations
        const bool fused = use_cluster && !mask_batched && iters >= 1 && getenv("PNPADMM_NO_FUSED_PROLOGUE") == nullptr;
        *launches_reconstruct = fused ? 1 + (p1 > 0 ? 1 : 0) + (p2 > 0 ? 6 + 1 + 2 * iters : 0) : 2 + solve;
    }
    return PNPADMM_OK;
}

size_t pnpadmm_workspace_bytes(int B, int N, int is_f64, int mask_batched) {
    if (B <= 0 || N <= 0) return 0;
    return ws_bytes_impl(B, N, is_f64 ? 8 : 4, mask_batched);
}

#define ST(s) ((cudaStream_t)(s))

int pnpadmm_acquire_f32(const float* img, const uint8_t* mask, const float* noise, float* y, int B, int N,
                        int mask_batched, int noise_batched, void* ws, size_t wsb, pnpadmm_stream_t s) {
    return acquire_impl<float>(img, mask, noise, y, B, N, mask_batched, noise_batched, 0, ws, wsb, ST(s));
}
int pnpadmm_acquire_f64(const double* img, const uint8_t* mask, const double* noise, double* y, int B, int N,
                        int mask_batched, int noise_batched, int spectrum_f32, void* ws, size_t wsb,
                        pnpadmm_stream_t s) {
    return acquire_impl<double>(img, mask, noise, y, B, N, mask_batched, noise_batched, spectrum_f32, ws, wsb, ST(s));
}
int pnpadmm_zero_filled_f32(const float* y, float* x0, int B, int N, void* ws, size_t wsb, pnpadmm_stream_t s) {
    return zero_filled_impl<float>(y, x0, B, N, ws, wsb, ST(s));
}
int pnpadmm_zero_filled_f64(const double* y, double* x0, int B, int N, void* ws, size_t wsb, pnpadmm_stream_t s) {
    return zero_filled_impl<double>(y, x0, B, N, ws, wsb, ST(s));
}
int pnpadmm_prepare_f32(const float* y, const uint8_t* mask, int B, int N, int mb, double reo, void* ws, size_t wsb,
                        pnpadmm_stream_t s) {
    return prepare_impl<float>(y, mask, B, N, mb, reo, ws, wsb, ST(s));
}
int pnpadmm_prepare_f64(const double* y, const uint8_t* mask, int B, int N, int mb, double reo, void* ws, size_t wsb,
                        pnpadmm_stream_t s) {
    return prepare_impl<double>(y, mask, B, N, mb, reo, ws, wsb, ST(s));
}
int pnpadmm_xupdate_f32(const float* z, const float* w, float* x, float* xpw, int B, int N, int mb, int kernel,
                        void* ws, size_t wsb, pnpadmm_stream_t s) {
    return xupdate_impl<float>(z, w, x, xpw, B, N, mb, kernel, ws, wsb, ST(s));
}
int pnpadmm_xupdate_f64(const double* z, const double* w, double* x, double* xpw, int B, int N, int mb, int kernel,
                        void* ws, size_t wsb, pnpadmm_stream_t s) {
    return xupdate_impl<double>(z, w, x, xpw, B, N, mb, kernel, ws, wsb, ST(s));
}
int pnpadmm_iterate_f32(float* x, float* z, float* w, int B, int N, int mb, int prox, int iters, double lambda1,
                        double reo, double alpha, double b, int kernel, void* ws, size_t wsb, pnpadmm_stream_t s) {
    return iterate_impl<float>(x, z, w, B, N, mb, prox, iters, lambda1, reo, alpha, b, kernel, ws, wsb, ST(s));
}
int pnpadmm_iterate_f64(double* x, double* z, double* w, int B, int N, int mb, int prox, int iters, double lambda1,
                        double reo, double alpha, double b, int kernel, void* ws, size_t wsb, pnpadmm_stream_t s) {
    return iterate_impl<double>(x, z, w, B, N, mb, prox, iters, lambda1, reo, alpha, b, kernel, ws, wsb, ST(s));
}
int pnpadmm_solve_f32(const float* y, const uint8_t* mask, float* x, float* z, float* w, int B, int N, int mb,
                      int prox, int iters, double lambda1, double reo, double alpha, double b, int kernel, void* ws,
                      size_t wsb, pnpadmm_stream_t s) {
    return solve_impl<float>(y, mask, x, z, w, B, N, mb, prox, iters, lambda1, reo, alpha, b, kernel, ws, wsb, ST(s));
}
int pnpadmm_solve_f64(const double* y, const uint8_t* mask, double* x, double* z, double* w, int B, int N, int mb,
                      int prox, int iters, double lambda1, double reo, double alpha, double b, int kernel, void* ws,
                      size_t wsb, pnpadmm_stream_t s) {
    return solve_impl<double>(y, mask, x, z, w, B, N, mb, prox, iters, lambda1, reo, alpha, b, kernel, ws, wsb, ST(s));
}

int pnpadmm_reconstruct_f32(const float* img, const uint8_t* img8, const uint8_t* mask, const float* noise, float* x, float* z,
                            float* w, int B, int N, int mb, int nb, int prox, int iters, double lambda1, double reo, double alpha,
                            double b, int kernel, void* ws, size_t wsb, pnpadmm_stream_t s) {
    return reconstruct_impl<float>(img, img8, mask, noise, x, z, w, B, N, mb, nb, prox, iters, lambda1, reo, alpha, b, kernel, ws, wsb, ST(s));
}
int pnpadmm_reconstruct_f64(const double* img, const uint8_t* img8, const uint8_t* mask, const double* noise, double* x, double* z,
                            double* w, int B, int N, int mb, int nb, int prox, int iters, double lambda1, double reo, double alpha,
                            double b, int kernel, void* ws, size_t wsb, pnpadmm_stream_t s) {
    return reconstruct_impl<double>(img, img8, mask, noise, x, z, w, B, N, mb, nb, prox, iters, lambda1, reo, alpha, b, kernel, ws, wsb, ST(s));
}

namespace {
// full k-space lines?  (host mask of the *_host entry points: 255 row compares of N bytes)
bool host_mask_row_separable(const uint8_t* m, int N) {
    for (int r = 1; r < N; ++r)
        for (int c = 0; c < N; ++c)
            if ((m[(size_t)r * N + c] != 0) != (m[c] != 0)) return false;
    return true;
}
int host_auto_kernel(int kernel, const uint8_t* h_mask, int N) {
    static const bool off = [] { const char* e = getenv("PNPADMM_NO_ROWSEP"); return e && atoi(e) != 0; }();
    return (!off && kernel == PNPADMM_KERNEL_AUTO && (N == 256 || N == 512 || N == 1024) && host_mask_row_separable(h_mask, N)) ? PNPADMM_KERNEL_ROWSEP : kernel;
}
}  // namespace

size_t pnpadmm_host_scratch_bytes(int B, int N) {
    if (B <= 0 || N <= 0) return 0;
    const size_t nn = (size_t)N * N, n = (size_t)B * nn;
    // img u8, mask u8, noise c64, img f32, y c64, x, z, w f32
    return align_up(n) + align_up(nn) + align_up(nn * 8) + align_up(n * 4) + align_up(n * 8) + 3 * align_up(n * 4);
}

int pnpadmm_reconstruct_host_f32(const uint8_t* h_img, const uint8_t* h_mask, const float* h_noise, float* h_x, int B,
                                 int N, int prox, int iters, double lambda1, double reo, double alpha, double b,
                                 int kernel, void* d_scratch, size_t scratch_bytes, void* ws, size_t wsb,
                                 pnpadmm_stream_t s) {
    if (!h_img || !h_mask || !h_noise || !h_x || B <= 0) return fail(PNPADMM_ERR_BAD_ARG, "reconstruct_host: NULL pointer or B <= 0");
    int rc = check_n(N, false); if (rc) return rc;
    if (!d_scratch || ((uintptr_t)d_scratch) % kAlign || scratch_bytes < pnpadmm_host_scratch_bytes(B, N))
        return fail(PNPADMM_ERR_WORKSPACE, "reconstruct_host: d_scratch NULL, misaligned or smaller than %zu bytes",
                    pnpadmm_host_scratch_bytes(B, N));
    DeviceState* d; rc = ensure_device(&d); if (rc) return rc;
    cudaStream_t st = ST(s);
    const size_t nn = (size_t)N * N, n = (size_t)B * nn;
    unsigned char* p = (unsigned char*)d_scratch;
    uint8_t* d_img8 = p; p += align_up(n);
    uint8_t* d_mask = p; p += align_up(nn);
    float* d_noise = (float*)p; p += align_up(nn * 8);
    float* d_img = (float*)p; p += align_up(n * 4);
    float* d_y = (float*)p; p += align_up(n * 8);
    float* d_x = (float*)p; p += align_up(n * 4);
    float* d_z = (float*)p; p += align_up(n * 4);
    float* d_w = (float*)p;
    CUDA_TRY(cudaMemcpyAsync(d_img8, h_img, n, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_mask, h_mask, nn, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_noise, h_noise, nn * 8, cudaMemcpyHostToDevice, st));
    (void)d_img; (void)d_y;   // kept in the scratch layout; the reconstruct path reads the uint8 images directly
    kernel = host_auto_kernel(kernel, h_mask, N);
    rc = reconstruct_impl<float>(nullptr, d_img8, d_mask, d_noise, d_x, d_z, d_w, B, N, 0, 0, prox, iters, lambda1, reo, alpha, b, kernel,
                                 ws, wsb, st);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(h_x, d_x, n * 4, cudaMemcpyDeviceToHost, st));
    return PNPADMM_OK;
}

// Pipelined variant: copies on their own streams, n_slots device slots, ordering by the events of a caller-owned
// pipeline object (pnpadmm_pipeline_create), so any number of pipelines can share a device.
size_t pnpadmm_host_pipeline_scratch_bytes(int B, int N, int n_slots) {
    if (B <= 0 || N <= 0 || n_slots < 2 || n_slots > PNPADMM_PIPELINE_MAX_SLOTS) return 0;
    const size_t nn = (size_t)N * N, n = (size_t)B * nn;
    // per slot: img u8, mask u8, noise c64, x f32;  shared by all slots (compute stream only): img f32, y c64, z, w
    return n_slots * (align_up(n) + align_up(nn) + align_up(nn * 8) + align_up(n * 4)) + align_up(n * 4) + align_up(n * 8) +
           2 * align_up(n * 4);
}

}  // extern "C"

namespace {
struct PipeKey {
    void* d_scratch; void* ws; size_t scratch_bytes, wsb;
    double lambda1, reo, alpha, b;
    int B, N, prox, iters, kernel, slot, n_slots, out_format;
};
struct PipeGraph {
    bool valid = false;
    PipeKey key;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    unsigned long long stamp = 0;
};
constexpr int kPipeGraphs = 8;
}  // namespace

// The object behind pnpadmm_pipeline_t: the ordering events of one pipeline and the captured compute sections
// (u8 -> unit, acquisition, zero-fill, prepare, iterations: one graph launch per step instead of ~25 driver calls).
struct pnpadmm_pipeline_s {
    int dev = 0, n_slots = 2;
    int out_format = PNPADMM_OUT_F32;
    cudaEvent_t in_ready[PNPADMM_PIPELINE_MAX_SLOTS], done[PNPADMM_PIPELINE_MAX_SLOTS], out_done[PNPADMM_PIPELINE_MAX_SLOTS];
    bool used[PNPADMM_PIPELINE_MAX_SLOTS];
    PipeGraph graphs[kPipeGraphs];
    PipeKey last_key[PNPADMM_PIPELINE_MAX_SLOTS];     // configuration of the previous call per slot
    bool have_last[PNPADMM_PIPELINE_MAX_SLOTS];
    unsigned long long clock = 0;
    std::mutex mu;
};

extern "C" {

int pnpadmm_pipeline_create(pnpadmm_pipeline_t* out, int n_slots) {
    if (!out) return fail(PNPADMM_ERR_BAD_ARG, "pipeline_create: NULL pointer");
    if (n_slots < 2 || n_slots > PNPADMM_PIPELINE_MAX_SLOTS)
        return fail(PNPADMM_ERR_BAD_ARG, "pipeline_create: n_slots=%d must be in [2, %d]", n_slots, PNPADMM_PIPELINE_MAX_SLOTS);
    DeviceState* d; int rc = ensure_device(&d); if (rc) return rc;
    pnpadmm_pipeline_s* p = new (std::nothrow) pnpadmm_pipeline_s();
    if (!p) return fail(PNPADMM_ERR_CUDA, "pipeline_create: out of host memory");
    CUDA_TRY(cudaGetDevice(&p->dev));
    p->n_slots = n_slots;
    for (int i = 0; i < PNPADMM_PIPELINE_MAX_SLOTS; ++i) { p->in_ready[i] = p->done[i] = p->out_done[i] = nullptr; p->used[i] = false; p->have_last[i] = false; }
    for (int i = 0; i < n_slots; ++i) {
        if (cudaEventCreateWithFlags(&p->in_ready[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->done[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->out_done[i], cudaEventDisableTiming) != cudaSuccess) {
            const cudaError_t e = cudaGetLastError();
            pnpadmm_pipeline_destroy(p);
            return fail(PNPADMM_ERR_CUDA, "pipeline_create: cudaEventCreate failed: %s", cudaGetErrorString(e));
        }
    }
    *out = p;
    return PNPADMM_OK;
}

int pnpadmm_pipeline_set_output(pnpadmm_pipeline_t p, int format) {
    if (!p) return fail(PNPADMM_ERR_BAD_ARG, "pipeline_set_output: NULL pipeline");
    if (format != PNPADMM_OUT_F32 && format != PNPADMM_OUT_U8) return fail(PNPADMM_ERR_BAD_ARG, "pipeline_set_output: unknown format %d", format);
    std::lock_guard<std::mutex> lk(p->mu);
    p->out_format = format;
    return PNPADMM_OK;
}

int pnpadmm_pipeline_destroy(pnpadmm_pipeline_t p) {
    if (!p) return PNPADMM_OK;
    for (int i = 0; i < PNPADMM_PIPELINE_MAX_SLOTS; ++i) {
        if (p->in_ready[i]) (void)cudaEventDestroy(p->in_ready[i]);
        if (p->done[i]) (void)cudaEventDestroy(p->done[i]);
        if (p->out_done[i]) (void)cudaEventDestroy(p->out_done[i]);
    }
    for (int i = 0; i < kPipeGraphs; ++i)
        if (p->graphs[i].valid) { (void)cudaGraphExecDestroy(p->graphs[i].exec); (void)cudaGraphDestroy(p->graphs[i].graph); }
    (void)cudaGetLastError();
    delete p;
    return PNPADMM_OK;
}

int pnpadmm_reconstruct_host_pipelined_f32(pnpadmm_pipeline_t pipe, const uint8_t* h_img, const uint8_t* h_mask,
                                           const float* h_noise, float* h_x, int B, int N, int prox, int iters, double lambda1,
                                           double reo, double alpha, double b, int kernel, void* d_scratch, size_t scratch_bytes,
                                           void* ws, size_t wsb, int slot, pnpadmm_stream_t compute, pnpadmm_stream_t h2d,
                                           pnpadmm_stream_t d2h) {
    if (!pipe) return fail(PNPADMM_ERR_BAD_ARG, "reconstruct_host_pipelined: NULL pipeline (pnpadmm_pipeline_create)");
    if (!h_img || !h_mask || !h_noise || !h_x || B <= 0) return fail(PNPADMM_ERR_BAD_ARG, "reconstruct_host_pipelined: NULL pointer or B <= 0");
    const int S = pipe->n_slots;
    if (slot < 0 || slot >= S) return fail(PNPADMM_ERR_BAD_ARG, "reconstruct_host_pipelined: slot must be in [0, %d)", S);
    int rc = check_n(N, false); if (rc) return rc;
    rc = check_prox(prox, iters, reo, b); if (rc) return rc;
    if (!d_scratch || ((uintptr_t)d_scratch) % kAlign || scratch_bytes < pnpadmm_host_pipeline_scratch_bytes(B, N, S))
        return fail(PNPADMM_ERR_WORKSPACE, "reconstruct_host_pipelined: d_scratch NULL, misaligned or smaller than %zu bytes",
                    pnpadmm_host_pipeline_scratch_bytes(B, N, S));
    DeviceState* d; rc = ensure_device(&d); if (rc) return rc;
    int dev = 0; CUDA_TRY(cudaGetDevice(&dev));
    if (dev != pipe->dev) return fail(PNPADMM_ERR_BAD_ARG, "reconstruct_host_pipelined: pipeline belongs to device %d, current device is %d", pipe->dev, dev);
    cudaStream_t sc = ST(compute), si = ST(h2d), so = ST(d2h);
    kernel = host_auto_kernel(kernel, h_mask, N);
    const size_t nn = (size_t)N * N, n = (size_t)B * nn;
    unsigned char* p = (unsigned char*)d_scratch;
    uint8_t* d_img8[PNPADMM_PIPELINE_MAX_SLOTS]; uint8_t* d_mask[PNPADMM_PIPELINE_MAX_SLOTS];
    float* d_noise[PNPADMM_PIPELINE_MAX_SLOTS]; float* d_x[PNPADMM_PIPELINE_MAX_SLOTS];
    for (int i = 0; i < S; ++i) {
        d_img8[i] = p; p += align_up(n);
        d_mask[i] = p; p += align_up(nn);
        d_noise[i] = (float*)p; p += align_up(nn * 8);
        d_x[i] = (float*)p; p += align_up(n * 4);
    }
    float* d_img = (float*)p; p += align_up(n * 4);
    float* d_y = (float*)p; p += align_up(n * 8);
    float* d_z = (float*)p; p += align_up(n * 4);
    float* d_w = (float*)p;
    std::lock_guard<std::mutex> lk(pipe->mu);
    // inputs of this slot: free once the compute of the previous call on the slot has finished (done[slot])
    if (pipe->used[slot]) CUDA_TRY(cudaStreamWaitEvent(si, pipe->done[slot], 0));
    CUDA_TRY(cudaMemcpyAsync(d_img8[slot], h_img, n, cudaMemcpyHostToDevice, si));
    CUDA_TRY(cudaMemcpyAsync(d_mask[slot], h_mask, nn, cudaMemcpyHostToDevice, si));
    CUDA_TRY(cudaMemcpyAsync(d_noise[slot], h_noise, nn * 8, cudaMemcpyHostToDevice, si));
    CUDA_TRY(cudaEventRecord(pipe->in_ready[slot], si));
    // compute: needs the inputs, and the slot's output buffer drained by the previous D2H
    CUDA_TRY(cudaStreamWaitEvent(sc, pipe->in_ready[slot], 0));
    if (pipe->used[slot]) CUDA_TRY(cudaStreamWaitEvent(sc, pipe->out_done[slot], 0));
    (void)d_img; (void)d_y;
    const bool out_u8 = pipe->out_format == PNPADMM_OUT_U8;
    auto enqueue_compute = [&]() -> int {
        const int r = reconstruct_impl<float>(nullptr, d_img8[slot], d_mask[slot], d_noise[slot], d_x[slot], d_z, d_w, B, N, 0, 0, prox, iters,
                                              lambda1, reo, alpha, b, kernel, ws, wsb, sc);
        if (r || !out_u8) return r;
        // img_E as the reference saves it (S1:133-138): uint8 gray levels; the slot's image buffer is free after the solve
        unit_to_u8_kernel<<<grid_1d(n, d->sm_count), 256, 0, sc>>>(d_x[slot], d_img8[slot], n);
        LAUNCH_CHECK("unit_to_u8_kernel");
        return PNPADMM_OK;
    };
    // The compute section has fixed arguments per (slot, parameters): replay it as one graph from the second sighting on.
    static const char* nograph = getenv("PNPADMM_NO_GRAPH");
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    bool graphable = !(nograph && atoi(nograph) != 0) && sc != nullptr && sc != cudaStreamLegacy &&
                     cudaStreamIsCapturing(sc, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone;
    (void)cudaGetLastError();
    PipeGraph* g = nullptr;
    if (graphable) {
        PipeKey key;
        memset(&key, 0, sizeof(key));
        key.d_scratch = d_scratch; key.ws = ws; key.scratch_bytes = scratch_bytes; key.wsb = wsb;
        key.lambda1 = lambda1; key.reo = reo; key.alpha = alpha; key.b = b;
        key.B = B; key.N = N; key.prox = prox; key.iters = iters; key.kernel = kernel; key.slot = slot; key.n_slots = S;
        key.out_format = pipe->out_format;
        PipeGraph* victim = nullptr;
        for (int i = 0; i < kPipeGraphs; ++i) {
            PipeGraph& e = pipe->graphs[i];
            if (e.valid && memcmp(&e.key, &key, sizeof(key)) == 0) { g = &e; break; }
            if (!victim || (victim->valid && (!e.valid || e.stamp < victim->stamp))) victim = &e;
        }
        if (!g && pipe->have_last[slot] && memcmp(&pipe->last_key[slot], &key, sizeof(key)) == 0) {
            // second call in a row with this configuration on this slot: capture the section
            g = victim;
            if (g->valid) { (void)cudaGraphExecDestroy(g->exec); (void)cudaGraphDestroy(g->graph); g->valid = false; }
            if (cudaStreamBeginCapture(sc, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                const int r = enqueue_compute();
                cudaGraph_t gr = nullptr;
                const cudaError_t e_end = cudaStreamEndCapture(sc, &gr);
                cudaGraphExec_t ex = nullptr;
                if (r == PNPADMM_OK && e_end == cudaSuccess && gr && cudaGraphInstantiate(&ex, gr, 0) == cudaSuccess) {
                    memcpy(&g->key, &key, sizeof(key)); g->graph = gr; g->exec = ex; g->valid = true;
                } else {
                    if (gr) (void)cudaGraphDestroy(gr);
                    (void)cudaGetLastError();
                    if (r != PNPADMM_OK) return r;
                    g = nullptr;                     // fall back to direct launches
                }
            } else {
                (void)cudaGetLastError();
                g = nullptr;
            }
        }
        memcpy(&pipe->last_key[slot], &key, sizeof(key)); pipe->have_last[slot] = true;
    }
    if (g) {
        g->stamp = ++pipe->clock;
        CUDA_TRY(cudaGraphLaunch(g->exec, sc));
    } else {
        rc = enqueue_compute(); if (rc) return rc;
    }
    CUDA_TRY(cudaEventRecord(pipe->done[slot], sc));
    CUDA_TRY(cudaStreamWaitEvent(so, pipe->done[slot], 0));
    if (out_u8) CUDA_TRY(cudaMemcpyAsync(h_x, d_img8[slot], n, cudaMemcpyDeviceToHost, so));
    else CUDA_TRY(cudaMemcpyAsync(h_x, d_x[slot], n * 4, cudaMemcpyDeviceToHost, so));
    CUDA_TRY(cudaEventRecord(pipe->out_done[slot], so));
    pipe->used[slot] = true;
    return PNPADMM_OK;
}

int pnpadmm_reconstruct_host_wait(pnpadmm_pipeline_t pipe, int slot) {
    if (!pipe) return fail(PNPADMM_ERR_BAD_ARG, "reconstruct_host_wait: NULL pipeline");
    if (slot < 0 || slot >= pipe->n_slots) return fail(PNPADMM_ERR_BAD_ARG, "reconstruct_host_wait: slot must be in [0, %d)", pipe->n_slots);
    if (!pipe->used[slot]) return PNPADMM_OK;   // nothing was ever enqueued on this slot
    CUDA_TRY(cudaEventSynchronize(pipe->out_done[slot]));
    return PNPADMM_OK;
}

int pnpadmm_soft_f32(const float* x, float* out, double c, size_t n, pnpadmm_stream_t s) { return soft_impl<float>(x, out, c, n, ST(s)); }
int pnpadmm_soft_f64(const double* x, double* out, double c, size_t n, pnpadmm_stream_t s) { return soft_impl<double>(x, out, c, n, ST(s)); }
int pnpadmm_cnc_combine_f32(const float* z, const float* x, const float* w, const float* sd, float* t, double alpha,
                            double coef, size_t n, pnpadmm_stream_t s) {
    return combine_impl<float>(z, x, w, sd, t, alpha, coef, n, ST(s));
}
int pnpadmm_cnc_combine_f64(const double* z, const double* x, const double* w, const double* sd, double* t,
                            double alpha, double coef, size_t n, pnpadmm_stream_t s) {
    return combine_impl<double>(z, x, w, sd, t, alpha, coef, n, ST(s));
}
int pnpadmm_dual_update_f32(float* x, float* z, float* w, int clamp01, size_t n, pnpadmm_stream_t s) { return dual_impl<float>(x, z, w, clamp01, n, ST(s)); }
int pnpadmm_dual_update_f64(double* x, double* z, double* w, int clamp01, size_t n, pnpadmm_stream_t s) { return dual_impl<double>(x, z, w, clamp01, n, ST(s)); }

int pnpadmm_metrics_f32(const float* x, const uint8_t* ref, int B, int N, int quantize, double* out, void* scratch,
                        size_t scratch_bytes, pnpadmm_stream_t s) {
    return metrics_impl<float>(x, ref, B, N, quantize, out, scratch, scratch_bytes, ST(s));
}
int pnpadmm_metrics_f64(const double* x, const uint8_t* ref, int B, int N, int quantize, double* out, void* scratch,
                        size_t scratch_bytes, pnpadmm_stream_t s) {
    return metrics_impl<double>(x, ref, B, N, quantize, out, scratch, scratch_bytes, ST(s));
}
size_t pnpadmm_metrics_scratch_bytes(int B) { return B > 0 ? sizeof(MetricsAcc) * (size_t)B : 0; }

int pnpadmm_measure_fp32_peak(double* flops, pnpadmm_stream_t s) {
    if (!flops) return fail(PNPADMM_ERR_BAD_ARG, "measure_fp32_peak: NULL pointer");
    DeviceState* d; int rc = ensure_device(&d); if (rc) return rc;
    cudaStream_t st = ST(s);
    const int blocks = d->sm_count * 8, threads = 256, iters = 4096;
    float* out = nullptr;
    CUDA_TRY(cudaMalloc(&out, sizeof(float) * blocks * threads));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    fma_peak_kernel<<<blocks, threads, 0, st>>>(out, 64, 1.0001f, 0.5f);   // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(e0, st));
        fma_peak_kernel<<<blocks, threads, 0, st>>>(out, iters, 1.0001f, 0.5f);
        CUDA_TRY(cudaEventRecord(e1, st));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    LAUNCH_CHECK("fma_peak_kernel");
    *flops = 2.0 * 64.0 * iters * (double)blocks * threads / (best * 1e-3);
    return PNPADMM_OK;
}

size_t pnpadmm_dncnn_activation_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return align_up((size_t)B * H * W * 64 * 2);
}
int pnpadmm_conv64_bf16(const void* in, void* out, const void* w, const float* bias, int B, int H, int W, int relu,
                        pnpadmm_stream_t s) {
    return conv64_impl(in, out, w, bias, B, H, W, relu, ST(s));
}
int pnpadmm_dncnn_forward_bf16(const float* x, float* out, int B, int cin, int H, int W, int n_mid, const float* w_head,
                               const float* b_head, const void* w_mid, const float* b_mid, const void* w_tail,
                               const float* b_tail, int residual, void* act0, void* act1, pnpadmm_stream_t s) {
    return dncnn_forward_impl(x, out, B, cin, H, W, n_mid, w_head, b_head, w_mid, b_mid, w_tail, b_tail, residual, act0, act1, ST(s));
}

int pnpadmm_conv64_dilated_bf16(const void* in, void* out, const void* w, const float* bias, int B, int H, int W, int relu, int dilation,
                                pnpadmm_stream_t s) {
    return conv64_impl(in, out, w, bias, B, H, W, relu, ST(s), dilation);
}
int pnpadmm_dncnn_forward_dilated_bf16(const float* x, float* out, int B, int cin, int H, int W, int n_mid, const int* dilation_mid,
                                       const float* w_head, const float* b_head, const void* w_mid, const float* b_mid,
                                       const void* w_tail, const float* b_tail, int residual, void* act0, void* act1,
                                       pnpadmm_stream_t s) {
    if (n_mid > 0 && !dilation_mid) return fail(PNPADMM_ERR_BAD_ARG, "dncnn_forward_dilated: dilation_mid is NULL");
    return dncnn_forward_impl(x, out, B, cin, H, W, n_mid, w_head, b_head, w_mid, b_mid, w_tail, b_tail, residual, act0, act1, ST(s),
                              dilation_mid);
}

int pnpadmm_ffdnet_forward_bf16(const float* x, float* out, int B, int H, int W, float sigma, int n_mid, const void* w_head,
                                const float* b_head, const void* w_mid, const float* b_mid, const void* w_tail, const float* b_tail,
                                void* act0, void* act1, pnpadmm_stream_t s) {
    return ffdnet_forward_impl(x, out, B, H, W, sigma, n_mid, w_head, b_head, w_mid, b_mid, w_tail, b_tail, act0, act1, ST(s));
}

// debug only (not part of include/pnpadmm.h): the work decomposition conv_geometry picks for a K5 layer on `sms` SMs (host logic,
// no device needed: tests/test_dncnn_fused.py checks its invariants on the CPU)
int pnpadmm_debug_conv_plan(int B, int H, int W, int dilation, int sms, int* xtiles, int* ystrips, int* items) {
    if (sms <= 0 || !xtiles || !ystrips || !items) return fail(PNPADMM_ERR_BAD_ARG, "debug_conv_plan: bad argument");
    tc::ConvParams p{};
    const int rc = conv_geometry(p, B, H, W, "debug_conv_plan", sms, dilation);
    if (rc) return rc;
    *xtiles = p.xtiles; *ystrips = p.ystrips; *items = p.items;
    return PNPADMM_OK;
}

// debug only (not part of include/pnpadmm.h): copy kernel with 16-byte accesses; `dst` / `src` may be pinned host memory (zero-copy
// over the host link: tools/pcie_probe.py compares it with the copy engines)
int pnpadmm_debug_copy(void* dst, const void* src, size_t bytes, int blocks, pnpadmm_stream_t s) {
    if (!dst || !src || (bytes & 15) || ((((uintptr_t)dst) | ((uintptr_t)src)) & 15)) return fail(PNPADMM_ERR_BAD_ARG, "debug_copy: 16-byte alignment");
    debug_copy_kernel<<<blocks > 0 ? blocks : 296, 256, 0, ST(s)>>>(static_cast<uint4*>(dst), static_cast<const uint4*>(src), bytes / 16);
    LAUNCH_CHECK("debug_copy_kernel");
    return PNPADMM_OK;
}

// debug only (not part of include/pnpadmm.h): the planner's per-device constants
int pnpadmm_debug_plan_constants(double* out4, int* calibrated) {
    DeviceState* d; int rc = ensure_device(&d); if (rc) return rc;
    out4[0] = d->tau1_us; out4[1] = d->k2_a_us; out4[2] = d->k2_b_us; out4[3] = d->k2_pro_us;
    if (calibrated) *calibrated = d->calibrated ? 1 : 0;
    return PNPADMM_OK;
}

// debug only (not part of include/pnpadmm.h): read and clear the wait-time counters of conv64_tc_kernel (PNPADMM_TC_DEBUG bit 256)
int pnpadmm_debug_tc_prof(unsigned long long* h_out8) {
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpyFromSymbol(h_out8, tc::g_tc_prof, sizeof(unsigned long long) * 8));
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyToSymbol(tc::g_tc_prof, z, sizeof(z)));
    return PNPADMM_OK;
}

}  // extern "C"
