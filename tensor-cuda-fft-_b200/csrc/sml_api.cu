// sml_api.cu -- C ABI of libspectral_mix_b200.so (see include/spectral_mix_b200.h).
// Host-side dispatch only: argument checks, plan selection, TMA descriptor encoding, kernel launches.
// No torch types, no allocation per call (a per-(device, T) twiddle table is created once).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

#include "../../include/spectral_mix_b200.h"
#include "sml_fast.cuh"
#include "sml_fast_ws.cuh"
#include "sml_generic.cuh"
#include "sml_wirtinger.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

#define SML_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) return fail("%s failed: %s", #expr, cudaGetErrorString(_e));    \
    } while (0)

inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------------
// per-device state
// ------------------------------------------------------------------------------------------------
struct DeviceState {
    int sm_count = 0;
    int cc_major = 0;
    std::map<int, sml::cf*> twiddles;   // T -> W_T^n table
};
std::mutex g_mu;
std::map<int, DeviceState> g_dev;

int device_state(DeviceState** out, int* dev_out) {
    int dev = 0;
    SML_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    DeviceState& st = g_dev[dev];
    if (st.sm_count == 0) {
        SML_CUDA(cudaDeviceGetAttribute(&st.sm_count, cudaDevAttrMultiProcessorCount, dev));
        SML_CUDA(cudaDeviceGetAttribute(&st.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
    }
    *out = &st;
    if (dev_out) *dev_out = dev;
    return 0;
}

// SML_DEBUG=1: a host-mapped record that a timed-out mbarrier wait fills before it traps (readable after the fault)
unsigned int* g_dbg_host = nullptr;
unsigned int* debug_record() {
    static std::once_flag once;
    static unsigned int* dev = nullptr;
    std::call_once(once, [] {
        const char* e = getenv("SML_DEBUG");
        if (e == nullptr || atoi(e) == 0) return;
        if (cudaHostAlloc(&g_dbg_host, 4096, cudaHostAllocMapped) != cudaSuccess) { g_dbg_host = nullptr; return; }
        memset(g_dbg_host, 0, 4096);
        if (cudaHostGetDevicePointer(&dev, g_dbg_host, 0) != cudaSuccess) dev = nullptr;
    });
    return dev;
}

int twiddle_table(DeviceState* st, int T, cudaStream_t stream, const sml::cf** out) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = st->twiddles.find(T);
    if (it == st->twiddles.end()) {
        sml::cf* tab = nullptr;
        SML_CUDA(cudaMalloc(&tab, sizeof(sml::cf) * (size_t)T));
        sml::twiddle_table_kernel<<<(T + 255) / 256, 256, 0, stream>>>(tab, T);
        count_launch();
        SML_CUDA(cudaGetLastError());
        SML_CUDA(cudaStreamSynchronize(stream));   // one-time: other streams may use the table next
        it = st->twiddles.emplace(T, tab).first;
    }
    *out = it->second;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct Plan {
    int path = SML_PATH_GENERIC;
    int k = 0;
    int NR = 0, KJ = 0, P = 0, M = 0, R = 0;
    int ctas_per_sm = 1;
    bool ws = false;   // warp-specialised 512-thread kernel (NR = 32 only)
};

inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

Plan make_plan(int T, int D, int F, int io_dtype) {
    Plan p;
    p.k = F < T / 2 ? F : T / 2;
    if (p.k < 1 || !is_pow2(T)) return p;
    const int esz = io_dtype == SML_DTYPE_BF16 ? 2 : 4;
    if ((D * esz) % 16 != 0) return p;   // TMA global strides must be multiples of 16 bytes (and D even)
    // smallest supported square sub-transform M = NR*NR >= 2k that still fits into T
    static const int kNR[3] = {8, 16, 32};
    static const int kP[3] = {32, 8, 4};   // channel pairs per CTA
    for (int i = 0; i < 3; ++i) {
        const int M = kNR[i] * kNR[i];
        if (M >= 2 * p.k && M <= T) {
            p.path = SML_PATH_FAST;
            p.NR = kNR[i];
            p.P = kP[i];
            p.M = M;
            p.R = T / M;
            p.ctas_per_sm = p.NR == 8 ? 2 : 3;   // CTAs per SM the kernel variant is compiled for (register budget)
            // TMA store address arithmetic and box coordinates stay in 32 bits
            const int need = (p.k + p.NR - 1) / p.NR;   // positive f2 columns that hold live bins
            // instantiated KJ values per NR (see launch_fast)
            if (p.NR == 32) p.KJ = need <= 8 ? 8 : need <= 12 ? 12 : 16;
            else if (p.NR == 16) p.KJ = need <= 4 ? 4 : 8;
            else p.KJ = 4;
            if (p.NR == 32 && p.KJ == 16) p.ctas_per_sm = 2;   // 64 accumulator registers: 3 CTAs/SM would spill
            if (const char* e = getenv("SML_FAST_CTAS")) {   // tuning knob for the NR=32, KJ=12 kernel: 2 or 3 CTAs per SM
                if (atoi(e) == 2 && p.NR == 32 && p.KJ == 12) p.ctas_per_sm = 2;
            }
            if (p.NR == 32) {   // M = 1024: warp-specialised kernel, 8 pairs (64-byte TMA rows) per CTA, one CTA per SM
                p.ws = false;   // experimental (opt-in) until it is parity-green on the GPU
                if (const char* e = getenv("SML_FAST_WS")) p.ws = atoi(e) != 0;   // tuning knob: 1 = warp-specialised kernel
                if (p.ws) { p.P = 8; p.ctas_per_sm = 1; }
            }
            return p;
        }
    }
    return p;
}

// ------------------------------------------------------------------------------------------------
// TMA descriptor (driver entry point resolved through the runtime: no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn* out) {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    static cudaError_t err = cudaSuccess;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (err == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    if (fn == nullptr) return fail("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(err));
    *out = fn;
    return 0;
}

// view of a (B, T, D) activation as the 4-D tensor {D, R, M, B}: t = R*m + r
int encode_act_map(CUtensorMap* map, const void* base, int B, int T, int D, int io_dtype, const Plan& p) {
    EncodeTiledFn enc;
    if (get_encode_fn(&enc)) return 1;
    const cuuint64_t esz = io_dtype == SML_DTYPE_BF16 ? 2 : 4;
    cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)p.R, (cuuint64_t)p.M, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)D * esz, (cuuint64_t)p.R * D * esz, (cuuint64_t)T * D * esz};
    const int boxrows = p.M < 256 ? p.M : 256;
    cuuint32_t box[4] = {(cuuint32_t)(2 * p.P), 1u, (cuuint32_t)boxrows, 1u};
    cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    CUresult r = enc(map, io_dtype == SML_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                     4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// fast-path launch
// ------------------------------------------------------------------------------------------------
template <int NR, int KJ, int P, int MINB, typename IO, bool BWD>
int launch_fast_inst(const CUtensorMap& map_in, const CUtensorMap& map_out, const sml::FastParams& prm, int grid,
                     cudaStream_t stream) {
    using C = sml::FastCfg<NR, P, IO>;
    auto kern = sml::sml_fast_kernel<NR, KJ, P, MINB, IO, BWD>;
    static std::once_flag once;   // one per instantiation
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] {
        attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
    });
    if (attr_err != cudaSuccess) return fail("cudaFuncSetAttribute(smem=%zu) failed: %s", C::SMEM_BYTES, cudaGetErrorString(attr_err));
    kern<<<grid, C::NT, C::SMEM_BYTES, stream>>>(map_in, map_out, prm);
    count_launch();
    SML_CUDA(cudaGetLastError());
    return 0;
}

template <int KJ, typename IO, bool BWD>
int launch_ws_inst(const CUtensorMap& map_in, const CUtensorMap& map_out, const sml::FastParams& prm, int grid,
                   cudaStream_t stream) {
    using C = sml::WsCfg<IO>;
    auto kern = sml::sml_ws_kernel<KJ, IO, BWD>;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [&] {
        attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES);
    });
    if (attr_err != cudaSuccess) return fail("cudaFuncSetAttribute(smem=%zu) failed: %s", C::SMEM_BYTES, cudaGetErrorString(attr_err));
    kern<<<grid, C::NT, C::SMEM_BYTES, stream>>>(map_in, map_out, prm);
    count_launch();
    SML_CUDA(cudaGetLastError());
    return 0;
}

template <typename IO, bool BWD>
int launch_fast(const Plan& p, const CUtensorMap& map_in, const CUtensorMap& map_out, const sml::FastParams& prm,
                int grid, cudaStream_t stream) {
    if (p.ws) {
        if (p.KJ == 8) return launch_ws_inst<8, IO, BWD>(map_in, map_out, prm, grid, stream);
        if (p.KJ == 12) return launch_ws_inst<12, IO, BWD>(map_in, map_out, prm, grid, stream);
        return launch_ws_inst<16, IO, BWD>(map_in, map_out, prm, grid, stream);
    }
#define SML_CASE(NR_, KJ_, P_, MINB_) \
    if (p.NR == NR_ && p.KJ == KJ_ && p.P == P_ && p.ctas_per_sm == MINB_) return launch_fast_inst<NR_, KJ_, P_, MINB_, IO, BWD>(map_in, map_out, prm, grid, stream);
    SML_CASE(32, 8, 4, 3)
    SML_CASE(32, 12, 4, 3)
    SML_CASE(32, 16, 4, 2)
    SML_CASE(32, 12, 4, 2)
    SML_CASE(16, 4, 8, 3)
    SML_CASE(16, 8, 8, 3)
    SML_CASE(8, 4, 32, 2)
#undef SML_CASE
    return fail("internal: no fast kernel for NR=%d KJ=%d P=%d", p.NR, p.KJ, p.P);
}

int check_common(const void* a, const void* b, int B, int T, int D, int F, int io_dtype) {
    if (a == nullptr || b == nullptr) return fail("null activation pointer");
    if (B < 1 || T < 1 || D < 1 || F < 1) return fail("invalid shape B=%d T=%d D=%d F=%d", B, T, D, F);
    if (io_dtype != SML_DTYPE_F32 && io_dtype != SML_DTYPE_BF16) return fail("unsupported io_dtype %d", io_dtype);
    if ((long long)B * T * D >= (1ll << 40)) return fail("tensor too large");
    return 0;
}

template <typename IO>
int forward_impl(const void* x, const float* w_re, const float* w_im, const float* bias, void* y, void* xlow, int B,
                 int T, int D, int F, int io_dtype, cudaStream_t stream) {
    DeviceState* st;
    if (device_state(&st, nullptr)) return 1;
    if (st->cc_major != 10) return fail("libspectral_mix_b200 is built for sm_100a only (device is sm_%d*)", st->cc_major * 10);
    const Plan p = make_plan(T, D, F, io_dtype);
    const sml::cf* gtab = nullptr;
    if (twiddle_table(st, T, stream, &gtab)) return 1;
    const float invT = 1.0f / (float)T;
    const bool aligned = ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0);
    if (p.path == SML_PATH_FAST && aligned) {
        CUtensorMap map, map_out;
        if (encode_act_map(&map, x, B, T, D, io_dtype, p)) return 1;
        if (encode_act_map(&map_out, y, B, T, D, io_dtype, p)) return 1;
        sml::FastParams prm{};
        prm.out = y; prm.w_re = w_re; prm.w_im = w_im; prm.bias = bias;
        prm.xlow = reinterpret_cast<sml::cf*>(xlow);
        prm.gtab = gtab;
        prm.B = B; prm.T = T; prm.D = D; prm.F = F; prm.k = p.k; prm.R = p.R;
        prm.ntd = (D + 2 * p.P - 1) / (2 * p.P);
        prm.ntiles = B * prm.ntd;
        prm.invT = invT;
        prm.dbg = debug_record();
        const int slots = st->sm_count * p.ctas_per_sm;
        const int grid = prm.ntiles < slots ? prm.ntiles : slots;
        return launch_fast<IO, false>(p, map, map_out, prm, grid, stream);
    }
    // generic path
    if (p.k > 0 && xlow == nullptr) return fail("generic path needs the xlow buffer (sml_xlow_bytes) as scratch");
    if (p.k > 0) {
        dim3 blk(32, 8), ga((D + 31) / 32, (p.k + 7) / 8, B), gs((D + 31) / 32, (T + 7) / 8, B);
        sml::generic_analysis_kernel<IO><<<ga, blk, 0, stream>>>((const IO*)x, (sml::cf*)xlow, gtab, T, D, p.k);
        sml::generic_synthesis_kernel<IO, false><<<gs, blk, 0, stream>>>((const sml::cf*)xlow, w_re, w_im, bias, (IO*)y, gtab, T, D, F, p.k, invT);
        count_launch(2);
    } else {
        dim3 blk(32, 8), gs((D + 31) / 32, (T + 7) / 8, B);
        sml::generic_synthesis_kernel<IO, false><<<gs, blk, 0, stream>>>(nullptr, w_re, w_im, bias, (IO*)y, gtab, T, D, F, 0, invT);
        count_launch();
    }
    SML_CUDA(cudaGetLastError());
    return 0;
}

template <typename IO>
int backward_impl(const void* g, const void* xlow, const float* w_re, const float* w_im, void* gx, float* gw_re,
                  float* gw_im, float* gb, void* ws, size_t ws_bytes, int B, int T, int D, int F, int io_dtype,
                  cudaStream_t stream) {
    DeviceState* st;
    if (device_state(&st, nullptr)) return 1;
    if (st->cc_major != 10) return fail("libspectral_mix_b200 is built for sm_100a only (device is sm_%d*)", st->cc_major * 10);
    const Plan p = make_plan(T, D, F, io_dtype);
    const bool want_grads = gw_re != nullptr;
    if (want_grads && (gw_im == nullptr || gb == nullptr)) return fail("gw_re, gw_im and gb must be given together");
    if (want_grads && p.k > 0 && xlow == nullptr) return fail("filter gradients need xlow saved by sml_forward");
    const sml::cf* gtab = nullptr;
    if (twiddle_table(st, T, stream, &gtab)) return 1;
    const float invT = 1.0f / (float)T;
    const bool aligned = ((uintptr_t)g % 16 == 0) && ((uintptr_t)gx % 16 == 0);
    if (p.path == SML_PATH_FAST && aligned) {
        if (want_grads) {
            SML_CUDA(cudaMemsetAsync(gw_re, 0, sizeof(float) * (size_t)D * F, stream));
            SML_CUDA(cudaMemsetAsync(gw_im, 0, sizeof(float) * (size_t)D * F, stream));
            SML_CUDA(cudaMemsetAsync(gb, 0, sizeof(float) * (size_t)D, stream));
        }
        CUtensorMap map, map_out;
        if (encode_act_map(&map, g, B, T, D, io_dtype, p)) return 1;
        if (encode_act_map(&map_out, gx, B, T, D, io_dtype, p)) return 1;
        sml::FastParams prm{};
        prm.out = gx; prm.w_re = w_re; prm.w_im = w_im; prm.bias = nullptr;
        prm.xlow = reinterpret_cast<sml::cf*>(const_cast<void*>(xlow));
        prm.gw_re = gw_re; prm.gw_im = gw_im; prm.gb = gb;
        prm.gtab = gtab;
        prm.B = B; prm.T = T; prm.D = D; prm.F = F; prm.k = p.k; prm.R = p.R;
        prm.ntd = (D + 2 * p.P - 1) / (2 * p.P);
        prm.ntiles = B * prm.ntd;
        prm.invT = invT;
        prm.dbg = debug_record();
        const int slots = st->sm_count * p.ctas_per_sm;
        const int grid = prm.ntiles < slots ? prm.ntiles : slots;
        return launch_fast<IO, true>(p, map, map_out, prm, grid, stream);
    }
    // generic path: G into workspace, then synthesis with conj(W) and the batch reduction
    const size_t need = sizeof(sml::cf) * (size_t)B * D * (size_t)p.k;
    if (p.k > 0 && (ws == nullptr || ws_bytes < need)) return fail("workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    dim3 blk(32, 8), gs((D + 31) / 32, (T + 7) / 8, B);
    if (p.k > 0) {
        dim3 ga((D + 31) / 32, (p.k + 7) / 8, B);
        sml::generic_analysis_kernel<IO><<<ga, blk, 0, stream>>>((const IO*)g, (sml::cf*)ws, gtab, T, D, p.k);
        count_launch();
    }
    sml::generic_synthesis_kernel<IO, true><<<gs, blk, 0, stream>>>((const sml::cf*)ws, w_re, w_im, nullptr, (IO*)gx, gtab, T, D, F, p.k, invT);
    count_launch();
    if (want_grads) {
        const long long n = (long long)D * F;
        sml::generic_filtergrad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const sml::cf*)ws, (const sml::cf*)xlow, gw_re, gw_im, gb, B, D, F, p.k, invT);
        count_launch();
        if (p.k == 0) {
            sml::generic_biasgrad_kernel<IO><<<(D + 127) / 128, 128, 0, stream>>>((const IO*)g, gb, (long long)B * T, D);
            count_launch();
        }
    }
    SML_CUDA(cudaGetLastError());
    return 0;
}

}   // namespace

// ------------------------------------------------------------------------------------------------
// exported C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int sml_abi_version(void) { return 1; }

const char* sml_last_error(void) { return g_err; }

unsigned long long sml_launch_count(void) { return g_launches.load(); }

int sml_debug_dump(void) {
    if (g_dbg_host == nullptr) return 0;
    const unsigned int n = g_dbg_host[0];
    fprintf(stderr, "sml_debug: %u timed-out mbarrier waits recorded\n", n);
    for (unsigned int i = 0; i < n && i < 64u; ++i) {
        const unsigned int* r = g_dbg_host + 8 + 8 * i;
        fprintf(stderr, "  tag=%u block=%u tid=%u parity=%u aux=%u\n", r[0], r[1], r[2], r[3], r[4]);
    }
    return (int)n;
}

int sml_plan(int B, int T, int D, int F, int io_dtype, int* path, int* M, int* R, int* k) {
    if (check_common(&B, &B, B, T, D, F, io_dtype)) return 1;
    const Plan p = make_plan(T, D, F, io_dtype);
    if (path) *path = p.path;
    if (M) *M = p.M;
    if (R) *R = p.R;
    if (k) *k = p.k;
    return 0;
}

size_t sml_xlow_bytes(int B, int T, int D, int F) {
    const int k = F < T / 2 ? F : T / 2;
    return sizeof(sml::cf) * (size_t)B * D * (size_t)(k > 0 ? k : 0);
}

size_t sml_workspace_bytes(int B, int T, int D, int F, int io_dtype) {
    const Plan p = make_plan(T, D, F, io_dtype);
    if (p.path == SML_PATH_FAST) return 0;
    return sml_xlow_bytes(B, T, D, F);
}

int sml_forward(const void* x, const float* w_re, const float* w_im, const float* bias, void* y, void* xlow_save,
                int B, int T, int D, int F, int io_dtype, void* stream) {
    if (check_common(x, y, B, T, D, F, io_dtype)) return 1;
    if (w_re == nullptr || w_im == nullptr) return fail("null filter pointer");
    g_err[0] = 0;
    if (io_dtype == SML_DTYPE_F32)
        return forward_impl<float>(x, w_re, w_im, bias, y, xlow_save, B, T, D, F, io_dtype, (cudaStream_t)stream);
    return forward_impl<__nv_bfloat16>(x, w_re, w_im, bias, y, xlow_save, B, T, D, F, io_dtype, (cudaStream_t)stream);
}

int sml_backward(const void* g, const void* xlow, const float* w_re, const float* w_im, void* gx, float* gw_re,
                 float* gw_im, float* gb, void* workspace, size_t workspace_bytes, int B, int T, int D, int F,
                 int io_dtype, void* stream) {
    if (check_common(g, gx, B, T, D, F, io_dtype)) return 1;
    if (w_re == nullptr || w_im == nullptr) return fail("null filter pointer");
    g_err[0] = 0;
    if (io_dtype == SML_DTYPE_F32)
        return backward_impl<float>(g, xlow, w_re, w_im, gx, gw_re, gw_im, gb, workspace, workspace_bytes, B, T, D, F,
                                    io_dtype, (cudaStream_t)stream);
    return backward_impl<__nv_bfloat16>(g, xlow, w_re, w_im, gx, gw_re, gw_im, gb, workspace, workspace_bytes, B, T, D,
                                        F, io_dtype, (cudaStream_t)stream);
}

int sml_wirtinger_mul_forward(const void* x, const void* w, void* out, long long B, long long N, void* stream) {
    if (!x || !w || !out) return fail("null pointer");
    if (B < 1 || N < 1) return fail("invalid shape B=%lld N=%lld", B, N);
    g_err[0] = 0;
    const long long total = B * N;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    sml::wirtinger_mul_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float2*)x, (const float2*)w, (float2*)out, B, N);
    count_launch();
    SML_CUDA(cudaGetLastError());
    return 0;
}

int sml_wirtinger_mul_backward(const void* g, const void* x, const void* w, void* gx, void* gw, long long B,
                               long long N, void* stream) {
    if (!g || !x || !w || !gx || !gw) return fail("null pointer");
    if (B < 1 || N < 1) return fail("invalid shape B=%lld N=%lld", B, N);
    g_err[0] = 0;
    sml::wirtinger_mul_bwd_kernel<<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        (const float2*)g, (const float2*)x, (const float2*)w, (float2*)gx, (float2*)gw, B, N);
    count_launch();
    SML_CUDA(cudaGetLastError());
    return 0;
}

int sml_wirtinger_filter_forward(const void* x_freq, const float* w_re, const float* w_im, void* out, int B, int T,
                                 int D, int F, void* stream) {
    if (!x_freq || !w_re || !w_im || !out) return fail("null pointer");
    if (B < 1 || T < 1 || D < 1 || F < 1) return fail("invalid shape B=%d T=%d D=%d F=%d", B, T, D, F);
    g_err[0] = 0;
    const int k = F < T / 2 ? F : T / 2;
    const long long total = (long long)B * T * D;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    sml::wirtinger_filter_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float2*)x_freq, w_re, w_im, (float2*)out, B, T, D, F, k);
    count_launch();
    SML_CUDA(cudaGetLastError());
    return 0;
}

int sml_wirtinger_filter_backward(const void* g, const void* x_freq, const float* w_re, const float* w_im, void* gx,
                                  float* gw_re, float* gw_im, int B, int T, int D, int F, void* stream) {
    if (!g || !x_freq || !w_re || !w_im || !gx || !gw_re || !gw_im) return fail("null pointer");
    if (B < 1 || T < 1 || D < 1 || F < 1) return fail("invalid shape B=%d T=%d D=%d F=%d", B, T, D, F);
    g_err[0] = 0;
    const int k = F < T / 2 ? F : T / 2;
    cudaStream_t s = (cudaStream_t)stream;
    SML_CUDA(cudaMemsetAsync(gw_re, 0, sizeof(float) * (size_t)D * F, s));
    SML_CUDA(cudaMemsetAsync(gw_im, 0, sizeof(float) * (size_t)D * F, s));
    dim3 blk(32, 8), grid((D + 31) / 32, (T + 7) / 8);
    sml::wirtinger_filter_bwd_kernel<<<grid, blk, 0, s>>>((const float2*)g, (const float2*)x_freq, w_re, w_im, (float2*)gx, gw_re, gw_im, B, T, D, F, k);
    count_launch();
    SML_CUDA(cudaGetLastError());
    return 0;
}

}   // extern "C"
