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
#include <tuple>
#include <utility>

#include "../../include/spectral_mix_b200.h"
#include "sml_host.h"
#include "sml_generic.cuh"
#include "sml_wirtinger.cuh"
#include "sml_block.cuh"

namespace {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

}   // namespace

namespace sml_host {
int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

const Knobs& knobs() {
    static const Knobs k = [] {
        Knobs v;
        if (const char* e = getenv("SML_FAST_CTAS")) v.fast_ctas = atoi(e);
        if (const char* e = getenv("SML_FAST_XB")) v.fast_xb = atoi(e);
        if (const char* e = getenv("SML_TC")) v.tc = atoi(e) != 0 ? 1 : 0;
        if (const char* e = getenv("SML_PDL")) v.pdl = atoi(e);   // 1: every kernel of the chain; 2: only the small batch-reduction kernel
        if (const char* e = getenv("SML_EXT_CTAS")) v.ext_ctas = atoi(e);
        if (const char* e = getenv("SML_SPLIT")) v.split = atoi(e);
        if (const char* e = getenv("SML_L2_HINT")) v.l2_hint = atoi(e) != 0 ? 1 : 0;
        return v;
    }();
    return k;
}

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
}   // namespace sml_host

namespace {
using sml_host::count_launch;
using sml_host::fail;
using sml_host::EncodeTiledFn;
using sml_host::get_encode_fn;
using sml_host::knobs;
using sml_host::launch_fast;
using sml_host::launch_fast_ext;
using sml_host::launch_fast_split;
using sml_host::Plan;

#define SML_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) return fail("%s failed: %s", #expr, cudaGetErrorString(_e));    \
    } while (0)


// ------------------------------------------------------------------------------------------------
// per-device state
// ------------------------------------------------------------------------------------------------
struct SplitScratch {   // pass splitting: partial bands + arrival counters, one per (device, stream)
    char* buf = nullptr;
    size_t cap = 0;
};
struct DeviceState {
    int sm_count = 0;
    int cc_major = 0;
    std::map<int, sml::cf*> twiddles;   // T -> W_T^n table
    std::map<cudaStream_t, SplitScratch> split;
};
std::mutex g_mu;
std::map<int, DeviceState> g_dev;

int device_state(DeviceState** out, int* dev_out) {
    int dev = 0;
    SML_CUDA(cudaGetDevice(&dev));
    // Bind the device's primary context to THIS thread before any driver-API call (cuTensorMapEncodeTiled returns
    // CUDA_ERROR_INVALID_CONTEXT otherwise): PyTorch's autograd worker threads reach sml_backward without ever
    // having made a context-binding runtime call of their own.
    static thread_local int bound_dev = -1;
    if (bound_dev != dev) {
        SML_CUDA(cudaFree(nullptr));
        bound_dev = dev;
    }
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
        if (cudaHostAlloc(&g_dbg_host, 8192, cudaHostAllocMapped) != cudaSuccess) { g_dbg_host = nullptr; return; }
        memset(g_dbg_host, 0, 8192);   // words 0..1023: mbarrier-timeout records; words 1024..: phase timeline of the tensor-core kernel
        if (cudaHostGetDevicePointer(&dev, g_dbg_host, 0) != cudaSuccess) dev = nullptr;
    });
    return dev;
}

int twiddle_table(DeviceState* st, int T, cudaStream_t stream, const sml::cf** out) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = st->twiddles.find(T);
    if (it == st->twiddles.end()) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone)
            return fail("first call for T=%d on this device builds a twiddle table: run the shape once before capturing a CUDA graph", T);
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
// pass splitting (sml_fast.cuh): when a launch has fewer work items than half the resident CTA slots -- long sequences
// with a small batch per rank: T = 128K holds 128 items per batch element for 296 CTA slots -- `split` CTAs share one item,
// each streaming R / split consecutive passes; the partial bands are exchanged through an L2-resident scratch block.
// ------------------------------------------------------------------------------------------------
int choose_split(int ntiles, int slots, int R) {
    if (knobs().split > 0) {
        int s = knobs().split;
        while (s > 1 && (R % s != 0)) s >>= 1;
        return s < 1 ? 1 : s;
    }
    int s = 1;
    // Measured on B200 (D = 1024: 128 items per batch element on 296 CTA slots; profiles/r02_pass_splitting.md): splitting pays only
    // while the units still fit into ONE round of resident CTAs -- B = 1, T = 128K: 0.967 ms unsplit, 0.614 (S = 2), 0.627 (4), 0.701
    // (8), 0.879 (16); with 256 or more items S = 1 wins (every member re-reads all S partial bands and repeats the mid phase).
    // So: the largest power of two that divides R, leaves at least 4 passes per CTA and keeps items * S within the resident slots.
    while ((long long)ntiles * (2 * s) <= (long long)slots && R % (2 * s) == 0 && R / (2 * s) >= 4 && s < 8) s *= 2;
    return s;
}
// scratch of a split launch: [units][2 KJ][threads] complex partial bands, then one arrival counter per work item (cleared here).
// Cached per (device, stream); grown on demand (like the twiddle tables: not while the stream is being captured).
int setup_split(DeviceState* st, const Plan& p, sml::FastParams* prm, int slots, cudaStream_t stream, bool ext = false) {
    prm->split = (p.NR == 32 && !ext) ? choose_split(prm->ntiles, slots, p.R) : 1;   // instantiated for the largest sub-transform only
    prm->xch = nullptr;
    prm->xflag = nullptr;
    if (prm->split <= 1) { prm->split = 1; return 0; }
    const size_t band = sizeof(sml::cf) * (size_t)prm->ntiles * prm->split * (2 * p.KJ) * (size_t)(p.NR * p.P);
    const size_t flags = sizeof(unsigned int) * (size_t)prm->ntiles;
    const size_t need = ((band + 255) & ~(size_t)255) + flags;
    char* buf = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        SplitScratch& sc = st->split[stream];
        if (sc.cap < need) {
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) {
                prm->split = 1;   // no allocation inside a capture: this launch (and the graph) runs unsplit -- same results
                return 0;
            }
            if (sc.buf) { SML_CUDA(cudaStreamSynchronize(stream)); cudaFree(sc.buf); sc.buf = nullptr; sc.cap = 0; }
            SML_CUDA(cudaMalloc(&sc.buf, need));
            sc.cap = need;
        }
        buf = sc.buf;
    }
    prm->xch = reinterpret_cast<sml::cf*>(buf);
    prm->xflag = reinterpret_cast<unsigned int*>(buf + ((band + 255) & ~(size_t)255));
    SML_CUDA(cudaMemsetAsync(prm->xflag, 0, flags, stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------

Plan make_plan(int T, int D, int F, int io_dtype) {
    Plan p;
    p.k = F < T / 2 ? F : T / 2;
    if (p.k < 1) return p;
    const int esz = io_dtype == SML_DTYPE_BF16 ? 2 : 4;
    if ((D * esz) % 16 != 0) return p;   // TMA global strides must be multiples of 16 bytes (and D even)
    // smallest supported square sub-transform M = NR*NR that divides T and holds the band: M >= 2k (one band column
    // per sub-bin), or failing that the "wide band" form k <= M (two band columns per sub-bin; covers the full
    // half-spectrum cases T = 2k such as T = 512 with embed >= 512 or T = 128, and 512 < k <= 1024, i.e. embed up to 2048)
    static const int kNR[3] = {8, 16, 32};
    static const int kP[3] = {32, 8, 4};   // channel pairs per CTA
    // pass 0: one band bin per sub-bin (M >= 2k); pass 1: two (M >= k); pass 2: up to four (2M >= k), small sub-transforms only
    for (int pass = 0; pass < 3; ++pass)
    for (int i = 0; i < (pass == 2 ? 2 : 3); ++i) {
        const int M = kNR[i] * kNR[i];
        if ((pass == 0 ? M >= 2 * p.k : pass == 1 ? M >= p.k : 2 * M >= p.k) && M <= T && T % M == 0) {   // any T = R * M
            p.path = SML_PATH_FAST;
            p.NR = kNR[i];
            p.P = kP[i];
            p.M = M;
            p.R = T / M;
            p.ctas_per_sm = p.NR == 8 ? 2 : 3;   // CTAs per SM the kernel variant is compiled for (register budget)
            // TMA store address arithmetic and box coordinates stay in 32 bits
            const int need = (p.k + p.NR - 1) / p.NR;   // positive f2 columns that hold live bins
            // instantiated KJ values per NR (see launch_fast)
            if (p.NR == 32) p.KJ = need <= 8 ? 8 : need <= 12 ? 12 : need <= 16 ? 16 : need <= 24 ? 24 : 32;
            else if (p.NR == 16) p.KJ = need <= 4 ? 4 : need <= 8 ? 8 : need <= 12 ? 12 : need <= 16 ? 16 : need <= 24 ? 24 : 32;
            else p.KJ = need <= 4 ? 4 : need <= 8 ? 8 : 16;
            if (p.NR == 16 && p.KJ == 32) p.ctas_per_sm = 2;   // 128 accumulator registers
            // 96+ accumulator registers: two CTAs per SM.  KJ = 16 (64 accumulator registers, embed 1024: BASELINE configs[2] and
            // the long-context sweep) still compiles to 166 registers without spills and runs 9 % faster with three CTAs per SM
            // (cfg-3 fp32 step 1.288 -> 1.181 ms, forward 0.56 -> 0.62 of the HBM roofline; SML_FAST_CTAS=2 restores two)
            if (p.NR == 32 && p.KJ >= 24) p.ctas_per_sm = 2;
            // ... on most shapes: at (32, 32768, 1024) three CTAs per SM are 13 % SLOWER (the 128 KB row stride of a pass camps on
            // DRAM channels with more rows in flight).  So this plan is tuned on first use per shape (tuned_ctas below).
            if (p.NR == 32 && p.KJ == 16) { p.tunable = knobs().fast_ctas == 0; if (knobs().fast_ctas == 2) p.ctas_per_sm = 2; }
            if (knobs().fast_ctas == 2 && p.NR == 32 && p.KJ == 12) p.ctas_per_sm = 2;   // tuning knob: 2 or 3 CTAs per SM
            // (P = 6, i.e. 48-byte rows and 2 CTAs of 6 warps per SM, was measured at 0.320 ms per forward launch against
            //  0.197 ms for P = 4: rows that are not a multiple of the 32-byte sector straddle sectors on loads and stores.)
            // TMA landing tiles: two (loads run two passes ahead) pay off for the small sub-transforms (measured at
            // (32, 8192, 256) fp32: 0.152 vs 0.167 ms per forward launch); at M = 1024 one tile is faster (bf16 cfg-2: 0.203 vs
            // 0.208 ms, cfg-3 fp32: 0.592 vs 0.597 ms) and for fp32 two do not fit next to three CTAs per SM anyway.
            p.xb = p.NR <= 16 ? 2 : 1;
            if (knobs().fast_xb) p.xb = knobs().fast_xb == 2 ? 2 : 1;   // tuning knob
            // bf16 I/O is compute-bound on the CUDA-core butterflies: eligible problems run the four DFT stages on the
            // tensor cores instead (sml_tc.cuh).  SML_TC=0 keeps the CUDA-core kernel.
            p.tc = sml_host::tc_eligible(T, D, p.k, io_dtype) && knobs().tc > 0;
            return p;
        }
    }
    return p;
}

// ------------------------------------------------------------------------------------------------
// first-use tuning of the CTAs per SM of a tunable plan (make_plan): both variants are launched on the caller's stream (one
// warm-up + one timed launch each, CUDA events), the faster one is cached per (device, shape, direction).  The launches write the
// same results, so tuning is invisible except for the one-time cost; inside a CUDA-graph capture the untuned default is used.
// ------------------------------------------------------------------------------------------------
struct TuneKey {
    int dev, B, T, D, F, io, bwd, grads;
    bool operator<(const TuneKey& o) const {
        return std::tie(dev, B, T, D, F, io, bwd, grads) < std::tie(o.dev, o.B, o.T, o.D, o.F, o.io, o.bwd, o.grads);
    }
};
std::map<TuneKey, int> g_tuned;

template <typename Run>
int tuned_ctas(const TuneKey& key, const Plan& p, cudaStream_t stream, Run&& run, int* choice) {
    *choice = p.ctas_per_sm;
    if (!p.tunable) return 0;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_tuned.find(key);
        if (it != g_tuned.end()) { *choice = it->second; return 0; }
    }
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) return 0;
    cudaEvent_t e0, e1;
    SML_CUDA(cudaEventCreate(&e0));
    SML_CUDA(cudaEventCreate(&e1));
    // warm both variants (module load, shared-memory attribute, scratch, TLB), then three alternating timed rounds, minimum each
    float ms[2] = {0.f, 0.f};
    int rc = 0;
    for (int c = 2; c <= 3 && !rc; ++c) {
        Plan q = p;
        q.ctas_per_sm = c;
        rc = run(q);
    }
    for (int round = 0; round < 3 && !rc; ++round)
        for (int c = 2; c <= 3 && !rc; ++c) {
            Plan q = p;
            q.ctas_per_sm = c;
            float t = 0.f;
            rc = cudaEventRecord(e0, stream) != cudaSuccess;
            if (!rc) rc = run(q);
            if (!rc) rc = cudaEventRecord(e1, stream) != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess ||
                          cudaEventElapsedTime(&t, e0, e1) != cudaSuccess;
            if (round == 0 || t < ms[c - 2]) ms[c - 2] = t;
        }
    if (rc) { cudaEventDestroy(e0); cudaEventDestroy(e1); return rc == 1 ? 1 : fail("plan tuning failed: %s", cudaGetErrorString(cudaGetLastError())); }
    const int best = ms[1] < ms[0] ? 3 : 2;
    if (getenv("SML_TUNE_VERBOSE"))
        fprintf(stderr, "[sml tune] B=%d T=%d D=%d F=%d io=%d bwd=%d grads=%d: 2 CTAs/SM %.4f ms, 3 CTAs/SM %.4f ms -> %d\n", key.B, key.T, key.D,
                key.F, key.io, key.bwd, key.grads, ms[0], ms[1], best);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    std::lock_guard<std::mutex> lk(g_mu);
    g_tuned[key] = best;
    *choice = best;
    return 0;
}

// view of a (B, T, D) activation as the 4-D tensor {D, R, M, B}: t = R*m + r
// rows (default T): the tensor holds only `rows` rows per batch element (a multiple of R) -- boxes that reach past them are
// zero-filled on load and clipped on store (row windows of the extended kernels)
int encode_act_map(CUtensorMap* map, const void* base, int B, int T, int D, int io_dtype, const Plan& p, int rows = -1) {
    EncodeTiledFn enc;
    if (get_encode_fn(&enc)) return 1;
    const cuuint64_t esz = io_dtype == SML_DTYPE_BF16 ? 2 : 4;
    if (rows < 0) rows = T;
    cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)p.R, (cuuint64_t)(rows / p.R), (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)D * esz, (cuuint64_t)p.R * D * esz, (cuuint64_t)rows * D * esz};
    const int boxrows = p.M < 256 ? p.M : 256;
    cuuint32_t box[4] = {(cuuint32_t)(2 * p.P), 1u, (cuuint32_t)boxrows, 1u};
    cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    CUresult r = enc(map, io_dtype == SML_DTYPE_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                     4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
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
    if (p.path == SML_PATH_FAST && aligned && p.tc) {
        sml_host::TcLaunch a;
        a.w_re = w_re; a.w_im = w_im; a.bias = bias;
        a.xlow = reinterpret_cast<sml::cf*>(xlow);
        a.B = B; a.T = T; a.D = D; a.F = F; a.k = p.k;
        a.dbg = debug_record();
        return sml_host::launch_tc<false>(x, y, a, st->sm_count, stream);
    }
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
        prm.l2_hint = knobs().l2_hint;
        auto run = [&](const Plan& q) -> int {
            sml::FastParams pq = prm;
            const int slots = st->sm_count * q.ctas_per_sm;
            if (setup_split(st, q, &pq, slots, stream)) return 1;
            const int units = pq.ntiles * pq.split;
            const int grid = units < slots ? units : slots;
            if (pq.split > 1) return launch_fast_split<IO, false>(q, map, map_out, pq, grid, stream);
            return launch_fast<IO, false>(q, map, map_out, pq, grid, stream);
        };
        Plan q = p;
        int dev = 0;
        cudaGetDevice(&dev);
        if (tuned_ctas(TuneKey{dev, B, T, D, F, io_dtype, 0, xlow != nullptr}, p, stream, run, &q.ctas_per_sm)) return 1;
        return run(q);
    }
    // generic path
    if (p.k > 0 && xlow == nullptr) return fail("generic path needs the xlow buffer (sml_xlow_bytes) as scratch");
    if (p.k > 0) {
        dim3 blk(256), ga((D + 63) / 64, (p.k + 63) / 64, B), gs((D + 63) / 64, (T + 63) / 64, B);
        sml::generic_analysis_kernel<IO><<<ga, blk, 0, stream>>>((const IO*)x, (sml::cf*)xlow, gtab, T, D, p.k);
        sml::generic_synthesis_kernel<IO, false><<<gs, blk, 0, stream>>>((const sml::cf*)xlow, w_re, w_im, bias, (IO*)y, gtab, T, D, F, p.k, invT);
        count_launch(2);
    } else {
        dim3 blk(256), gs((D + 63) / 64, (T + 63) / 64, B);
        sml::generic_synthesis_kernel<IO, false><<<gs, blk, 0, stream>>>(nullptr, w_re, w_im, bias, (IO*)y, gtab, T, D, F, 0, invT);
        count_launch();
    }
    SML_CUDA(cudaGetLastError());
    return 0;
}

// batch reduction of the per-batch filter/bias gradient terms (+ the fused multimem all-reduce when flat_mc is given)
int launch_filtergrad_reduce(const sml::cf* gpart, const float* gbpart, float* gw_re, float* gw_im, float* gb, int B, int D, int F,
                             int k, float* flat_mc, float* flat_next, cudaStream_t stream) {
    const bool vec4 = F % 4 == 0 && k % 4 == 0 && ((uintptr_t)gw_re % 16 == 0) && ((uintptr_t)gw_im % 16 == 0) && ((uintptr_t)gpart % 16 == 0) &&
                      (flat_mc == nullptr || (((uintptr_t)flat_mc % 16 == 0) && ((uintptr_t)flat_next % 16 == 0)));
    if (vec4) {
        const long long n = (long long)D * (F / 4);
        SML_CUDA(sml_host::launch_pdl_if(knobs().pdl >= 1, sml::filtergrad_reduce4_kernel, dim3((unsigned)((n + 63) / 64)), dim3(64, 4), 0, stream,
                                      reinterpret_cast<const float2*>(gpart), gbpart, gw_re, gw_im, gb, B, D, F, k, flat_mc, flat_next));
    } else {
        const long long n = (long long)D * ((F + 1) / 2);
        SML_CUDA(sml_host::launch_pdl_if(knobs().pdl >= 1, sml::filtergrad_reduce_kernel, dim3((unsigned)((n + 63) / 64)), dim3(64, 4), 0, stream,
                                      reinterpret_cast<const float2*>(gpart), gbpart, gw_re, gw_im, gb, B, D, F, k, flat_mc, flat_next));
    }
    count_launch();
    return 0;
}

template <typename IO>
int backward_impl(const void* g, const void* xlow, const float* w_re, const float* w_im, void* gx, float* gw_re,
                  float* gw_im, float* gb, void* ws, size_t ws_bytes, int B, int T, int D, int F, int io_dtype,
                  cudaStream_t stream, float* flat_mc = nullptr, float* flat_next = nullptr) {
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
    if (flat_mc != nullptr) {
        if (!(p.path == SML_PATH_FAST && aligned)) return fail("the fused reduce + all-reduce needs the fused kernels (T a multiple of 64, 16-byte aligned activations)");
        if (!want_grads || flat_next == nullptr) return fail("the fused reduce + all-reduce needs gw_re/gw_im/gb and the next-parity buffer");
        if (gw_im != gw_re + (size_t)D * F || gb != gw_re + 2 * (size_t)D * F)
            return fail("the fused reduce + all-reduce needs gw_re | gw_im | gb back to back in one flat buffer");
    }
    if (p.path == SML_PATH_FAST && aligned) {
        const size_t part_bytes = sizeof(sml::cf) * (size_t)B * D * (size_t)p.k;
        const size_t need_ws = part_bytes + sizeof(float) * (size_t)B * D;
        if (want_grads && (ws == nullptr || ws_bytes < need_ws))
            return fail("workspace too small: need %zu bytes (sml_workspace_bytes), got %zu", need_ws, ws_bytes);
        sml::cf* const gpart = want_grads ? reinterpret_cast<sml::cf*>(ws) : nullptr;
        float* const gbpart = want_grads ? reinterpret_cast<float*>(static_cast<char*>(ws) + part_bytes) : nullptr;
        if (p.tc) {
            sml_host::TcLaunch a;
            a.w_re = w_re; a.w_im = w_im;
            a.xlow = reinterpret_cast<sml::cf*>(const_cast<void*>(xlow));
            a.gw_re = gw_re; a.gpart = gpart; a.gbpart = gbpart;
            a.B = B; a.T = T; a.D = D; a.F = F; a.k = p.k;
            a.dbg = debug_record();
            if (sml_host::launch_tc<true>(g, gx, a, st->sm_count, stream)) return 1;
            if (want_grads && launch_filtergrad_reduce(gpart, gbpart, gw_re, gw_im, gb, B, D, F, p.k, flat_mc, flat_next, stream)) return 1;
            return 0;
        }
        CUtensorMap map, map_out;
        if (encode_act_map(&map, g, B, T, D, io_dtype, p)) return 1;
        if (encode_act_map(&map_out, gx, B, T, D, io_dtype, p)) return 1;
        sml::FastParams prm{};
        prm.out = gx; prm.w_re = w_re; prm.w_im = w_im; prm.bias = nullptr;
        prm.xlow = reinterpret_cast<sml::cf*>(const_cast<void*>(xlow));
        prm.gw_re = gw_re;
        prm.gpart = want_grads ? reinterpret_cast<sml::cf*>(ws) : nullptr;
        prm.gbpart = want_grads ? reinterpret_cast<float*>(static_cast<char*>(ws) + part_bytes) : nullptr;
        prm.gtab = gtab;
        prm.B = B; prm.T = T; prm.D = D; prm.F = F; prm.k = p.k; prm.R = p.R;
        prm.ntd = (D + 2 * p.P - 1) / (2 * p.P);
        prm.ntiles = B * prm.ntd;
        prm.invT = invT;
        prm.dbg = debug_record();
        prm.l2_hint = knobs().l2_hint;
        auto run = [&](const Plan& q) -> int {
            sml::FastParams pq = prm;
            const int slots = st->sm_count * q.ctas_per_sm;
            if (setup_split(st, q, &pq, slots, stream)) return 1;
            const int units = pq.ntiles * pq.split;
            const int grid = units < slots ? units : slots;
            return pq.split > 1 ? launch_fast_split<IO, true>(q, map, map_out, pq, grid, stream) : launch_fast<IO, true>(q, map, map_out, pq, grid, stream);
        };
        Plan q = p;
        int dev = 0;
        cudaGetDevice(&dev);
        if (tuned_ctas(TuneKey{dev, B, T, D, F, io_dtype, 1, want_grads}, p, stream, run, &q.ctas_per_sm)) return 1;
        if (run(q)) return 1;
        if (want_grads && launch_filtergrad_reduce(prm.gpart, prm.gbpart, gw_re, gw_im, gb, B, D, F, p.k, flat_mc, flat_next, stream)) return 1;
        return 0;
    }
    // generic path: G into workspace, then synthesis with conj(W) and the batch reduction
    const size_t need = sizeof(sml::cf) * (size_t)B * D * (size_t)p.k;
    if (p.k > 0 && (ws == nullptr || ws_bytes < need)) return fail("workspace too small: need %zu bytes, got %zu", need, ws_bytes);
    dim3 blk(256), gs((D + 63) / 64, (T + 63) / 64, B);
    if (p.k > 0) {
        dim3 ga((D + 63) / 64, (p.k + 63) / 64, B);
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


// ------------------------------------------------------------------------------------------------
// extended fused kernels: block prologue / epilogue around the transform (sml_ext)
// ------------------------------------------------------------------------------------------------
struct ExtGeom {
    int T_in, in_row0, T_out, out_row0;
};
// plan of the extended kernels: the fused plan with two CTAs per SM at the largest sub-transform (64 more registers for the
// prefetched row statistics) and two landing tiles wherever they fit (the residual rows are prefetched through them)
Plan make_plan_ext(int T, int D, int F, int io_dtype) {
    Plan p = make_plan(T, D, F, io_dtype);
    if (p.path != SML_PATH_FAST) return p;
    p.tc = false;
    if (p.NR == 32) p.ctas_per_sm = 2;
    p.xb = 2;
    if (p.NR == 32 && p.KJ <= 12 && knobs().ext_ctas == 3) { p.ctas_per_sm = 3; p.xb = 1; }   // tuning knob SML_EXT_CTAS
    if (knobs().fast_xb) p.xb = knobs().fast_xb == 2 ? 2 : 1;
    return p;
}
int ext_check(const Plan& p, int T, int D, const sml_ext* e, ExtGeom* g) {
    if (p.path != SML_PATH_FAST) return fail("the extended entry points need the fused kernels (T a multiple of 64 that holds the band, 16-byte rows)");
    g->T_in = e && e->T_in > 0 ? e->T_in : T;
    g->T_out = e && e->T_out > 0 ? e->T_out : T;
    g->in_row0 = e ? e->in_row0 : 0;
    g->out_row0 = e ? e->out_row0 : 0;
    if (g->in_row0 < 0 || g->out_row0 < 0 || g->in_row0 + g->T_in > T || g->out_row0 + g->T_out > T)
        return fail("row windows [%d, %d) / [%d, %d) do not fit the transform length %d", g->in_row0, g->in_row0 + g->T_in, g->out_row0,
                    g->out_row0 + g->T_out, T);
    if (g->out_row0 != 0)
        return fail("out_row0 must be 0: TMA stores cannot start at a negative coordinate; shift the output by a phase ramp on the "
                    "filter instead, w[d,f] *= exp(2 pi i f out_row0 / T), w_nyq *= (-1)^out_row0");
    if (g->T_in % p.R != 0 || g->T_out % p.R != 0)
        return fail("T_in = %d and T_out = %d must be multiples of R = %d passes for this plan (pad the buffers)", g->T_in, g->T_out, p.R);
    if (e && e->w_nyq != nullptr) {
        if (T % 2 != 0 || p.k != T / 2 || p.k != p.NR * p.KJ)   // the kernel finds the bin at band column KJ of the f1 = 0 lanes
            return fail("w_nyq needs a full half-spectrum filter (F >= T/2, T even) on a plan that carries the bin T/2");
    }
    return 0;
}

template <typename IO, bool BWD>
int ext_impl(const void* in, const float* w_re, const float* w_im, const float* bias, void* out, void* xlow, float* gw_re,
             float* gw_im, float* gb, void* ws, size_t ws_bytes, int B, int T, int D, int F, int io_dtype, const sml_ext* e,
             cudaStream_t stream) {
    DeviceState* st;
    if (device_state(&st, nullptr)) return 1;
    if (st->cc_major != 10) return fail("libspectral_mix_b200 is built for sm_100a only (device is sm_%d*)", st->cc_major * 10);
    const Plan p = make_plan_ext(T, D, F, io_dtype);
    ExtGeom g;
    if (ext_check(p, T, D, e, &g)) return 1;
    if (((uintptr_t)in % 16) || ((uintptr_t)out % 16) || (e && e->residual && ((uintptr_t)e->residual % 16)))
        return fail("the extended entry points need 16-byte aligned activations");
    const bool want_grads = BWD && gw_re != nullptr;
    if (want_grads && (gw_im == nullptr || gb == nullptr)) return fail("gw_re, gw_im and gb must be given together");
    if (want_grads && xlow == nullptr) return fail("filter gradients need xlow saved by sml_forward_ext");
    const size_t part_bytes = sizeof(sml::cf) * (size_t)B * D * (size_t)p.k;
    const size_t need_ws = part_bytes + sizeof(float) * (size_t)B * D;
    if (want_grads && (ws == nullptr || ws_bytes < need_ws))
        return fail("workspace too small: need %zu bytes (sml_workspace_bytes), got %zu", need_ws, ws_bytes);
    const sml::cf* gtab = nullptr;
    if (twiddle_table(st, T, stream, &gtab)) return 1;
    // forward: in = x (input window), out = y (output window); backward: in = g (output window), out = gx (input window)
    const int rows_in = BWD ? g.T_out : g.T_in, row0_in = BWD ? g.out_row0 : g.in_row0;
    const int rows_out = BWD ? g.T_in : g.T_out, row0_out = BWD ? g.in_row0 : g.out_row0;
    CUtensorMap map_in, map_out, map_res;
    if (encode_act_map(&map_in, in, B, T, D, io_dtype, p, rows_in)) return 1;
    if (encode_act_map(&map_out, out, B, T, D, io_dtype, p, rows_out)) return 1;
    const bool res = !BWD && e && e->residual != nullptr;
    if (res) { if (encode_act_map(&map_res, e->residual, B, T, D, io_dtype, p, rows_out)) return 1; }
    else map_res = map_out;
    sml::FastParams prm{};
    prm.out = out; prm.w_re = w_re; prm.w_im = w_im; prm.bias = BWD ? nullptr : bias;
    prm.xlow = reinterpret_cast<sml::cf*>(xlow);
    prm.gw_re = BWD ? gw_re : nullptr;
    prm.gpart = want_grads ? reinterpret_cast<sml::cf*>(ws) : nullptr;
    prm.gbpart = want_grads ? reinterpret_cast<float*>(static_cast<char*>(ws) + part_bytes) : nullptr;
    prm.gtab = gtab;
    prm.B = B; prm.T = T; prm.D = D; prm.F = F; prm.k = p.k; prm.R = p.R;
    prm.ntd = (D + 2 * p.P - 1) / (2 * p.P);
    prm.ntiles = B * prm.ntd;
    prm.invT = 1.0f / (float)T;
    prm.dbg = debug_record();
        prm.l2_hint = knobs().l2_hint;
    prm.stats = (!BWD && e) ? reinterpret_cast<const float2*>(e->row_stats) : nullptr;
    prm.scale = e ? e->chan_scale : nullptr;
    prm.wnyq = e ? e->w_nyq : nullptr;
    prm.sb_re = (!BWD && e) ? e->sb_re : nullptr;
    prm.sb_im = (!BWD && e) ? e->sb_im : nullptr;
    prm.sb_nyq = (!BWD && e && e->w_nyq) ? e->sb_nyq : nullptr;
    if (prm.sb_re != nullptr && prm.sb_im == nullptr) return fail("sb_re and sb_im must be given together");
    prm.xnyq = e ? e->x_nyq : nullptr;
    prm.gnyqpart = (BWD && e) ? e->g_nyq : nullptr;
    prm.d_core = (BWD && e) ? e->d_core : nullptr;
    prm.d_q = (BWD && e && e->q_re && e->q_im) ? e->d_q : nullptr;
    prm.q_re = e ? e->q_re : nullptr;
    prm.q_im = e ? e->q_im : nullptr;
    prm.q_nyq = e ? e->q_nyq : nullptr;
    if (e && e->h_re != nullptr) {   // rank-one filter mode
        if (e->h_im == nullptr || e->chan == nullptr) return fail("rank-one filter mode needs h_re, h_im and chan");
        prm.h_re = e->h_re; prm.h_im = e->h_im; prm.h_nyq = e->h_nyq; prm.chan = e->chan; prm.bg = e->bg;
        prm.hpart = BWD ? reinterpret_cast<sml::cf*>(e->hpart) : nullptr;
        if (prm.hpart != nullptr && xlow == nullptr) return fail("hpart needs xlow saved by sml_forward_ext");
        if (want_grads) return fail("rank-one filter mode returns the filter gradient through hpart: pass gw_re = gw_im = gb = NULL");
    } else if (w_re == nullptr || w_im == nullptr) {
        return fail("null filter pointer");
    }
    if (prm.d_core != nullptr && xlow == nullptr) return fail("d_core needs xlow saved by sml_forward_ext");
    if (prm.q_re != nullptr && prm.q_im == nullptr) prm.q_re = nullptr;
    prm.res = res ? 1 : 0;
    if (row0_out != 0) return fail("the backward of a problem with in_row0 != 0 is not supported (its output rows would start at a shifted row)");
    prm.in_q = row0_in / p.R; prm.in_r = row0_in % p.R;
    const int slots = st->sm_count * p.ctas_per_sm;
    if (setup_split(st, p, &prm, slots, stream, true)) return 1;
    const int units = prm.ntiles * prm.split;
    const int grid = units < slots ? units : slots;
    if (launch_fast_ext<IO, BWD>(p, map_in, map_out, map_res, prm, grid, stream)) return 1;
    if (want_grads && launch_filtergrad_reduce(prm.gpart, prm.gbpart, gw_re, gw_im, gb, B, D, F, p.k, nullptr, nullptr, stream)) return 1;
    return 0;
}

template <typename IO>
int ln_stats_impl(const void* x, void* stats, int B, int T, int T_in, int in_row0, int D, float eps, cudaStream_t stream) {
    const long long nrows = (long long)B * T;
    const bool vec = ((uintptr_t)x % 16 == 0) && (D % sml::Vec16<IO>::N == 0);
    const unsigned blocks = (unsigned)((nrows + 7) / 8);
    if (vec) sml::ln_stats_kernel<IO, true><<<blocks, 256, 0, stream>>>((const IO*)x, (float2*)stats, nrows, T, T_in, in_row0, D, eps);
    else sml::ln_stats_kernel<IO, false><<<blocks, 256, 0, stream>>>((const IO*)x, (float2*)stats, nrows, T, T_in, in_row0, D, eps);
    count_launch();
    SML_CUDA(cudaGetLastError());
    return 0;
}
template <typename IO>
int ln_backward_impl(const void* gh, const void* x, const void* stats, const void* gres, const float* cadd, void* gx, int B, int T,
                     int T_in, int in_row0, int D, cudaStream_t stream) {
    const long long nrows = (long long)B * T_in;
    const bool vec = ((uintptr_t)x % 16 == 0) && ((uintptr_t)gh % 16 == 0) && ((uintptr_t)gx % 16 == 0) &&
                     (gres == nullptr || (uintptr_t)gres % 16 == 0) && (D % sml::Vec16<IO>::N == 0);
    const unsigned blocks = (unsigned)((nrows + 7) / 8);
    if (vec) sml::ln_backward_kernel<IO, true><<<blocks, 256, 0, stream>>>((const IO*)gh, (const IO*)x, (const float2*)stats, (const IO*)gres, cadd, (IO*)gx, nrows, T, T_in, in_row0, D);
    else sml::ln_backward_kernel<IO, false><<<blocks, 256, 0, stream>>>((const IO*)gh, (const IO*)x, (const float2*)stats, (const IO*)gres, cadd, (IO*)gx, nrows, T, T_in, in_row0, D);
    count_launch();
    SML_CUDA(cudaGetLastError());
    return 0;
}


// ------------------------------------------------------------------------------------------------
// host-buffer pipeline: fwd+bwd over HOST tensors, chunked along the batch axis so that the H2D copy of chunk i+1,
// the kernels of chunk i and the D2H copy of chunk i-1 overlap (three streams, NBUF device slots).
// ------------------------------------------------------------------------------------------------
__global__ void accumulate_kernel(float* __restrict__ total, const float* __restrict__ part, size_t n, int first) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        total[i] = first ? part[i] : total[i] + part[i];
}

struct HostPipe {
    static constexpr int NBUF = 3;
    cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[NBUF] = {}, ev_cmp[NBUF] = {}, ev_out[NBUF] = {};   // g landed, kernels done, gx left
    cudaEvent_t ev_x[NBUF] = {}, ev_fwd[NBUF] = {};                       // x landed, forward done
    char* act[NBUF][4] = {};   // x, g, y, gx
    char* xlow[NBUF] = {};
    char* ws[NBUF] = {};
    float* params = nullptr;       // [w_re | w_im | bias]
    float* grad_part = nullptr;    // [gw_re | gw_im | gb] of one chunk
    float* grad_total = nullptr;
    size_t cap_act = 0, cap_xlow = 0, cap_ws = 0, cap_par = 0;
    bool ready = false;
};
std::mutex g_pipe_mu;
std::map<int, HostPipe> g_pipes;

void pipe_free(HostPipe& hp) {   // drops every device buffer and resets the capacities (streams and events stay)
    for (int i = 0; i < HostPipe::NBUF; ++i) {
        for (int j = 0; j < 4; ++j) { if (hp.act[i][j]) cudaFree(hp.act[i][j]); hp.act[i][j] = nullptr; }
        if (hp.xlow[i]) cudaFree(hp.xlow[i]);
        if (hp.ws[i]) cudaFree(hp.ws[i]);
        hp.xlow[i] = hp.ws[i] = nullptr;
    }
    if (hp.params) cudaFree(hp.params);
    if (hp.grad_part) cudaFree(hp.grad_part);
    if (hp.grad_total) cudaFree(hp.grad_total);
    hp.params = hp.grad_part = hp.grad_total = nullptr;
    hp.cap_act = hp.cap_xlow = hp.cap_ws = hp.cap_par = 0;
}

int pipe_reserve_impl(HostPipe& hp, size_t act_bytes, size_t xlow_bytes, size_t ws_bytes, size_t par_floats);
int pipe_reserve(HostPipe& hp, size_t act_bytes, size_t xlow_bytes, size_t ws_bytes, size_t par_floats) {
    const int rc = pipe_reserve_impl(hp, act_bytes, xlow_bytes, ws_bytes, par_floats);
    if (rc != 0) {   // a cudaMalloc failed half way: never leave NULL slots behind a capacity that says they exist
        cudaDeviceSynchronize();
        pipe_free(hp);
    }
    return rc;
}

int pipe_reserve_impl(HostPipe& hp, size_t act_bytes, size_t xlow_bytes, size_t ws_bytes, size_t par_floats) {
    if (!hp.ready) {
        SML_CUDA(cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking));
        SML_CUDA(cudaStreamCreateWithFlags(&hp.s_cmp, cudaStreamNonBlocking));
        SML_CUDA(cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking));
        for (int i = 0; i < HostPipe::NBUF; ++i) {
            SML_CUDA(cudaEventCreateWithFlags(&hp.ev_in[i], cudaEventDisableTiming));
            SML_CUDA(cudaEventCreateWithFlags(&hp.ev_cmp[i], cudaEventDisableTiming));
            SML_CUDA(cudaEventCreateWithFlags(&hp.ev_out[i], cudaEventDisableTiming));
            SML_CUDA(cudaEventCreateWithFlags(&hp.ev_x[i], cudaEventDisableTiming));
            SML_CUDA(cudaEventCreateWithFlags(&hp.ev_fwd[i], cudaEventDisableTiming));
        }
        hp.ready = true;
    }
    auto grow = [](char** p, size_t* cap, size_t need) -> cudaError_t {
        if (need <= *cap && *p != nullptr) return cudaSuccess;
        if (*p) cudaFree(*p);
        *p = nullptr;
        cudaError_t e = cudaMalloc(p, need > 0 ? need : 256);
        if (e == cudaSuccess) *cap = need;
        return e;
    };
    if (act_bytes > hp.cap_act || hp.act[0][0] == nullptr) {
        SML_CUDA(cudaDeviceSynchronize());
        for (int i = 0; i < HostPipe::NBUF; ++i)
            for (int j = 0; j < 4; ++j) {
                size_t cap = 0;
                SML_CUDA(grow(&hp.act[i][j], &cap, act_bytes));
            }
        hp.cap_act = act_bytes;
    }
    if (xlow_bytes > hp.cap_xlow || hp.xlow[0] == nullptr) {
        SML_CUDA(cudaDeviceSynchronize());
        for (int i = 0; i < HostPipe::NBUF; ++i) { size_t cap = 0; SML_CUDA(grow(&hp.xlow[i], &cap, xlow_bytes)); }
        hp.cap_xlow = xlow_bytes;
    }
    if (ws_bytes > hp.cap_ws || hp.ws[0] == nullptr) {
        SML_CUDA(cudaDeviceSynchronize());
        for (int i = 0; i < HostPipe::NBUF; ++i) { size_t cap = 0; SML_CUDA(grow(&hp.ws[i], &cap, ws_bytes)); }
        hp.cap_ws = ws_bytes;
    }
    if (par_floats > hp.cap_par || hp.params == nullptr) {
        SML_CUDA(cudaDeviceSynchronize());
        size_t c0 = 0, c1 = 0, c2 = 0;
        SML_CUDA(grow(reinterpret_cast<char**>(&hp.params), &c0, par_floats * sizeof(float)));
        SML_CUDA(grow(reinterpret_cast<char**>(&hp.grad_part), &c1, par_floats * sizeof(float)));
        SML_CUDA(grow(reinterpret_cast<char**>(&hp.grad_total), &c2, par_floats * sizeof(float)));
        hp.cap_par = par_floats;
    }
    return 0;
}

template <typename IO>
int fwd_bwd_host_body(HostPipe& hp, const void* x, const void* g, const float* w_re, const float* w_im, const float* bias, void* y,
                      void* gx, float* gw_re, float* gw_im, float* gb, int B, int T, int D, int F, int io_dtype,
                      int chunk_batch);

template <typename IO>
int fwd_bwd_host_impl(const void* x, const void* g, const float* w_re, const float* w_im, const float* bias, void* y,
                      void* gx, float* gw_re, float* gw_im, float* gb, int B, int T, int D, int F, int io_dtype,
                      int chunk_batch) {
    int dev = 0;
    SML_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_pipe_mu);   // one host-pipeline call per device at a time
    HostPipe& hp = g_pipes[dev];
    const int rc = fwd_bwd_host_body<IO>(hp, x, g, w_re, w_im, bias, y, gx, gw_re, gw_im, gb, B, T, D, F, io_dtype, chunk_batch);
    if (rc != 0 && hp.ready) {
        // an error in the middle of the pipeline: asynchronous copies into the caller's host buffers may still be in
        // flight -- drain all three streams before reporting the failure (the caller may free y / gx right away)
        cudaStreamSynchronize(hp.s_in);
        cudaStreamSynchronize(hp.s_cmp);
        cudaStreamSynchronize(hp.s_out);
    }
    return rc;
}

template <typename IO>
int fwd_bwd_host_body(HostPipe& hp, const void* x, const void* g, const float* w_re, const float* w_im, const float* bias, void* y,
                      void* gx, float* gw_re, float* gw_im, float* gb, int B, int T, int D, int F, int io_dtype,
                      int chunk_batch) {
    const size_t esz = sizeof(IO);
    const size_t row_bytes = (size_t)T * D * esz;                 // one batch element
    int cb = chunk_batch;
    if (cb <= 0) {   // about 32 MB of activations per chunk and at least 4 chunks when the batch allows it
        cb = (int)((32u << 20) / row_bytes);
        if (cb < 1) cb = 1;
        if (cb > (B + 3) / 4) cb = (B + 3) / 4;
        if (cb < 1) cb = 1;
    }
    if (cb > B) cb = B;
    const size_t nW = (size_t)D * F, nP = 2 * nW + D;
    const bool want_grads = gw_re != nullptr;
    const bool fast_plan = make_plan(T, D, F, io_dtype).path == SML_PATH_FAST;
    if (want_grads && (gw_im == nullptr || gb == nullptr)) return fail("gw_re, gw_im and gb must be given together");
    if (pipe_reserve(hp, (size_t)cb * row_bytes, sml_xlow_bytes(cb, T, D, F), sml_workspace_bytes(cb, T, D, F, io_dtype), nP))
        return 1;
    float* d_wre = hp.params;
    float* d_wim = hp.params + nW;
    float* d_bias = bias ? hp.params + 2 * nW : nullptr;
    SML_CUDA(cudaMemcpyAsync(d_wre, w_re, nW * sizeof(float), cudaMemcpyHostToDevice, hp.s_cmp));
    SML_CUDA(cudaMemcpyAsync(d_wim, w_im, nW * sizeof(float), cudaMemcpyHostToDevice, hp.s_cmp));
    if (bias) SML_CUDA(cudaMemcpyAsync(d_bias, bias, (size_t)D * sizeof(float), cudaMemcpyHostToDevice, hp.s_cmp));
    const int nchunks = (B + cb - 1) / cb;
    for (int i = 0; i < nchunks; ++i) {
        const int s = i % HostPipe::NBUF;
        const int b0 = i * cb, nb = (B - b0 < cb) ? B - b0 : cb;
        const size_t off = (size_t)b0 * row_bytes, bytes = (size_t)nb * row_bytes;
        char *dx = hp.act[s][0], *dg = hp.act[s][1], *dy = hp.act[s][2], *dgx = hp.act[s][3];
        // H2D: the slot's x/g are free once the kernels of chunk i - NBUF have run.  x and g get their own events so the
        // forward starts as soon as x is in, and y leaves while the backward still runs (shorter pipeline fill and drain).
        if (i >= HostPipe::NBUF) SML_CUDA(cudaStreamWaitEvent(hp.s_in, hp.ev_cmp[s], 0));
        SML_CUDA(cudaMemcpyAsync(dx, (const char*)x + off, bytes, cudaMemcpyHostToDevice, hp.s_in));
        SML_CUDA(cudaEventRecord(hp.ev_x[s], hp.s_in));
        SML_CUDA(cudaMemcpyAsync(dg, (const char*)g + off, bytes, cudaMemcpyHostToDevice, hp.s_in));
        SML_CUDA(cudaEventRecord(hp.ev_in[s], hp.s_in));
        // forward: wait for x and for the slot's y/gx to have left (D2H of chunk i - NBUF)
        SML_CUDA(cudaStreamWaitEvent(hp.s_cmp, hp.ev_x[s], 0));
        if (i >= HostPipe::NBUF) SML_CUDA(cudaStreamWaitEvent(hp.s_cmp, hp.ev_out[s], 0));
        if (forward_impl<IO>(dx, d_wre, d_wim, d_bias, dy, want_grads || !fast_plan ? hp.xlow[s] : nullptr, nb, T, D, F,
                             io_dtype, hp.s_cmp))
            return 1;
        SML_CUDA(cudaEventRecord(hp.ev_fwd[s], hp.s_cmp));
        SML_CUDA(cudaStreamWaitEvent(hp.s_out, hp.ev_fwd[s], 0));
        SML_CUDA(cudaMemcpyAsync((char*)y + off, dy, bytes, cudaMemcpyDeviceToHost, hp.s_out));
        // backward: wait for g
        SML_CUDA(cudaStreamWaitEvent(hp.s_cmp, hp.ev_in[s], 0));
        float* pg = want_grads ? hp.grad_part : nullptr;
        if (backward_impl<IO>(dg, want_grads ? hp.xlow[s] : nullptr, d_wre, d_wim, dgx, pg, pg ? pg + nW : nullptr,
                              pg ? pg + 2 * nW : nullptr, hp.ws[s], hp.cap_ws, nb, T, D, F, io_dtype, hp.s_cmp))
            return 1;
        if (want_grads) {
            accumulate_kernel<<<148, 256, 0, hp.s_cmp>>>(hp.grad_total, hp.grad_part, nP, i == 0);
            count_launch();
        }
        SML_CUDA(cudaEventRecord(hp.ev_cmp[s], hp.s_cmp));
        SML_CUDA(cudaStreamWaitEvent(hp.s_out, hp.ev_cmp[s], 0));
        SML_CUDA(cudaMemcpyAsync((char*)gx + off, dgx, bytes, cudaMemcpyDeviceToHost, hp.s_out));
        SML_CUDA(cudaEventRecord(hp.ev_out[s], hp.s_out));
    }
    if (want_grads) {   // s_out already waits for the last chunk's kernels
        SML_CUDA(cudaMemcpyAsync(gw_re, hp.grad_total, nW * sizeof(float), cudaMemcpyDeviceToHost, hp.s_out));
        SML_CUDA(cudaMemcpyAsync(gw_im, hp.grad_total + nW, nW * sizeof(float), cudaMemcpyDeviceToHost, hp.s_out));
        SML_CUDA(cudaMemcpyAsync(gb, hp.grad_total + 2 * nW, (size_t)D * sizeof(float), cudaMemcpyDeviceToHost, hp.s_out));
    }
    SML_CUDA(cudaStreamSynchronize(hp.s_out));
    SML_CUDA(cudaStreamSynchronize(hp.s_cmp));
    SML_CUDA(cudaGetLastError());
    return 0;
}

}   // namespace

// ------------------------------------------------------------------------------------------------
// exported C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int sml_abi_version(void) { return 3; }

const char* sml_last_error(void) { return g_err; }

unsigned long long sml_launch_count(void) { return g_launches.load(); }

int sml_debug_dump(void) {
    if (g_dbg_host == nullptr) return 0;
    const unsigned int n = g_dbg_host[0];
    fprintf(stderr, "sml_debug: %u timed-out mbarrier waits recorded\n", n);
    {   // phase timeline of the last tensor-core launch (CTA 0): start, analysis done, mid done, synthesis done -- per work item
        const unsigned long long* tl = reinterpret_cast<const unsigned long long*>(g_dbg_host + 1024);
        for (int it = 0; it < 8 && tl[4 * it] != 0; ++it)
            fprintf(stderr, "  tc timeline item %d: analysis %.2f us, mid %.2f us, synthesis %.2f us%s\n", it, (tl[4 * it + 1] - tl[4 * it]) * 1e-3,
                    (tl[4 * it + 2] - tl[4 * it + 1]) * 1e-3, (tl[4 * it + 3] - tl[4 * it + 2]) * 1e-3,
                    it > 0 ? "" : "  (first item: includes the cold start)");
        const unsigned long long* ta = reinterpret_cast<const unsigned long long*>(g_dbg_host + 1024 + 64);
        for (int it = 0; it < 8 && tl[4 * it] != 0; ++it)
            fprintf(stderr, "  tc timeline item %d, warp 4 (us): E1 wait-x %.2f wait-mma %.2f work %.2f | EA wait %.2f work %.2f | EB wait %.2f work %.2f | mid+other %.2f\n",
                    it, ta[8 * it] * 1e-3, ta[8 * it + 1] * 1e-3, ta[8 * it + 2] * 1e-3, ta[8 * it + 3] * 1e-3, ta[8 * it + 4] * 1e-3,
                    ta[8 * it + 5] * 1e-3, ta[8 * it + 6] * 1e-3, ta[8 * it + 7] * 1e-3);
    }
    for (unsigned int t = 0; t < 31u; ++t) {
        const unsigned int cnt = g_dbg_host[1 + t];
        if (cnt == 0) continue;
        fprintf(stderr, "  tag %u: %u waits\n", t, cnt);
        for (unsigned int i = 0; i < cnt && i < 4u; ++i) {
            const unsigned int* r = g_dbg_host + 32 + 8 * (4 * t + i);
            fprintf(stderr, "    block=%u tid=%u parity=%u aux=%u\n", r[1], r[2], r[3], r[4]);
        }
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
    // fast path: per-batch-element filter-gradient terms (B,D,k) complex64 + bias-gradient terms (B,D) fp32;
    // generic path: the low-band spectrum of g.  (Not needed when no filter gradient is requested on the fast path.)
    if (p.path == SML_PATH_FAST) return sml_xlow_bytes(B, T, D, F) + sizeof(float) * (size_t)B * D;
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

int sml_backward_allreduce(const void* g, const void* xlow, const float* w_re, const float* w_im, void* gx, float* gw_re,
                           float* gw_im, float* gb, void* workspace, size_t workspace_bytes, int B, int T, int D, int F,
                           int io_dtype, void* flat_multicast, void* flat_next, void* stream) {
    if (check_common(g, gx, B, T, D, F, io_dtype)) return 1;
    if (w_re == nullptr || w_im == nullptr) return fail("null filter pointer");
    if (flat_multicast == nullptr) return fail("null multicast pointer");
    g_err[0] = 0;
    if (io_dtype == SML_DTYPE_F32)
        return backward_impl<float>(g, xlow, w_re, w_im, gx, gw_re, gw_im, gb, workspace, workspace_bytes, B, T, D, F, io_dtype,
                                    (cudaStream_t)stream, (float*)flat_multicast, (float*)flat_next);
    return backward_impl<__nv_bfloat16>(g, xlow, w_re, w_im, gx, gw_re, gw_im, gb, workspace, workspace_bytes, B, T, D, F,
                                        io_dtype, (cudaStream_t)stream, (float*)flat_multicast, (float*)flat_next);
}

int sml_ext_supported(int B, int T, int D, int F, int io_dtype, const sml_ext* ext) {
    if (check_common(&B, &B, B, T, D, F, io_dtype)) return 1;
    g_err[0] = 0;
    const Plan p = make_plan_ext(T, D, F, io_dtype);
    ExtGeom g;
    return ext_check(p, T, D, ext, &g);
}

int sml_ext_hpart_rows(int B, int T, int D, int F, int io_dtype) {
    const Plan p = make_plan_ext(T, D, F, io_dtype);
    if (p.path != SML_PATH_FAST || B < 1) return 0;
    return B * ((D + 2 * p.P - 1) / (2 * p.P));
}

int sml_forward_ext(const void* x, const float* w_re, const float* w_im, const float* bias, void* y, void* xlow_save, int B,
                    int T, int D, int F, int io_dtype, const sml_ext* ext, void* stream) {
    if (check_common(x, y, B, T, D, F, io_dtype)) return 1;
    g_err[0] = 0;
    if (io_dtype == SML_DTYPE_F32)
        return ext_impl<float, false>(x, w_re, w_im, bias, y, xlow_save, nullptr, nullptr, nullptr, nullptr, 0, B, T, D, F, io_dtype, ext,
                                      (cudaStream_t)stream);
    return ext_impl<__nv_bfloat16, false>(x, w_re, w_im, bias, y, xlow_save, nullptr, nullptr, nullptr, nullptr, 0, B, T, D, F, io_dtype,
                                          ext, (cudaStream_t)stream);
}

int sml_backward_ext(const void* g, const void* xlow, const float* w_re, const float* w_im, void* gx, float* gw_re, float* gw_im,
                     float* gb, void* workspace, size_t workspace_bytes, int B, int T, int D, int F, int io_dtype,
                     const sml_ext* ext, void* stream) {
    if (check_common(g, gx, B, T, D, F, io_dtype)) return 1;
    g_err[0] = 0;
    if (io_dtype == SML_DTYPE_F32)
        return ext_impl<float, true>(g, w_re, w_im, nullptr, gx, const_cast<void*>(xlow), gw_re, gw_im, gb, workspace, workspace_bytes, B,
                                     T, D, F, io_dtype, ext, (cudaStream_t)stream);
    return ext_impl<__nv_bfloat16, true>(g, w_re, w_im, nullptr, gx, const_cast<void*>(xlow), gw_re, gw_im, gb, workspace,
                                         workspace_bytes, B, T, D, F, io_dtype, ext, (cudaStream_t)stream);
}

int sml_ln_stats(const void* x, void* stats, int B, int T, int T_in, int in_row0, int D, float eps, int io_dtype, void* stream) {
    if (x == nullptr || stats == nullptr) return fail("null pointer");
    if (B < 1 || T < 1 || D < 1 || T_in < 1 || in_row0 < 0 || in_row0 + T_in > T) return fail("invalid shape B=%d T=%d T_in=%d in_row0=%d D=%d", B, T, T_in, in_row0, D);
    if (io_dtype != SML_DTYPE_F32 && io_dtype != SML_DTYPE_BF16) return fail("unsupported io_dtype %d", io_dtype);
    g_err[0] = 0;
    if (io_dtype == SML_DTYPE_F32) return ln_stats_impl<float>(x, stats, B, T, T_in, in_row0, D, eps, (cudaStream_t)stream);
    return ln_stats_impl<__nv_bfloat16>(x, stats, B, T, T_in, in_row0, D, eps, (cudaStream_t)stream);
}

int sml_ln_backward(const void* gh, const void* x, const void* stats, const void* g_res, const float* chan_add, void* gx, int B,
                    int T, int T_in, int in_row0, int D, int io_dtype, void* stream) {
    if (gh == nullptr || x == nullptr || stats == nullptr || gx == nullptr) return fail("null pointer");
    if (B < 1 || T < 1 || D < 1 || T_in < 1 || in_row0 < 0 || in_row0 + T_in > T) return fail("invalid shape B=%d T=%d T_in=%d in_row0=%d D=%d", B, T, T_in, in_row0, D);
    if (io_dtype != SML_DTYPE_F32 && io_dtype != SML_DTYPE_BF16) return fail("unsupported io_dtype %d", io_dtype);
    g_err[0] = 0;
    if (io_dtype == SML_DTYPE_F32) return ln_backward_impl<float>(gh, x, stats, g_res, chan_add, gx, B, T, T_in, in_row0, D, (cudaStream_t)stream);
    return ln_backward_impl<__nv_bfloat16>(gh, x, stats, g_res, chan_add, gx, B, T, T_in, in_row0, D, (cudaStream_t)stream);
}

int sml_spectral_ema_scan(const void* chunks, const void* state_in, const float* rho, const float* theta, void* state_out, int B,
                          int S, int F, int mode, void* stream) {
    if (chunks == nullptr || rho == nullptr || theta == nullptr || state_out == nullptr) return fail("null pointer");
    if (B < 1 || S < 0 || F < 1) return fail("invalid shape B=%d S=%d F=%d", B, S, F);
    if (mode != 0 && mode != 1) return fail("unknown SpectralEMA mode %d (0 = aligned, 1 = polar)", mode);
    g_err[0] = 0;
    const long long n = (long long)B * F;
    sml::spectral_ema_scan_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const float2*)chunks, (const float2*)state_in, rho, theta, (float2*)state_out, B, S, F, mode);
    count_launch();
    SML_CUDA(cudaGetLastError());
    return 0;
}

int sml_fwd_bwd_host(const void* x, const void* g, const float* w_re, const float* w_im, const float* bias, void* y,
                     void* gx, float* gw_re, float* gw_im, float* gb, int B, int T, int D, int F, int io_dtype,
                     int chunk_batch) {
    if (check_common(x, y, B, T, D, F, io_dtype)) return 1;
    if (g == nullptr || gx == nullptr) return fail("null activation pointer");
    if (w_re == nullptr || w_im == nullptr) return fail("null filter pointer");
    g_err[0] = 0;
    if (io_dtype == SML_DTYPE_F32)
        return fwd_bwd_host_impl<float>(x, g, w_re, w_im, bias, y, gx, gw_re, gw_im, gb, B, T, D, F, io_dtype, chunk_batch);
    return fwd_bwd_host_impl<__nv_bfloat16>(x, g, w_re, w_im, bias, y, gx, gw_re, gw_im, gb, B, T, D, F, io_dtype, chunk_batch);
}

int sml_host_release(void) {
    int dev = 0;
    SML_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_pipe_mu);
    auto it = g_pipes.find(dev);
    if (it == g_pipes.end()) return 0;
    HostPipe& hp = it->second;
    SML_CUDA(cudaDeviceSynchronize());
    pipe_free(hp);
    return 0;
}

int sml_release(void) {
    // constant tables (twiddles, DFT matrices) of every device this process has used; they are rebuilt on the next call
    SML_CUDA(cudaDeviceSynchronize());
    {
        std::lock_guard<std::mutex> lk(g_mu);
        for (auto& dv : g_dev) {
            for (auto& kv : dv.second.twiddles) cudaFree(kv.second);
            dv.second.twiddles.clear();
            for (auto& kv : dv.second.split) cudaFree(kv.second.buf);
            dv.second.split.clear();
        }
    }
    return sml_host::tc_release_tables();
}

int sml_wirtinger_mul_forward(const void* x, const void* w, void* out, long long B, long long N, void* stream) {
    if (!x || !w || !out) return fail("null pointer");
    if (B < 1 || N < 1) return fail("invalid shape B=%lld N=%lld", B, N);
    g_err[0] = 0;
    const long long nb = ((N + 1) / 2 + 255) / 256;
    // enough batch slices to fill the machine (148 SMs x 8 CTAs) when N alone does not
    long long by = (148 * 8 + nb - 1) / nb;
    if (by > B) by = B;
    if (by > 65535) by = 65535;
    sml::wirtinger_mul_fwd_kernel<<<dim3((unsigned)nb, (unsigned)by), 256, 0, (cudaStream_t)stream>>>((const float2*)x, (const float2*)w, (float2*)out, B, N);
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
    dim3 blk(32, 8), grid((D + 31) / 32, (T + 7) / 8, 1);
    long long bz = (148ll * 8 + (long long)grid.x * grid.y - 1) / ((long long)grid.x * grid.y);
    grid.z = (unsigned)(bz > B ? B : (bz > 65535 ? 65535 : bz));
    if ((T + 7) / 8 > 65535) return fail("T too large for this entry point");
    sml::wirtinger_filter_fwd_kernel<<<grid, blk, 0, (cudaStream_t)stream>>>((const float2*)x_freq, w_re, w_im, (float2*)out, B, T, D, F, k);
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
    const int rows = T > F ? T : F;      // rows f < T carry gx, rows f < F carry the gradient columns (zero beyond k)
    dim3 blk(32, 8), grid((D + 31) / 32, (rows + 7) / 8);
    if ((rows + 7) / 8 > 65535) return fail("T / F too large for this entry point");
    sml::wirtinger_filter_bwd_kernel<<<grid, blk, 0, s>>>((const float2*)g, (const float2*)x_freq, w_re, w_im, (float2*)gx, gw_re, gw_im, B, T, D, F, k);
    count_launch();
    SML_CUDA(cudaGetLastError());
    return 0;
}

}   // extern "C"
