// sml_inst_tc.cu -- host side of the tensor-core (tcgen05) bf16 kernels of sml_tc.cuh: constant tables (DFT-64 matrix, band
// DFT matrix, inter-stage twiddles) built once per (device, T), TMA descriptors, launch.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "sml_host.h"
#include "sml_tc.cuh"

namespace sml_host {

#define SML_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) return fail("%s failed: %s", #expr, cudaGetErrorString(_e));    \
    } while (0)

namespace {

struct TcTables {
    unsigned char* b1 = nullptr;   // shared by every T of the device
    unsigned char* b2 = nullptr;
    float2* tw = nullptr;
};
std::mutex g_tc_mu;
std::map<std::pair<int, int>, TcTables> g_tc_tables;   // (device, T)
std::map<int, unsigned char*> g_tc_b1;                 // device -> B1 image

uint16_t bf16_rn(double v) {
    const float f = (float)v;
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
uint32_t sw128(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }   // Swizzle<3,4,3>: 16-byte chunk ^= 128-byte row (mod 8)

// B1[slot][m1], slot 0 = Re S_0, slot 1 = Re S_32, slot 2 f1 + c = Re / Im of S_f1 (f1 = 1..31): cos(2 pi f1 m1 / 64), -sin(...)
void build_b1(std::vector<unsigned char>& img) {
    img.assign(8192, 0);
    const double PI = 3.14159265358979323846;
    for (int s = 0; s < 64; ++s) {
        const int f1 = s == 0 ? 0 : s == 1 ? 32 : s / 2;
        const bool imag = s >= 2 && (s & 1);
        for (int m1 = 0; m1 < 64; ++m1) {
            const double ang = 2.0 * PI * (double)((f1 * m1) % 64) / 64.0;
            const uint16_t h = bf16_rn(imag ? -sin(ang) : cos(ang));
            memcpy(&img[sw128((uint32_t)(s * 128 + m1 * 2))], &h, 2);
        }
    }
}
// B2[(q,c')][(n,c)], f2 = q - 8, theta = 2 pi n f2 / N2:  (re,re) cos, (re,im) sin, (im,re) -sin, (im,im) cos; atoms of 64 columns
void build_b2(std::vector<unsigned char>& img, int N2) {
    const int NA = (N2 + 31) / 32;
    img.assign((size_t)NA * 4096, 0);
    const double PI = 3.14159265358979323846;
    for (int q = 0; q < 16; ++q)
        for (int n = 0; n < N2; ++n) {
            const long long m = ((long long)n * (q - 8)) % N2;
            const double ang = 2.0 * PI * (double)((m + N2) % N2) / (double)N2;
            const double cs = cos(ang), sn = sin(ang);
            for (int cr = 0; cr < 2; ++cr)
                for (int cc = 0; cc < 2; ++cc) {
                    const double v = cr == 0 ? (cc == 0 ? cs : sn) : (cc == 0 ? -sn : cs);
                    const int row = 2 * q + cr, kk = 2 * n + cc;
                    const uint16_t h = bf16_rn(v);
                    memcpy(&img[(size_t)(kk / 64) * 4096 + sw128((uint32_t)(row * 128 + (kk % 64) * 2))], &h, 2);
                }
        }
}
// tw[n][j] = W_T^{n j} (j >= 1), W_T^{32 n} (j = 0), as (cos, -sin)
void build_tw(std::vector<float2>& tw, int T, int N2) {
    tw.resize((size_t)N2 * 32);
    const double PI = 3.14159265358979323846;
    for (int n = 0; n < N2; ++n)
        for (int j = 0; j < 32; ++j) {
            const long long e = ((long long)n * (j == 0 ? 32 : j)) % T;
            const double ang = 2.0 * PI * (double)e / (double)T;
            tw[(size_t)n * 32 + j] = make_float2((float)cos(ang), (float)(-sin(ang)));
        }
}

int tc_tables(int T, cudaStream_t stream, TcTables* out) {
    int dev = 0;
    SML_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tc_mu);
    auto it = g_tc_tables.find({dev, T});
    if (it == g_tc_tables.end()) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone)
            return fail("first call for T=%d on this device builds constant tables: run the shape once before capturing a CUDA graph", T);
        TcTables t;
        const int N2 = T / 64;
        auto b1it = g_tc_b1.find(dev);
        if (b1it == g_tc_b1.end()) {
            std::vector<unsigned char> img;
            build_b1(img);
            unsigned char* p = nullptr;
            SML_CUDA(cudaMalloc(&p, img.size()));
            SML_CUDA(cudaMemcpy(p, img.data(), img.size(), cudaMemcpyHostToDevice));
            b1it = g_tc_b1.emplace(dev, p).first;
        }
        t.b1 = b1it->second;
        std::vector<unsigned char> img2;
        build_b2(img2, N2);
        SML_CUDA(cudaMalloc(&t.b2, img2.size()));
        SML_CUDA(cudaMemcpy(t.b2, img2.data(), img2.size(), cudaMemcpyHostToDevice));
        std::vector<float2> tw;
        build_tw(tw, T, N2);
        SML_CUDA(cudaMalloc(&t.tw, tw.size() * sizeof(float2)));
        SML_CUDA(cudaMemcpy(t.tw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
        it = g_tc_tables.emplace(std::make_pair(dev, T), t).first;
    }
    *out = it->second;
    return 0;
}

// (B, T, D) bf16 activation as the 4-D tensor {d, m1, n, b}, t = N2*m1 + n; box = 32 channels x 64 m1 x 4 n: the tile lands as
// [n][m1][d], 64-byte rows.  Loads use SWIZZLE_64B (inner box = swizzle span; SWIZZLE_128B would pad the 64-byte rows to 128
// bytes, measured by tools/microbench/umma_probe.cu test 5) -- the canonical MN-major SW64 operand layout of tcgen05.
int encode_tc_map(CUtensorMap* map, const void* base, int B, int T, int D, bool swizzle) {
    EncodeTiledFn enc;
    if (get_encode_fn(&enc)) return 1;
    const cuuint64_t N2 = (cuuint64_t)T / 64;
    cuuint64_t dims[4] = {(cuuint64_t)D, 64, N2, (cuuint64_t)B};
    cuuint64_t strides[3] = {N2 * (cuuint64_t)D * 2, (cuuint64_t)D * 2, (cuuint64_t)T * D * 2};
    cuuint32_t box[4] = {32u, 64u, 4u, 1u};
    cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (tensor-core path) failed with CUresult %d", (int)r);
    return 0;
}

}   // namespace

bool tc_eligible(int T, int D, int k, int io_dtype) {
    return io_dtype == SML_DTYPE_BF16 && D % 32 == 0 && T % 512 == 0 && T >= 512 && T <= 16384 && k >= 1 && k <= 512 && k <= T / 2;
}

template <bool BWD, int MODE>
int launch_tc_inst(const CUtensorMap& map_in, const sml::TcParams& prm, int grid, size_t smem_bytes, cudaStream_t stream) {
    auto kern = sml::sml_tc_kernel<BWD, MODE>;
    static std::atomic<unsigned long long> attr_done{0};   // per kernel instantiation: one bit per device ordinal
    int dev = 0;
    SML_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(attr_done.load(std::memory_order_acquire) & bit)) {
        SML_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_done.fetch_or(bit, std::memory_order_release);
    }
    SML_CUDA(launch_pdl(kern, dim3(grid), dim3(sml::tc::THREADS), smem_bytes, stream, map_in, prm));
    count_launch();
    return 0;
}

template <bool BWD>
int launch_tc(const void* in, void* out, const TcLaunch& a, int sm_count, cudaStream_t stream) {
    TcTables tabs;
    if (tc_tables(a.T, stream, &tabs)) return 1;
    CUtensorMap map_in;
    if (encode_tc_map(&map_in, in, a.B, a.T, a.D, true)) return 1;
    sml::TcParams prm{};
    prm.w_re = a.w_re; prm.w_im = a.w_im; prm.bias = a.bias;
    prm.xlow = a.xlow; prm.gw_re = a.gw_re; prm.gpart = a.gpart; prm.gbpart = a.gbpart;
    prm.b1_img = tabs.b1; prm.b2_img = tabs.b2; prm.tw = tabs.tw;
    prm.B = a.B; prm.T = a.T; prm.D = a.D; prm.F = a.F; prm.k = a.k;
    prm.N2 = a.T / 64;
    prm.ntd = a.D / 32;
    prm.nitems = a.B * prm.ntd;
    prm.out = out;
    prm.invT = 1.0f / (float)a.T;
    prm.dbg = a.dbg;
    // as many x-tile landing slots as fit into 227 KB of shared memory (the analysis phase is TMA-latency bound)
    const uint32_t fixed = sml::tc::smem_map(prm.N2, 0).total;
    int nslot = (int)((227u * 1024u - fixed) / (16384u + 1024u));
    if (nslot > sml::tc::MAX_SLOT) nslot = sml::tc::MAX_SLOT;
    if (nslot < 4) return fail("internal: tensor-core path without room for its x tiles (T=%d)", a.T);
    prm.nslot = nslot;
    static const int cfence = [] { const char* e = getenv("SML_TC_CFENCE"); return e ? atoi(e) : 0; }();
    prm.consumer_fence = cfence;
    const size_t smem_bytes = sml::tc::smem_map(prm.N2, nslot).total;
    const int grid = prm.nitems < sm_count ? prm.nitems : sm_count;
    // bring-up aid: SML_TC_DUMP=<file> writes the intermediates of work item 0 of a FORWARD launch (tools/tc_dump_check.py)
    static const char* dump_path = getenv("SML_TC_DUMP");
    if constexpr (!BWD) {
    if (dump_path != nullptr) {
        const size_t dump_floats = (size_t)prm.N2 * 2048 * 2 + 2 * 36864;
        SML_CUDA(cudaMalloc(&prm.dump, dump_floats * sizeof(float)));
        SML_CUDA(cudaMemset(prm.dump, 0, dump_floats * sizeof(float)));
        if (launch_tc_inst<BWD, 1>(map_in, prm, grid, smem_bytes, stream)) return 1;
        SML_CUDA(cudaStreamSynchronize(stream));
        std::vector<float> host(dump_floats);
        SML_CUDA(cudaMemcpy(host.data(), prm.dump, dump_floats * sizeof(float), cudaMemcpyDeviceToHost));
        if (FILE* f = fopen(dump_path, "wb")) { fwrite(host.data(), sizeof(float), dump_floats, f); fclose(f); }
        cudaFree(prm.dump);
        return 0;
    }
    if (prm.dbg != nullptr) return launch_tc_inst<BWD, 2>(map_in, prm, grid, smem_bytes, stream);   // SML_DEBUG=1: phase timing of CTA 0
    }
    return launch_tc_inst<BWD, 0>(map_in, prm, grid, smem_bytes, stream);
}

template int launch_tc<false>(const void*, void*, const TcLaunch&, int, cudaStream_t);
template int launch_tc<true>(const void*, void*, const TcLaunch&, int, cudaStream_t);

int tc_release_tables() {
    std::lock_guard<std::mutex> lk(g_tc_mu);
    for (auto& kv : g_tc_tables) {
        if (kv.second.b2) cudaFree(kv.second.b2);
        if (kv.second.tw) cudaFree(kv.second.tw);
    }
    g_tc_tables.clear();
    for (auto& kv : g_tc_b1) cudaFree(kv.second);
    g_tc_b1.clear();
    return 0;
}

}   // namespace sml_host
