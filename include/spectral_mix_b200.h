/* spectral_mix_b200.h -- C ABI of libspectral_mix_b200.so (sm_100a only, no CPU fallback).
 *
 * Drop-in boundary for ONE hot path of fricker2025-star/Tensor-Cuda-FFT- ("FFT-Tensor"):
 * fft_tensor.spectral_layers.SpectralMixingLayer forward + backward and the Wirtinger filter multiply.
 * The reference has no FFI of its own for this path (its boundary is the nn.Module; the only native
 * convention is the absent `fft_tensor_cuda` extension, setup.py:21-50, tensor.py:13-18), so each entry
 * point below cites the reference Python lines it replaces.  All paths are relative to /root/reference.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device, owned by the caller, contiguous
 *     (sml_fwd_bwd_host alone takes HOST pointers); activations should be 16-byte aligned: unaligned ones are served
 *     by the generic kernels, which always need the xlow and workspace buffers;
 *   - activations x/y/g/gx are (B, T, D) row-major, element type given by io_dtype;
 *   - filter parameters and their gradients are fp32 (D, F) row-major, bias (D,)   (spectral_layers.py:57-61);
 *   - k = min(F, T/2) live bins (spectral_layers.py:94); xlow is the library's own layout (B, D, k) complex64;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it, no internal sync.  Exception: the first
 *     call for a new (device, T) allocates and fills constant tables and synchronises once -- warm a shape up before
 *     capturing it in a CUDA graph (a first call on a capturing stream fails with a message saying so);
 *   - return value 0 = ok; non-zero = error, message from sml_last_error() (thread local);
 *     an unsupported argument is an error, never a silent fallback.
 */
#ifndef SPECTRAL_MIX_B200_H
#define SPECTRAL_MIX_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SML_DTYPE_F32 0
#define SML_DTYPE_BF16 1

#define SML_PATH_FAST 1    /* fused streamed band-limited FFT kernel (T a multiple of the sub-transform length M) */
#define SML_PATH_GENERIC 2 /* direct band-limited DFT kernels (any T, D, F) */

/* ABI version of this header (bumped on any signature change). */
int sml_abi_version(void);

/* Last error message of the calling thread ("" if none). */
const char* sml_last_error(void);

/* Which kernels a given problem maps to.  path: SML_PATH_*; M: sub-transform length (fast path), R = T/M passes. */
int sml_plan(int B, int T, int D, int F, int io_dtype, int* path, int* M, int* R, int* k);

/* Bytes of the saved low-band spectrum X_low = fft(x)[:, :k, :] in library layout (B, D, k) complex64. */
size_t sml_xlow_bytes(int B, int T, int D, int F);

/* Scratch bytes sml_backward needs (fast path: only when filter gradients are requested). */
size_t sml_workspace_bytes(int B, int T, int D, int F, int io_dtype);

/* Forward.  Replaces SpectralMixingLayer.forward steps 1-4, spectral_layers.py:88-116:
 *   y = Re(ifft(lowpass(fft(x, dim=1) * complex(w_re, w_im)[:, :k].T), dim=1)) + bias
 * bias may be NULL (no bias add).  xlow_save may be NULL on the fast path when no filter gradient will be
 * needed; when non-NULL it receives X_low for sml_backward.  (Dropout, :118, stays in the host module.) */
int sml_forward(const void* x, const float* w_re, const float* w_im, const float* bias, void* y,
                void* xlow_save, int B, int T, int D, int F, int io_dtype, void* stream);

/* Backward.  Replaces the autograd graph of spectral_layers.py:88-116, which equals
 * WirtingerGradient.backward, wirtinger_ops.py:53-82:
 *   gx    = Re(ifft(lowpass(fft(g) * conj(W))))                       (dL/dx)
 *   gW    = (1/T) sum_b fft(g)[b,:k,:] * conj(X_low[b])  -> gw_re = Re, gw_im = Im, columns >= k zero
 *   gb[d] = sum_{b,t} g[b,t,d]
 * gw_re/gw_im/gb are OVERWRITTEN (not accumulated).  Pass gw_re = gw_im = gb = NULL (and xlow = NULL) to get
 * gx only.  workspace: sml_workspace_bytes() bytes (on the fast path it may be NULL when no filter gradient is
 * requested).  The batch reduction of gW and gb is a deterministic two-phase sum (no atomics). */
int sml_backward(const void* g, const void* xlow, const float* w_re, const float* w_im, void* gx,
                 float* gw_re, float* gw_im, float* gb, void* workspace, size_t workspace_bytes, int B, int T,
                 int D, int F, int io_dtype, void* stream);

/* Backward whose batch reduction IS the data-parallel all-reduce (the one collective of the path: the filter/bias gradient
 * sum over ranks, the cross-rank continuation of wirtinger_ops.py:77-80; the reference itself has no distributed code).
 * Same as sml_backward, except that gw_re | gw_im | gb must be the three blocks of ONE flat fp32 buffer (gw_im = gw_re + D*F,
 * gb = gw_re + 2*D*F) that lives in NVLink symmetric memory on every rank of the job, and
 *   flat_multicast = the MULTICAST alias of that flat buffer (NVSwitch multicast object mapped on all ranks),
 *   flat_next      = this rank's LOCAL copy of the flat buffer of the other parity (cleared here for the next step).
 * The reduction kernel pushes its sums with multimem.red.add: the switch adds them into every rank's copy.  The caller
 * (1) keeps two flat buffers and alternates them step by step, both zero before the first use, (2) runs a cross-rank barrier
 * on the stream after the backward of every layer that shares the buffer; after it gw_* / gb hold the sum over all ranks.
 * Fused kernels only (an unsupported shape is an error); not bitwise reproducible (the switch picks the summation order). */
int sml_backward_allreduce(const void* g, const void* xlow, const float* w_re, const float* w_im, void* gx,
                           float* gw_re, float* gw_im, float* gb, void* workspace, size_t workspace_bytes, int B,
                           int T, int D, int F, int io_dtype, void* flat_multicast, void* flat_next, void* stream);

/* Forward + backward over HOST buffers (pinned memory recommended: pageable memory serialises the copies).
 * Same math as sml_forward followed by sml_backward, for hosts whose activations live in CPU memory: the batch is cut
 * into chunks of `chunk_batch` elements (0 = choose) and the host->device copy of chunk i+1, the two kernels of
 * chunk i and the device->host copy of chunk i-1 overlap on three internal streams.  All pointers are HOST pointers;
 * x, g, y, gx: (B, T, D) in io_dtype; w_re, w_im, gw_re, gw_im: (D, F) fp32; bias, gb: (D,) fp32.  bias may be NULL;
 * gw_re = gw_im = gb = NULL skips the filter gradient.  Synchronous: returns when every output is in host memory.
 * Runs on the current CUDA device; device staging buffers are cached between calls. */
int sml_fwd_bwd_host(const void* x, const void* g, const float* w_re, const float* w_im, const float* bias, void* y,
                     void* gx, float* gw_re, float* gw_im, float* gb, int B, int T, int D, int F, int io_dtype,
                     int chunk_batch);

/* Frees the device staging buffers sml_fwd_bwd_host caches on the current device (they are re-created on the next call). */
int sml_host_release(void);

/* Frees the constant tables the library keeps per (device, T) -- twiddles, DFT matrices of the tensor-core path -- on every
 * device (synchronises the current device first; the next call for a shape rebuilds its tables).  Long-running hosts that
 * sweep many sequence lengths (the reference's generate() grows T by one per step, byte_spectral_model.py:163-208) call
 * this to bound the cache. */
int sml_release(void);

/* Wirtinger filter multiply on an already transformed tensor.
 * Replaces WirtingerGradient.forward / .backward, wirtinger_ops.py:34-50 / :53-82.
 * x, g, out, gx: (B, N) complex64 (interleaved re,im); w, gw: (N,) complex64 broadcast over B.
 *   out = x * w ;  gx = g * conj(w) ;  gw[n] = sum_b g[b,n] * conj(x[b,n]) */
int sml_wirtinger_mul_forward(const void* x, const void* w, void* out, long long B, long long N, void* stream);
int sml_wirtinger_mul_backward(const void* g, const void* x, const void* w, void* gx, void* gw, long long B,
                               long long N, void* stream);

/* Low-pass Wirtinger spectral filter on a (B, T, D) complex64 spectrum.
 * Replaces WirtingerSpectralFilter.forward, wirtinger_ops.py:170-203, and its backward:
 *   out[b,f,d] = f < k ? x[b,f,d] * W[d,f] : 0
 *   gx[b,f,d]  = f < k ? g[b,f,d] * conj(W[d,f]) : 0 ;  gW[d,f] = sum_b g[b,f,d] conj(x[b,f,d]) (cols >= k zero) */
int sml_wirtinger_filter_forward(const void* x_freq, const float* w_re, const float* w_im, void* out, int B,
                                 int T, int D, int F, void* stream);
int sml_wirtinger_filter_backward(const void* g, const void* x_freq, const float* w_re, const float* w_im,
                                  void* gx, float* gw_re, float* gw_im, int B, int T, int D, int F,
                                  void* stream);

/* Number of kernel launches issued by this library in the calling process so far (bench.py's gpu_launches). */
unsigned long long sml_launch_count(void);

/* Diagnostics (SML_DEBUG=1 in the environment): prints the record left by a kernel whose mbarrier wait timed out
 * (the kernel traps instead of hanging the GPU) to stderr; returns the number of records. */
int sml_debug_dump(void);

#ifdef __cplusplus
}
#endif
#endif /* SPECTRAL_MIX_B200_H */
