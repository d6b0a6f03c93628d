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

/* ---------------------------------------------------------------------------------------------------------------------
 * Extended entry points: the transform with the hosting block's prologue / epilogue fused in (fused kernels only: an
 * unsupported shape is an error and the caller composes the unfused calls instead).  They serve
 *   - SpectralMLPBlock.forward  `x + spectral_mix(norm1(x))`                      fft_tensor/spectral_layers.py:161, :185
 *   - FixedSpectralBlock.forward: pre-LayerNorm, zero-padded causal FFT convolution with a full half-spectrum multiplier
 *     k_freq[f] * gain[c] * sigmoid(gate_freq)[f] * sigmoid(gate_ctx)[b,c] * mask[f], irfft, keep [:T], + residual
 *                                                                                 fft_lm/train_fixed_full.py:498-555
 *   - overlap_save_block_update (chunked inference)                               scripts/generate_chunked_overlap_save.py:78-177
 * All fields of sml_ext are optional (NULL / 0 = feature off; T_in = T_out = 0 means "T rows").  Geometry is that of the
 * FORWARD call; sml_backward_ext takes the same struct and swaps the roles (g has the output geometry, gx the input one).
 */
typedef struct sml_ext {
    const void* row_stats;    /* (B, T) float2 {mean, rstd} per TRANSFORM row, {0,0} on zero-padding rows (sml_ln_stats):
                                 the kernel transforms x^ = (x - mean) * rstd.  Forward only. */
    const void* residual;     /* (B, T_out, D) io_dtype: added to the output rows.  Forward only. */
    const float* chan_scale;  /* (B, D): factor on the filtered spectrum of (batch element, channel) */
    const float* w_nyq;       /* (D,): real weight of the bin T/2 (needs F >= T/2; bins 1..T/2-1 keep the layer's
                                 Re(ifft(.)) convention, i.e. an irfft multiplier H enters as w = 2 H there, w[0] = H[0]) */
    const float* sb_re;       /* (D, F) + (D, F) + (D,): "spectral bias" added to X*W (and to the bin T/2) before chan_scale --  */
    const float* sb_im;       /*   how a LayerNorm beta enters a zero-padded transform: beta[d] * rfft(rect_T_in)[f] * w0[d,f].  */
    const float* sb_nyq;      /*   Forward only; all three or none (sb_nyq only with w_nyq).                                   */
    float* x_nyq;             /* (B, D): spectrum at the bin T/2; written by the forward, read by the backward */
    float* g_nyq;             /* (B, D): backward only: per-batch-element gradient terms of w_nyq (sum over B outside) */
    float* d_core;            /* (B, D): backward only: (1/T) sum_f Re(conj(G) X W) incl. the bin T/2 (rank-one mode: against H instead of W)   */
    float* d_q;               /* (B, D): backward only: (1/T) sum_f Re(conj(G) Q).  Together the gradient of chan_scale without a pass over  */
                              /*   y: dL/dchan_scale[b,c] = d_core (+ bg[c] * d_q when the spectral bias is bg[c] * Q[f]).  Need xlow.       */
    const float* q_re;        /* (F,) Q, rank-one mode only: the d_q terms of the backward and the forward's spectral bias bg[d] * Q[f]      */
    const float* q_im;        /* (F,) */
    const float* q_nyq;       /* (1,) device scalar: Q at the bin T/2 */
    /* Rank-one filter mode (h_re != NULL; w_re / w_im of the call may then be NULL): w[d,f] = chan[d] * (h_re, h_im)[f] and, in the
     * forward, sb[d,f] = bg[d] * (q_re, q_im)[f] are formed inside the kernel -- the multiplier of FixedSpectralBlock is
     * gain[c] * H[f], there is no (D, F) array to read.  w_nyq / sb_nyq stay (D,) vectors.  Backward: the filter gradient is
     * contracted over the channels in the kernel, hpart[item, f] = sum_{c in item} chan[c] gW[b,c,f] with items = B * ceil(D / 2P)
     * rows of k complex bins (sum the rows for dL/dH; sml_ext_hpart_rows gives the row count) -- no (B, D, k) spill, no workspace;
     * d_core is then taken against H instead of W: dL/dchan[c] = sum_b chan_scale * d_core, dL/dchan_scale = chan * d_core + bg * d_q. */
    const float* h_re;        /* (F,) */
    const float* h_im;        /* (F,) */
    const float* h_nyq;       /* (1,) device scalar: H at the bin T/2 (real), for d_core */
    const float* chan;        /* (D,) */
    const float* bg;          /* (D,) or NULL */
    void* hpart;              /* (sml_ext_hpart_rows, k) complex64, backward only */
    int T_in, in_row0;        /* x holds T_in rows; x row i is transform row in_row0 + i; the other transform rows are zero */
    int T_out, out_row0;      /* y holds T_out rows = transform rows 0 .. T_out-1, the rest is dropped.  out_row0 must be 0 (TMA
                                 stores cannot start at a negative coordinate): an output window starting at row o is the phase
                                 ramp w[d,f] *= exp(2 pi i f o / T), w_nyq *= (-1)^o on the filter.  sml_backward_ext
                                 likewise needs in_row0 == 0. */
} sml_ext;

/* 0 if sml_forward_ext / sml_backward_ext can run this problem (fused plan, row windows compatible with it); else an error
 * with the reason in sml_last_error(). */
int sml_ext_supported(int B, int T, int D, int F, int io_dtype, const sml_ext* ext);

/* Rows of sml_ext.hpart for a problem (work items of the fused plan: B * channel tiles); 0 if the plan is not fused. */
int sml_ext_hpart_rows(int B, int T, int D, int F, int io_dtype);

/* sml_forward / sml_backward with the extensions above.  xlow then holds the spectrum of the normalised, zero-padded
 * input; gw_re / gw_im / gb are the gradients of the arrays that were passed in (the host maps them back onto the block's
 * own parameters). */
int sml_forward_ext(const void* x, const float* w_re, const float* w_im, const float* bias, void* y, void* xlow_save,
                    int B, int T, int D, int F, int io_dtype, const sml_ext* ext, void* stream);
int sml_backward_ext(const void* g, const void* xlow, const float* w_re, const float* w_im, void* gx, float* gw_re,
                     float* gw_im, float* gb, void* workspace, size_t workspace_bytes, int B, int T, int D, int F,
                     int io_dtype, const sml_ext* ext, void* stream);

/* LayerNorm row statistics in the layout sml_ext.row_stats expects: stats (B, T) float2, transform row t <- input row
 * t - in_row0 of x (B, T_in, D), {0,0} outside.  eps as in torch.nn.LayerNorm (biased variance). */
int sml_ln_stats(const void* x, void* stats, int B, int T, int T_in, int in_row0, int D, float eps, int io_dtype,
                 void* stream);
/* LayerNorm backward (no affine part: gamma / beta fold into the filter) fused with the skip connection's gradient:
 *   gx = rstd * (gh' - mean_d(gh') - x^ * mean_d(gh' * x^)) + g_res ;  gh' = gh + chan_add[b, :]
 * gh = dL/dx^ (B, T_in, D); g_res (B, T_in, D) nullable; chan_add (B, D) fp32 nullable: a per-(batch element, channel)
 * constant on dL/dx^ (the gradient of a mean over the rows: FixedSpectralBlock's pooled context, train_fixed_full.py:531). */
int sml_ln_backward(const void* gh, const void* x, const void* stats, const void* g_res, const float* chan_add, void* gx,
                    int B, int T, int T_in, int in_row0, int D, int io_dtype, void* stream);

/* SpectralEMA.scan, fft_lm/spectral_ssm.py:107-125 (update: :78-105): chunks (B, S, F) complex64, state_in (B, F)
 * complex64 or NULL (zeros), rho / theta (F,) fp32, state_out (B, F) complex64.  mode 0 = "aligned", 1 = "polar". */
int sml_spectral_ema_scan(const void* chunks, const void* state_in, const float* rho, const float* theta, void* state_out,
                          int B, int S, int F, int mode, void* stream);

/* Number of kernel launches issued by this library in the calling process so far (bench.py's gpu_launches). */
unsigned long long sml_launch_count(void);

/* Diagnostics (SML_DEBUG=1 in the environment): prints the record left by a kernel whose mbarrier wait timed out
 * (the kernel traps instead of hanging the GPU) to stderr; returns the number of records. */
int sml_debug_dump(void);

#ifdef __cplusplus
}
#endif
#endif /* SPECTRAL_MIX_B200_H */
