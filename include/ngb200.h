/*
 * ngb200.h -- C ABI of libngb200.so, the B200 (sm_100a) kernel library behind neograd's tensor-op hot path.
 *
 * The reference (pranftw/neograd v0.0.4) has no FFI: its extension point is the Python `Operation`
 * protocol (README.md:44-52, neograd/autograd/ops/operation.py:83-117).  Each entry point below replaces the
 * NumPy expression inside one `Operation.forward()` / grad_fn body; the citation after each declaration is the
 * reference line whose arithmetic it reproduces (paths relative to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer to float32 unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void*; calls enqueue work and never synchronise the stream;
 *   - the library never takes ownership of caller memory; it owns one internal scratch arena per device
 *     (split-K partials, reduction partials), sized by ngb_init();
 *   - return value: 0 = NGB_OK, otherwise an ngb_status; ngb_last_error() gives the message (thread local);
 *   - shapes are int64 arrays, row-major (C order), exactly NumPy's `.shape`;
 *   - conv / pool outputs and upstream gradients use "neograd order": output position (p,q) of the textbook
 *     (Ho,Wo) map is stored at flat index q*Ho+p of the (Ho,Wo)-shaped result (conv.py:39-47 vs 149-151).
 */
#ifndef NGB200_H_
#define NGB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGB_ABI_VERSION 1

typedef void* ngb_stream_t;

typedef enum {
  NGB_OK = 0,
  NGB_ERR_INVALID = 1,      /* bad argument (maps to ValueError on the Python side) */
  NGB_ERR_CUDA = 2,         /* CUDA runtime / driver failure */
  NGB_ERR_UNSUPPORTED = 3,  /* shape outside what the kernels implement */
  NGB_ERR_SCRATCH = 4       /* internal scratch arena too small (call ngb_init with a larger size) */
} ngb_status;

typedef enum { NGB_ADD = 0, NGB_SUB = 1, NGB_MUL = 2, NGB_DIV = 3, NGB_POW = 4, NGB_SECOND = 5 } ngb_binary_op;
typedef enum { NGB_EXP = 0, NGB_LOG = 1, NGB_RELU = 2, NGB_LEAKY_RELU = 3, NGB_SIGMOID = 4, NGB_TANH = 5 } ngb_unary_op;

/* GEMM engine selection for ngb_gemm */
typedef enum {
  NGB_GEMM_AUTO = 0,   /* tcgen05 TF32 when the shape/alignment qualifies, else SIMT fp32 */
  NGB_GEMM_SIMT = 1,   /* fp32 FMA on CUDA cores (exact fp32 products) */
  NGB_GEMM_TCGEN05 = 2 /* force the tcgen05/TMEM/TMA TF32 kernel; NGB_ERR_UNSUPPORTED if not eligible */
} ngb_gemm_engine;

/* ---- library ---------------------------------------------------------------------------------------------- */
int ngb_abi_version(void);
const char* ngb_last_error(void);
/* Allocates the per-device scratch arena (bytes; 0 = default 64 MiB). Idempotent; must run before graph capture. */
int ngb_init(int64_t scratch_bytes);
int ngb_shutdown(void);
/* Debug aid: synchronises the device and reports (NGB_ERR_CUDA + ngb_last_error) if a tcgen05 producer / MMA / epilogue
 * warp gave up waiting on a pipeline barrier since the last call (such waits are bounded so a bug cannot hang the GPU). */
int ngb_debug_check_pipelines(void);
/* Non-blocking health check for the normal training path (no synchronisation: reads host-mapped words the kernels write):
 * NGB_ERR_CUDA + ngb_last_error() if a tcgen05 pipeline wait gave up (the kernel left early, its output is incomplete) or
 * the data-parallel peer exchange failed (ngb_xgpu_status).  neograd_b200 polls it every 16 optimizer steps and every 64
 * replays of a captured step. */
int ngb_runtime_status(void);
/* Host only (no GPU): the bank-conflict-searched shared-memory layout of the small-channel conv weight-gradient tile for a
 * layer: out[0..5] = row stride, column stride, plane pitch, taps-kh-fastest flag, wavefronts per gather load x 1000 with
 * that layout and with the natural one. */
int ngb_debug_wgrad_layout(int64_t C, int64_t H, int64_t W, int64_t KH, int64_t KW, int64_t pad, int64_t stride, int* out);
/* Pre-sizes the grow-on-demand workspace the tensor-core convolutions stage their channels-last operand copies in
 * (N*C*H*W*4 bytes of the largest staged tensor).  Only needed before CUDA-graph capture, where growing is not allowed. */
int ngb_workspace_reserve(int64_t bytes);
/* Lanes: the host may drive up to 4 streams concurrently (neograd_b200 runs the weight / bias gradients of
 * Tensor.backward() -- tensor.py:93-100, whose results nothing downstream consumes -- beside the data-gradient chain).
 * Kernels of different lanes can overlap on the GPU, so each lane owns its scratch arena and conv workspace.
 * ngb_set_lane selects the lane of the CALLING THREAD's subsequent launches and returns the previous one (-1 on error);
 * the first selection of a lane allocates its arena, so select every lane once before a CUDA-graph capture. */
int ngb_set_lane(int lane);
/* Hint for the persistent kernels: corun = 1 -> the calling thread's next launches are expected to overlap another
 * lane's kernels (they then leave part of every SM free); 0 (default) -> full occupancy.  Returns the previous value. */
int ngb_set_corun(int corun);
/* Process-wide precision mode: what NGB_GEMM_AUTO (and the conv dispatcher) resolve to.  NGB_GEMM_AUTO (default) = TF32
 * tensor cores whenever the shape qualifies; NGB_GEMM_SIMT = exact-fp32 CUDA-core kernels everywhere (parity runs,
 * grad_check); NGB_GEMM_TCGEN05 is not a valid default.  Returns the previous value. */
int ngb_set_default_engine(int engine);
/* Which engine ngb_gemm(NGB_GEMM_AUTO) would pick: returns NGB_GEMM_SIMT or NGB_GEMM_TCGEN05. */
int ngb_gemm_query(int transA, int transB, int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc);
/* Number of kernels this library has launched in this process (for bench.py's gpu_launches). */
int64_t ngb_launch_count(void);

/* ---- Dot: neograd/autograd/ops/basics.py:180-207 ---------------------------------------------------------- */
/* C[M,N] (row-major, ldc) = op(A)[M,K] * op(B)[K,N] (+ C if accumulate).  transA=0: A is M x K (lda); transA=1: A is
 * stored K x M.  transB likewise.  fwd: np.dot(A,B) basics.py:194; dgrad: np.dot(ug, B.T) :206 (transB=1);
 * wgrad: np.dot(A.T, ug) :207 (transA=1). */
int ngb_gemm(int engine, int transA, int transB, int64_t M, int64_t N, int64_t K,
             const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
             int accumulate, ngb_stream_t stream);

/* ---- Linear layers in one kernel per pass: neograd/nn/layers/misc.py:63 (+ the activation layer that follows) ------ */
/* `act` is an ngb_unary_op (NGB_RELU, NGB_LEAKY_RELU with slope `alpha`, NGB_SIGMOID, NGB_TANH) or -1 for none.
 * x (batch, n_in), W (n_in, n_out), b (n_out) or NULL, all row-major and dense.
 * forward: z = x.W + b (misc.py:63: dot basics.py:194 + broadcast add basics.py:20) and a = act(z) (activations.py:20, 205,
 * 55, 87) from the same accumulators; either output pointer may be NULL (not both). */
int ngb_linear_fwd(const float* x, const float* W, const float* b, float* z, float* a, int act, float alpha,
                   int64_t batch, int64_t n_in, int64_t n_out, ngb_stream_t stream);
/* data gradient: dx (+)= (ug . W^T) * act'(zprev)  -- basics.py:206 followed by the backward of the activation that
 * produced this layer's input from zprev (activations.py:31, 216, 63, 99); zprev (batch, n_in) may be NULL (no mask). */
int ngb_linear_dgrad(const float* ug, const float* W, const float* zprev, float* dx, int act, float alpha,
                     int64_t batch, int64_t n_in, int64_t n_out, int accumulate, ngb_stream_t stream);
/* weight and bias gradient from one pass over ug: dW (+)= x^T . ug (basics.py:207), db (+)= column sums of ug (the
 * `unbroadcast` of the bias add: autograd/utils.py:75 via tensor.py:93-100); db may be NULL. */
int ngb_linear_wgrad(const float* x, const float* ug, float* dW, float* db, int64_t batch, int64_t n_in, int64_t n_out,
                     int accumulate_w, int accumulate_b, ngb_stream_t stream);
/* classifier head: z = x.W + b, loss = SoftmaxCE(z, targets) along the class axis (nn/loss.py:153-157) and
 * dz = (softmax(z) - targets) * inv_num_examples (loss.py:170 with upstream gradient 1) in one kernel; z may be NULL.
 * Layers with at most 32 outputs (ngb_linear_softmax_ce_serves tells; NGB_ERR_UNSUPPORTED otherwise). */
int ngb_linear_softmax_ce(const float* x, const float* W, const float* b, const float* targets, float* z, float* loss,
                          float* dz, int64_t batch, int64_t n_in, int64_t n_out, float inv_num_examples, float epsilon,
                          ngb_stream_t stream);
int ngb_linear_softmax_ce_serves(int64_t batch, int64_t n_in, int64_t n_out);

/* ---- Conv3D / Conv2D: neograd/autograd/ops/conv.py:121-343 ------------------------------------------------- */
/* x (N,C,H,W), w (O,C,KH,KW), b (O) -> y (N,O,Ho,Wo) in neograd order.  neograd's single-channel Conv2D
 * (conv.py:131-152) is the C=O=1 case with a one-element bias. */
int ngb_conv_fprop(const float* x, const float* w, const float* b, float* y,
                   int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int64_t KH, int64_t KW,
                   int64_t pad, int64_t stride, ngb_stream_t stream);                       /* conv.py:245-269 */
int ngb_conv_dgrad(const float* ug, const float* w, float* dx,
                   int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int64_t KH, int64_t KW,
                   int64_t pad, int64_t stride, int accumulate, ngb_stream_t stream);       /* conv.py:299-309 */
int ngb_conv_wgrad(const float* ug, const float* x, float* dw,
                   int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int64_t KH, int64_t KW,
                   int64_t pad, int64_t stride, int accumulate, ngb_stream_t stream);       /* conv.py:311-321 */
int ngb_conv_bgrad(const float* ug, float* db, int64_t N, int64_t O, int64_t HoWo,
                   int accumulate, ngb_stream_t stream);                                    /* conv.py:323-327, 198-199 */

/* Conv -> ReLU -> MaxPool(2x2, stride 2) blocks, forward in ONE kernel (conv.py:266-297 + activations.py:20 + conv.py:384-402):
 * y_pool / idx as ngb_maxpool_relu_fwd would produce them from ngb_conv_fprop's output, which is never written.  fp32 FMAs.
 * First-layer geometries only (C = 1, pad 0, stride 1, square even output); otherwise NGB_ERR_UNSUPPORTED. */
int ngb_conv_relu_pool_fwd(const float* x, const float* w, const float* b, float* y_pool, uint8_t* idx, int64_t N, int64_t C,
                           int64_t H, int64_t W, int64_t O, int64_t KH, int64_t KW, int64_t pad, int64_t stride,
                           ngb_stream_t stream);

/* 1 when ngb_conv_relu_pool_fwd serves the geometry (a caller that defers the convolution until its consumer is known asks first) */
int ngb_conv_relu_pool_fwd_serves(int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int64_t KH, int64_t KW,
                                  int64_t pad, int64_t stride);

/* Conv -> ReLU -> MaxPool(2x2, stride 2) blocks, backward: the gradient of the convolution's output is three quarters
 * zeros (one cell per pooling window) and never has to exist in HBM.  ngb_conv_wgrad_pooled computes conv.py:311-321 from
 * the pool's upstream gradient ug_pool, the pool's output y_pool, its argmax indices idx and the conv output z (= the ReLU's
 * input): mask and scatter happen in the kernel's shared memory.  ngb_maxpool_relu_mask writes the compact masked gradient
 * gm[e] = ug_pool[e] * relu'(z at the argmax); the bias gradient (conv.py:323-327) is the sum of gm, computed without
 * writing it by ngb_conv_bgrad_pooled (Ho x Wo = the conv output plane).
 * idx must be the bytes ngb_maxpool_relu_fwd wrote: bits 0-1 = argmax cell (kh*2 + kw), bit 2 = relu'(z at the argmax);
 * the conv-gradient kernels take the mask from those bytes (y_pool and z are read by ngb_maxpool_relu_mask and
 * ngb_conv_bgrad_pooled only, and may be NULL everywhere else).
 * NGB_ERR_UNSUPPORTED -> materialise the gradient with ngb_maxpool_relu_bwd and use ngb_conv_wgrad. */
int ngb_maxpool_relu_mask(const float* ug_pool, const float* y_pool, const float* z, float* gm, int64_t outer, int64_t H,
                          int64_t W, ngb_stream_t stream);
int ngb_conv_bgrad_pooled(const float* ug_pool, const float* y_pool, const float* z, float* db, int64_t N, int64_t O,
                          int64_t Ho, int64_t Wo, int accumulate, ngb_stream_t stream);
/* same sum from ug_pool and the argmax bytes alone (bit 2 = relu'(z at the argmax)); HpWp = pooled positions per plane */
int ngb_conv_bgrad_pooled_idx(const float* ug_pool, const uint8_t* idx, float* db, int64_t N, int64_t O, int64_t HpWp,
                              int accumulate, ngb_stream_t stream);
int ngb_conv_dgrad_pooled(const float* ug_pool, const float* y_pool, const uint8_t* idx, const float* z, const float* w,
                          float* dx, int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int64_t KH, int64_t KW,
                          int64_t pad, int64_t stride, int accumulate, ngb_stream_t stream);    /* conv.py:299-309 */
int ngb_conv_wgrad_pooled(const float* ug_pool, const float* y_pool, const uint8_t* idx, const float* z, const float* x,
                          float* dw, int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int64_t KH, int64_t KW,
                          int64_t pad, int64_t stride, int accumulate, ngb_stream_t stream);
/* dw and db (conv.py:311-327) in one call; db may be NULL.  First layers (C <= 3) get both from ONE pass over the pooled
 * gradient: only the argmax cell of a window carries gradient, so the weight gradient is C*KH*KW fp32 FMAs per (window, o). */
int ngb_conv_wbgrad_pooled(const float* ug_pool, const float* y_pool, const uint8_t* idx, const float* z, const float* x,
                           float* dw, float* db, int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int64_t KH,
                           int64_t KW, int64_t pad, int64_t stride, int accumulate_w, int accumulate_b, ngb_stream_t stream);
/* 1 when ngb_conv_wbgrad_pooled serves the geometry with ONE kernel, else 0 (it then runs the two separate kernels). */
int ngb_conv_wbgrad_pooled_is_fused(int64_t N, int64_t C, int64_t H, int64_t W, int64_t O, int64_t KH, int64_t KW,
                                    int64_t pad, int64_t stride);

/* ---- MaxPool2D / MaxPool3D: neograd/autograd/ops/conv.py:362-542 ------------------------------------------- */
/* x (outer,H,W) with outer = N or N*C.  y and idx (uint8, position of the FIRST maximum inside the zero-padded
 * window, row-major in the window; np.argmax conv.py:424/523) are (outer,Ho,Wo) in neograd order.  idx may be NULL. */
int ngb_maxpool_fwd(const float* x, float* y, uint8_t* idx, int64_t outer, int64_t H, int64_t W,
                    int64_t KH, int64_t KW, int64_t pad, int64_t stride, ngb_stream_t stream);  /* conv.py:384-402 */
/* dx (outer,H,W): each cell takes the value the LAST window covering it (fragment order: W outer, H inner) assigned,
 * i.e. ug if the cell is that window's argmax else 0 (assignment semantics conv.py:427/526); uncovered cells = 0. */
int ngb_maxpool_bwd(const float* ug, const uint8_t* idx, float* dx, int64_t outer, int64_t H, int64_t W,
                    int64_t KH, int64_t KW, int64_t pad, int64_t stride, int accumulate,
                    ngb_stream_t stream);                                                   /* conv.py:404-431 */
/* ReLU -> MaxPool pairs (activations.py:20 followed by conv.py:384-402) without the ReLU's output ever touching HBM:
 *   fwd: y / idx = maxpool(max(0, x));    bwd: dx (+)= (xm >= 0) * maxpool_bwd(ug, idx)   (xm = the ReLU's input).
 * Only 2x2 / stride 2 / pad 0 on 16-byte aligned rows (even H for bwd); other geometries return NGB_ERR_UNSUPPORTED
 * and the caller runs the two ops separately.  Results are bit-identical to the separate kernels, except that the fused
 * forward also records the ReLU gradient at the argmax in bit 2 of every idx byte (1 = gradient flows: the maximum is
 * positive, or the window was clamped and its first cell is exactly zero or above -- activations.py:31).  Every kernel that
 * reads the idx of a pool with at most four cells ignores that bit; the pooled conv-gradient kernels above use it. */
int ngb_maxpool_relu_fwd(const float* x, float* y, uint8_t* idx, int64_t outer, int64_t H, int64_t W,
                         int64_t KH, int64_t KW, int64_t pad, int64_t stride, ngb_stream_t stream);
int ngb_maxpool_relu_bwd(const float* ug, const uint8_t* idx, const float* xm, float* dx, int64_t outer, int64_t H,
                         int64_t W, int64_t KH, int64_t KW, int64_t pad, int64_t stride, int accumulate,
                         ngb_stream_t stream);

/* ---- broadcasting elementwise ops + unbroadcast: basics.py:6-339, utils.py:34-78, tensor.py:93-100 -------- */
/* out = a (op) b with NumPy broadcasting.  A NULL operand pointer means "use the scalar immediate". */
int ngb_ew_binary_fwd(int op, const float* a, float a_scalar, const int64_t* a_shape, int a_rank,
                      const float* b, float b_scalar, const int64_t* b_shape, int b_rank,
                      float* out, ngb_stream_t stream);                                     /* basics.py:20,63,107,150,313 */
/* Fused: g = dLocal(op, side; a, b) * ug over the broadcast shape, summed over the axes along which operand `side`
 * was broadcast (utils.py:73-78), written (accumulate=0) or added (accumulate=1) to grad (operand-shaped). */
int ngb_ew_binary_bwd(int op, int side, const float* ug,
                      const float* a, float a_scalar, const int64_t* a_shape, int a_rank,
                      const float* b, float b_scalar, const int64_t* b_shape, int b_rank,
                      float* grad, int accumulate, ngb_stream_t stream);                    /* basics.py:32-33,76-77,119-120,163-164,325-327 */
int ngb_ew_unary_fwd(int op, float alpha, const float* x, float* out, int64_t n, ngb_stream_t stream);
                                                                                            /* basics.py:236,274; activations.py:20,54,86,205 */
int ngb_ew_unary_bwd(int op, float alpha, const float* x, const float* ug, float* out, int64_t n,
                     int accumulate, ngb_stream_t stream);                                  /* basics.py:246,284; activations.py:31,62-63,94-95,216 */
/* out = sum of x over the axes whose bit is set in axis_mask (bit d = axis d); out has the remaining dims. */
int ngb_reduce_sum(const float* x, const int64_t* shape, int rank, uint32_t axis_mask, float* out,
                   int accumulate, ngb_stream_t stream);                                    /* basics.py:369; utils.py:75 */

/* ---- small dense helpers used by the host engine ----------------------------------------------------------- */
int ngb_fill(float* dst, float value, int64_t n, ngb_stream_t stream);
int ngb_axpy(float* dst, const float* src, float alpha, int64_t n, ngb_stream_t stream);  /* dst += alpha*src; tensor.py:301 */
int ngb_transpose2d(const float* src, float* dst, int64_t rows, int64_t cols, ngb_stream_t stream); /* basics.py:417 */
int ngb_copy(float* dst, const float* src, int64_t n, ngb_stream_t stream);                /* device-to-device, async */
/* dst = np.transpose(src, perm) materialised C-contiguous; shape is src's shape, rank <= 6. */
int ngb_permute(const float* src, float* dst, const int64_t* shape, const int* perm, int rank, ngb_stream_t stream);

/* ---- SoftmaxCE: neograd/nn/loss.py:124-172 ----------------------------------------------------------------- */
/* logits/targets viewed as (outer, L, inner); softmax along L.  loss (1 float) = -(1/num_examples) * sum t*log(p+eps). */
int ngb_softmax_ce_fwd(const float* logits, const float* targets, int64_t outer, int64_t L, int64_t inner,
                       float inv_num_examples, float epsilon, float* loss, ngb_stream_t stream);   /* loss.py:153-157 */
/* dlogits = (ug[0] * inv_num_examples) * (softmax(logits) - targets); ug is a device scalar (may be NULL = 1). */
int ngb_softmax_ce_bwd(const float* logits, const float* targets, const float* ug, int64_t outer, int64_t L,
                       int64_t inner, float inv_num_examples, float* dlogits, int accumulate,
                       ngb_stream_t stream);                                                /* loss.py:167-170 */

/* ---- Softmax layer: neograd/nn/activations.py:105-180 ------------------------------------------------------- */
/* x viewed as (outer, L, inner); y = exp(x - max) / sum along L (activations.py:161-174). */
int ngb_softmax_fwd(const float* x, int64_t outer, int64_t L, int64_t inner, float* y, ngb_stream_t stream);
/* dx = Jacobian(softmax(x)) . ug per slice (activations.py:131-158) = s * (ug - sum_L(ug * s)). */
int ngb_softmax_bwd(const float* x, const float* ug, int64_t outer, int64_t L, int64_t inner, float* dx,
                    int accumulate, ngb_stream_t stream);

/* BCE in one pass: cost = (-1/num_examples) * sum(o * log(t + eps) + (1 - o) * log(1 - t + eps)) -- nn/loss.py:83-84,
 * with the reference's argument roles as they are (outputs multiply the logs of the targets).  o, t: n elements each. */
int ngb_bce_fwd(const float* outputs, const float* targets, int64_t n, float inv_num_examples, float epsilon, float* loss,
                ngb_stream_t stream);
/* its gradients for an upstream gradient ug (device scalar, NULL = 1): d_outputs and / or d_targets (NULL = not wanted) */
int ngb_bce_bwd(const float* outputs, const float* targets, const float* ug, int64_t n, float inv_num_examples,
                float epsilon, float* d_outputs, int accumulate_o, float* d_targets, int accumulate_t, ngb_stream_t stream);

/* ---- optimizers: neograd/nn/optim.py ------------------------------------------------------------------------ */
/* One fused pass over a flat parameter arena.  g is multiplied by grad_scale first (1/world for data parallel).
 * Hyper-parameters are doubles (the reference computes in float64): 1-beta and the bias corrections are formed in
 * double on the host and only then rounded to fp32.  `t` is the iteration count AFTER the increment of optim.py:169. */
int ngb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                  double epsilon, int64_t t, double grad_scale, ngb_stream_t stream);       /* optim.py:168-179, 191-197 */
/* CUDA-graph friendly variant: the iteration count lives in device memory, so a captured training step replays with the
 * right bias correction.  t_dev points to TWO int32: [0] = steps taken so far, [1] = scratch counter (0 between calls).
 * tick = 1: this launch is step t_dev[0] + 1 and stores it when it completes (optim.py:169 folded into the update);
 * tick = 2: same step number, nothing stored (an optimizer that updates several ranges of one arena passes 2 on all but
 * the last range); tick = 0: the step number is t_dev[0] as it stands. */
int ngb_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                      double epsilon, int32_t* t_dev, int tick, double grad_scale, ngb_stream_t stream);
int ngb_sgd_step(float* p, const float* g, int64_t n, double lr, double grad_scale, ngb_stream_t stream);  /* optim.py:42-47 */
int ngb_momentum_step(float* p, const float* g, float* m, int64_t n, double lr, double beta, double grad_scale,
                      ngb_stream_t stream);                                                 /* optim.py:63-70, 86-92 */
int ngb_rmsprop_step(float* p, const float* g, float* r, int64_t n, double lr, double beta, double epsilon,
                     double grad_scale, ngb_stream_t stream);                               /* optim.py:112-118, 134-140 */

/* ---- input pipeline on the device (new) -----------------------------------------------------------
 * A batch crosses PCIe in its source format (uint8 pixels, int32 class ids) and is expanded in HBM. */
/* dst[i] = src[i] * scale + shift  (e.g. scale = 1/(255 std), shift = -mean/std) */
int ngb_prep_u8_affine(const uint8_t* src, float* dst, int64_t n, float scale, float shift, ngb_stream_t stream);
/* per-row standardisation (x - mean) / std, population std: examples/digits.ipynb cell 5; src is uint8 or float32 */
int ngb_prep_standardize_rows(const void* src, int src_is_u8, float* dst, int64_t rows, int64_t cols, ngb_stream_t stream);
/* dst (rows, classes) = np.eye(classes)[labels]: examples/digits.ipynb cell 6 */
int ngb_prep_one_hot(const int32_t* labels, float* dst, int64_t rows, int64_t classes, ngb_stream_t stream);
/* standard-normal fill (Philox4x32-10 + Box-Muller; the role of np.random.randn in examples/gans/basic.ipynb cell 9).
 * Element i depends only on (seed, call offset, i).  offset_dev != NULL: the offset is read from -- and then advanced in --
 * device memory (a captured graph draws fresh noise on every replay); else `offset` is used. */
int ngb_randn(float* dst, int64_t n, uint64_t seed, uint32_t* offset_dev, uint32_t offset, ngb_stream_t stream);

/* ---- data parallel over NVLink peer memory (new) -------------------------------------------------------
 * The gradient exchange FUSED with the optimizer update (one launch per optimizer step and parameter range; replaces the
 * reference's single-process optim.py:42-47 / 63-92 / 112-140 / 168-197 under data parallelism).  Every rank owns a receive
 * buffer [2 parities][world][slot_stride] floats and flag words [world][max_blocks] in memory its peers can store to
 * (recv_ptrs_dev / flag_ptrs_dev: DEVICE arrays of `world` device pointers, entry r = rank r's buffer).  The kernel pushes
 * the range [off, off + n) of `grad` into slot [rank] of every peer, signals, waits for the peers' pushes (bounded by
 * timeout_s; <= 0 means 60 s), sums the slots in rank order (own gradient in place), multiplies by grad_scale and updates
 * params / state.  epoch_dev: two uint32 (exchanges completed, scratch counter).  kind: 0 Adam (state1 = m, state2 = v,
 * t_dev / tick as ngb_adam_step_dev), 1 GD, 2 Momentum (state1; beta1 = beta), 3 RMSProp (state1; beta1 = beta).
 * A failed exchange applies NO update and makes every later call a no-op; ngb_xgpu_status reports it.
 * Low-latency protocol: when the slots have room for it (2 * (off + n) <= slot_stride) and n <= 151552, every float travels as
 * one 8-byte {value, epoch} word that validates itself on arrival -- no system fence, no flag round trip (the flag words are
 * not used): the exchange then costs one NVLink store latency instead of ~25 us. */
int ngb_xgpu_exchange_update(const void* recv_ptrs_dev, const void* flag_ptrs_dev, int rank, int world, int max_blocks,
                             int64_t slot_stride, int64_t off, int64_t n, const float* grad, float* params, float* state1,
                             float* state2, uint32_t* epoch_dev, int kind, double lr, double beta1, double beta2,
                             double epsilon, int32_t* t_dev, int tick, double grad_scale, double timeout_s,
                             ngb_stream_t stream);
int ngb_xgpu_blocks(int64_t n, int max_blocks);   /* blocks the launch above uses for n floats (same on every rank) */
int ngb_xgpu_status(void);  /* non-blocking: NGB_ERR_CUDA once an exchange has timed out (reads a host-mapped word) */
int ngb_xgpu_check(void);   /* same, after synchronising the device (tests / tools) */
int ngb_xgpu_reset(void);   /* clears the sticky failure (tests that provoke a timeout) */

#ifdef __cplusplus
}
#endif
#endif /* NGB200_H_ */
