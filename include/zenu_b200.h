/* zenu_b200.h — native C ABI of the B200 (sm_100a) backend for ZeNu's CNN-training hot path.
 *
 * This is the drop-in boundary: a flat extern "C" surface in the style of zenu-cuda-kernel-sys
 * (reference: zenu-cuda-kernel-sys/kernel/kernel.h:1-18, bindgen'd by zenu-cuda-kernel-sys/build.rs:37-53)
 * that replaces, for this path only, what zenu-cuda reaches through cuDNN-frontend
 * (zenu-cudnn-frontend-wrapper-sys/cudnn_frontend_wrapper/include/cudnn_frontend_wrapper.h:100-186),
 * legacy cuDNN BatchNorm (zenu-cuda/src/cudnn/batch_norm.rs:51-91,206-249,380-414), cuBLAS GEMM
 * (zenu-cuda/src/cublas/mod.rs:84-160) and the hand kernels of zenu-cuda-kernel-sys.
 * The symbol-compatible shims for those two old interfaces are declared in
 * zenu_kernel_compat.h and zenu_cudnn_frontend_compat.h; INTEGRATION.md shows the Rust bindings.
 *
 * Conventions
 *  - every entry returns int: 0 = ZB_OK, otherwise a zb_status; zb_last_error() gives the message
 *    (reference convention replaced: kernel-sys returns void / unchecked, the FE wrapper an enum).
 *  - all pointers are DEVICE pointers unless a parameter is named host_*; the callee never allocates
 *    or frees caller tensors (reference ownership: Matrix<Owned<T>> owns every buffer,
 *    zenu-matrix/src/device/mod.rs:33-69).  Scratch lives inside zb_ctx.
 *  - sizes are int64_t (the reference uses int and overflows above 2^31 elements).
 *  - work is enqueued on the ctx stream; nothing synchronises the device implicitly.
 *  - dtype: ZB_F32 or ZB_F64 (the reference's only element types, zenu-matrix/src/num.rs:44-86).
 *  - layout: ZB_NCHW = the reference contract (activations NCHW, filters KCRS);
 *            ZB_NHWC = the backend's native layout (activations NHWC, filters KRSC), zero-copy.
 *  - math:   ZB_MATH_TF32  tcgen05 kind::tf32 tensor cores, fp32 accumulate   (rel. tol 1e-3)
 *            ZB_MATH_TF32X3 3xTF32 split on the same tensor-core kernels (hi/lo operand split, three accumulating
 *                          passes lo*hi + hi*lo + hi*hi)                         (rel. tol 1e-5)
 *            ZB_MATH_FP32  FFMA (f32) / DFMA (f64) SIMT kernels                (rel. tol 1e-5)
 *            ZB_MATH_DEFAULT = the ctx default (TF32 for f32; f64 always runs DFMA).
 */
#ifndef ZENU_B200_H
#define ZENU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zb_ctx zb_ctx;

typedef enum {
  ZB_OK = 0,
  ZB_ERR_INVALID = 1,      /* bad argument / shape mismatch (reference: shape_check panics, nn/conv/shape_check.rs:4-53) */
  ZB_ERR_CUDA = 2,         /* CUDA runtime / driver failure */
  ZB_ERR_UNSUPPORTED = 3,  /* valid request this build cannot serve (never a silent CPU fallback) */
  ZB_ERR_TIMEOUT = 4,      /* a device-side barrier wait timed out (protocol bug guard) */
  ZB_ERR_NCCL = 5
} zb_status;

typedef enum { ZB_F32 = 0, ZB_F64 = 1 } zb_dtype;
/* ZB_NCHW_X: conv fprop / wgrad only — the input x is NCHW (how the reference hands a batch to the first layer), the
 * filter is KRSC and y / dy are NHWC.  Served for C <= 4 stems on the TF32 path (the NCHW -> NHWC4 repack is part of the
 * conv's own staging pass); other shapes return ZB_ERR_UNSUPPORTED and the caller converts the layout first. */
typedef enum { ZB_NCHW = 0, ZB_NHWC = 1, ZB_NCHW_X = 2 } zb_layout;
typedef enum { ZB_MATH_DEFAULT = 0, ZB_MATH_TF32 = 1, ZB_MATH_TF32X3 = 2, ZB_MATH_FP32 = 3 } zb_math_mode;
typedef enum { ZB_OP_ADD = 0, ZB_OP_SUB = 1, ZB_OP_MUL = 2, ZB_OP_DIV = 3 } zb_binary_op;

/* ---- context ------------------------------------------------------------------------------------
 * Replaces the process-global Mutex<ZenuCudaState{cublas,cudnn,stream,mempool}> (zenu-cuda/src/lib.rs:18-112).
 * `stream` may be NULL (the library creates its own non-blocking stream) or a cudaStream_t to share. */
int zb_ctx_create(zb_ctx** out, int device, void* stream);
int zb_ctx_destroy(zb_ctx* ctx);
int zb_ctx_synchronize(zb_ctx* ctx);
int zb_ctx_set_math(zb_ctx* ctx, int math_mode);
/* BatchNorm epsilon used by every zb_bn2d_* call of this ctx.  Default 1e-10, the reference's constant on both of its paths
 * (zenu-matrix/src/nn/batch_norm.rs:296, zenu-cuda/src/cudnn/batch_norm.rs:82); the cuDNN-frontend BatchNorm shim sets the value its
 * descriptor was created with around each call. */
int zb_ctx_set_bn_epsilon(zb_ctx* ctx, double eps);
double zb_ctx_bn_epsilon(zb_ctx* ctx);
void* zb_ctx_stream(zb_ctx* ctx);
/* Side context: a child ctx on its own non-blocking stream with its own scratch arena, for work that may overlap the parent's stream
 * (the host model runs conv wgrad there: its result is only needed by the optimizer, and a tensor-bound wgrad co-resides with the
 * HBM-bound BatchNorm-backward kernels of the next layer).  Created on first use, destroyed with the parent, included in
 * zb_ctx_synchronize / zb_ctx_check / zb_ctx_launch_count.  zb_ctx_fork makes the side stream wait for everything enqueued on the
 * parent so far; zb_ctx_join makes the parent wait for everything enqueued on the side stream.  Both are event based (no host sync)
 * and capturable.  Buffers a side kernel reads must stay alive (and unwritten by the parent stream) until the join. */
zb_ctx* zb_ctx_side(zb_ctx* ctx);
int zb_ctx_fork(zb_ctx* ctx);
int zb_ctx_join(zb_ctx* ctx);
/* Non-zero status if any kernel since the last call hit a device-side timeout. Synchronises. */
int zb_ctx_check(zb_ctx* ctx);
/* Number of kernels this ctx has launched so far (bench.py reports the delta as gpu_launches). */
unsigned long long zb_ctx_launch_count(zb_ctx* ctx);
/* Optional per-op timing with CUDA events on the ctx stream (used by bench.py for the live roofline figures).
 * cls: 0 = tensor-core conv/GEMM launches (work = algorithmic FLOPs), 1 = BatchNorm ops (work = algorithmic
 * bytes), 2 = other elementwise ops (bytes).  enable clears earlier records; read synchronises the stream. */
int zb_ctx_profile_enable(zb_ctx* ctx, int enable);
int zb_ctx_profile_read(zb_ctx* ctx, int cls, int64_t* ops, double* total_ms, double* work);
const char* zb_last_error(void);
const char* zb_version(void);

/* ---- convolution --------------------------------------------------------------------------------
 * Replaces conv_fwd / conv_bkwd_data / conv_bkwd_weight (zenu-matrix/src/nn/conv/mod.rs:21-81;
 * Nvidia impl nn/conv/nvidia.rs:16-122 -> zenu-cuda/src/cudnn/graph_conv.rs:40-268 -> conv.cpp:29-187).
 * Shapes come from the live tensors (fixes SURVEY S3: the reference Conv2d layer bakes a 32x32 dummy). */
typedef struct {
  int64_t n, c, h, w;          /* input  N, C_in, H, W   */
  int64_t k, kh, kw;           /* filter C_out, k_h, k_w */
  int64_t pad_h, pad_w, stride_h, stride_w, dil_h, dil_w;
} zb_conv2d_desc;

int64_t zb_conv_out_size(int64_t in, int64_t k, int64_t pad, int64_t stride, int64_t dil);
/* y = conv(x, w) (+ bias[k] when bias != NULL, fused in the epilogue; reference: separate conv_bias_add pass) */
int zb_conv2d_fprop(zb_ctx* ctx, int dtype, int layout, int math, const zb_conv2d_desc* d, const void* x,
                    const void* w, const void* bias, void* y);
/* Conv fprop that also emits the per-channel statistics a following BatchNorm2d (training) needs, accumulated in the
 * conv epilogue while the output tile is still in registers: stat_partial [rows][2][k] = partial sums of (y - shift[k]) and
 * (y - shift[k])^2 over the pixels each row covers (NHWC output, f32, tensor-core paths).  `stat_partial` must hold
 * zb_conv2d_bnstats_rows(ctx) x 2 x k floats; `shift` [k] is any per-channel offset (the BN running mean is a good one), the
 * same pointer is handed to zb_bn2d_fwd_train_prestats.  *stat_rows = rows written, or 0 when this shape / math mode cannot
 * fuse them (the conv result is complete either way; the caller then runs the plain zb_bn2d_fwd_train). */
int zb_conv2d_bnstats_rows(zb_ctx* ctx);
int zb_conv2d_fprop_bnstats(zb_ctx* ctx, int dtype, int layout, int math, const zb_conv2d_desc* d, const void* x,
                            const void* w, const void* bias, void* y, const void* shift, void* stat_partial,
                            int64_t* stat_rows);
int zb_conv2d_dgrad(zb_ctx* ctx, int dtype, int layout, int math, const zb_conv2d_desc* d, const void* dy,
                    const void* w, void* dx);
/* dx += dgrad(dy, w): gradient fan-in (zenu-autograd/src/lib.rs:480-481 `grad + old`) folded into the dgrad epilogue */
int zb_conv2d_dgrad_acc(zb_ctx* ctx, int dtype, int layout, int math, const zb_conv2d_desc* d, const void* dy,
                        const void* w, void* dx);
/* dx = dgrad(dy, w) + mask(dx): as zb_conv2d_dgrad_acc, with the OLD dx taken through a 1-bit-per-element mask (bit i = element i of
 * dx in its layout counts, else it is 0).  The fused BN+add+ReLU backward (zb_bn2d_bwd_mask with dres == NULL) leaves the residual
 * branch's gradient as the pair (dy, ReLU bits) instead of writing the masked tensor; the convolution whose dgrad accumulates into
 * that branch applies the mask in its epilogue.  old_bits == NULL is zb_conv2d_dgrad_acc. */
int zb_conv2d_dgrad_acc_masked(zb_ctx* ctx, int dtype, int layout, int math, const zb_conv2d_desc* d, const void* dy,
                               const void* w, void* dx, const void* old_bits);
/* out[i] = bit i of `bits` ? x[i] : 0 (materialises such a masked gradient for a consumer that cannot mask on the fly; out may be x) */
int zb_mask_apply(zb_ctx* ctx, int dtype, const void* x, const void* bits, void* out, int64_t n);
int zb_conv2d_wgrad(zb_ctx* ctx, int dtype, int layout, int math, const zb_conv2d_desc* d, const void* dy,
                    const void* x, void* dw);
/* ---- plan description (parity-test support: which kernel variant serves a shape) ------------------------------------------
 * zb_conv2d_plan_describe runs the SAME dispatch and host planners as zb_conv2d_{fprop,dgrad,wgrad} in a dry mode (nothing is
 * launched, allocated or written; no tensor needs to exist) and writes one "kernel<variant> key=value ...;" segment per launch the
 * call would make into buf (NUL-terminated, truncated to cap).  Values that depend on the batch size rather than on the kernel variant
 * carry a '~' prefix, so a test can assert that the variant exercised at a small batch is the one the benchmarked batch takes.
 * Returns the buffer size needed (call with buf == NULL to size it), or -(zb_status) on failure.
 * zb_ctx_plan_trace(ctx, 1) records the same segments for REAL calls made by this thread; zb_ctx_plan_trace_read returns and clears them. */
typedef enum { ZB_PLAN_FPROP = 0, ZB_PLAN_DGRAD = 1, ZB_PLAN_WGRAD = 2 } zb_plan_op;
typedef enum { ZB_PLAN_BNSTATS = 1, ZB_PLAN_BIAS = 2, ZB_PLAN_ACCUMULATE = 4 } zb_plan_flags;
int64_t zb_conv2d_plan_describe(zb_ctx* ctx, int op, int dtype, int layout, int math, const zb_conv2d_desc* d, int flags, char* buf,
                                int64_t cap);
int zb_ctx_plan_trace(zb_ctx* ctx, int enable);
int64_t zb_ctx_plan_trace_read(zb_ctx* ctx, char* buf, int64_t cap);

/* conv2d_bias_add / conv2d_bias_bkwd (nn/conv/mod.rs:84-107; kernels array_array.cu:49-74,
 * conv2d_bkwd_data.cu:121-195 — the latter is wrong for N>1 in the reference; this one is not). */
int zb_conv2d_bias_add(zb_ctx* ctx, int dtype, int layout, const void* x, const void* bias, void* y, int64_t n,
                       int64_t k, int64_t h, int64_t w);
int zb_conv2d_bias_bwd(zb_ctx* ctx, int dtype, int layout, const void* dy, void* dbias, int64_t n, int64_t k,
                       int64_t h, int64_t w);

/* ---- batch normalisation ------------------------------------------------------------------------
 * Replaces BatchNormalization::{batch_norm_2d_forward_train, batch_norm_2d_backward,
 * bach_norm_2d_forward_inference} (zenu-matrix/src/nn/batch_norm.rs:130-280).  eps = 1e-10, momentum
 * weights the OLD running stat, running variance unbiased, saved = mean and 1/sqrt(var+eps).
 * Fusions (not in the reference, which runs relu / add as separate kernels):
 *   relu != 0        y = max(bn(x) [+ residual], 0)
 *   residual != NULL y = bn(x) + residual
 * saved_mean / saved_inv_std may be NULL in fwd; in bwd NULL means "recompute from x" (batch_norm.rs:355-368). */
int zb_bn2d_fwd_train(zb_ctx* ctx, int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w,
                      double momentum, const void* x, const void* scale, const void* bias, void* running_mean,
                      void* running_var, void* saved_mean, void* saved_inv_std, void* y, const void* residual,
                      int relu);
/* Same as zb_bn2d_fwd_train with the statistics pass replaced by the partial sums zb_conv2d_fprop_bnstats produced
 * (x is that conv's output; NHWC f32): x is read once instead of twice. */
int zb_bn2d_fwd_train_prestats(zb_ctx* ctx, int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w,
                               double momentum, const void* x, const void* scale, const void* bias, void* running_mean,
                               void* running_var, void* saved_mean, void* saved_inv_std, void* y, const void* residual,
                               int relu, const void* stat_partial, int64_t stat_rows, const void* shift);
/* Superset of the two forwards above (NHWC f32): stat_partial may be NULL (statistics computed here), relu_mask may be NULL.
 * relu_mask (relu != 0, c % 32 == 0): zb_bn2d_relu_mask_words(n,c,h,w) 32-bit words, bit i = (output element i > 0) in NHWC
 * order.  zb_bn2d_bwd_mask then takes the ReLU mask from it instead of re-reading the forward output: the fused
 * BN+add+ReLU backward moves 6 tensor passes instead of 7 when dres receives the masked gradient, and 5 when dres is NULL: the
 * residual branch then consumes (dy, relu_mask) directly (zb_conv2d_dgrad_acc_masked, or this same call as the shortcut BatchNorm's
 * backward -- "the gradient is dy masked by these bits" is all zb_bn2d_bwd_mask means). */
int64_t zb_bn2d_relu_mask_words(int64_t n, int64_t c, int64_t h, int64_t w);
int zb_bn2d_fwd_train_fused(zb_ctx* ctx, int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w,
                            double momentum, const void* x, const void* scale, const void* bias, void* running_mean,
                            void* running_var, void* saved_mean, void* saved_inv_std, void* y, const void* residual,
                            int relu, const void* stat_partial, int64_t stat_rows, const void* shift, void* relu_mask);
int zb_bn2d_bwd_mask(zb_ctx* ctx, int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w, const void* x,
                     const void* dy, const void* scale, const void* saved_mean, const void* saved_inv_std, void* dx,
                     void* dscale, void* dbias, const void* relu_mask, void* dres);
int zb_bn2d_fwd_infer(zb_ctx* ctx, int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w, const void* x,
                      const void* scale, const void* bias, const void* mean, const void* var, void* y);
/* dy is the gradient w.r.t. the (possibly fused) output.  When the forward fused relu, pass the forward
 * OUTPUT in y (mask = y > 0); dres (optional) receives the gradient of the residual input (= masked dy). */
int zb_bn2d_bwd(zb_ctx* ctx, int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w, const void* x,
                const void* dy, const void* scale, const void* saved_mean, const void* saved_inv_std, void* dx,
                void* dscale, void* dbias, const void* y, void* dres);
/* Backward of y = max(bn(x), 0) (fused BN+ReLU, no residual) that never reads y: the mask y > 0 is recomputed from x,
 * scale, bias and the saved statistics with the forward's exact rounding sequence (20 instead of 28 bytes / element). */
int zb_bn2d_relu_bwd(zb_ctx* ctx, int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w, const void* x,
                     const void* dy, const void* scale, const void* bias, const void* saved_mean,
                     const void* saved_inv_std, void* dx, void* dscale, void* dbias);

/* Network stem: BatchNorm2d(train) + ReLU + max_pool2d(3x3, stride 2, pad 1) as one pass each way, NHWC f32.  Results of the
 * forward (pooled values, winner codes in the zb_maxpool2d_fwd_idx format, running / saved statistics) are those of
 * zb_bn2d_fwd_train_fused(relu = 1) followed by zb_maxpool2d_fwd_idx, without writing the BatchNorm output; the backward equals
 * zb_maxpool2d_bwd_idx followed by zb_bn2d_relu_bwd without materialising the un-pooled gradient.  Replaces the node sequence
 * batch_norm_2d -> relu -> max_pool_2d of zenu-autograd (src/nn/batch_norm.rs:107-119, src/activation/relu.rs,
 * src/nn/pool2d.rs:149-201).  ZB_ERR_UNSUPPORTED for other windows / dtypes / layouts or when C / 4 is not a power of two <= 256:
 * the caller composes the separate entry points then.  stat_partial / stat_rows / shift: as zb_bn2d_fwd_train_fused (NULL / 0 /
 * NULL: statistics are reduced here). */
int zb_bn2d_relu_maxpool_fwd_train(zb_ctx* ctx, int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w, int64_t k,
                                   int64_t stride, int64_t pad, double momentum, const void* x, const void* scale, const void* bias,
                                   void* running_mean, void* running_var, void* saved_mean, void* saved_inv_std, void* y_pool,
                                   void* pool_idx, const void* stat_partial, int64_t stat_rows, const void* shift);
int zb_bn2d_relu_maxpool_bwd(zb_ctx* ctx, int dtype, int layout, int64_t n, int64_t c, int64_t h, int64_t w, int64_t k, int64_t stride,
                             int64_t pad, const void* x, const void* dy_pool, const void* pool_idx, const void* scale, const void* bias,
                             const void* saved_mean, const void* saved_inv_std, void* dx, void* dscale, void* dbias);

/* ---- GEMM / Linear ------------------------------------------------------------------------------
 * Replaces Gemm::gemm_unchecked (zenu-matrix/src/operation/mul.rs:12-29,113-147 -> cublas{S,D}gemm_v2_64,
 * zenu-cuda/src/cublas/mod.rs:84-160).  Row-major: C[m,n] = alpha*op(A)*op(B) + beta*C. */
int zb_gemm(zb_ctx* ctx, int dtype, int math, int trans_a, int trans_b, int64_t m, int64_t n, int64_t k,
            double alpha, const void* a, int64_t lda, const void* b, int64_t ldb, double beta, void* c, int64_t ldc);
/* Linear layer (zenu-layer/src/layers/linear.rs:22-31): y[b,out] = x[b,in] * W^T + bias, W is [out,in].
 * No materialised transposes (the reference makes four, zenu-autograd/src/functions/transpose.rs:24-32). */
int zb_linear_fwd(zb_ctx* ctx, int dtype, int math, const void* x, const void* w, const void* bias, void* y,
                  int64_t batch, int64_t in_f, int64_t out_f);
int zb_linear_bwd(zb_ctx* ctx, int dtype, int math, const void* x, const void* w, const void* dy, void* dx, void* dw,
                  void* dbias, int64_t batch, int64_t in_f, int64_t out_f);

/* ---- elementwise / reductions / copies ------------------------------------------------------------
 * relu, relu_backward_mask: ReluOps (zenu-matrix/src/operation/relu.rs:12-29; kernels activations.cu:3-41).
 * relu_bwd fuses mask*dy (reference: mask kernel + mul kernel). */
int zb_relu(zb_ctx* ctx, int dtype, const void* x, void* y, double alpha, int64_t n);
int zb_relu_backward_mask(zb_ctx* ctx, int dtype, const void* x, void* mask, double alpha, int64_t n);
int zb_relu_bwd(zb_ctx* ctx, int dtype, const void* x, const void* dy, void* dx, double alpha, int64_t n);
/* c = a op b (same shape), c = a op scalar: AddOps..DivOps (operation/basic_operations.rs:32-276). c may alias a. */
int zb_binary(zb_ctx* ctx, int dtype, int op, const void* a, const void* b, void* c, int64_t n);
int zb_binary_scalar(zb_ctx* ctx, int dtype, int op, const void* a, double scalar, void* c, int64_t n);
/* c[r, j] = a[r, j] op b[j]: numpy-style trailing broadcast of a [cols] rhs (with_clousers.rs:108-240);
 * one launch instead of one kernel per row. */
int zb_binary_bcast_rows(zb_ctx* ctx, int dtype, int op, const void* a, const void* b, void* c, int64_t rows,
                         int64_t cols);
/* out[j] = sum_r a[r, j]  (Matrix::sum(axis 0), operation/sum.rs:9-31: bias / gamma / beta gradients) */
int zb_sum_rows(zb_ctx* ctx, int dtype, const void* a, void* out, int64_t rows, int64_t cols);
/* Axis reductions of a contiguous tensor viewed as [outer][len][inner] -> [outer][inner]: Matrix::sum(axis) (operation/sum.rs:9-31, a
 * loop of `len` add_assign launches there), Matrix::mean(axis) (operation/mean.rs:8-20: sum / len) and Matrix::variance(axis)
 * (operation/var.rs:18-26: biased, the mean of squared differences from the mean; mean_out, optional, also receives the mean).
 * keep_dim is a view concern of the caller: the output has outer * inner elements either way. */
int zb_sum_axis(zb_ctx* ctx, int dtype, const void* a, void* out, int64_t outer, int64_t len, int64_t inner);
int zb_mean_axis(zb_ctx* ctx, int dtype, const void* a, void* out, int64_t outer, int64_t len, int64_t inner);
int zb_variance_axis(zb_ctx* ctx, int dtype, const void* a, void* out, void* mean_out, int64_t outer, int64_t len, int64_t inner);
/* sum_to (operation/sum.rs:35-92; autograd node zenu-autograd/src/functions/sum_to.rs): reduce src to a shape it broadcasts from --
 * shapes are right-aligned, missing leading axes and axes whose target extent is 1 are summed.  Host shape arrays, <= 8 axes. */
int zb_sum_to(zb_ctx* ctx, int dtype, const void* src, const int64_t* src_shape, int src_ndim, void* dst, const int64_t* dst_shape,
              int dst_ndim);
int zb_fill(zb_ctx* ctx, int dtype, void* x, double value, int64_t n); /* zeros(): reference scales by 0 (NaN-unsafe) */
int zb_copy(zb_ctx* ctx, int dtype, const void* src, void* dst, int64_t n);
/* Strided element copy, dst[sum_k i_k * dst_strides[k]] = src[sum_k i_k * src_strides[k]] over a <= 8-axis index space (element
 * strides, host arrays): what copy_from / to_default_stride / transpose-materialise reach through CopyBlas::copy_raw
 * (operation/copy_from.rs:9-55: one cublas{S,D}copy per contiguous run).  Dense-to-dense calls take the vectorised zb_copy. */
int zb_copy_strided(zb_ctx* ctx, int dtype, const void* src, void* dst, int ndim, const int64_t* shape, const int64_t* src_strides,
                    const int64_t* dst_strides);
/* layout transforms (replace transpose_by_index_new_matrix + to_default_stride copies) */
int zb_nchw_to_nhwc(zb_ctx* ctx, int dtype, const void* src, void* dst, int64_t n, int64_t c, int64_t h, int64_t w);
int zb_nhwc_to_nchw(zb_ctx* ctx, int dtype, const void* src, void* dst, int64_t n, int64_t c, int64_t h, int64_t w);

/* ---- "next" rows: pooling and the loss head ----------------------------------
 * max-pool: zenu-matrix/src/nn/pool2d.rs:52-190 (CPU semantics: zero padding takes part in the max,
 * first maximum wins).  Global average pool is absent in the reference.
 * softmax_xent: zenu-autograd/src/loss/cross_entropy.rs:12-23 fused with its backward. */
int zb_maxpool2d_fwd(zb_ctx* ctx, int dtype, int layout, const void* x, void* y, int64_t n, int64_t c, int64_t h,
                     int64_t w, int64_t kh, int64_t kw, int64_t sh, int64_t sw, int64_t ph, int64_t pw);
int zb_maxpool2d_bwd(zb_ctx* ctx, int dtype, int layout, const void* x, const void* dy, void* dx, int64_t n,
                     int64_t c, int64_t h, int64_t w, int64_t kh, int64_t kw, int64_t sh, int64_t sw, int64_t ph,
                     int64_t pw);
/* Indexed variant (NHWC, C % 4 == 0): forward also writes idx[N,P,Q,C] (uint8: winning tap r*kw+s, 255 = padding won);
 * backward gathers from (dy, idx) instead of re-reading x: no memset, no atomics, deterministic. */
int zb_maxpool2d_fwd_idx(zb_ctx* ctx, int dtype, int layout, const void* x, void* y, void* idx, int64_t n, int64_t c,
                         int64_t h, int64_t w, int64_t kh, int64_t kw, int64_t sh, int64_t sw, int64_t ph, int64_t pw);
int zb_maxpool2d_bwd_idx(zb_ctx* ctx, int dtype, int layout, const void* dy, const void* idx, void* dx, int64_t n,
                         int64_t c, int64_t h, int64_t w, int64_t kh, int64_t kw, int64_t sh, int64_t sw, int64_t ph,
                         int64_t pw);
int zb_gap_fwd(zb_ctx* ctx, int dtype, int layout, const void* x, void* y, int64_t n, int64_t c, int64_t hw);
int zb_gap_bwd(zb_ctx* ctx, int dtype, int layout, const void* dy, void* dx, int64_t n, int64_t c, int64_t hw);
/* loss (device scalar) = -(1/B) sum t*log(softmax(z)); dz (optional) = (softmax(z)*sum_j t - t)/B */
int zb_softmax_xent(zb_ctx* ctx, int dtype, const void* z, const void* t, void* loss, void* dz, int64_t batch,
                    int64_t classes);

/* ---- optimizers -----------------------------------------------------------------------------------
 * Replace Optimizer::update (zenu-optimizer/src/sgd.rs:20-30, adam.rs:19-58, adamw.rs:20-69): one fused
 * kernel over a flat parameter bucket instead of >=2 / >=10 kernels + temporaries per tensor.
 * grad_scale multiplies the gradient first (1/world after a data-parallel sum-allreduce). */
int zb_sgd_step(zb_ctx* ctx, int dtype, void* param, const void* grad, double lr, double grad_scale, int64_t n);
/* step_t is the 1-based step AFTER the increment; weight_decay is applied (AdamW, decoupled) when decay != 0 */
int zb_adam_step(zb_ctx* ctx, int dtype, void* param, const void* grad, void* m, void* v, double lr, double beta1,
                 double beta2, double eps, double weight_decay, int decay, int64_t step_t, double grad_scale,
                 int64_t n);

/* The same update with the bias corrections 1/(1-beta1^t), 1/(1-beta2^t) read from a device table instead of being kernel
 * arguments, so that a captured step (zb_model_set_graph) stays valid as t advances: entry i of `table` ([count][2], dtype) holds
 * the corrections of step first_step + i (computed on the host with the reference's arithmetic, adam.rs:23-25) and the kernel
 * uses entry *step_index (device int32); zb_adam_advance adds 1 to it after the step's last launch. */
int zb_adam_table_fill(zb_ctx* ctx, int dtype, double beta1, double beta2, int64_t first_step, int64_t count, void* table);
int zb_adam_step_table(zb_ctx* ctx, int dtype, void* param, const void* grad, void* m, void* v, double lr, double beta1,
                       double beta2, double eps, double weight_decay, int decay, const void* table,
                       const int32_t* step_index, double grad_scale, int64_t n);
int zb_adam_advance(zb_ctx* ctx, int32_t* step_index);

/* ---- data parallel (new; the reference has no multi-GPU path) -------------------------
 * One process per GPU.  Rank 0 calls zb_dp_unique_id (128 bytes, host), ships it to the others through
 * whatever rendezvous the host has (torch.distributed / a file), then every rank calls zb_dp_init.
 * zb_dp_allreduce_sum enqueues ncclAllReduce(sum) on the ctx comm stream, ordered after everything already
 * enqueued on the compute stream; zb_dp_wait makes the compute stream wait for it. */
/* Bucket plan of the flat parameter / gradient buffers (pure host logic, callable without a GPU): parameters in forward
 * order with kind 0 weight / 1 bias / 2 buffer; buckets of ~bucket_bytes are filled from the LAST parameter backwards
 * (bucket 0 is complete first during backward); offsets are in elements, 16-byte aligned, weights before biases inside a
 * bucket; buffers (kind 2) get offsets in their own area and bucket -1. */
int zb_dp_plan_buckets(const int64_t* numel, const int* kind, int n, int64_t bucket_bytes, int elem_size, int* bucket_out,
                       int64_t* offset_out, int* num_buckets, int64_t* total_elems, int64_t* buffer_elems);
int zb_dp_unique_id(zb_ctx* ctx, void* host_id128);
int zb_dp_init(zb_ctx* ctx, const void* host_id128, int rank, int world);
int zb_dp_allreduce_sum(zb_ctx* ctx, int dtype, void* buf, int64_t n);
int zb_dp_wait(zb_ctx* ctx);
/* Polls ncclCommGetAsyncError (also done by zb_dp_wait once per step and by zb_ctx_check): a failure that surfaced asynchronously --
 * a peer died, a link went down -- aborts the communicator, so blocked collectives are released, and returns ZB_ERR_NCCL from then on. */
int zb_dp_check(zb_ctx* ctx);
int zb_dp_rank(zb_ctx* ctx);
int zb_dp_world(zb_ctx* ctx);

/* ---- input pipeline -------------------------------------------------------------------------------------------------
 * The batch crosses PCIe as uint8 (NHWC as decoded, or NCHW) and is expanded on the device into the model's NCHW input:
 * dst[n][c][h][w] = (src / 255 - mean[c]) / std[c] (host_mean / host_std: c host doubles, NULL = 0 / 1); int32 labels become
 * the one-hot rows the reference's cross_entropy expects.  Replaces the per-sample Vec<Variable> + CPU concat + synchronous
 * f32 cudaMemcpy of zenu/src/dataset.rs:74-100. */
int zb_input_u8_to_float(zb_ctx* ctx, int dtype, int src_layout, const void* src_u8, void* dst_nchw, int64_t n, int64_t c,
                         int64_t h, int64_t w, const double* host_mean, const double* host_std);
int zb_onehot(zb_ctx* ctx, int dtype, const void* labels_i32, void* out, int64_t n, int64_t classes);

/* Library-owned staging around those two kernels: `slots` (2..8) sets of pinned host buffers (uint8 images in src_layout + int32
 * labels) and their device twins, a copy stream, and per-slot events.  The decoder fills slot s's pinned buffers
 * (zb_input_stage_host_buffers; zb_input_stage_host_sync first when the slot is being reused), zb_input_stage_submit enqueues the
 * host->device copy of the BYTES on the copy stream (it overlaps the training step of the previous batch), zb_input_stage_wait makes
 * the ctx compute stream wait for that copy, expands the batch on the device and returns the model's inputs: x [n][c][h][w] and
 * one-hot targets [n][classes] in `dtype` (device pointers, stable per slot).  Nothing synchronises the host except host_sync.
 * Replaces zenu/src/dataset.rs:74-100 (per-sample Variables + CPU concat) + the synchronous f32 copy of Matrix::to::<Nvidia>()
 * (zenu-matrix/src/matrix.rs:139-160,486). */
typedef struct zb_input_stage zb_input_stage;
int zb_input_stage_create(zb_ctx* ctx, int dtype, int src_layout, int64_t n, int64_t c, int64_t h, int64_t w, int64_t classes,
                          const double* host_mean, const double* host_std, int slots, zb_input_stage** out);
int zb_input_stage_destroy(zb_input_stage* st);
int64_t zb_input_stage_h2d_bytes(const zb_input_stage* st); /* bytes one submit moves across PCIe */
int zb_input_stage_host_buffers(zb_input_stage* st, int slot, void** images_u8, void** labels_i32);
int zb_input_stage_host_sync(zb_input_stage* st, int slot);
int zb_input_stage_submit(zb_input_stage* st, int slot);
int zb_input_stage_wait(zb_input_stage* st, int slot, void** x_nchw, void** targets_onehot);

/* ---- host model API (layers / tape / optimizer above the op ABI) -----------------------------------
 * C face of the C++ host side (zenu_b200/csrc/host): Module::call + Variable::backward + Optimizer::update
 * (reference: zenu-layer/src/lib.rs:21-51, zenu-autograd/src/lib.rs:413-420, zenu-optimizer/src/lib.rs:8-10)
 * for the architectures the configs name: "small_cnn" (zenu/examples/cifar10.rs:29-69), "resnet18", "resnet50"
 * (torchvision v1.5 topology composed from Conv2d / BatchNorm2d / Linear).
 * Inputs are NCHW (reference contract); parameters are exposed as device pointers with reference names
 * ("conv1.conv2d.filter", "bn1.batch_norm_2d.scale", ...).  Conv filters are stored KRSC ([K,R,S,C]).
 * fused != 0 uses the fused BN+ReLU(+residual) nodes; fused == 0 the reference's separate nodes. */
typedef struct zb_model zb_model;
typedef enum { ZB_OPT_SGD = 0, ZB_OPT_ADAM = 1, ZB_OPT_ADAMW = 2 } zb_optimizer_kind;
typedef enum { ZB_PARAM_WEIGHT = 0, ZB_PARAM_BIAS = 1, ZB_PARAM_BUFFER = 2 } zb_param_kind;

int zb_model_create(zb_ctx* ctx, const char* arch, int dtype, int num_classes, int fused, uint64_t seed,
                    int64_t bucket_bytes, zb_model** out);
int zb_model_destroy(zb_model* m);
int zb_model_param_count(zb_model* m);
/* shape gets up to 4 extents; data / grad are device pointers (grad is NULL for buffers) */
int zb_model_param_info(zb_model* m, int index, char* name, int name_cap, int64_t* shape, int* ndim, int* kind,
                        void** data, void** grad);
int zb_model_set_train(zb_model* m, int train);
int zb_model_set_optimizer(zb_model* m, int kind, double lr, double beta1, double beta2, double eps,
                           double weight_decay);
/* logits_out: device [batch, classes] */
int zb_model_forward(zb_model* m, const void* x_nchw, int64_t batch, int64_t c, int64_t h, int64_t w,
                     void* logits_out);
/* forward + loss + backward; gradients are left in the flat gradient buffer; bucket allreduces are enqueued on
 * the comm stream as each bucket completes when the ctx is data parallel.  loss_dev: device scalar. */
int zb_model_forward_backward(zb_model* m, const void* x_nchw, const void* targets_onehot, int64_t batch, int64_t c,
                              int64_t h, int64_t w, void* loss_dev);
int zb_model_update(zb_model* m); /* Optimizer::update */
/* forward_backward + update; if host_loss != NULL the loss is copied back (synchronises the stream) */
int zb_model_train_step(zb_model* m, const void* x_nchw, const void* targets_onehot, int64_t batch, int64_t c,
                        int64_t h, int64_t w, void* loss_dev, double* host_loss);
/* zb_model_train_step without waiting for it: the step is enqueued and its loss is copied, behind it on the compute stream, into a
 * pinned two-slot ring.  zb_model_loss_wait(m, age, &v) blocks until the step enqueued `age` calls ago (0 = the latest, 1 = the one
 * before it) has completed and returns its loss.  A loop that logs every step's loss waits with age 1 right after enqueueing, so the
 * device never idles while the host stages the next batch (the reference's loop reads the loss synchronously each step:
 * zenu/examples/mnist.rs:138 `loss.get_data().asum()`). */
int zb_model_train_step_async(zb_model* m, const void* x_nchw, const void* targets_onehot, int64_t batch, int64_t c,
                              int64_t h, int64_t w, void* loss_dev);
int zb_model_loss_wait(zb_model* m, int age, double* host_loss);
/* CUDA-graph replay of zb_model_train_step (the reference rebuilds and walks its Rc<RefCell> tape every step,
 * zenu-autograd/src/lib.rs:220-237): after two eager steps the step is captured from the compute stream once per distinct
 * (buffers, shape) signature and replayed, NCCL bucket allreduces included.  Needs a ctx whose compute stream can be captured
 * (not the legacy default stream); otherwise, and while per-node profiling is on, steps keep running eagerly. */
int zb_model_set_graph(zb_model* m, int enable);
/* Run every conv wgrad of the backward sweep on the ctx's side stream (zb_ctx_side), overlapping the BatchNorm-backward / dgrad
 * chain of the following layers; results are bit-identical.  Off by default: measured neutral-to-slower on B200. */
int zb_model_set_wgrad_overlap(zb_model* m, int enable);
int zb_model_graph_count(zb_model* m); /* step graphs captured so far (0: every step so far ran eagerly) */
/* Per-node timing (CUDA events on the compute stream around every tape node, forward and backward).  dump writes one
 * line per distinct node key: "key\tcount\ttotal_ms\talgorithmic_flops\talgorithmic_bytes\n" (sums over count) and returns
 * the buffer size needed (call with buf == NULL to size it). */
int zb_model_profile_enable(zb_model* m, int enable);
int64_t zb_model_profile_dump(zb_model* m, char* buf, int64_t cap);
/* bytes currently held by the model's caching allocator (activations + parameters) */
int64_t zb_model_bytes_reserved(zb_model* m);

/* ---- model files (reference: zenu::save_model / load_model, zenu/src/lib.rs:26-67) -------------------------------
 * bincode 1.3.3 image of HashMap<String, Variable>: per entry {key, shape, stride, data, data_type "f32"|"f64", ptr_offset}
 * (zenu-matrix/src/impl_serde.rs:11-40) in the reference's layouts (filters [K,C,R,S], conv bias [1,K,1,1]).
 * zb_model_load mirrors load_model: a key the model does not have is an error, parameters the file does not name are
 * left untouched.  The zb_ckpt_* functions are the host-only reader / writer underneath (no GPU needed): all pointers
 * are HOST pointers; data returned by zb_ckpt_entry is dense row-major and owned by the zb_ckpt. */
typedef struct zb_ckpt zb_ckpt;
int zb_model_save(zb_model* m, const char* path);
int zb_model_load(zb_model* m, const char* path);
/* Training-state files (DP resume): the same container with, next to the parameters, the optimizer state the reference
 * keeps in memory only (zenu-optimizer/src/adam.rs:9-17,62-89: step, m and v as HashMap<String, Variable> keyed by parameter name):
 * "optimizer.step" (scalar: updates done so far) and, for Adam / AdamW, "optimizer.m.<param>" / "optimizer.v.<param>" in the
 * parameter's own reference layout.  zb_model_load_state restores all of it (set the optimizer kind and its hyper-parameters with
 * zb_model_set_optimizer first; they are configuration, not state): training resumed from a state file continues bit-identically
 * to the uninterrupted run.  A plain model file is also accepted (parameters only). */
int zb_model_save_state(zb_model* m, const char* path);
int zb_model_load_state(zb_model* m, const char* path);
int zb_ckpt_write(const char* path, int dtype, int n, const char* const* names, const int* ndims,
                  const int64_t* const* shapes, const void* const* host_data);
int zb_ckpt_open(const char* path, zb_ckpt** out);
int zb_ckpt_count(const zb_ckpt* ck);
int zb_ckpt_entry(const zb_ckpt* ck, int index, char* name, int name_cap, int64_t* shape, int* ndim, int* dtype,
                  const void** host_data, int64_t* numel);
int zb_ckpt_close(zb_ckpt* ck);

#ifdef __cplusplus
}
#endif
#endif /* ZENU_B200_H */
