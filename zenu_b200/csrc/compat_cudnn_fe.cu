// compat_cudnn_fe.cu — conv descriptor shim with the reference's 4-call shape
// (cudnn_frontend_wrapper.h:100-186; implementation it replaces: src/conv.cpp:29-187, src/i_graph_desc.h:18-66).
// There is no graph to build: "check" validates the geometry and "execute" calls the native conv entry points
// on NCHW / KCRS tensors.  Runs on the legacy default stream like the reference's cuDNN handle did.
#include "../../include/zenu_cudnn_frontend_compat.h"
#include "common.cuh"

namespace zb {
zb_ctx* compat_ctx();

struct ConvShim {
  int dtype;
  zb_conv2d_desc d;
  bool contiguous;
};

static bool default_strides(const CudnnTensorShapeStride* s) {
  int64_t expect = 1;
  for (int i = static_cast<int>(s->num_dims) - 1; i >= 0; --i) {
    if (s->dims[i] != 1 && s->strides[i] != expect) return false;
    expect *= s->dims[i];
  }
  return true;
}

static CudnnFrontendError_t make_shim(ConvShim** out, CudnnFrontendDataType_t dt, const CudnnTensorShapeStride* x,
                                      const CudnnTensorShapeStride* w, const CudnnTensorShapeStride* y, const ConvInfo* info) {
  if (!out || !x || !w || !y || !info) return INVALID_VALUE;
  if (dt != DATA_TYPE_FLOAT && dt != DATA_TYPE_DOUBLE) return NOT_SUPPORTED;  // reference: FLOAT or DOUBLE only (conv.cpp:64-65)
  if (x->num_dims != 4 || w->num_dims != 4 || y->num_dims != 4 || info->num_dims != 2) return NOT_SUPPORTED;
  ConvShim* s = new ConvShim();
  s->dtype = (dt == DATA_TYPE_FLOAT) ? ZB_F32 : ZB_F64;
  s->d.n = x->dims[0]; s->d.c = x->dims[1]; s->d.h = x->dims[2]; s->d.w = x->dims[3];
  s->d.k = w->dims[0]; s->d.kh = w->dims[2]; s->d.kw = w->dims[3];
  s->d.pad_h = info->padding[0]; s->d.pad_w = info->padding[1];
  s->d.stride_h = info->stride[0]; s->d.stride_w = info->stride[1];
  s->d.dil_h = info->dilation[0]; s->d.dil_w = info->dilation[1];
  s->contiguous = default_strides(x) && default_strides(w) && default_strides(y);
  const bool shapes_ok = w->dims[1] == x->dims[1] && y->dims[0] == x->dims[0] && y->dims[1] == w->dims[0] &&
                         y->dims[2] == zb_conv_out_size(s->d.h, s->d.kh, s->d.pad_h, s->d.stride_h, s->d.dil_h) &&
                         y->dims[3] == zb_conv_out_size(s->d.w, s->d.kw, s->d.pad_w, s->d.stride_w, s->d.dil_w);
  if (!shapes_ok) { delete s; return INVALID_VALUE; }
  *out = s;
  return SUCCESS;
}
static CudnnFrontendError_t status_of(int rc) {
  if (rc == ZB_OK) return SUCCESS;
  printf("zenu_b200: %s\n", zb_last_error());  // the reference prints the cuDNN message to stdout (i_graph_desc.h:60-64)
  return rc == ZB_ERR_INVALID ? INVALID_VALUE : rc == ZB_ERR_UNSUPPORTED ? NOT_SUPPORTED : FAILURE;
}
}  // namespace zb

using namespace zb;

namespace zb {
// BatchNorm descriptor shim (reference: src/batchnorm.cpp:166-281).  X is a 4-D tensor [N, C, H, W] whose strides say where the
// channel dimension lives: default strides = NCHW, channel stride 1 = NHWC (both contiguous); anything else is NOT_SUPPORTED.
struct BnShim {
  int dtype;
  int layout;      // ZB_NCHW / ZB_NHWC, -1 = strides this library does not serve
  int64_t n, c, h, w;
  double eps, momentum;
  bool training, backward;
};
static int bn_layout(const CudnnTensorShapeStride* s) {
  if (default_strides(s)) return ZB_NCHW;
  const int64_t n = s->dims[0], c = s->dims[1], h = s->dims[2], w = s->dims[3];
  const bool nhwc = (c == 1 || s->strides[1] == 1) && (w == 1 || s->strides[3] == c) && (h == 1 || s->strides[2] == w * c) &&
                    (n == 1 || s->strides[0] == h * w * c);
  return nhwc ? ZB_NHWC : -1;
}
static CudnnFrontendError_t make_bn_shim(BnShim** out, CudnnFrontendDataType_t dt, const CudnnTensorShapeStride* shape, double eps,
                                         double momentum, bool training, bool backward) {
  if (!out || !shape) return INVALID_VALUE;
  if (dt != DATA_TYPE_FLOAT && dt != DATA_TYPE_DOUBLE) return NOT_SUPPORTED;
  if (shape->num_dims != 4) return NOT_SUPPORTED;   // reference: "only supports BN1D or BN2D" as 4-D shapes (batchnorm.cpp:19-29)
  for (int i = 0; i < 4; ++i)
    if (shape->dims[i] <= 0) return INVALID_VALUE;
  BnShim* s = new BnShim();
  s->dtype = dt == DATA_TYPE_FLOAT ? ZB_F32 : ZB_F64;
  s->layout = bn_layout(shape);
  s->n = shape->dims[0]; s->c = shape->dims[1]; s->h = shape->dims[2]; s->w = shape->dims[3];
  s->eps = eps; s->momentum = momentum; s->training = training; s->backward = backward;
  *out = s;
  return SUCCESS;
}
// runs `call` with the ctx epsilon set to the descriptor's
template <typename F>
static CudnnFrontendError_t with_eps(zb_ctx* ctx, double eps, F&& call) {
  const double keep = zb_ctx_bn_epsilon(ctx);
  zb_ctx_set_bn_epsilon(ctx, eps);
  const int rc = call();
  zb_ctx_set_bn_epsilon(ctx, keep);
  return status_of(rc);
}
}  // namespace zb

extern "C" {

CudnnFrontendError_t create_batch_norm_descriptor(BatchNormDescriptor** desc, CudnnFrontendDataType_t dt, const CudnnTensorShapeStride* shape,
                                                  float epsilon, float momentum, bool is_training) {
  return make_bn_shim(reinterpret_cast<BnShim**>(desc), dt, shape, epsilon, momentum, is_training, false);
}
void batch_norm_desc_debug(BatchNormDescriptor* desc) {
  const BnShim* s = reinterpret_cast<const BnShim*>(desc);
  if (!s) return;
  printf("BatchNorm %s X [%lld, %lld, %lld, %lld] %s %s epsilon %g momentum %g running stats %d\n", s->backward ? "backward" : "forward",
         static_cast<long long>(s->n), static_cast<long long>(s->c), static_cast<long long>(s->h), static_cast<long long>(s->w),
         s->layout == ZB_NCHW ? "NCHW" : s->layout == ZB_NHWC ? "NHWC" : "unsupported strides", s->dtype == ZB_F32 ? "float" : "double",
         s->eps, s->momentum, s->training ? 1 : 0);
}
CudnnFrontendError_t check_graph(BatchNormDescriptor* desc, void*) {
  return (desc && reinterpret_cast<BnShim*>(desc)->layout >= 0) ? SUCCESS : NOT_SUPPORTED;
}
CudnnFrontendError_t get_workspace_size(BatchNormDescriptor* desc, int64_t* ws) {
  if (!desc || !ws) return INVALID_VALUE;
  *ws = 0;
  return SUCCESS;
}
CudnnFrontendError_t execute_batch_norm_forward_training(BatchNormDescriptor* desc, BatchNormExecutionBuffers* b, void*, void*) {
  ZB_API_RANGE();
  BnShim* s = reinterpret_cast<BnShim*>(desc);
  zb_ctx* ctx = compat_ctx();
  if (!s || !b || !ctx || s->backward) return INVALID_VALUE;
  if (s->layout < 0) return NOT_SUPPORTED;
  void* rm = nullptr;
  void* rv = nullptr;
  if (s->training) {   // has_running_stats (batchnorm.cpp:146-149,196-201): next = (1 - momentum) * prev + momentum * batch
    if (!b->prev_running_mean || !b->prev_running_var || !b->next_running_mean || !b->next_running_var) return INVALID_VALUE;
    const size_t bytes = static_cast<size_t>(s->c) * (s->dtype == ZB_F32 ? 4 : 8);
    if (b->next_running_mean != b->prev_running_mean &&
        cudaMemcpyAsync(b->next_running_mean, b->prev_running_mean, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(zb_ctx_stream(ctx))) != cudaSuccess)
      return FAILURE;
    if (b->next_running_var != b->prev_running_var &&
        cudaMemcpyAsync(b->next_running_var, b->prev_running_var, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(zb_ctx_stream(ctx))) != cudaSuccess)
      return FAILURE;
    rm = b->next_running_mean;
    rv = b->next_running_var;
  }
  return with_eps(ctx, s->eps, [&]() {
    // zb_bn2d_fwd_train weights the OLD running statistic with its momentum argument (the reference CPU convention)
    return zb_bn2d_fwd_train(ctx, s->dtype, s->layout, s->n, s->c, s->h, s->w, 1.0 - s->momentum, b->X, b->scale, b->bias, rm, rv, b->mean,
                             b->inv_variance, b->Y, nullptr, 0);
  });
}
void destroy_batch_norm_descriptor(BatchNormDescriptor* desc) { delete reinterpret_cast<BnShim*>(desc); }

CudnnFrontendError_t create_batch_norm_backward_data_descriptor(BatchNormBkwdDescriptor** desc, CudnnFrontendDataType_t dt,
                                                                const CudnnTensorShapeStride* shape) {
  return make_bn_shim(reinterpret_cast<BnShim**>(desc), dt, shape, 0.0, 0.0, true, true);
}
CudnnFrontendError_t check_backward_data_graph(BatchNormBkwdDescriptor* desc, void*) {
  return (desc && reinterpret_cast<BnShim*>(desc)->layout >= 0) ? SUCCESS : NOT_SUPPORTED;
}
CudnnFrontendError_t get_backward_data_workspace_size(BatchNormBkwdDescriptor* desc, int64_t* ws) {
  if (!desc || !ws) return INVALID_VALUE;
  *ws = 0;
  return SUCCESS;
}
CudnnFrontendError_t execute_batch_norm_backward_data(BatchNormBkwdDescriptor* desc, BatchNormBkwdExecutionBuffers* b, void*, void*) {
  ZB_API_RANGE();
  BnShim* s = reinterpret_cast<BnShim*>(desc);
  zb_ctx* ctx = compat_ctx();
  if (!s || !b || !ctx || !s->backward) return INVALID_VALUE;
  if (s->layout < 0) return NOT_SUPPORTED;
  if (!b->mean || !b->inv_variance) return INVALID_VALUE;   // the graph is built with set_saved_mean_and_inv_variance (batchnorm.cpp:232-234)
  // dscale / dbias come back in the io data type (the reference tags them FLOAT for every io type, batchnorm.cpp:238-239)
  return status_of(zb_bn2d_bwd(ctx, s->dtype, s->layout, s->n, s->c, s->h, s->w, b->X, b->DY, b->scale, b->mean, b->inv_variance, b->DX,
                               b->dscale, b->dbias, nullptr, nullptr));
}
void destroy_batch_norm_backward_data_descriptor(BatchNormBkwdDescriptor* desc) { delete reinterpret_cast<BnShim*>(desc); }

CudnnFrontendError_t create_conv_descriptor(ConvDescriptor** desc, CudnnFrontendDataType_t dt, CudnnTensorShapeStride* x,
                                            CudnnTensorShapeStride* w, CudnnTensorShapeStride* y, ConvInfo* info) {
  return make_shim(reinterpret_cast<ConvShim**>(desc), dt, x, w, y, info);
}
CudnnFrontendError_t check_conv_graph(ConvDescriptor* desc, void*) {
  return (desc && reinterpret_cast<ConvShim*>(desc)->contiguous) ? SUCCESS : NOT_SUPPORTED;
}
CudnnFrontendError_t get_conv_workspace_size(ConvDescriptor* desc, int64_t* ws) {
  if (!desc || !ws) return INVALID_VALUE;
  *ws = 0;
  return SUCCESS;
}
CudnnFrontendError_t execute_conv_forward(ConvDescriptor* desc, ConvBufers* b, void*, void*) {
  ZB_API_RANGE();
  ConvShim* s = reinterpret_cast<ConvShim*>(desc);
  zb_ctx* ctx = compat_ctx();
  if (!s || !b || !ctx) return INVALID_VALUE;
  return status_of(zb_conv2d_fprop(ctx, s->dtype, ZB_NCHW, ZB_MATH_DEFAULT, &s->d, b->X, b->filter, nullptr, b->Y));
}
void destroy_conv_descriptor(ConvDescriptor* desc) { delete reinterpret_cast<ConvShim*>(desc); }

CudnnFrontendError_t create_conv_backward_data_descriptor(ConvBkwdDataDescriptor** desc, CudnnFrontendDataType_t dt,
                                                          CudnnTensorShapeStride* dy, CudnnTensorShapeStride* w,
                                                          CudnnTensorShapeStride* dx, ConvInfo* info) {
  return make_shim(reinterpret_cast<ConvShim**>(desc), dt, dx, w, dy, info);
}
CudnnFrontendError_t check_conv_backward_data_graph(ConvBkwdDataDescriptor* desc, void*) {
  return (desc && reinterpret_cast<ConvShim*>(desc)->contiguous) ? SUCCESS : NOT_SUPPORTED;
}
CudnnFrontendError_t get_conv_backward_data_workspace_size(ConvBkwdDataDescriptor* desc, int64_t* ws) {
  if (!desc || !ws) return INVALID_VALUE;
  *ws = 0;
  return SUCCESS;
}
CudnnFrontendError_t execute_conv_backward_data(ConvBkwdDataDescriptor* desc, ConvBkwdDataBuffers* b, void*, void*) {
  ZB_API_RANGE();
  ConvShim* s = reinterpret_cast<ConvShim*>(desc);
  zb_ctx* ctx = compat_ctx();
  if (!s || !b || !ctx) return INVALID_VALUE;
  return status_of(zb_conv2d_dgrad(ctx, s->dtype, ZB_NCHW, ZB_MATH_DEFAULT, &s->d, b->DY, b->filter, b->DX));
}
void destroy_conv_backward_data_descriptor(ConvBkwdDataDescriptor* desc) { delete reinterpret_cast<ConvShim*>(desc); }

CudnnFrontendError_t create_conv_backward_filter_descriptor(ConvBkwdFilterDescriptor** desc, CudnnFrontendDataType_t dt,
                                                            CudnnTensorShapeStride* x, CudnnTensorShapeStride* dy,
                                                            CudnnTensorShapeStride* dw, ConvInfo* info) {
  return make_shim(reinterpret_cast<ConvShim**>(desc), dt, x, dw, dy, info);
}
CudnnFrontendError_t check_conv_backward_filter_graph(ConvBkwdFilterDescriptor* desc, void*) {
  return (desc && reinterpret_cast<ConvShim*>(desc)->contiguous) ? SUCCESS : NOT_SUPPORTED;
}
CudnnFrontendError_t get_conv_backward_filter_workspace_size(ConvBkwdFilterDescriptor* desc, int64_t* ws) {
  if (!desc || !ws) return INVALID_VALUE;
  *ws = 0;
  return SUCCESS;
}
CudnnFrontendError_t execute_conv_backward_filter(ConvBkwdFilterDescriptor* desc, ConvBkwdFilterBuffers* b, void*, void*) {
  ZB_API_RANGE();
  ConvShim* s = reinterpret_cast<ConvShim*>(desc);
  zb_ctx* ctx = compat_ctx();
  if (!s || !b || !ctx) return INVALID_VALUE;
  return status_of(zb_conv2d_wgrad(ctx, s->dtype, ZB_NCHW, ZB_MATH_DEFAULT, &s->d, b->DY, b->X, b->DW));
}
void destroy_conv_backward_filter_descriptor(ConvBkwdFilterDescriptor* desc) { delete reinterpret_cast<ConvShim*>(desc); }

}  // extern "C"
