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

extern "C" {

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
  ConvShim* s = reinterpret_cast<ConvShim*>(desc);
  zb_ctx* ctx = compat_ctx();
  if (!s || !b || !ctx) return INVALID_VALUE;
  return status_of(zb_conv2d_wgrad(ctx, s->dtype, ZB_NCHW, ZB_MATH_DEFAULT, &s->d, b->DY, b->X, b->DW));
}
void destroy_conv_backward_filter_descriptor(ConvBkwdFilterDescriptor* desc) { delete reinterpret_cast<ConvShim*>(desc); }

}  // extern "C"
