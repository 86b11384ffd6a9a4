/* zenu_cudnn_frontend_compat.h — shim with the 4-call shape of the reference's cuDNN-frontend wrapper for
 * convolution (zenu-cudnn-frontend-wrapper-sys/cudnn_frontend_wrapper/include/cudnn_frontend_wrapper.h:12-31,
 * 100-186): create_* / check_*_graph / get_*_workspace_size / execute_*, same struct layouts, same enum.
 * `cudnnHandle_t*` is carried as an opaque `void*` and ignored: there is no cuDNN behind this library.
 * Workspace size is always 0 (scratch is owned by the library).  Added: destroy_* — the reference never
 * frees its descriptors (cudnn_frontend_wrapper.cpp:13,39,66,91,118 `new` without a matching export).
 * Tensors are NCHW / KCRS with the default (contiguous) strides the reference asserts
 * (zenu-matrix/src/nn/conv/interface.rs:270-281); other strides return NOT_SUPPORTED.
 * The wrapper's BatchNorm graph entry points (cudnn_frontend_wrapper.h:33-98) are mirrored too: the reference marks them
 * experimental and never calls them at run time (zenu-cuda/src/cudnn/graph_batchnorm.rs:1), but graph_batchnorm.rs:5-11 is a compiled
 * `pub mod` that imports all nine, so zenu-cuda does not link without them.  They are served by zb_bn2d_fwd_train / zb_bn2d_bwd with
 * cuDNN-frontend's conventions translated: `momentum` weights the NEW statistic there (next = (1 - m) * prev + m * batch; the
 * reference's legacy path weights the old one), epsilon is the descriptor's, prev_/next_ running statistics are separate buffers,
 * peer_stats_* (multi-GPU BN) are ignored, X may be NCHW- or NHWC-strided.
 */
#ifndef ZENU_CUDNN_FRONTEND_COMPAT_H
#define ZENU_CUDNN_FRONTEND_COMPAT_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum { SUCCESS = 0, FAILURE = 1, INVALID_VALUE = 2, NOT_SUPPORTED = 3 } CudnnFrontendError_t;
typedef enum { DATA_TYPE_HALF = 0, DATA_TYPE_FLOAT = 1, DATA_TYPE_DOUBLE = 2 } CudnnFrontendDataType_t;

typedef struct {
  size_t num_dims;
  int64_t dims[8];
  int64_t strides[8];
} CudnnTensorShapeStride;

/* ---- BatchNorm (cudnn_frontend_wrapper.h:33-98) ---- */
typedef struct BatchNormDescriptor BatchNormDescriptor;
typedef struct {
  void* X;
  void* mean;
  void* inv_variance;
  void* scale;
  void* bias;
  void* peer_stats_0;
  void* peer_stats_1;
  void* prev_running_mean;
  void* prev_running_var;
  void* next_running_mean;
  void* next_running_var;
  void* Y;
} BatchNormExecutionBuffers;
CudnnFrontendError_t create_batch_norm_descriptor(BatchNormDescriptor** desc, CudnnFrontendDataType_t data_type,
                                                  const CudnnTensorShapeStride* shape, float epsilon, float momentum, bool is_training);
void batch_norm_desc_debug(BatchNormDescriptor* desc);
CudnnFrontendError_t check_graph(BatchNormDescriptor* desc, void* handle);
CudnnFrontendError_t get_workspace_size(BatchNormDescriptor* desc, int64_t* workspace_size);
CudnnFrontendError_t execute_batch_norm_forward_training(BatchNormDescriptor* desc, BatchNormExecutionBuffers* buffers, void* workspace,
                                                         void* handle);
void destroy_batch_norm_descriptor(BatchNormDescriptor* desc);

typedef struct BatchNormBkwdDescriptor BatchNormBkwdDescriptor;
typedef struct {
  void* X;
  void* DY;
  void* scale;
  void* mean;
  void* inv_variance;
  void* dscale;
  void* dbias;
  void* DX;
  void* peer_stats_0;
  void* peer_stats_1;
} BatchNormBkwdExecutionBuffers;
CudnnFrontendError_t create_batch_norm_backward_data_descriptor(BatchNormBkwdDescriptor** desc, CudnnFrontendDataType_t data_type,
                                                                const CudnnTensorShapeStride* shape);
CudnnFrontendError_t check_backward_data_graph(BatchNormBkwdDescriptor* desc, void* handle);
CudnnFrontendError_t get_backward_data_workspace_size(BatchNormBkwdDescriptor* desc, int64_t* workspace_size);
CudnnFrontendError_t execute_batch_norm_backward_data(BatchNormBkwdDescriptor* desc, BatchNormBkwdExecutionBuffers* buffers,
                                                      void* workspace, void* handle);
void destroy_batch_norm_backward_data_descriptor(BatchNormBkwdDescriptor* desc);

/* ---- convolution (cudnn_frontend_wrapper.h:100-186) ---- */
typedef struct { void* X; void* filter; void* Y; } ConvBufers;
typedef struct { int64_t padding[2]; int64_t stride[2]; int64_t dilation[2]; int64_t num_dims; } ConvInfo;
typedef struct { void* DY; void* filter; void* DX; } ConvBkwdDataBuffers;
typedef struct { void* X; void* DY; void* DW; } ConvBkwdFilterBuffers;

typedef struct ConvDescriptor ConvDescriptor;
typedef struct ConvBkwdDataDescriptor ConvBkwdDataDescriptor;
typedef struct ConvBkwdFilterDescriptor ConvBkwdFilterDescriptor;

CudnnFrontendError_t create_conv_descriptor(ConvDescriptor** desc, CudnnFrontendDataType_t data_type,
                                            CudnnTensorShapeStride* x_shape, CudnnTensorShapeStride* w_shape,
                                            CudnnTensorShapeStride* y_shape, ConvInfo* info);
CudnnFrontendError_t check_conv_graph(ConvDescriptor* desc, void* handle);
CudnnFrontendError_t get_conv_workspace_size(ConvDescriptor* desc, int64_t* workspace_size);
CudnnFrontendError_t execute_conv_forward(ConvDescriptor* desc, ConvBufers* buffers, void* workspace, void* handle);
void destroy_conv_descriptor(ConvDescriptor* desc);

CudnnFrontendError_t create_conv_backward_data_descriptor(ConvBkwdDataDescriptor** desc, CudnnFrontendDataType_t data_type,
                                                          CudnnTensorShapeStride* dy_shape, CudnnTensorShapeStride* w_shape,
                                                          CudnnTensorShapeStride* dx_shape, ConvInfo* info);
CudnnFrontendError_t check_conv_backward_data_graph(ConvBkwdDataDescriptor* desc, void* handle);
CudnnFrontendError_t get_conv_backward_data_workspace_size(ConvBkwdDataDescriptor* desc, int64_t* workspace_size);
CudnnFrontendError_t execute_conv_backward_data(ConvBkwdDataDescriptor* desc, ConvBkwdDataBuffers* buffers, void* workspace,
                                                void* handle);
void destroy_conv_backward_data_descriptor(ConvBkwdDataDescriptor* desc);

CudnnFrontendError_t create_conv_backward_filter_descriptor(ConvBkwdFilterDescriptor** desc, CudnnFrontendDataType_t data_type,
                                                            CudnnTensorShapeStride* x_shape, CudnnTensorShapeStride* dy_shape,
                                                            CudnnTensorShapeStride* dw_shape, ConvInfo* info);
CudnnFrontendError_t check_conv_backward_filter_graph(ConvBkwdFilterDescriptor* desc, void* handle);
CudnnFrontendError_t get_conv_backward_filter_workspace_size(ConvBkwdFilterDescriptor* desc, int64_t* workspace_size);
CudnnFrontendError_t execute_conv_backward_filter(ConvBkwdFilterDescriptor* desc, ConvBkwdFilterBuffers* buffers,
                                                  void* workspace, void* handle);
void destroy_conv_backward_filter_descriptor(ConvBkwdFilterDescriptor* desc);

#ifdef __cplusplus
}
#endif
#endif
