/* zenu_kernel_compat.h — symbol-compatible shim for zenu-cuda-kernel-sys.
 *
 * Same names and signatures as the reference's kernel.h aggregate
 * (zenu-cuda-kernel-sys/kernel/{activations.h:7-11, array_array.h:7-28, array_scalar.h:7-112,
 * conv2d_bkwd_data.h:7-17, memory_access.h:7-11, element_wise.h:7-9}), so zenu-cuda/src/kernel/{mod,activation}.rs
 * link unchanged against libzenu_b200.so.  All pointers are device pointers, sizes/strides are int element
 * counts, work runs on the legacy default stream exactly like the reference kernels.
 * Unit-stride calls take the 128-bit vectorised kernels of the native ABI; other strides use a strided kernel.
 * Differences from the reference, all deliberate:
 *   - conv2d_bias_bkwd_* indexes NCHW correctly for N > 1 (reference bug);
 *   - conv_bias_add_* does not cudaDeviceSynchronize() after the launch (array_array.cu:65,73);
 *   - array_clip_* reads its input (the reference swaps in/out indices, array_scalar.cu:9);
 *   - array_max_idx_* treats `a` as the device pointer it is (the reference cudaMemcpy's it as host memory).
 */
#ifndef ZENU_KERNEL_COMPAT_H
#define ZENU_KERNEL_COMPAT_H
#ifdef __cplusplus
extern "C" {
#endif

#define ZENU_COMPAT_DECL_TYPE(T, SFX)                                                                              \
  void relu_##SFX(T* input, T* output, T alpha, int size, int input_stride, int output_stride);                   \
  void relu_backward_mask_##SFX(T* input, T* mask, T alpha, int size, int input_stride, int mask_stride);         \
  void array_array_add_##SFX(T* a, int stride_a, T* b, int stride_b, T* c, int stride_c, int n);                  \
  void array_array_sub_##SFX(T* a, int stride_a, T* b, int stride_b, T* c, int stride_c, int n);                  \
  void array_array_mul_##SFX(T* a, int stride_a, T* b, int stride_b, T* c, int stride_c, int n);                  \
  void array_array_div_##SFX(T* a, int stride_a, T* b, int stride_b, T* c, int stride_c, int n);                  \
  void array_array_add_assign_##SFX(T* a, int stride_a, T* b, int stride_b, int n);                               \
  void array_array_sub_assign_##SFX(T* a, int stride_a, T* b, int stride_b, int n);                               \
  void array_array_mul_assign_##SFX(T* a, int stride_a, T* b, int stride_b, int n);                               \
  void array_array_div_assign_##SFX(T* a, int stride_a, T* b, int stride_b, int n);                               \
  void conv_bias_add_##SFX(const T* input, T* output, int channel_stride, const T* bias, int bias_size,           \
                           int total_elements);                                                                    \
  void conv2d_bias_bkwd_##SFX(const T* dOut, T* dbias, int N, int C, int H, int W);                               \
  void array_scalar_add_##SFX(T* a, int size, int stride_a, T scalar, T* out, int stride_out);                    \
  void array_scalar_sub_##SFX(T* a, int size, int stride_a, T scalar, T* out, int stride_out);                    \
  void array_scalar_mul_##SFX(T* a, int size, int stride_a, T scalar, T* out, int stride_out);                    \
  void array_scalar_div_##SFX(T* a, int size, int stride_a, T scalar, T* out, int stride_out);                    \
  void array_scalar_add_assign_##SFX(T* a, int size, int stride, T scalar);                                       \
  void array_scalar_sub_assign_##SFX(T* a, int size, int stride, T scalar);                                       \
  void array_scalar_mul_assign_##SFX(T* a, int size, int stride, T scalar);                                       \
  void array_scalar_div_assign_##SFX(T* a, int size, int stride, T scalar);                                       \
  void array_scalar_pointer_add_##SFX(T* a, int size, int stride_a, T* scalar, T* out, int stride_out);           \
  void array_scalar_pointer_sub_##SFX(T* a, int size, int stride_a, T* scalar, T* out, int stride_out);           \
  void array_scalar_pointer_mul_##SFX(T* a, int size, int stride_a, T* scalar, T* out, int stride_out);           \
  void array_scalar_pointer_div_##SFX(T* a, int size, int stride_a, T* scalar, T* out, int stride_out);           \
  void array_scalar_pointer_add_assign_##SFX(T* a, int size, int stride, T* scalar);                              \
  void array_scalar_pointer_sub_assign_##SFX(T* a, int size, int stride, T* scalar);                              \
  void array_scalar_pointer_mul_assign_##SFX(T* a, int size, int stride, T* scalar);                              \
  void array_scalar_pointer_div_assign_##SFX(T* a, int size, int stride, T* scalar);                              \
  void array_clip_##SFX(T* input, T* output, int size, int stride_in, int stride_out, T min, T max);              \
  void array_clip_assign_##SFX(T* input, int size, int stride, T min, T max);                                     \
  void array_clip_backward_##SFX(T* input, T* mask, T max, T min, int size, int stride_in, int stride_mask);      \
  void array_clip_backward_assign_##SFX(T* input, T max, T min, int size, int stride);                            \
  void array_pow_##SFX(T* a, int size, int stride_a, T scalar, T* out, int stride_o);                             \
  void array_pow_assign_##SFX(T* a, int size, int stride, T scalar);                                              \
  void memory_access_##SFX(T* array, int offset, T* result);                                                      \
  void memory_set_##SFX(T* array, int offset, T value);                                                           \
  void array_max_idx_##SFX(T* a, int size, int stride, int* out);

#define ZENU_COMPAT_DECL_UNARY(NAME)                                                      \
  void array_##NAME##_float(float* a, int size, int stride_in, float* out, int stride_out);     \
  void array_##NAME##_double(double* a, int size, int stride_in, double* out, int stride_out);  \
  void array_##NAME##_assign_float(float* a, int size, int stride);                             \
  void array_##NAME##_assign_double(double* a, int size, int stride);

ZENU_COMPAT_DECL_TYPE(float, float)
ZENU_COMPAT_DECL_TYPE(double, double)
ZENU_COMPAT_DECL_UNARY(sin) ZENU_COMPAT_DECL_UNARY(cos) ZENU_COMPAT_DECL_UNARY(tan)
ZENU_COMPAT_DECL_UNARY(asin) ZENU_COMPAT_DECL_UNARY(acos) ZENU_COMPAT_DECL_UNARY(atan)
ZENU_COMPAT_DECL_UNARY(sinh) ZENU_COMPAT_DECL_UNARY(cosh) ZENU_COMPAT_DECL_UNARY(tanh)
ZENU_COMPAT_DECL_UNARY(abs) ZENU_COMPAT_DECL_UNARY(sqrt) ZENU_COMPAT_DECL_UNARY(exp) ZENU_COMPAT_DECL_UNARY(log)

#ifdef __cplusplus
}
#endif
#endif
