//! Replacement for `impl ReluOps for Nvidia` (zenu-matrix/src/operation/relu.rs:99-167).  Unit-stride calls (what
//! Matrix::relu / relu_backward_mask issue for contiguous tensors, :169-243) take the vectorised native kernels; strided calls
//! keep going through the symbol-compatible kernel-sys shim (relu_float / relu_double ..., include/zenu_kernel_compat.h), which
//! zenu-cuda/src/kernel/activation.rs:8-81 links unchanged.
use zenu_b200_sys as sys;
use zenu_cuda::kernel::activation::{relu as relu_strided, relu_backward_mask as relu_backward_mask_strided};

use crate::{device::nvidia::{b200, Nvidia}, num::Num};

use super::relu::ReluOps;

impl ReluOps for Nvidia {
    fn relu<T: Num>(input: *const T, output: *mut T, alpha: T, size: usize, input_stride: usize, output_stride: usize) {
        if input_stride == 1 && output_stride == 1 {
            b200::check(unsafe {
                sys::zb_relu(b200::ctx(), b200::dtype::<T>(), input.cast(), output.cast(), alpha.to_f64().unwrap(), i64::try_from(size).unwrap())
            });
        } else if T::is_f32() {
            relu_strided(input.cast_mut().cast::<f32>(), output.cast(), alpha.to_f32().unwrap(), size, input_stride, output_stride);
        } else {
            relu_strided(input.cast_mut().cast::<f64>(), output.cast(), alpha.to_f64().unwrap(), size, input_stride, output_stride);
        }
    }

    fn relu_backward_mask<T: Num>(input: *const T, mask: *mut T, alpha: T, size: usize, input_stride: usize, mask_stride: usize) {
        if input_stride == 1 && mask_stride == 1 {
            b200::check(unsafe {
                sys::zb_relu_backward_mask(b200::ctx(), b200::dtype::<T>(), input.cast(), mask.cast(), alpha.to_f64().unwrap(),
                                           i64::try_from(size).unwrap())
            });
        } else if T::is_f32() {
            relu_backward_mask_strided(input.cast_mut().cast::<f32>(), mask.cast(), alpha.to_f32().unwrap(), size, input_stride, mask_stride);
        } else {
            relu_backward_mask_strided(input.cast_mut().cast::<f64>(), mask.cast(), alpha.to_f64().unwrap(), size, input_stride, mask_stride);
        }
    }
}
