//! Replacement for the `impl $name for Nvidia` arm of `impl_basic_op_trait!` (zenu-matrix/src/operation/basic_operations.rs:184-231):
//! AddOps / SubOps / MulOps / DivOps.  Contiguous operands (every stride 1: residual adds, gradient fan-in `grad + old`
//! (zenu-autograd/src/lib.rs:480-481), SGD's `grad * lr`) go to the vectorised native kernels; anything strided keeps using the
//! reference's own zenu-cuda wrappers, which now resolve to the symbol-compatible shim in libzenu_b200.so.
//!
//! `$op` is ZB_OP_ADD / SUB / MUL / DIV; `$gpu_*` are the zenu_cuda::kernel functions the macro already receives.
#[cfg(feature = "nvidia")]
macro_rules! impl_basic_op_nvidia {
    ($name:ident, $op:expr, $gpu_array:ident, $gpu_array_assign:ident, $gpu_scalar:ident, $gpu_scalar_assign:ident,
     $gpu_scalar_ptr:ident, $gpu_scalar_assign_ptr:ident) => {
        impl $name for Nvidia {
            fn array_array<T: Num>(to: *mut T, lhs: *const T, rhs: *const T, num_elm: usize, to_stride: usize, lhs_stride: usize, rhs_stride: usize) {
                if to_stride == 1 && lhs_stride == 1 && rhs_stride == 1 {
                    b200::check(unsafe {
                        sys::zb_binary(b200::ctx(), b200::dtype::<T>(), $op, lhs.cast(), rhs.cast(), to.cast(), i64::try_from(num_elm).unwrap())
                    });
                } else {
                    $gpu_array(to, lhs, rhs, num_elm, to_stride, lhs_stride, rhs_stride);
                }
            }

            fn array_assign<T: Num>(to: *mut T, rhs: *const T, num_elm: usize, to_stride: usize, rhs_stride: usize) {
                if to_stride == 1 && rhs_stride == 1 {
                    // c may alias a (include/zenu_b200.h: zb_binary)
                    b200::check(unsafe {
                        sys::zb_binary(b200::ctx(), b200::dtype::<T>(), $op, to.cast_const().cast(), rhs.cast(), to.cast(), i64::try_from(num_elm).unwrap())
                    });
                } else {
                    $gpu_array_assign(to, rhs, num_elm, to_stride, rhs_stride);
                }
            }

            fn scalar<T: Num>(to: *mut T, lhs: *const T, rhs: T, num_elm: usize, to_stride: usize, lhs_stride: usize) {
                if to_stride == 1 && lhs_stride == 1 {
                    b200::check(unsafe {
                        sys::zb_binary_scalar(b200::ctx(), b200::dtype::<T>(), $op, lhs.cast(), rhs.to_f64().unwrap(), to.cast(),
                                              i64::try_from(num_elm).unwrap())
                    });
                } else {
                    $gpu_scalar(to, lhs, rhs, num_elm, to_stride, lhs_stride);
                }
            }

            fn scalar_assign<T: Num>(to: *mut T, rhs: T, num_elm: usize, to_stride: usize) {
                if to_stride == 1 {
                    b200::check(unsafe {
                        sys::zb_binary_scalar(b200::ctx(), b200::dtype::<T>(), $op, to.cast_const().cast(), rhs.to_f64().unwrap(), to.cast(),
                                              i64::try_from(num_elm).unwrap())
                    });
                } else {
                    $gpu_scalar_assign(to, rhs, num_elm, to_stride);
                }
            }

            // scalar held in device memory (array_scalar_pointer_*): unchanged, served by the shim
            fn scalar_ptr<T: Num>(to: *mut T, lhs: *const T, scalar: *const T, to_stride: usize, lhs_stride: usize, num_elm: usize) {
                $gpu_scalar_ptr(to, lhs, scalar, num_elm, to_stride, lhs_stride);
            }

            fn scalar_assign_ptr<T: Num>(to: *mut T, scalar: *const T, num_elm: usize, to_stride: usize) {
                $gpu_scalar_assign_ptr(to, scalar, num_elm, to_stride);
            }
        }
    };
}
