//! Replacement for `impl BatchNormalization for Nvidia` (zenu-matrix/src/nn/batch_norm.rs:169-280), which wraps
//! cudnnBatchNormalizationForwardTraining / Backward / ForwardInference (zenu-cuda/src/cudnn/batch_norm.rs:51-91,206-249,380-414).
//!
//! `momentum` arrives in the reference's convention (weight of the OLD running statistic, batch_norm.rs:310-311); the cuDNN path
//! converts it with `1. - momentum` (:183) -- zb_bn2d_fwd_train takes the reference convention directly, so it is passed through.
//! Epsilon is the library default 1e-10 (= batch_norm.rs:296 and zenu-cuda/src/cudnn/batch_norm.rs:82).  The `device_batch_norm*`
//! descriptor arguments are ignored: nothing is cached per shape.
use zenu_b200_sys as sys;

use crate::{
    device::nvidia::{b200, Nvidia},
    dim::DimDyn,
    matrix::{Matrix, Ref},
    num::Num,
};

use super::{BatchNorm2dBackwardConfig, BatchNorm2dConfig, BatchNorm2dInferenceConfig, BatchNormalization};

fn nchw(s: DimDyn) -> (i64, i64, i64, i64) {
    let i = |v: usize| i64::try_from(v).unwrap();
    (i(s[0]), i(s[1]), i(s[2]), i(s[3]))
}

impl BatchNormalization for Nvidia {
    fn batch_norm_2d_forward_train<T: Num>(
        momentum: f64,
        x: Matrix<Ref<&T>, DimDyn, Self>,
        y: Matrix<Ref<&mut T>, DimDyn, Self>,
        scale: Matrix<Ref<&T>, DimDyn, Self>,
        bias: Matrix<Ref<&T>, DimDyn, Self>,
        mean: Matrix<Ref<&mut T>, DimDyn, Self>,
        variance: Matrix<Ref<&mut T>, DimDyn, Self>,
        saving_mean: Option<Matrix<Ref<&mut T>, DimDyn, Self>>,
        saving_inv_variance: Option<Matrix<Ref<&mut T>, DimDyn, Self>>,
        _device_batch_norm: &Option<BatchNorm2dConfig<T>>,
    ) {
        let (n, c, h, w) = nchw(x.shape());
        let opt = |m: Option<Matrix<Ref<&mut T>, DimDyn, Self>>| m.map_or(std::ptr::null_mut(), |m| m.as_mut_ptr().cast());
        b200::check(unsafe {
            sys::zb_bn2d_fwd_train(
                b200::ctx(), b200::dtype::<T>(), sys::ZB_NCHW, n, c, h, w, momentum,
                x.as_ptr().cast(), scale.as_ptr().cast(), bias.as_ptr().cast(),
                mean.as_mut_ptr().cast(), variance.as_mut_ptr().cast(), opt(saving_mean), opt(saving_inv_variance),
                y.as_mut_ptr().cast(), std::ptr::null(), 0,
            )
        });
    }

    fn batch_norm_2d_backward<T: Num>(
        x: Matrix<Ref<&T>, DimDyn, Self>,
        y_grad: Matrix<Ref<&T>, DimDyn, Self>,
        x_grad: Matrix<Ref<&mut T>, DimDyn, Self>,
        scale: Matrix<Ref<&T>, DimDyn, Self>,
        scale_grad: Matrix<Ref<&mut T>, DimDyn, Self>,
        bias_grad: Matrix<Ref<&mut T>, DimDyn, Self>,
        saving_mean: Option<Matrix<Ref<&T>, DimDyn, Self>>,
        saving_inv_variance: Option<Matrix<Ref<&T>, DimDyn, Self>>,
        _device_batch_norm_backward: &Option<BatchNorm2dBackwardConfig<T>>,
    ) {
        let (n, c, h, w) = nchw(x.shape());
        // None = "recompute the batch statistics from x" (batch_norm.rs:355-368), which zb_bn2d_bwd does for NULL
        let opt = |m: Option<Matrix<Ref<&T>, DimDyn, Self>>| m.map_or(std::ptr::null(), |m| m.as_ptr().cast());
        b200::check(unsafe {
            sys::zb_bn2d_bwd(
                b200::ctx(), b200::dtype::<T>(), sys::ZB_NCHW, n, c, h, w,
                x.as_ptr().cast(), y_grad.as_ptr().cast(), scale.as_ptr().cast(), opt(saving_mean), opt(saving_inv_variance),
                x_grad.as_mut_ptr().cast(), scale_grad.as_mut_ptr().cast(), bias_grad.as_mut_ptr().cast(),
                std::ptr::null(), std::ptr::null_mut(),
            )
        });
    }

    fn bach_norm_2d_forward_inference<T: Num>(
        x: Matrix<Ref<&T>, DimDyn, Self>,
        y: Matrix<Ref<&mut T>, DimDyn, Self>,
        scale: Matrix<Ref<&T>, DimDyn, Self>,
        bias: Matrix<Ref<&T>, DimDyn, Self>,
        mean: Matrix<Ref<&T>, DimDyn, Self>,
        variance: Matrix<Ref<&T>, DimDyn, Self>,
        _device_batch_norm_inference: &Option<BatchNorm2dInferenceConfig<T>>,
    ) {
        let (n, c, h, w) = nchw(x.shape());
        b200::check(unsafe {
            sys::zb_bn2d_fwd_infer(
                b200::ctx(), b200::dtype::<T>(), sys::ZB_NCHW, n, c, h, w,
                x.as_ptr().cast(), scale.as_ptr().cast(), bias.as_ptr().cast(), mean.as_ptr().cast(), variance.as_ptr().cast(),
                y.as_mut_ptr().cast(),
            )
        });
    }
}
