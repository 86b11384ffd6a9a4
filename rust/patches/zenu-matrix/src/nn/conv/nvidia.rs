//! Replacement for zenu-matrix/src/nn/conv/nvidia.rs:16-155 (`impl ConvFwd / ConvBkwdData / ConvBkwdFilter / ConvBias for Nvidia`).
//!
//! Before: validate the cached config, rebuild the cuDNN-frontend graph when the batch changed, allocate a workspace per call
//! (NNCache), execute the graph (zenu-cuda/src/cudnn/graph_conv.rs:94-111 -> conv.cpp:71-81).
//! After:  one plan-less call per op with the geometry taken from the LIVE matrices (this also fixes SURVEY S3: the layer's config
//! is built from a 32x32 dummy input and the graph asserted C/H/W equal to it), no workspace, no descriptor to leak.
//! `config` is kept in the signatures so the traits (interface.rs:37-89) and every caller stay unchanged; only padding / stride /
//! dilation are read from it.
use zenu_b200_sys as sys;

use crate::{
    device::nvidia::{b200, Nvidia},
    dim::{DimDyn, DimTrait},
    matrix::{Matrix, Ref},
    num::Num,
};

use super::interface::{
    ConvBias, ConvBkwdData, ConvBkwdDataConfig, ConvBkwdFilter, ConvBkwdFilterConfig, ConvConfigInner, ConvFwd, ConvFwdConfig,
};

/// zb_conv2d_desc from the input [N,C,H,W] and filter [K,C,kh,kw] shapes (NCHW / KCRS, default strides: interface.rs:270-281).
fn desc(x: DimDyn, w: DimDyn, inner: &ConvConfigInner) -> sys::zb_conv2d_desc {
    let i = |v: usize| i64::try_from(v).unwrap();
    sys::zb_conv2d_desc {
        n: i(x[0]), c: i(x[1]), h: i(x[2]), w: i(x[3]),
        k: i(w[0]), kh: i(w[2]), kw: i(w[3]),
        pad_h: i(inner.padding[0]), pad_w: i(inner.padding[1]),
        stride_h: i(inner.stride[0]), stride_w: i(inner.stride[1]),
        dil_h: i(inner.dilation[0]), dil_w: i(inner.dilation[1]),
    }
}

impl ConvFwd for Nvidia {
    fn conv_fwd<T: Num>(
        input: Matrix<Ref<&T>, DimDyn, Self>,
        weight: Matrix<Ref<&T>, DimDyn, Self>,
        output: Matrix<Ref<&mut T>, DimDyn, Self>,
        config: &mut ConvFwdConfig<T>,
    ) {
        let d = desc(input.shape(), weight.shape(), &config.inner);
        b200::check(unsafe {
            sys::zb_conv2d_fprop(
                b200::ctx(), b200::dtype::<T>(), sys::ZB_NCHW, sys::ZB_MATH_DEFAULT, &d,
                input.as_ptr().cast(), weight.as_ptr().cast(), std::ptr::null(), output.as_mut_ptr().cast(),
            )
        });
    }
}

impl ConvBkwdData for Nvidia {
    fn conv_bkwd_data<T: Num>(
        dy: Matrix<Ref<&T>, DimDyn, Self>,
        filter: Matrix<Ref<&T>, DimDyn, Self>,
        dx: Matrix<Ref<&mut T>, DimDyn, Self>,
        config: &mut ConvBkwdDataConfig<T>,
    ) {
        let d = desc(dx.shape(), filter.shape(), &config.inner);
        b200::check(unsafe {
            sys::zb_conv2d_dgrad(
                b200::ctx(), b200::dtype::<T>(), sys::ZB_NCHW, sys::ZB_MATH_DEFAULT, &d,
                dy.as_ptr().cast(), filter.as_ptr().cast(), dx.as_mut_ptr().cast(),
            )
        });
    }
}

impl ConvBkwdFilter for Nvidia {
    fn conv_bkwd_filter<T: Num>(
        dy: Matrix<Ref<&T>, DimDyn, Self>,
        x: Matrix<Ref<&T>, DimDyn, Self>,
        dw: Matrix<Ref<&mut T>, DimDyn, Self>,
        config: &mut ConvBkwdFilterConfig<T>,
    ) {
        let d = desc(x.shape(), dw.shape(), &config.inner);
        b200::check(unsafe {
            sys::zb_conv2d_wgrad(
                b200::ctx(), b200::dtype::<T>(), sys::ZB_NCHW, sys::ZB_MATH_DEFAULT, &d,
                dy.as_ptr().cast(), x.as_ptr().cast(), dw.as_mut_ptr().cast(),
            )
        });
    }
}

impl ConvBias for Nvidia {
    fn conv2d_bias<T: Num>(
        input: Matrix<Ref<&T>, DimDyn, Self>,
        bias: Matrix<Ref<&T>, DimDyn, Self>,
        output: Matrix<Ref<&mut T>, DimDyn, Self>,
    ) {
        let s = input.shape();
        let i = |v: usize| i64::try_from(v).unwrap();
        // bias is [1, K, 1, 1] (zenu-layer/src/layers/conv2d.rs:99); no cudaDeviceSynchronize after the launch (array_array.cu:65,73)
        debug_assert_eq!(bias.shape().num_elm(), s[1]);
        b200::check(unsafe {
            sys::zb_conv2d_bias_add(
                b200::ctx(), b200::dtype::<T>(), sys::ZB_NCHW, input.as_ptr().cast(), bias.as_ptr().cast(),
                output.as_mut_ptr().cast(), i(s[0]), i(s[1]), i(s[2]), i(s[3]),
            )
        });
    }

    fn conv2d_bias_bkwd<T: Num>(d_output: Matrix<Ref<&T>, DimDyn, Self>, bias: Matrix<Ref<&mut T>, DimDyn, Self>) {
        let s = d_output.shape();
        let i = |v: usize| i64::try_from(v).unwrap();
        // correct for N > 1 (the reference kernel indexes NCHW as [C][N*H*W], conv2d_bkwd_data.cu:134,144-146)
        b200::check(unsafe {
            sys::zb_conv2d_bias_bwd(
                b200::ctx(), b200::dtype::<T>(), sys::ZB_NCHW, d_output.as_ptr().cast(), bias.as_mut_ptr().cast(),
                i(s[0]), i(s[1]), i(s[2]), i(s[3]),
            )
        });
    }
}
