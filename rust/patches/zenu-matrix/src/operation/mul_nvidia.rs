//! Replacement for `impl Gemm for Nvidia` (zenu-matrix/src/operation/mul.rs:112-147).
//!
//! Before: cublas{S,D}gemm_v2_64 with A / B and (m, n) swapped to express a row-major C in cuBLAS' column-major world (:145), i32
//! extents, a handle that never gets a stream or a math mode (true-FP32 SGEMM).
//! After:  zb_gemm is row-major as written -- C[m,n] = alpha * op(A) * op(B) + beta * C with lda / ldb / ldc the row pitches --
//! on the tcgen05 TF32 path (or 3xTF32 / FFMA per zb_ctx_set_math; f64 always DFMA), i64 extents.
use zenu_b200_sys as sys;

use crate::{
    device::nvidia::{b200, Nvidia},
    matrix_blas::BlasTrans,
    num::Num,
};

use super::mul::Gemm;

impl Gemm for Nvidia {
    #[expect(clippy::many_single_char_names, clippy::similar_names)]
    fn gemm_unchecked<T: Num>(
        transa: BlasTrans, transb: BlasTrans, m: usize, n: usize, k: usize,
        alpha: T, a: *const T, lda: usize, b: *const T, ldb: usize, beta: T, c: *mut T, ldc: usize,
    ) {
        let t = |tr: BlasTrans| i32::from(!matches!(tr, BlasTrans::None));   // real types: Conjugate == Ordinary
        let i = |v: usize| i64::try_from(v).unwrap();
        b200::check(unsafe {
            sys::zb_gemm(
                b200::ctx(), b200::dtype::<T>(), sys::ZB_MATH_DEFAULT, t(transa), t(transb), i(m), i(n), i(k),
                alpha.to_f64().unwrap(), a.cast(), i(lda), b.cast(), i(ldb), beta.to_f64().unwrap(), c.cast(), i(ldc),
            )
        });
    }
}
