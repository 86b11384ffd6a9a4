//! Bodies of `Adam::update` (zenu-optimizer/src/adam.rs:19-58) and `AdamW::update` (adamw.rs:20-69) with the data-parallel hook and
//! the fused kernel.  Before: ~10 Matrix ops and as many temporaries per parameter.  After: one kernel per parameter.
//! `step_t` is the 1-based step after the increment (adam.rs:21-25: bias corrections beta.powf(step)); AdamW decays the tensors
//! returned by `weights()` only (adamw.rs:28,61-65) with the decoupled form `p -= lr * wd * p`.
use zenu_b200_sys as sys;
use zenu_layer::Parameters;
use zenu_matrix::{device::{nvidia::b200, Device}, num::Num};

use crate::dp::allreduce_mean_grads;

#[allow(clippy::too_many_arguments)]
pub fn adam_update_nvidia<T: Num, D: Device, P: Parameters<T, D>>(
    parameters: &P, m: &std::collections::HashMap<String, zenu_autograd::Variable<T, D>>,
    v: &std::collections::HashMap<String, zenu_autograd::Variable<T, D>>, step_t: usize, lr: T, beta1: T, beta2: T, eps: T,
    weight_decay: Option<T>,
) {
    let params = parameters.parameters();
    allreduce_mean_grads(&params);                           // <- the "inside update_parameters" hook
    let decayed = weight_decay.map(|_| parameters.weights());
    for (key, data) in &params {
        let Some(grad) = data.get_grad() else { continue };   // e.g. BatchNorm running statistics
        let decay = decayed.as_ref().is_some_and(|w| w.contains_key(key));
        let n = i64::try_from(data.get_data().shape().num_elm()).unwrap();
        b200::check(unsafe {
            sys::zb_adam_step(
                b200::ctx(), b200::dtype::<T>(), data.get_data_mut().as_ptr().cast_mut().cast(), grad.get_data().as_ptr().cast(),
                m[key].get_data_mut().as_ptr().cast_mut().cast(), v[key].get_data_mut().as_ptr().cast_mut().cast(),
                lr.to_f64().unwrap(), beta1.to_f64().unwrap(), beta2.to_f64().unwrap(), eps.to_f64().unwrap(),
                weight_decay.map_or(0.0, |w| w.to_f64().unwrap()), i32::from(decay), i64::try_from(step_t).unwrap(), 1.0, n,
            )
        });
    }
}
