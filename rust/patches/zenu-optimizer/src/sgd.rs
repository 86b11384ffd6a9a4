//! zenu-optimizer/src/sgd.rs:20-30 with the data-parallel hook and the fused update.
//! Before: per parameter `tmp = grad * lr` (alloc + kernel) then `data -= tmp` (kernel).
//! After:  gradients averaged over the ranks first (no-op for one process), then one fused kernel per parameter (`p -= lr * g`).
use zenu_b200_sys as sys;
use zenu_layer::Parameters;
use zenu_matrix::{device::{nvidia::b200, Device}, num::Num};

use crate::{dp::allreduce_mean_grads, Optimizer};

pub struct SGD<T: Num, D: Device> {
    pub learning_rate: T,
    _device: std::marker::PhantomData<D>,
}

impl<T: Num, D: Device> SGD<T, D> {
    pub fn new(learning_rate: T) -> Self {
        Self { learning_rate, _device: std::marker::PhantomData }
    }
}

impl<T: Num, D: Device, P: Parameters<T, D>> Optimizer<T, D, P> for SGD<T, D> {
    fn update(&self, parameters: &P) {
        let params = parameters.parameters();
        allreduce_mean_grads(&params);                       // <- the "inside update_parameters" hook
        for data in params.values() {
            if let Some(grad) = data.get_grad() {
                if D::is_nvidia() {
                    let n = i64::try_from(data.get_data().shape().num_elm()).unwrap();
                    b200::check(unsafe {
                        sys::zb_sgd_step(b200::ctx(), b200::dtype::<T>(), data.get_data_mut().as_ptr().cast_mut().cast(),
                                         grad.get_data().as_ptr().cast(), self.learning_rate.to_f64().unwrap(), 1.0, n)
                    });
                } else {
                    let update_data = grad.get_data().to_ref() * self.learning_rate;
                    let mut data = data.get_data_mut();
                    let mut data = data.to_ref_mut();
                    data -= update_data;
                }
            }
        }
    }
}
