//! Data-parallel gradient exchange for the reference's optimizers: the hook BASELINE.json calls "inside update_parameters" sits at
//! the top of `Optimizer::update` (zenu-optimizer/src/sgd.rs:20-30, adam.rs:19-58, adamw.rs:20-69).
//!
//! One process per GPU (`Variable` is `Rc<RefCell>`, the runtime is single-device): rank 0 creates the NCCL id with
//! `zb_dp_unique_id`, ships the 128 bytes to the other ranks by whatever rendezvous the launcher offers (here: a file named by
//! ZENU_DP_ID_FILE on a shared filesystem), every rank calls `zb_dp_init`.  `allreduce_mean_grads` then sums each gradient over the
//! ranks on the library's communication stream and scales by 1 / world.
//!
//! This per-tensor form keeps the reference's one-Variable-per-parameter storage.  The native host stack (zb_model_*) shows the
//! faster layout: parameters and gradients in flat ~25 MB buckets (zb_dp_plan_buckets), one ncclAllReduce per bucket enqueued as
//! soon as backward has produced its last gradient, 1 / world folded into the fused optimizer kernel (grad_scale).
use std::collections::HashMap;
use std::io::{Read, Write};

use zenu_autograd::Variable;
use zenu_b200_sys as sys;
use zenu_matrix::{device::{nvidia::{b200, Nvidia}, Device}, num::Num};

pub fn init_from_env() {
    let rank: i32 = std::env::var("RANK").ok().and_then(|v| v.parse().ok()).unwrap_or(0);
    let world: i32 = std::env::var("WORLD_SIZE").ok().and_then(|v| v.parse().ok()).unwrap_or(1);
    let mut id = [0u8; 128];
    if world > 1 {
        let path = std::env::var("ZENU_DP_ID_FILE").expect("ZENU_DP_ID_FILE (shared path for the NCCL unique id)");
        if rank == 0 {
            b200::check(unsafe { sys::zb_dp_unique_id(b200::ctx(), id.as_mut_ptr().cast()) });
            let tmp = format!("{path}.tmp");
            std::fs::File::create(&tmp).unwrap().write_all(&id).unwrap();
            std::fs::rename(&tmp, &path).unwrap();
        } else {
            loop {
                if let Ok(mut f) = std::fs::File::open(&path) {
                    if f.read_exact(&mut id).is_ok() {
                        break;
                    }
                }
                std::thread::sleep(std::time::Duration::from_millis(50));
            }
        }
    }
    b200::check(unsafe { sys::zb_dp_init(b200::ctx(), id.as_ptr().cast(), rank, world) });
}

/// Sum every gradient over the ranks and divide by the world size.  Parameters are visited in sorted key order: the reference
/// iterates a HashMap, whose order differs between processes, and collectives must be issued in the same order on every rank.
pub fn allreduce_mean_grads<T: Num, D: Device>(parameters: &HashMap<String, Variable<T, D>>) {
    let world = unsafe { sys::zb_dp_world(b200::ctx()) };
    if world <= 1 || std::any::TypeId::of::<D>() != std::any::TypeId::of::<Nvidia>() {
        return;
    }
    let mut keys: Vec<&String> = parameters.keys().collect();
    keys.sort();
    let mut reduced = Vec::new();
    for k in keys {
        if let Some(grad) = parameters[k].get_grad() {
            let g = grad.get_data_mut();
            let n = i64::try_from(g.shape().num_elm()).unwrap();
            b200::check(unsafe { sys::zb_dp_allreduce_sum(b200::ctx(), b200::dtype::<T>(), g.as_ptr().cast_mut().cast(), n) });
            reduced.push((grad, n));
        }
    }
    b200::check(unsafe { sys::zb_dp_wait(b200::ctx()) });
    let inv = 1.0 / f64::from(world);
    for (grad, n) in reduced {
        let g = grad.get_data_mut();
        let p = g.as_ptr().cast_mut().cast();
        b200::check(unsafe { sys::zb_binary_scalar(b200::ctx(), b200::dtype::<T>(), sys::ZB_OP_MUL, p, inv, p, n) });
    }
}
