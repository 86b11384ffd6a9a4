//! Process-wide handle of the B200 backend.  Replaces `ZENU_CUDA_STATE: Lazy<Mutex<ZenuCudaState{cublas, cudnn, stream, mempool}>>`
//! (zenu-cuda/src/lib.rs:18-112) for the ops that moved to libzenu_b200: one `zb_ctx` per process / rank, no mutex around the calls
//! (the host graph is `Rc<RefCell>`, i.e. single-threaded already), no cuBLAS / cuDNN handles.
//!
//! Add to zenu-matrix/src/device/nvidia/mod.rs:  `pub mod b200;`  (this file as device/nvidia/b200.rs).
use std::ffi::CStr;
use std::os::raw::c_int;
use std::ptr;
use std::sync::OnceLock;

use zenu_b200_sys as sys;

use crate::num::Num;

pub struct B200Ctx(pub *mut sys::zb_ctx);
// The context is only ever used from the thread that owns the (non-Send) Variables; OnceLock needs the marker traits.
unsafe impl Send for B200Ctx {}
unsafe impl Sync for B200Ctx {}

static CTX: OnceLock<B200Ctx> = OnceLock::new();

/// Device ordinal: `LOCAL_RANK` when launched one process per GPU (the data-parallel layout), else 0 like the reference
/// (zenu-cuda/src/runtime/mod.rs:127,210 hard-codes device 0).
pub fn ctx() -> *mut sys::zb_ctx {
    CTX.get_or_init(|| {
        let device: c_int = std::env::var("LOCAL_RANK").ok().and_then(|v| v.parse().ok()).unwrap_or(0);
        let mut raw: *mut sys::zb_ctx = ptr::null_mut();
        // stream = cudaStreamLegacy (0x1): the ops that have NOT moved (pooling, dropout, RNN through cuDNN) still run on the legacy
        // default stream, so everything stays ordered without extra synchronisation.
        let rc = unsafe { sys::zb_ctx_create(&mut raw, device, 0x1 as *mut _) };
        check(rc);
        B200Ctx(raw)
    })
    .0
}

/// `success_or_panic` of zenu-cuda/src/cudnn/graph_utils.rs:31-36, for zb_status.
#[track_caller]
pub fn check(rc: c_int) {
    if rc != sys::ZB_OK {
        let msg = unsafe { CStr::from_ptr(sys::zb_last_error()) }.to_string_lossy().into_owned();
        panic!("zenu_b200 status {rc}: {msg}");
    }
}

/// zb_dtype of a `Num` (the reference dispatches on TypeId / is_f32(), zenu-matrix/src/num.rs:44-86).
pub fn dtype<T: Num>() -> c_int {
    if T::is_f32() {
        sys::ZB_F32
    } else {
        sys::ZB_F64
    }
}
