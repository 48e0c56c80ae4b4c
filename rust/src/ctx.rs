//! One `spl_ctx` per host thread (spl.h); a non-zero status becomes a panic — the reference's
//! convention (`assert!` / `assert_eq!`, never `Result`).
use std::ffi::CStr;
use std::os::raw::c_int;

use crate::ffi::*;

pub(crate) struct Ctx {
    raw: *mut spl_ctx,
}

impl Ctx {
    fn new(device: c_int) -> Self {
        let mut raw = std::ptr::null_mut();
        let st = unsafe { spl_ctx_create(device, std::ptr::null_mut(), &mut raw) };
        if st != SPL_OK {
            panic!("spl_ctx_create failed ({st}): a CUDA device is required, there is no CPU fallback");
        }
        Ctx { raw }
    }
    pub(crate) fn raw(&self) -> *mut spl_ctx { self.raw }
    /// Panics with the library's message unless `status` is `SPL_OK`.
    pub(crate) fn check(&self, status: c_int) {
        if status != SPL_OK {
            let msg = unsafe { CStr::from_ptr(spl_last_error(self.raw)) }.to_string_lossy().into_owned();
            panic!("{msg}");
        }
    }
}

impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { spl_ctx_destroy(self.raw) };
    }
}

thread_local! {
    static CTX: Ctx = Ctx::new(0);
}

/// Runs `f` with this thread's context.
pub(crate) fn with_ctx<R>(f: impl FnOnce(&Ctx) -> R) -> R {
    CTX.with(|c| f(c))
}
