//! Page-locked host vectors for `matvec_into`: with pinned x and y the library pipelines the upload, the
//! product and the download over row chunks (`spl_spmv_host`); ordinary slices take one staged copy each way.
//! An extension: the reference has no dense vectors (its product is `&A * &X` with X an n x 1 matrix).

use std::ffi::c_void;
use std::ops::{Deref, DerefMut};

use crate::ffi::{spl_host_alloc, spl_host_free, SPL_OK};
use crate::scalar::Scalar;

/// A fixed-length vector of `T` in page-locked host memory (`spl_host_alloc`), zero-initialised.
pub struct PinnedVec<T: Scalar> {
    ptr: *mut T,
    len: usize,
}

// plain host memory owned by this value
unsafe impl<T: Scalar> Send for PinnedVec<T> {}
unsafe impl<T: Scalar> Sync for PinnedVec<T> {}

impl<T: Scalar> PinnedVec<T> {
    /// `len` zeros.  Panics when the driver cannot pin that much memory.
    pub fn zeros(len: usize) -> Self {
        let mut p: *mut c_void = std::ptr::null_mut();
        let status = unsafe { spl_host_alloc((len * std::mem::size_of::<T>()) as u64, &mut p) };
        assert!(status == SPL_OK && !p.is_null(), "spl_host_alloc failed (no CUDA device, or out of pinnable memory)");
        let v = PinnedVec { ptr: p as *mut T, len };
        unsafe { std::ptr::write_bytes(v.ptr, 0, len) };
        v
    }

    /// A pinned copy of `x`.
    pub fn from_slice(x: &[T]) -> Self {
        let mut v = Self::zeros(x.len());
        v.copy_from_slice(x);
        v
    }
}

impl<T: Scalar> Deref for PinnedVec<T> {
    type Target = [T];
    fn deref(&self) -> &[T] { unsafe { std::slice::from_raw_parts(self.ptr, self.len) } }
}

impl<T: Scalar> DerefMut for PinnedVec<T> {
    fn deref_mut(&mut self) -> &mut [T] { unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) } }
}

impl<T: Scalar> Drop for PinnedVec<T> {
    fn drop(&mut self) { unsafe { spl_host_free(self.ptr as *mut c_void) }; }
}
