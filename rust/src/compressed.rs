//! What `CsrMatrix` and `CscMatrix` share: the device handle (`spl_mat`, immutable after creation,
//! freed on drop) and the lazily downloaded host mirror behind `rowptr()` / `colind()` / `values()`
//! (exactly sized `Vec`s with `usize` indices, as the reference's tests assert on `capacity()`).
use std::marker::PhantomData;
use std::os::raw::{c_int, c_void};
use std::sync::atomic::{AtomicBool, Ordering};
use std::sync::OnceLock;

use crate::coo::CooMatrix;
use crate::ctx::with_ctx;
use crate::dok::DokMatrix;
use crate::ffi::*;
use crate::scalar::Scalar;

pub(crate) struct Host<T> {
    pub ptr: Vec<usize>,
    pub ind: Vec<usize>,
    pub val: Vec<T>,
}

pub(crate) struct Compressed<T: Scalar> {
    pub nrows: usize,
    pub ncols: usize,
    pub nnz: usize,
    format: c_int,
    raw: *mut spl_mat,
    host: OnceLock<Host<T>>,
    dirty: AtomicBool,           // values_mut() was handed out: store the values back before device work
    _t: PhantomData<T>,
}

// The reference's CsrMatrix / CscMatrix are plain Vecs: Send + Sync, and `&`-operations from several
// threads on one matrix are legal.  They stay legal here: the device handle is immutable after creation,
// every thread works through its own context (stream), and the library orders those streams against
// the matrix itself (a `ready` event recorded at creation, one event per foreign stream that used it,
// waited for before the blocks are freed: include/spl.h, "sharing across contexts").  The host mirror
// is filled once behind a OnceLock; the write-back flag is atomic and only set through `&mut self`.
unsafe impl<T: Scalar> Send for Compressed<T> {}
unsafe impl<T: Scalar> Sync for Compressed<T> {}

impl<T: Scalar> Compressed<T> {
    /// Takes ownership of a handle returned by the library.
    pub fn adopt(raw: *mut spl_mat) -> Self {
        let (mut format, mut dtype) = (0, 0);
        let (mut nrows, mut ncols, mut nnz) = (0u64, 0u64, 0u64);
        unsafe { spl_mat_info(raw, &mut format, &mut dtype, &mut nrows, &mut ncols, &mut nnz) };
        debug_assert_eq!(dtype, T::DTYPE);
        Compressed { nrows: nrows as usize, ncols: ncols as usize, nnz: nnz as usize, format, raw,
                     host: OnceLock::new(), dirty: AtomicBool::new(false), _t: PhantomData }
    }

    /// `CsrMatrix::new` / `CscMatrix::new` (src/csr.rs:137-164, src/csc.rs:137-164): validating;
    /// the library reports which assertion failed and the message becomes the panic.
    pub fn new(format: c_int, nrows: usize, ncols: usize, ptr: Vec<usize>, ind: Vec<usize>, val: Vec<T>) -> Self {
        let mut raw = std::ptr::null_mut();
        with_ctx(|c| c.check(unsafe {
            spl_mat_from_compressed(c.raw(), format, T::DTYPE, nrows as u64, ncols as u64,
                ptr.len() as u64, ptr.as_ptr() as *const u64, ind.len() as u64, ind.as_ptr() as *const u64,
                val.len() as u64, val.as_ptr() as *const c_void, &mut raw)
        }));
        let m = Self::adopt(raw);
        let _ = m.host.set(Host { ptr, ind, val });      // the caller's vectors are the host mirror
        m
    }

    pub fn eye(format: c_int, size: usize) -> Self {
        assert!(size > 0);
        let mut raw = std::ptr::null_mut();
        with_ctx(|c| c.check(unsafe { spl_mat_eye(c.raw(), format, T::DTYPE, size as u64, &mut raw) }));
        Self::adopt(raw)
    }

    /// From<&CooMatrix>: the triplets are already on the device (streamed while pushed).
    pub fn from_coo(format: c_int, coo: &CooMatrix<T>) -> Self {
        let mut raw = std::ptr::null_mut();
        with_ctx(|c| c.check(unsafe { spl_mat_from_coo_builder(c.raw(), coo.raw(), format, 1, 1, &mut raw) }));
        Self::adopt(raw)
    }

    /// From<&DokMatrix>: keys unique, explicit zeros kept (dedup = 0, dropzero = 0).
    pub fn from_dok(format: c_int, dok: &DokMatrix<T>) -> Self {
        let (rows, cols, vals) = dok.triplets();
        let mut raw = std::ptr::null_mut();
        with_ctx(|c| c.check(unsafe {
            spl_mat_from_coo(c.raw(), format, T::DTYPE, dok.nrows() as u64, dok.ncols() as u64, vals.len() as u64,
                rows.as_ptr() as *const u64, cols.as_ptr() as *const u64, vals.as_ptr() as *const c_void,
                0, 0, &mut raw)
        }));
        Self::adopt(raw)
    }

    pub fn raw(&self) -> *const spl_mat {
        self.flush();
        self.raw
    }

    fn flush(&self) {
        if self.dirty.swap(false, Ordering::AcqRel) {
            let h = self.host.get().expect("values_mut implies a host mirror");
            with_ctx(|c| c.check(unsafe { spl_mat_set_values(c.raw(), self.raw, h.val.as_ptr() as *const c_void) }));
        }
    }

    pub fn host(&self) -> &Host<T> {
        self.host.get_or_init(|| {
            let nmajor = if self.format == SPL_CSR { self.nrows } else { self.ncols };
            // exactly sized: the reference's tests assert capacity() == len()
            let mut ptr = vec![0usize; nmajor + 1];
            let mut ind = vec![0usize; self.nnz];
            let mut val = vec![T::zero(); self.nnz];
            with_ctx(|c| c.check(unsafe {
                spl_mat_download(c.raw(), self.raw, ptr.as_mut_ptr() as *mut u64, ind.as_mut_ptr() as *mut u64,
                                 val.as_mut_ptr() as *mut c_void)
            }));
            Host { ptr, ind, val }
        })
    }

    /// `values_mut()` (src/csr.rs:270-272): structure kept, values written back lazily.
    pub fn values_mut(&mut self) -> &mut [T] {
        self.host();
        self.dirty.store(true, Ordering::Release);
        &mut self.host.get_mut().unwrap().val
    }

    /// Pointer and index arrays shared, values mutable: what `iter_mut` walks.
    pub fn host_parts_mut(&mut self) -> (&[usize], &[usize], &mut [T]) {
        self.host();
        let h = self.host.get_mut().unwrap();
        (&h.ptr, &h.ind, &mut h.val)
    }

    pub fn unary(&self, f: unsafe extern "C" fn(*mut spl_ctx, *const spl_mat, *mut *mut spl_mat) -> c_int) -> Self {
        let mut raw = std::ptr::null_mut();
        with_ctx(|c| c.check(unsafe { f(c.raw(), self.raw(), &mut raw) }));
        Self::adopt(raw)
    }

    pub fn binary(&self, rhs: &Self,
                  f: unsafe extern "C" fn(*mut spl_ctx, *const spl_mat, *const spl_mat, *mut *mut spl_mat) -> c_int) -> Self {
        let mut raw = std::ptr::null_mut();
        with_ctx(|c| c.check(unsafe { f(c.raw(), self.raw(), rhs.raw(), &mut raw) }));
        Self::adopt(raw)
    }

    pub fn convert(&self, format: c_int) -> Self {
        let mut raw = std::ptr::null_mut();
        with_ctx(|c| c.check(unsafe { spl_mat_convert(c.raw(), self.raw(), format, &mut raw) }));
        Self::adopt(raw)
    }

    /// Storage-order triplets (src/coo.rs:629-705).
    pub fn to_triplets(&self) -> (Vec<usize>, Vec<usize>, Vec<T>) {
        let (mut rows, mut cols, mut vals) = (vec![0usize; self.nnz], vec![0usize; self.nnz], vec![T::zero(); self.nnz]);
        with_ctx(|c| c.check(unsafe {
            spl_mat_to_coo(c.raw(), self.raw(), rows.as_mut_ptr() as *mut u64, cols.as_mut_ptr() as *mut u64,
                           vals.as_mut_ptr() as *mut c_void)
        }));
        (rows, cols, vals)
    }

    /// Stored entries [start, start + count) in storage order, read from the device (`spl_mat_read_entries`):
    /// the chunk an iterator holds instead of a host copy of the whole matrix.
    pub fn read_entries(&self, start: usize, count: usize) -> (Vec<usize>, Vec<usize>, Vec<T>) {
        let (mut rows, mut cols, mut vals) = (vec![0usize; count], vec![0usize; count], vec![T::zero(); count]);
        with_ctx(|c| c.check(unsafe {
            spl_mat_read_entries(c.raw(), self.raw(), start as u64, count as u64, rows.as_mut_ptr() as *mut u64,
                                 cols.as_mut_ptr() as *mut u64, vals.as_mut_ptr() as *mut c_void)
        }));
        (rows, cols, vals)
    }

    /// Row (CSR) or column (CSC) of the stored entry at position `p`, from the host mirror's pointer array.
    pub fn major_of(ptr: &[usize], p: usize) -> usize {
        ptr.partition_point(|&q| q <= p) - 1
    }

    /// y = A x with host vectors: the dense form of `&A * &X`, X n x 1.  On a CSC matrix the first product
    /// builds (and keeps) the CSR form on the device.  Pageable slices take the driver's staged copies;
    /// vectors allocated with `spl_host_alloc` (pinned) are pipelined chunk by chunk.
    pub fn matvec(&self, x: &[T]) -> Vec<T> {
        assert_eq!(self.ncols, x.len());
        let mut y = vec![T::zero(); self.nrows];
        with_ctx(|c| c.check(unsafe {
            spl_spmv_host(c.raw(), self.raw(), x.as_ptr() as *const c_void, y.as_mut_ptr() as *mut c_void)
        }));
        y
    }

    /// `matvec` into a caller-owned vector.  With `PinnedVec`s for both (page-locked memory) the upload, the
    /// product and the download are pipelined over row chunks; any other slices give the same result through
    /// one staged copy each way.
    pub fn matvec_into(&self, x: &[T], y: &mut [T]) {
        assert_eq!(self.ncols, x.len());
        assert_eq!(self.nrows, y.len());
        with_ctx(|c| c.check(unsafe {
            spl_spmv_host(c.raw(), self.raw(), x.as_ptr() as *const c_void, y.as_mut_ptr() as *mut c_void)
        }));
    }
}

impl<T: Scalar> std::fmt::Debug for Compressed<T> {
    /// `#[derive(Debug)]` in the reference (src/csr.rs:65): dims and the three arrays.
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        let h = self.host();
        f.debug_struct(if self.format == SPL_CSR { "CsrMatrix" } else { "CscMatrix" })
            .field("nrows", &self.nrows).field("ncols", &self.ncols)
            .field(if self.format == SPL_CSR { "rowptr" } else { "colptr" }, &h.ptr)
            .field(if self.format == SPL_CSR { "colind" } else { "rowind" }, &h.ind)
            .field("values", &h.val).finish()
    }
}

impl<T: Scalar> Drop for Compressed<T> {
    fn drop(&mut self) {
        with_ctx(|c| unsafe { spl_mat_free(c.raw(), self.raw) });
    }
}
