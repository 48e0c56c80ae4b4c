//! `CooMatrix<T>` (reference: src/coo.rs:52-57): insertion-ordered triplets, duplicates allowed.
//! Storage is an `spl_coo`: three pinned host arrays (row, col, value) owned by the library and
//! streamed to the device chunk by chunk while they are filled, so `CsrMatrix::from(&coo)` finds the
//! triplets already in HBM.  `get()` still hands out `(&usize, &usize, &T)` (src/coo.rs:386-390).
use std::marker::PhantomData;
use std::os::raw::c_void;

use crate::csc::CscMatrix;
use crate::csr::CsrMatrix;
use crate::ctx::with_ctx;
use crate::dok::DokMatrix;
use crate::ffi::*;
use crate::scalar::Scalar;

pub struct CooMatrix<T: Scalar> {
    nrows: usize,
    ncols: usize,
    raw: *mut spl_coo,
    _t: PhantomData<T>,
}

// The storage is plain memory behind a unique handle; `&mut self` guards every mutation.
unsafe impl<T: Scalar> Send for CooMatrix<T> {}
unsafe impl<T: Scalar> Sync for CooMatrix<T> {}

impl<T: Scalar> CooMatrix<T> {
    /// src/coo.rs:104-112
    pub fn new(nrows: usize, ncols: usize) -> Self {
        Self::with_capacity(nrows, ncols, 0)
    }

    /// src/coo.rs:162-170
    pub fn with_capacity(nrows: usize, ncols: usize, capacity: usize) -> Self {
        assert!(nrows > 0);
        assert!(ncols > 0);
        let mut raw = std::ptr::null_mut();
        with_ctx(|c| {
            c.check(unsafe { spl_coo_create(c.raw(), T::DTYPE, nrows as u64, ncols as u64, capacity as u64, &mut raw) })
        });
        CooMatrix { nrows, ncols, raw, _t: PhantomData }
    }

    /// src/coo.rs:127-139
    pub fn eye(size: usize) -> Self {
        assert!(size > 0);
        let mut m = Self::with_capacity(size, size, size);
        for i in 0..size {
            m.push(i, i, T::one());
        }
        m
    }

    /// src/coo.rs:204-220
    pub fn with_entries<I: IntoIterator<Item = (usize, usize, T)>>(nrows: usize, ncols: usize, entries: I) -> Self {
        let mut m = Self::new(nrows, ncols);
        m.extend(entries);
        m
    }

    /// src/coo.rs:254-288
    pub fn with_triplets(nrows: usize, ncols: usize, rowind: &[usize], colind: &[usize], values: &[T]) -> Self {
        assert_eq!(rowind.len(), values.len());
        assert_eq!(colind.len(), values.len());
        let mut m = Self::with_capacity(nrows, ncols, values.len());
        // usize is 64 bit on every target the library supports: the slices go down as they are
        let st = unsafe {
            spl_coo_extend(m.raw, values.len() as u64, rowind.as_ptr() as *const u64, colind.as_ptr() as *const u64,
                           values.as_ptr() as *const c_void)
        };
        m.check(st);
        m
    }

    pub fn nrows(&self) -> usize { self.nrows }
    pub fn ncols(&self) -> usize { self.ncols }
    pub fn shape(&self) -> (usize, usize) { (self.nrows, self.ncols) }
    /// src/coo.rs:349-351
    pub fn length(&self) -> usize { unsafe { spl_coo_len(self.raw) as usize } }
    /// src/coo.rs:366-368
    pub fn capacity(&self) -> usize { unsafe { spl_coo_capacity(self.raw) as usize } }

    fn host(&self) -> (*const usize, *const usize, *const T) {
        let (mut r, mut c, mut v) = (std::ptr::null(), std::ptr::null(), std::ptr::null());
        unsafe { spl_coo_host_ptrs(self.raw, &mut r, &mut c, &mut v) };
        (r as *const usize, c as *const usize, v as *const T)
    }

    /// src/coo.rs:386-390
    pub fn get(&self, index: usize) -> Option<(&usize, &usize, &T)> {
        if index >= self.length() {
            return None;
        }
        let (r, c, v) = self.host();
        // borrowed from the pinned arrays: valid until the next `&mut self` call
        unsafe { Some((&*r.add(index), &*c.add(index), &*v.add(index))) }
    }

    /// src/coo.rs:431-435: panics unless `row < nrows` and `col < ncols`
    pub fn push(&mut self, row: usize, col: usize, value: T) {
        let st = unsafe { spl_coo_push(self.raw, row as u64, col as u64, &value as *const T as *const c_void) };
        self.check(st);
    }

    /// src/coo.rs:450-452
    pub fn pop(&mut self) -> Option<(usize, usize, T)> {
        let n = self.length();
        if n == 0 {
            return None;
        }
        let last = self.get(n - 1).map(|(r, c, v)| (*r, *c, *v));
        let st = unsafe { spl_coo_truncate(self.raw, (n - 1) as u64) };
        self.check(st);
        last
    }

    /// src/coo.rs:470-472
    pub fn clear(&mut self) {
        let st = unsafe { spl_coo_truncate(self.raw, 0) };
        self.check(st);
    }

    /// src/coo.rs:491-495
    pub fn iter(&self) -> impl Iterator<Item = (&usize, &usize, &T)> + '_ {
        let (r, c, v) = self.host();
        (0..self.length()).map(move |i| unsafe { (&*r.add(i), &*c.add(i), &*v.add(i)) })
    }

    /// src/coo.rs:538-545
    pub fn transpose(&self) -> Self {
        let mut t = Self::with_capacity(self.ncols, self.nrows, self.length());
        for (r, c, v) in self.iter() {
            t.push(*c, *r, *v);
        }
        t
    }

    pub(crate) fn raw(&self) -> *mut spl_coo { self.raw }

    fn check(&self, status: std::os::raw::c_int) {
        if status != SPL_OK {
            let msg = unsafe { std::ffi::CStr::from_ptr(spl_coo_last_error(self.raw)) }.to_string_lossy().into_owned();
            panic!("{msg}");
        }
    }
}

impl<T: Scalar> Drop for CooMatrix<T> {
    fn drop(&mut self) {
        unsafe { spl_coo_free(self.raw) };
    }
}

/// src/coo.rs:566-573: every entry is asserted before any is stored
impl<T: Scalar> Extend<(usize, usize, T)> for CooMatrix<T> {
    fn extend<I: IntoIterator<Item = (usize, usize, T)>>(&mut self, iter: I) {
        let (mut rows, mut cols, mut vals) = (Vec::new(), Vec::new(), Vec::new());
        for (r, c, v) in iter {
            rows.push(r);
            cols.push(c);
            vals.push(v);
        }
        let st = unsafe {
            spl_coo_extend(self.raw, vals.len() as u64, rows.as_ptr() as *const u64, cols.as_ptr() as *const u64,
                           vals.as_ptr() as *const c_void)
        };
        self.check(st);
    }
}

/// src/coo.rs:629-705: storage-order expansion of the compressed arrays (`spl_mat_to_coo`)
impl<T: Scalar> From<&CsrMatrix<T>> for CooMatrix<T> {
    fn from(m: &CsrMatrix<T>) -> Self {
        let (rows, cols, vals) = m.inner().to_triplets();
        CooMatrix::with_triplets(m.nrows(), m.ncols(), &rows, &cols, &vals)
    }
}

impl<T: Scalar> From<&CscMatrix<T>> for CooMatrix<T> {
    fn from(m: &CscMatrix<T>) -> Self {
        let (rows, cols, vals) = m.inner().to_triplets();
        CooMatrix::with_triplets(m.nrows(), m.ncols(), &rows, &cols, &vals)
    }
}

/// src/coo.rs:729-749
impl<T: Scalar> From<&DokMatrix<T>> for CooMatrix<T> {
    fn from(dok: &DokMatrix<T>) -> Self {
        CooMatrix::with_entries(dok.nrows(), dok.ncols(), dok.iter().map(|(r, c, v)| (r, c, *v)))
    }
}
