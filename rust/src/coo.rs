//! `CooMatrix<T>` (reference: src/coo.rs:52-57): insertion-ordered triplets, duplicates allowed.
//! Storage is an `spl_coo`: three pinned host arrays (row, col, value) owned by the library and
//! streamed to the device chunk by chunk while they are filled, so `CsrMatrix::from(&coo)` finds the
//! triplets already in HBM.  Accessors hand out references into the pinned arrays, exactly as the
//! reference hands out references into its `Vec<(usize, usize, T)>`.
use std::fmt;
use std::marker::PhantomData;
use std::ops::{Add, Neg, Sub};
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

// The storage is plain memory behind a unique handle: `&mut self` guards every mutation of the host
// arrays, and the one operation that takes `&self` and touches shared state — the conversion's flush
// of the last partial chunk — is serialised inside the library (a mutex per builder), so two threads
// may convert the same `&CooMatrix` at once, as with the reference's `Vec`-backed type.
unsafe impl<T: Scalar> Send for CooMatrix<T> {}
unsafe impl<T: Scalar> Sync for CooMatrix<T> {}

/// Immutable entries iterator created by [`CooMatrix::iter`] (src/coo.rs:60-64, 605-611).
#[derive(Clone, Debug)]
pub struct Iter<'iter, T> {
    rows: std::slice::Iter<'iter, usize>,
    cols: std::slice::Iter<'iter, usize>,
    vals: std::slice::Iter<'iter, T>,
}

/// Mutable entries iterator created by [`CooMatrix::iter_mut`] (src/coo.rs:66-70, 613-619).
#[derive(Debug)]
pub struct IterMut<'iter, T> {
    rows: std::slice::Iter<'iter, usize>,
    cols: std::slice::Iter<'iter, usize>,
    vals: std::slice::IterMut<'iter, T>,
}

/// Move entries iterator created by `CooMatrix::into_iter` (src/coo.rs:72-75, 621-627).
#[derive(Debug)]
pub struct IntoIter<T> {
    iter: std::vec::IntoIter<(usize, usize, T)>,
}

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
        let idx: Vec<usize> = (0..size).collect();
        Self::from_slices(size, size, &idx, &idx, &vec![T::one(); size])
    }

    /// src/coo.rs:204-220
    pub fn with_entries<I>(nrows: usize, ncols: usize, entries: I) -> Self
    where
        I: IntoIterator<Item = (usize, usize, T)>,
    {
        let mut m = Self::new(nrows, ncols);
        m.extend(entries);
        m
    }

    /// src/coo.rs:254-288 — any three iterables, as in the reference
    pub fn with_triplets<R, C, V>(nrows: usize, ncols: usize, rowind: R, colind: C, values: V) -> Self
    where
        R: IntoIterator<Item = usize>,
        C: IntoIterator<Item = usize>,
        V: IntoIterator<Item = T>,
    {
        assert!(nrows > 0);
        assert!(ncols > 0);
        let rowind: Vec<_> = rowind.into_iter().collect();
        let colind: Vec<_> = colind.into_iter().collect();
        let values: Vec<_> = values.into_iter().collect();
        assert!(rowind.len() == values.len());
        assert!(colind.len() == values.len());
        Self::from_slices(nrows, ncols, &rowind, &colind, &values)
    }

    /// Bulk form: the slices go to the library as they are (usize is 64 bit on every supported target);
    /// `spl_coo_extend` asserts every index before it stores any.
    pub(crate) fn from_slices(nrows: usize, ncols: usize, rowind: &[usize], colind: &[usize], values: &[T]) -> Self {
        let m = Self::with_capacity(nrows, ncols, values.len());
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

    /// The three pinned arrays as slices of the stored length.
    fn slices(&self) -> (&[usize], &[usize], &[T]) {
        let n = self.length();
        if n == 0 {
            return (&[], &[], &[]);
        }
        let (mut r, mut c, mut v) = (std::ptr::null(), std::ptr::null(), std::ptr::null());
        unsafe {
            spl_coo_host_ptrs(self.raw, &mut r, &mut c, &mut v);
            (std::slice::from_raw_parts(r as *const usize, n), std::slice::from_raw_parts(c as *const usize, n),
             std::slice::from_raw_parts(v as *const T, n))
        }
    }

    /// Values writable: what the copy stream already took from `first` on is sent again at the next
    /// conversion (`spl_coo_invalidate`).
    fn slices_mut(&mut self, first: usize) -> (&[usize], &[usize], &mut [T]) {
        let n = self.length();
        if n == 0 {
            return (&[], &[], &mut []);
        }
        let st = unsafe { spl_coo_invalidate(self.raw, first as u64) };
        self.check(st);
        let (mut r, mut c, mut v) = (std::ptr::null(), std::ptr::null(), std::ptr::null());
        unsafe {
            spl_coo_host_ptrs(self.raw, &mut r, &mut c, &mut v);
            (std::slice::from_raw_parts(r as *const usize, n), std::slice::from_raw_parts(c as *const usize, n),
             std::slice::from_raw_parts_mut(v as *mut T, n))
        }
    }

    /// src/coo.rs:386-390
    pub fn get(&self, index: usize) -> Option<(&usize, &usize, &T)> {
        let (r, c, v) = self.slices();
        if index < v.len() { Some((&r[index], &c[index], &v[index])) } else { None }
    }

    /// src/coo.rs:408-412
    pub fn get_mut(&mut self, index: usize) -> Option<(&usize, &usize, &mut T)> {
        if index >= self.length() {
            return None;
        }
        let (r, c, v) = self.slices_mut(index);
        Some((&r[index], &c[index], &mut v[index]))
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
    pub fn iter(&self) -> Iter<T> {
        let (r, c, v) = self.slices();
        Iter { rows: r.iter(), cols: c.iter(), vals: v.iter() }
    }

    /// src/coo.rs:514-518
    pub fn iter_mut(&mut self) -> IterMut<T> {
        let (r, c, v) = self.slices_mut(0);
        IterMut { rows: r.iter(), cols: c.iter(), vals: v.iter_mut() }
    }

    /// src/coo.rs:538-545
    pub fn transpose(&self) -> Self {
        let (r, c, v) = self.slices();
        Self::from_slices(self.ncols, self.nrows, c, r, v)
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

/// `#[derive(Clone)]` in the reference (src/coo.rs:52): a second builder with the same entries.
impl<T: Scalar> Clone for CooMatrix<T> {
    fn clone(&self) -> Self {
        let (r, c, v) = self.slices();
        let m = Self::from_slices(self.nrows, self.ncols, r, c, v);
        let st = unsafe { spl_coo_reserve(m.raw, self.capacity() as u64) };
        m.check(st);
        m
    }
}

/// `#[derive(Debug)]` in the reference (src/coo.rs:52): same field names, entries as tuples.
impl<T: Scalar> fmt::Debug for CooMatrix<T> {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        let entries: Vec<(usize, usize, T)> = self.iter().map(|(r, c, v)| (r, c, *v)).collect();
        f.debug_struct("CooMatrix").field("nrows", &self.nrows).field("ncols", &self.ncols).field("entries", &entries).finish()
    }
}

/// src/coo.rs:548-574: every entry is asserted before any is stored
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

/// src/coo.rs:576-603
impl<T: Scalar> IntoIterator for CooMatrix<T> {
    type Item = (usize, usize, T);
    type IntoIter = IntoIter<T>;
    fn into_iter(self) -> Self::IntoIter {
        let entries: Vec<(usize, usize, T)> = self.iter().map(|(r, c, v)| (r, c, *v)).collect();
        IntoIter { iter: entries.into_iter() }
    }
}

impl<'iter, T> Iterator for Iter<'iter, T> {
    type Item = (usize, usize, &'iter T);
    fn next(&mut self) -> Option<Self::Item> {
        Some((*self.rows.next()?, *self.cols.next()?, self.vals.next()?))
    }
}

impl<'iter, T> Iterator for IterMut<'iter, T> {
    type Item = (usize, usize, &'iter mut T);
    fn next(&mut self) -> Option<Self::Item> {
        Some((*self.rows.next()?, *self.cols.next()?, self.vals.next()?))
    }
}

impl<T: Scalar> Iterator for IntoIter<T> {
    type Item = (usize, usize, T);
    fn next(&mut self) -> Option<Self::Item> {
        self.iter.next()
    }
}

/// src/coo.rs:629-705: storage-order expansion of the compressed arrays (`spl_mat_to_coo`)
impl<T: Scalar> From<&CsrMatrix<T>> for CooMatrix<T> {
    fn from(m: &CsrMatrix<T>) -> Self {
        let (rows, cols, vals) = m.inner().to_triplets();
        CooMatrix::from_slices(m.nrows(), m.ncols(), &rows, &cols, &vals)
    }
}
impl<T: Scalar> From<CsrMatrix<T>> for CooMatrix<T> {
    fn from(m: CsrMatrix<T>) -> Self { Self::from(&m) }
}
impl<T: Scalar> From<&CscMatrix<T>> for CooMatrix<T> {
    fn from(m: &CscMatrix<T>) -> Self {
        let (rows, cols, vals) = m.inner().to_triplets();
        CooMatrix::from_slices(m.nrows(), m.ncols(), &rows, &cols, &vals)
    }
}
impl<T: Scalar> From<CscMatrix<T>> for CooMatrix<T> {
    fn from(m: CscMatrix<T>) -> Self { Self::from(&m) }
}

/// src/coo.rs:707-749
impl<T: Scalar> From<&DokMatrix<T>> for CooMatrix<T> {
    fn from(dok: &DokMatrix<T>) -> Self {
        CooMatrix::with_entries(dok.nrows(), dok.ncols(), dok.iter().map(|(r, c, v)| (r, c, *v)))
    }
}
impl<T: Scalar> From<DokMatrix<T>> for CooMatrix<T> {
    fn from(dok: DokMatrix<T>) -> Self { Self::from(&dok) }
}

/// src/coo.rs:751-770: the entries of both operands, lhs first (duplicates are summed at conversion)
impl<T: Scalar> Add for &CooMatrix<T> {
    type Output = CooMatrix<T>;
    fn add(self, rhs: Self) -> Self::Output {
        assert_eq!(self.shape(), rhs.shape());
        let mut out = CooMatrix::with_capacity(self.nrows, self.ncols, self.length() + rhs.length());
        out.extend(self.iter().map(|(r, c, v)| (r, c, *v)));
        out.extend(rhs.iter().map(|(r, c, v)| (r, c, *v)));
        out
    }
}

/// src/coo.rs:772-791
impl<T: Scalar> Sub for &CooMatrix<T> {
    type Output = CooMatrix<T>;
    fn sub(self, rhs: Self) -> Self::Output {
        assert_eq!(self.shape(), rhs.shape());
        let mut out = CooMatrix::with_capacity(self.nrows, self.ncols, self.length() + rhs.length());
        out.extend(self.iter().map(|(r, c, v)| (r, c, *v)));
        out.extend(rhs.iter().map(|(r, c, v)| (r, c, -*v)));
        out
    }
}

/// src/coo.rs:793-804
impl<T: Scalar> Neg for &CooMatrix<T> {
    type Output = CooMatrix<T>;
    fn neg(self) -> Self::Output {
        CooMatrix::with_entries(self.nrows, self.ncols, self.iter().map(|(r, c, v)| (r, c, -*v)))
    }
}
