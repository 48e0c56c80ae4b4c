//! `CscMatrix<T>` (reference: src/csc.rs:65-72 and src/csc/conv/*, src/csc/ops/*): the column-major twin of
//! `CsrMatrix<T>`, on the same device handle type.
use std::ops::{Add, Mul, Neg, Sub};

use crate::compressed::Compressed;
use crate::coo::CooMatrix;
use crate::csr::CsrMatrix;
use crate::dok::DokMatrix;
use crate::ffi::*;
use crate::scalar::Scalar;

#[derive(Debug)]
pub struct CscMatrix<T: Scalar>(Compressed<T>);

/// Immutable entries iterator created by [`CscMatrix::iter`] (src/csc.rs:81-85, 442-448): storage order.
#[derive(Clone, Debug)]
pub struct Iter<'iter, T> {
    ptr: &'iter [usize],
    ind: &'iter [usize],
    val: &'iter [T],
    major: usize,
    pos: usize,
}

/// Mutable entries iterator created by [`CscMatrix::iter_mut`] (src/csc.rs:87-91, 450-456).
#[derive(Debug)]
pub struct IterMut<'iter, T> {
    ptr: &'iter [usize],
    ind: &'iter [usize],
    val: std::slice::IterMut<'iter, T>,
    major: usize,
    pos: usize,
}

/// Move entries iterator created by `CscMatrix::into_iter` (src/csc.rs:93-96, 458-464): the matrix is read
/// from the device chunk by chunk (`spl_mat_read_entries`), never as one host copy of nnz tuples.
#[derive(Debug)]
pub struct IntoIter<T: Scalar> {
    matrix: CscMatrix<T>,
    next: usize,
    chunk: std::vec::IntoIter<(usize, usize, T)>,
}

const ITER_CHUNK: usize = 1 << 16;

impl<T: Scalar> CscMatrix<T> {
    /// src/csc.rs:137-164 — panics like the reference on the first failing assertion
    pub fn new(nrows: usize, ncols: usize, colptr: Vec<usize>, rowind: Vec<usize>, values: Vec<T>) -> Self {
        CscMatrix(Compressed::new(SPL_CSC, nrows, ncols, colptr, rowind, values))
    }
    /// src/csc.rs:179-188
    pub fn eye(size: usize) -> Self { CscMatrix(Compressed::eye(SPL_CSC, size)) }

    pub fn nrows(&self) -> usize { self.0.nrows }
    pub fn ncols(&self) -> usize { self.0.ncols }
    pub fn shape(&self) -> (usize, usize) { (self.0.nrows, self.0.ncols) }
    /// src/csc.rs:287-289
    pub fn nnz(&self) -> usize { self.0.nnz }
    /// src/csc.rs:228-258
    pub fn colptr(&self) -> &[usize] { &self.0.host().ptr }
    pub fn rowind(&self) -> &[usize] { &self.0.host().ind }
    pub fn values(&self) -> &[T] { &self.0.host().val }
    /// src/csc.rs:270-272
    pub fn values_mut(&mut self) -> &mut [T] { self.0.values_mut() }

    /// src/csc.rs:303-316: (row, col, value) in storage order
    pub fn iter(&self) -> Iter<T> {
        let h = self.0.host();
        Iter { ptr: &h.ptr, ind: &h.ind, val: &h.val, major: 0, pos: 0 }
    }

    /// src/csc.rs:330-343 (the reference loops over the wrong dimension on non-square matrices,
    /// SURVEY.md appendix A; this walks every stored entry).  The values are written back to the device
    /// before the next device operation, like `values_mut`.
    pub fn iter_mut(&mut self) -> IterMut<T> {
        self.0.values_mut();                                   // marks the mirror dirty, downloads it if needed
        let (ptr, ind, val) = self.0.host_parts_mut();
        IterMut { ptr, ind, val: val.iter_mut(), major: 0, pos: 0 }
    }

    /// src/csc.rs:358-406
    pub fn transpose(&self) -> Self { CscMatrix(self.0.unary(spl_mat_transpose)) }

    /// Extension: y = A x with dense host vectors (`&A * &X`, X n x 1, src/csc/ops/mul.rs:5-61); the first
    /// product builds the CSR form of the matrix on the device and keeps it.
    pub fn matvec(&self, x: &[T]) -> Vec<T> { self.0.matvec(x) }
    /// `matvec` into `y`; pipelined when both are `PinnedVec`s.
    pub fn matvec_into(&self, x: &[T], y: &mut [T]) { self.0.matvec_into(x, y) }

    pub(crate) fn inner(&self) -> &Compressed<T> { &self.0 }
    pub(crate) fn wrap(c: Compressed<T>) -> Self { CscMatrix(c) }
}

/// src/csc/conv/coo.rs:3-116 (owned form :118-122)
impl<T: Scalar> From<&CooMatrix<T>> for CscMatrix<T> {
    fn from(coo: &CooMatrix<T>) -> Self { CscMatrix(Compressed::from_coo(SPL_CSC, coo)) }
}
impl<T: Scalar> From<CooMatrix<T>> for CscMatrix<T> {
    fn from(coo: CooMatrix<T>) -> Self { Self::from(&coo) }
}
/// src/csc/conv/dok.rs:3-76 (:78-82)
impl<T: Scalar> From<&DokMatrix<T>> for CscMatrix<T> {
    fn from(dok: &DokMatrix<T>) -> Self { CscMatrix(Compressed::from_dok(SPL_CSC, dok)) }
}
impl<T: Scalar> From<DokMatrix<T>> for CscMatrix<T> {
    fn from(dok: DokMatrix<T>) -> Self { Self::from(&dok) }
}
/// src/csc/conv/csr.rs:3-53 (:55-59)
impl<T: Scalar> From<&CsrMatrix<T>> for CscMatrix<T> {
    fn from(csc: &CsrMatrix<T>) -> Self { CscMatrix(csc.inner().convert(SPL_CSC)) }
}
impl<T: Scalar> From<CsrMatrix<T>> for CscMatrix<T> {
    fn from(csc: CsrMatrix<T>) -> Self { Self::from(&csc) }
}

/// src/csc/ops/add.rs:5-70 — shapes asserted equal (SPL_ERR_SHAPE -> panic)
impl<T: Scalar> Add for &CscMatrix<T> {
    type Output = CscMatrix<T>;
    fn add(self, rhs: Self) -> Self::Output { CscMatrix(self.0.binary(&rhs.0, spl_mat_add)) }
}
/// src/csc/ops/sub.rs:5-70
impl<T: Scalar> Sub for &CscMatrix<T> {
    type Output = CscMatrix<T>;
    fn sub(self, rhs: Self) -> Self::Output { CscMatrix(self.0.binary(&rhs.0, spl_mat_sub)) }
}
/// src/csc/ops/mul.rs:5-61 — `self.ncols() == rhs.nrows()` asserted
impl<T: Scalar> Mul for &CscMatrix<T> {
    type Output = CscMatrix<T>;
    fn mul(self, rhs: Self) -> Self::Output { CscMatrix(self.0.binary(&rhs.0, spl_mat_mul)) }
}
/// src/csc/ops/neg.rs:5-18
impl<T: Scalar> Neg for &CscMatrix<T> {
    type Output = CscMatrix<T>;
    fn neg(self) -> Self::Output { CscMatrix(self.0.unary(spl_mat_neg)) }
}

/// src/csc.rs:409-440
impl<T: Scalar> IntoIterator for CscMatrix<T> {
    type Item = (usize, usize, T);
    type IntoIter = IntoIter<T>;
    fn into_iter(self) -> Self::IntoIter {
        IntoIter { matrix: self, next: 0, chunk: Vec::new().into_iter() }
    }
}

impl<'iter, T> Iterator for Iter<'iter, T> {
    type Item = (usize, usize, &'iter T);
    fn next(&mut self) -> Option<Self::Item> {
        if self.pos >= self.val.len() {
            return None;
        }
        while self.ptr[self.major + 1] <= self.pos {
            self.major += 1;                                   // skips empty cols
        }
        let p = self.pos;
        self.pos += 1;
        let (ind, val) = (self.ind, self.val);                 // the slices outlive the iterator borrow
        Some((ind[p], self.major, &val[p]))
    }
}

impl<'iter, T> Iterator for IterMut<'iter, T> {
    type Item = (usize, usize, &'iter mut T);
    fn next(&mut self) -> Option<Self::Item> {
        let v = self.val.next()?;
        while self.ptr[self.major + 1] <= self.pos {
            self.major += 1;
        }
        let p = self.pos;
        self.pos += 1;
        Some((self.ind[p], self.major, v))
    }
}

impl<T: Scalar> Iterator for IntoIter<T> {
    type Item = (usize, usize, T);
    fn next(&mut self) -> Option<Self::Item> {
        if let Some(e) = self.chunk.next() {
            return Some(e);
        }
        let nnz = self.matrix.nnz();
        if self.next >= nnz {
            return None;
        }
        let count = ITER_CHUNK.min(nnz - self.next);
        let (rows, cols, vals) = self.matrix.inner().read_entries(self.next, count);
        self.next += count;
        self.chunk = rows.into_iter().zip(cols).zip(vals).map(|((r, c), v)| (r, c, v)).collect::<Vec<_>>().into_iter();
        self.chunk.next()
    }
}
