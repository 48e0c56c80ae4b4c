//! `CscMatrix<T>` (reference: src/csc.rs:65-72 and src/csc/conv/*, src/csc/ops/*): the column-major twin of
//! `CsrMatrix<T>`, on the same device handle type.
use std::ops::{Add, Mul, Neg, Sub};

use crate::compressed::Compressed;
use crate::coo::CooMatrix;
use crate::csr::CsrMatrix;
use crate::dok::DokMatrix;
use crate::ffi::*;
use crate::scalar::Scalar;

pub struct CscMatrix<T: Scalar>(Compressed<T>);

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

    /// src/csc.rs:303-316: (row, col, value) in storage order (column by column)
    pub fn iter(&self) -> impl Iterator<Item = (usize, usize, &T)> + '_ {
        let h = self.0.host();
        (0..self.0.ncols).flat_map(move |c| (h.ptr[c]..h.ptr[c + 1]).map(move |p| (h.ind[p], c, &h.val[p])))
    }

    /// src/csc.rs:358-406
    pub fn transpose(&self) -> Self { CscMatrix(self.0.unary(spl_mat_transpose)) }

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
