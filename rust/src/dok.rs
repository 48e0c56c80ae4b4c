//! `DokMatrix<T>` (reference: src/dok.rs:53-58): unordered map of unique keys, a host-side builder.
//! Its conversions into CSR/CSC are the assembly kernels with `dedup = 0, dropzero = 0`
//! (src/csr/conv/dok.rs:3-76: no zero drop — explicit zeros survive); everything else is O(len) host
//! work on the `HashMap`, as in the reference.
use std::collections::HashMap;
use std::ops::{Add, AddAssign, Neg, Sub, SubAssign};

use crate::coo::CooMatrix;
use crate::csc::CscMatrix;
use crate::csr::CsrMatrix;
use crate::scalar::Scalar;

#[derive(Clone, Debug)]
pub struct DokMatrix<T: Scalar> {
    nrows: usize,
    ncols: usize,
    entries: HashMap<(usize, usize), T>,
}

/// Immutable entries iterator created by [`DokMatrix::iter`] (src/dok.rs:61-65, 616-622).
#[derive(Clone, Debug)]
pub struct Iter<'iter, T> {
    iter: std::collections::hash_map::Iter<'iter, (usize, usize), T>,
}

/// Mutable entries iterator created by [`DokMatrix::iter_mut`] (src/dok.rs:67-71, 624-630).
#[derive(Debug)]
pub struct IterMut<'iter, T> {
    iter: std::collections::hash_map::IterMut<'iter, (usize, usize), T>,
}

/// Move entries iterator created by `DokMatrix::into_iter` (src/dok.rs:73-76, 632-638).
#[derive(Debug)]
pub struct IntoIter<T> {
    iter: std::collections::hash_map::IntoIter<(usize, usize), T>,
}

impl<T: Scalar> DokMatrix<T> {
    /// src/dok.rs:105-113
    pub fn new(nrows: usize, ncols: usize) -> Self {
        assert!(nrows > 0);
        assert!(ncols > 0);
        DokMatrix { nrows, ncols, entries: HashMap::new() }
    }

    /// src/dok.rs:128-135
    pub fn eye(size: usize) -> Self {
        assert!(size > 0);
        DokMatrix { nrows: size, ncols: size, entries: (0..size).map(|i| ((i, i), T::one())).collect() }
    }

    /// src/dok.rs:163-171
    pub fn with_capacity(nrows: usize, ncols: usize, capacity: usize) -> Self {
        assert!(nrows > 0);
        assert!(ncols > 0);
        DokMatrix { nrows, ncols, entries: HashMap::with_capacity(capacity) }
    }

    /// src/dok.rs:205-221: later entries of a cell replace earlier ones
    pub fn with_entries<I>(nrows: usize, ncols: usize, entries: I) -> Self
    where
        I: IntoIterator<Item = (usize, usize, T)>,
    {
        assert!(nrows > 0);
        assert!(ncols > 0);
        let entries: Vec<_> = entries.into_iter().collect();
        for (row, col, _) in &entries {
            assert!(*row < nrows);
            assert!(*col < ncols);
        }
        DokMatrix { nrows, ncols, entries: entries.into_iter().map(|(r, c, v)| ((r, c), v)).collect() }
    }

    /// src/dok.rs:255-289
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
        for row in rowind.iter() {
            assert!(*row < nrows);
        }
        for col in colind.iter() {
            assert!(*col < ncols);
        }
        let mut entries = HashMap::with_capacity(values.len());
        for (idx, value) in values.into_iter().enumerate() {
            entries.insert((rowind[idx], colind[idx]), value);
        }
        DokMatrix { nrows, ncols, entries }
    }

    pub fn nrows(&self) -> usize { self.nrows }
    pub fn ncols(&self) -> usize { self.ncols }
    pub fn shape(&self) -> (usize, usize) { (self.nrows, self.ncols) }
    /// src/dok.rs:350-352
    pub fn length(&self) -> usize { self.entries.len() }
    /// src/dok.rs:367-369
    pub fn capacity(&self) -> usize { self.entries.capacity() }
    /// src/dok.rs:393-395
    pub fn contains(&self, row: usize, col: usize) -> bool { self.entries.contains_key(&(row, col)) }
    /// src/dok.rs:416-418
    pub fn get(&self, row: usize, col: usize) -> Option<&T> { self.entries.get(&(row, col)) }
    /// src/dok.rs:439-441
    pub fn get_mut(&mut self, row: usize, col: usize) -> Option<&mut T> { self.entries.get_mut(&(row, col)) }

    /// src/dok.rs:462-466
    pub fn insert(&mut self, row: usize, col: usize, value: T) -> Option<T> {
        assert!(row < self.nrows);
        assert!(col < self.ncols);
        self.entries.insert((row, col), value)
    }

    /// src/dok.rs:484-486
    pub fn clear(&mut self) { self.entries.clear() }

    /// src/dok.rs:503-507
    pub fn iter(&self) -> Iter<T> { Iter { iter: self.entries.iter() } }

    /// src/dok.rs:524-528
    pub fn iter_mut(&mut self) -> IterMut<T> { IterMut { iter: self.entries.iter_mut() } }

    /// src/dok.rs:547-558
    pub fn transpose(&self) -> Self {
        DokMatrix { nrows: self.ncols, ncols: self.nrows,
                    entries: self.entries.iter().map(|(&(r, c), &v)| ((c, r), v)).collect() }
    }

    /// SoA dump for the assembly call (the result does not depend on the map's iteration order).
    pub(crate) fn triplets(&self) -> (Vec<usize>, Vec<usize>, Vec<T>) {
        let n = self.entries.len();
        let (mut rows, mut cols, mut vals) = (Vec::with_capacity(n), Vec::with_capacity(n), Vec::with_capacity(n));
        for (&(r, c), &v) in &self.entries {
            rows.push(r);
            cols.push(c);
            vals.push(v);
        }
        (rows, cols, vals)
    }
}

/// src/dok.rs:561-587: every entry is asserted before any is stored
impl<T: Scalar> Extend<(usize, usize, T)> for DokMatrix<T> {
    fn extend<I: IntoIterator<Item = (usize, usize, T)>>(&mut self, iter: I) {
        let entries: Vec<_> = iter.into_iter().collect();
        for (row, col, _) in &entries {
            assert!(*row < self.nrows);
            assert!(*col < self.ncols);
        }
        self.entries.extend(entries.into_iter().map(|(r, c, v)| ((r, c), v)));
    }
}

/// src/dok.rs:589-614
impl<T: Scalar> IntoIterator for DokMatrix<T> {
    type Item = (usize, usize, T);
    type IntoIter = IntoIter<T>;
    fn into_iter(self) -> Self::IntoIter { IntoIter { iter: self.entries.into_iter() } }
}

impl<'iter, T> Iterator for Iter<'iter, T> {
    type Item = (usize, usize, &'iter T);
    fn next(&mut self) -> Option<Self::Item> { self.iter.next().map(|((r, c), v)| (*r, *c, v)) }
}

impl<'iter, T> Iterator for IterMut<'iter, T> {
    type Item = (usize, usize, &'iter mut T);
    fn next(&mut self) -> Option<Self::Item> { self.iter.next().map(|((r, c), v)| (*r, *c, v)) }
}

impl<T: Scalar> Iterator for IntoIter<T> {
    type Item = (usize, usize, T);
    fn next(&mut self) -> Option<Self::Item> { self.iter.next().map(|((r, c), v)| (r, c, v)) }
}

/// src/dok.rs:640-668: `*entry.or_default() += value` in insertion order; no zero drop
impl<T: Scalar> From<&CooMatrix<T>> for DokMatrix<T> {
    fn from(coo: &CooMatrix<T>) -> Self {
        let mut map = HashMap::with_capacity(coo.length());
        for (row, col, value) in coo.iter() {
            *map.entry((row, col)).or_default() += *value;
        }
        DokMatrix { nrows: coo.nrows(), ncols: coo.ncols(), entries: map }
    }
}
impl<T: Scalar> From<CooMatrix<T>> for DokMatrix<T> {
    fn from(coo: CooMatrix<T>) -> Self { Self::from(&coo) }
}

/// src/dok.rs:676-700: the stored entries of the compressed matrix, read from the device in chunks
impl<T: Scalar> From<&CscMatrix<T>> for DokMatrix<T> {
    fn from(csc: &CscMatrix<T>) -> Self {
        DokMatrix { nrows: csc.nrows(), ncols: csc.ncols(), entries: csc.iter().map(|(r, c, &v)| ((r, c), v)).collect() }
    }
}
impl<T: Scalar> From<CscMatrix<T>> for DokMatrix<T> {
    fn from(csc: CscMatrix<T>) -> Self { Self::from(&csc) }
}
/// src/dok.rs:702-720, 771-775
impl<T: Scalar> From<&CsrMatrix<T>> for DokMatrix<T> {
    fn from(csr: &CsrMatrix<T>) -> Self {
        DokMatrix { nrows: csr.nrows(), ncols: csr.ncols(), entries: csr.iter().map(|(r, c, &v)| ((r, c), v)).collect() }
    }
}
impl<T: Scalar> From<CsrMatrix<T>> for DokMatrix<T> {
    fn from(csr: CsrMatrix<T>) -> Self { Self::from(&csr) }
}

/// src/dok.rs:722-736: lhs entries, then `or_default() += rhs` (dims of lhs, no shape assertion)
impl<T: Scalar> Add for &DokMatrix<T> {
    type Output = DokMatrix<T>;
    fn add(self, rhs: Self) -> Self::Output {
        let mut entries = self.entries.clone();
        for (&(row, col), &val) in rhs.entries.iter() {
            entries.entry((row, col)).or_default().add_assign(val);
        }
        DokMatrix { nrows: self.nrows, ncols: self.ncols, entries }
    }
}

/// src/dok.rs:738-752
impl<T: Scalar> Sub for &DokMatrix<T> {
    type Output = DokMatrix<T>;
    fn sub(self, rhs: Self) -> Self::Output {
        let mut entries = self.entries.clone();
        for (&(row, col), &val) in rhs.entries.iter() {
            entries.entry((row, col)).or_default().sub_assign(val);
        }
        DokMatrix { nrows: self.nrows, ncols: self.ncols, entries }
    }
}

/// src/dok.rs:754-769
impl<T: Scalar> Neg for &DokMatrix<T> {
    type Output = DokMatrix<T>;
    fn neg(self) -> Self::Output {
        DokMatrix { nrows: self.nrows, ncols: self.ncols,
                    entries: self.entries.iter().map(|(&(row, col), &val)| ((row, col), -val)).collect() }
    }
}
