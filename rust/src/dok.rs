//! `DokMatrix<T>` (reference: src/dok.rs:53-58): unordered map of unique keys, a host-side builder.
//! Its conversions into CSR/CSC are the assembly kernels with `dedup = 0, dropzero = 0`
//! (src/csr/conv/dok.rs:3-76: no zero drop — explicit zeros survive).
use std::collections::HashMap;

use crate::coo::CooMatrix;
use crate::scalar::Scalar;

pub struct DokMatrix<T: Scalar> {
    nrows: usize,
    ncols: usize,
    entries: HashMap<(usize, usize), T>,
}

impl<T: Scalar> DokMatrix<T> {
    /// src/dok.rs:105-113
    pub fn new(nrows: usize, ncols: usize) -> Self {
        assert!(nrows > 0);
        assert!(ncols > 0);
        DokMatrix { nrows, ncols, entries: HashMap::new() }
    }

    pub fn with_entries<I: IntoIterator<Item = (usize, usize, T)>>(nrows: usize, ncols: usize, entries: I) -> Self {
        let mut m = Self::new(nrows, ncols);
        for (r, c, v) in entries {
            m.insert(r, c, v);
        }
        m
    }

    pub fn nrows(&self) -> usize { self.nrows }
    pub fn ncols(&self) -> usize { self.ncols }
    pub fn length(&self) -> usize { self.entries.len() }
    pub fn get(&self, row: usize, col: usize) -> Option<&T> { self.entries.get(&(row, col)) }

    /// src/dok.rs:462-466
    pub fn insert(&mut self, row: usize, col: usize, value: T) -> Option<T> {
        assert!(row < self.nrows);
        assert!(col < self.ncols);
        self.entries.insert((row, col), value)
    }

    pub fn iter(&self) -> impl Iterator<Item = (usize, usize, &T)> + '_ {
        self.entries.iter().map(|(&(r, c), v)| (r, c, v))
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

/// src/dok.rs:640-668: `*entry.or_default() += value` in insertion order; no zero drop
impl<T: Scalar> From<&CooMatrix<T>> for DokMatrix<T> {
    fn from(coo: &CooMatrix<T>) -> Self {
        let mut m = DokMatrix::new(coo.nrows(), coo.ncols());
        for (r, c, v) in coo.iter() {
            *m.entries.entry((*r, *c)).or_default() += *v;
        }
        m
    }
}
