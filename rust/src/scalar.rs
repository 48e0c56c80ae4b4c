//! `Scalar` (reference: src/scalar.rs:55-57): implemented for `f32` and `f64` only; the shim adds
//! the dtype tag the ABI dispatches on.
use std::fmt::Debug;
use std::ops::{Add, AddAssign, Mul, Neg, Sub, SubAssign};
use std::os::raw::c_int;

use crate::ffi::{SPL_F32, SPL_F64};

pub trait Scalar:
    Copy + Debug + Default + PartialEq + Send + Sync + 'static
    + Add<Output = Self> + Sub<Output = Self> + Mul<Output = Self> + Neg<Output = Self>
    + AddAssign + SubAssign
{
    const DTYPE: c_int;
    fn zero() -> Self;
    fn one() -> Self;
}

impl Scalar for f32 {
    const DTYPE: c_int = SPL_F32;
    fn zero() -> Self { 0.0 }
    fn one() -> Self { 1.0 }
}

impl Scalar for f64 {
    const DTYPE: c_int = SPL_F64;
    fn zero() -> Self { 0.0 }
    fn one() -> Self { 1.0 }
}
