//! Sparse linear algebra library: the crate's public API with the data-parallel hot path on the
//! GPU (`libspalinalg_b200.so`, C ABI in `include/spl.h`).  Same types, methods, panics and results
//! as spalinalg v0.0.2; CSR/CSC matrices live in device memory and materialise host mirrors on
//! first use of `rowptr()` / `colind()` / `values()`.

pub mod coo;
pub mod csc;
pub mod csr;
pub mod dok;
pub mod pinned;
pub mod scalar;

mod compressed;
mod ctx;
mod ffi;

pub use coo::CooMatrix;
pub use csc::CscMatrix;
pub use csr::CsrMatrix;
pub use dok::DokMatrix;
pub use pinned::PinnedVec;
pub use scalar::Scalar;
