/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see orc_impl.inc).  Parity is PINNED:
 * tests/test_oracle.py checks every function against all hot-path golden
 * vectors of the reference's own unit tests and doctests,
 * committed under tests/golden/reference_goldens.json with file:line cites.
 * The reference itself (Rust) cannot be built in this image (no rustc/cargo),
 * so oracle/_ref does not exist; bench.py reports cpu_baseline.kind = "port".
 *
 * Build: make -C oracle   (gcc -O3 -ffp-contract=off -fno-fast-math)
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define T float
#define FN(x) x##_f32
#include "orc_impl.inc"
#undef T
#undef FN

#define T double
#define FN(x) x##_f64
#include "orc_impl.inc"
#undef T
#undef FN

/*
 * CsrMatrix::new / CscMatrix::new validation, src/csr.rs:144-156 and
 * src/csc.rs:144-156.  Returns 0 when the reference would construct the
 * matrix, else the 1-based ordinal of the first assertion that panics:
 *  1 nmajor/nrows>0  2 ncols>0  3 ptr.len()==n+1  4 ptr[0]==0
 *  5 ind.len()==ptr[n]  6 val.len()==ptr[n]  7 ptr non-decreasing
 *  8 indices in range  9 indices strictly increasing inside a segment
 * (nrows/ncols are the matrix dims; major_is_row picks CSR or CSC.)
 */
int orc_validate_compressed(int major_is_row, size_t nrows, size_t ncols,
                            size_t ptr_len, const size_t *ptr,
                            size_t ind_len, const size_t *ind, size_t val_len)
{
    size_t nmajor = major_is_row ? nrows : ncols;
    size_t nminor = major_is_row ? ncols : nrows;
    if (!(nrows > 0)) return 1;
    if (!(ncols > 0)) return 2;
    if (ptr_len != nmajor + 1) return 3;
    if (ptr[0] != 0) return 4;
    if (ind_len != ptr[nmajor]) return 5;
    if (val_len != ptr[nmajor]) return 6;
    for (size_t m = 0; m < nmajor; ++m) if (!(ptr[m] <= ptr[m + 1])) return 7;
    for (size_t p = 0; p < ind_len; ++p) if (!(ind[p] < nminor)) return 8;
    for (size_t m = 0; m < nmajor; ++m)
        for (size_t p = ptr[m]; p + 1 < ptr[m + 1]; ++p)
            if (!(ind[p] < ind[p + 1])) return 9;
    return 0;
}
