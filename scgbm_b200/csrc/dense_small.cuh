// dense_small.cuh -- K6 interfaces (fp64 tall-skinny + small dense, all on device)
#pragma once
#include <algorithm>
#include "common.cuh"

namespace scgbm {

// grow-only scratch for the small dense kernels
struct SmallWs {
  DevBuf<double> partial;  // ts_gram stage-1 partials
  DevBuf<double> eigV;     // Jacobi eigenvector accumulator when not in smem
  DevBuf<double> small;    // p x b style scratch
  void need_partial(int64_t n) { if (partial.n < n) partial.alloc(n + n / 4); }
  void need_eigV(int64_t n) { if (eigV.n < n) eigV.alloc(n); }
  void need_small(int64_t n) { if (small.n < n) small.alloc(n + n / 4); }
};

// C[p x q] = beta*C + alpha * A^T B      (A: n x p, B: n x q; deterministic)
void ts_gram(Ctx& ctx, SmallWs& ws, int64_t n, int p, int q, const double* A, int64_t lda,
             const double* B, int64_t ldb, double* C, int ldc, double alpha, double beta);

// C[n x q] = beta*C + alpha * A[n x p] S[p x q]
void ts_mul(Ctx& ctx, int64_t n, int p, int q, const double* A, int64_t lda, const double* S,
            int lds, double* C, int64_t ldc, double alpha, double beta);

// symmetric eigendecomposition, eigenvalues descending; A destroyed
void eigh_desc(Ctx& ctx, SmallWs& ws, int n, double* A, double* w, double* Vout);

void copy_mat(Ctx& ctx, int64_t n, int q, const double* src, int64_t lds, double* dst, int64_t ldd);

// K <- orthonormal basis of span(K) minus span(Qb); rank-deficient directions -> 0
void sym_orth(Ctx& ctx, SmallWs& ws, int64_t n, int b, double* K, int64_t ldk, const double* Qb,
              int64_t ldq, int pb, double* tmp_nb, bool robust, int* chol_flag);

}  // namespace scgbm
