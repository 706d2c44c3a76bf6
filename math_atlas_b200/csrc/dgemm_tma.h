// Fast DGEMM path (TMA + DMMA).  See dgemm_tma.cu.
#pragma once
#include <cstddef>
namespace ab {
// true if the TMA path can take this problem (16-byte aligned bases, even leading dimensions).
bool dgemm_tma_eligible(int M, int N, int K, const double* A, size_t lda, const double* B, size_t ldb);
// C(cmode triangle) = alpha*op(A)*op(B) + beta*C on device pointers; ta/tb = 1 for transposed operands.
// Returns false (nothing launched) when not eligible.
bool dgemm_tma(int ta, int tb, int M, int N, int K, double alpha, const double* A, size_t lda, const double* B,
               size_t ldb, double beta, double* C, size_t ldc, int cmode, const int* abort_flag = nullptr);
// Size of the persistent DGEMM grid for the calling thread: 0 = one CTA per SM (default); a
// smaller cap leaves SMs free for a kernel running beside the update (LU/Cholesky look-ahead).
void set_gemm_cta_cap(int cap);
// SMs (bit i of mask4 = SM i) the calling thread's DGEMM launches must stay off, nullptr = none.
// Used by the LU/Cholesky look-ahead so that the panel's thread-block cluster finds a free GPC.
void set_gemm_reserved(const unsigned long long* mask4);
// Split-K for small problems is applied only while this is on (the cblas / F77 / ATL GEMM entry points turn it on around
// their launch); see pick_shape in dgemm_tma.cu.
void set_gemm_splitk_allowed(bool on);
int gemm_cta_cap();
}  // namespace ab
