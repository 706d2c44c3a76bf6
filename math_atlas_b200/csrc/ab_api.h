// Device-level (all pointers are device pointers, work is enqueued on ab::current_stream(),
// nothing synchronises) level-3 BLAS and LAPACK drivers.  The C ABI in capi.cu wraps these with
// argument checking, host<->device staging and the final stream synchronisation.
// T is float, double, float2 (c) or double2 (z).  Enum values are the CBLAS ones (common.cuh).
#pragma once
#include "common.cuh"

namespace ab {

// C <- alpha*op(A)*op(B) + beta*C         (reference ATL_Xgemm, Clint/atlas-ammm.base:17707-17806)
// abort_flag: optional device int; when non-zero at kernel start the update is skipped.
template <typename T>
void gemm(int TA, int TB, int M, int N, int K, T alpha, const T* A, size_t lda, const T* B, size_t ldb, T beta, T* C,
          size_t ldc, const int* abort_flag = nullptr);

// C(uplo) <- alpha*op(A)*op(A)^T + beta*C  (reference ATL_Xsyrk, Clint/l3micro.base:2688)
template <typename T>
void syrk(int uplo, int trans, int N, int K, T alpha, const T* A, size_t lda, T beta, T* C, size_t ldc,
          const int* abort_flag = nullptr);

// C(uplo) <- alpha*op(A)*op(A)^H + beta*C, alpha/beta real, Im(diag)=0 (reference ATL_Xherk, -DConj_)
template <typename T>
void herk(int uplo, int trans, int N, int K, typename ST<T>::real alpha, const T* A, size_t lda,
          typename ST<T>::real beta, T* C, size_t ldc, const int* abort_flag = nullptr);

// B <- alpha*op(A)*B or alpha*B*op(A), A triangular   (reference ATL_Xtrmm, Clint/trmm-micro.base:1942-2157)
template <typename T>
void trmm(int side, int uplo, int trans, int diag, int M, int N, T alpha, const T* A, size_t lda, T* B, size_t ldb);

// solve op(A)*X = alpha*B or X*op(A) = alpha*B in place (reference ATL_Xtrsm, Clint/trsm-micro.base:350-424)
template <typename T>
void trsm(int side, int uplo, int trans, int diag, int M, int N, T alpha, const T* A, size_t lda, T* B, size_t ldb,
          const int* abort_flag = nullptr);

// C <- alpha*A*B + beta*C (Left) or alpha*B*A + beta*C (Right), A symmetric (herm=0) or Hermitian
// (herm=1) stored in one triangle          (reference ATL_Xsymm/ATL_Xhemm, Clint/l3micro.base)
template <typename T>
void symm(int side, int uplo, int herm, int M, int N, T alpha, const T* A, size_t lda, const T* B, size_t ldb, T beta, T* C, size_t ldc);

// C(uplo) <- alpha*op(A)*op(B)^T + alpha*op(B)*op(A)^T + beta*C    (reference ATL_Xsyr2k)
template <typename T>
void syr2k(int uplo, int trans, int N, int K, T alpha, const T* A, size_t lda, const T* B, size_t ldb, T beta, T* C, size_t ldc);

// C(uplo) <- alpha*op(A)*op(B)^H + conj(alpha)*op(B)*op(A)^H + beta*C, beta real   (reference ATL_Xher2k)
template <typename T>
void her2k(int uplo, int trans, int N, int K, T alpha, const T* A, size_t lda, const T* B, size_t ldb, typename ST<T>::real beta,
           T* C, size_t ldc);

// Row interchanges (reference ATL_laswp, Clint/atlas-lp.base:1053); piv is a DEVICE array here.
template <typename T>
void laswp(int N, T* A, size_t lda, int K1, int K2, const int* d_piv, int inci);
// column block b (blk_cols wide) takes the interchanges K1 + b*k1_step .. K2-1 (one launch; multi-GPU LU, lapack.cu)
template <typename T>
void laswp_block_steps(int N, T* A, size_t lda, int K1, int K2, const int* d_piv, int blk_cols, int k1_step);

// Column-major recursive LU with partial pivoting (reference ATL_getrfC, Clint/atlas-lp.base:1497-1617).
// d_ipiv: device int[min(M,N)], 0-based; d_info: device int, set to the 1-based index of the first
// exactly-zero pivot (0 if none).  Everything stays on the stream.
template <typename T>
void getrf(int M, int N, T* A, size_t lda, int* d_ipiv, int* d_info);

// Column-major recursive Cholesky (reference ATL_potrfL/U, Clint/atlas-lp.base:675-968).
// d_info: device int (must be zero on entry), set to the 1-based index of the first non-positive
// pivot; all work after the failing block is skipped on the device, like the reference's early return.
template <typename T>
void potrf(int uplo, int N, T* A, size_t lda, int* d_info);

// ---- QR (qr.cu): A = Q R, LAPACK storage (R above, reflector vectors below the diagonal), tau on the
// device, min(M,N) entries (reference ATL_geqrf, Students/qr.base:594)
template <typename T>
void geqrf(int M, int N, T* A, size_t lda, T* d_tau);
// Bt (cols x rows) <- transpose (conj = conjugate transpose) of B (rows x cols)
template <typename T>
void transpose_conj(int rows, int cols, const T* B, size_t ldb, T* Bt, size_t ldbt, bool conj);

// ---- solve / inverse layer (lapack_inv.cu) ---------------------------------
// B <- solution of op(A) X = B from getrf's factors; order/trans as ATL_getrs (atlas-lp.base:1631)
template <typename T>
void getrs(int order, int trans, int N, int NRHS, const T* A, size_t lda, const int* d_ipiv, T* B, size_t ldb);
// B <- solution of A X = B from potrf's factor (ATL_potrs, atlas-lp.base:969)
template <typename T>
void potrs(int order, int uplo, int N, int NRHS, const T* A, size_t lda, T* B, size_t ldb);
// *d_info <- 1-based index of the first zero diagonal element of a NonUnit triangle (ATL_trtri, :2341)
template <typename T>
void trtri_check(int diag, int N, const T* A, size_t lda, int* d_info);
// column-major triangular inverse in place (ATL_trtriCL/CU, :1902-2199); call trtri_check first
template <typename T>
void trtri(int uplo, int diag, int N, T* A, size_t lda);
// column-major U*U^H (Upper) or L^H*L (Lower) in place (ATL_lauumCU/CL, :221-311)
template <typename T>
void lauum(int uplo, int N, T* A, size_t lda);
// the part of ATL_getriC (:135-201) after trtri(Upper): inv(A)*L = inv(U), then column interchanges
template <typename T>
void getri_finish(int N, T* A, size_t lda, const int* d_ipiv);
template <typename T>
void conj_inplace(int M, int N, T* B, size_t ldb);

// large GEMMs launched by this host thread stay off n SMs (<= 16, one GPC) while n > 0 (see lapack.cu)
void lu_reserve_sms(int n);
// 1 if an M-row LU panel of type pre (s/d/c/z) runs on the cluster leaf, i.e. can be factored beside an update GEMM that keeps a GPC free
int lu_panel_overlaps(char pre, int M);
// diagnostic: copies the 3 checksums per block step of the last getrf run with ATLB200_LU_CHECKSUM=1; returns the step count
int lu_checksums(double* host_out, int max_steps);

// Grow-only per-thread device workspace (freed by free_workspace()).
void* workspace(size_t bytes, int slot = 0);
void keep_pool_cached();   // cudaMallocAsync pool keeps its memory between calls
void free_workspace();
long long launch_count();
void reset_launch_count();

}  // namespace ab
