// Fast ZGEMM path (TMA + DMMA on interleaved complex data).  See zgemm_tma.cu.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>
namespace ab {
// C(cmode triangle) = alpha*op(A)*op(B) + beta*C on device pointers.  ta/tb = 1: operand stored
// transposed; ca/cb = 1: operand conjugated; herk = 1: force Im(C(i,i)) = 0.
// Returns false (nothing launched) when the problem cannot take this path.
bool zgemm_tma(int ta, int ca, int tb, int cb, int M, int N, int K, double2 alpha, const double2* A, size_t lda, const double2* B,
               size_t ldb, double2 beta, double2* C, size_t ldc, int cmode, int herk, const int* abort_flag = nullptr);
}  // namespace ab
