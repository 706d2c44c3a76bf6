// Fast SGEMM path (exact FP32 FFMA, 128x128x8 tiles).  See sgemm_ffma.cu.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>
namespace ab {
// C(cmode triangle) = alpha*op(A)*op(B) + beta*C on device pointers; ta/tb = 1 for transposed operands.
bool sgemm_ffma(int ta, int tb, int M, int N, int K, float alpha, const float* A, size_t lda, const float* B, size_t ldb, float beta,
                float* C, size_t ldc, int cmode, const int* abort_flag = nullptr, const int* run_flag = nullptr);
// complex float: ta/tb transposed, ca/cb conjugated, herk forces Im(C(i,i)) = 0 (cgemm_ffma.cu)
bool cgemm_ffma(int ta, int ca, int tb, int cb, int M, int N, int K, float2 alpha, const float2* A, size_t lda, const float2* B,
                size_t ldb, float2 beta, float2* C, size_t ldc, int cmode, int herk, const int* abort_flag = nullptr,
                const int* run_flag = nullptr);
}  // namespace ab

// SGEMM on the tcgen05 tensor cores with the 3xTF32 error-compensated split (sgemm_tc.cu); returns
// false without doing anything when the packed operands would exceed the workspace budget.
namespace ab {
bool sgemm_tc(int ta, int tb, int M, int N, int K, float alpha, const float* A, size_t lda, const float* B, size_t ldb, float beta,
              float* C, size_t ldc, int cmode, const int* abort_flag);
bool cgemm_tc(int ta, int ca, int tb, int cb, int M, int N, int K, float2 alpha, const float2* A, size_t lda, const float2* B, size_t ldb,
              float2 beta, float2* C, size_t ldc, int cmode, int herk, const int* abort_flag);
}
