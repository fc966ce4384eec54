#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
namespace dsp {
cudaError_t zscore_fit(const double* x, int64_t n, int d, int pairwise, double* mean, double* std, cudaStream_t st);
cudaError_t zscore_apply(const double* x, int64_t n, int d, const double* mean, double* std, double* out, cudaStream_t st);
cudaError_t zscore_apply_f32(const float* x, int64_t n, int d, const double* mean, const double* std, double* out, cudaStream_t st);
cudaError_t widen_f32(const float* in, int64_t n, double* out, cudaStream_t st);
}  // namespace dsp
