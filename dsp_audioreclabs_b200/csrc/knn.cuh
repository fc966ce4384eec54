// KNN launchers (internal to libdspfront).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dsp {

constexpr int kKnnCand = 8;    // fp32 candidates kept per query by the scan
constexpr int kKnnMaxK = 16;   // largest supported n_neighbors (certificate needs k < kKnnCand, else rescan)

int knn_padded_dim(int d);     // 16 / 32 / 64, or 0 when d is too large for the tiled scan
cudaError_t knn_pack(const double* train, int64_t n, int d, int dp, float* train32, float* tnorm_max,
                     cudaStream_t st);
cudaError_t knn_scan(int dp, const float* train32, int64_t n, const double* q, int64_t m, int d,
                     int* cand_idx, float* cand_worst, float* qnorm, cudaStream_t st);
cudaError_t knn_rerank(const double* train, const float* train32, int dp, int64_t n, const double* q,
                       int64_t m, int d, int k, int64_t index_base, const int32_t* labels,
                       const int* cand_idx, const float* cand_worst, const float* qnorm,
                       float tnorm_max_host, int64_t* nbr_idx, double* nbr_sqdist, int32_t* nbr_label,
                       int32_t* redo_list, int32_t* redo_count, cudaStream_t st);
cudaError_t knn_redo_all(int32_t* redo_list, int32_t* redo_count, int64_t m, cudaStream_t st);
cudaError_t knn_rescan(const double* train, int64_t n, const double* q, int d, int k, int64_t index_base,
                       const int32_t* labels, const int32_t* redo_list, const int32_t* redo_count,
                       int64_t* nbr_idx, double* nbr_sqdist, int32_t* nbr_label, int sm_count,
                       cudaStream_t st);
cudaError_t knn_vote(const int32_t* nbr_label, int64_t m, int k, int32_t* out, cudaStream_t st);
cudaError_t knn_merge_vote(const double* cd, const int64_t* ci, const int32_t* cl, int r, int64_t m,
                           int k, int32_t* labels_out, int64_t* idx_out, double* dist_out,
                           cudaStream_t st);

}  // namespace dsp
