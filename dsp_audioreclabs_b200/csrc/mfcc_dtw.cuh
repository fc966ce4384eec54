// MFCC + DTW launchers (internal to libdspfront; csrc/mfcc_dtw.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dsp {

struct MfccArgs {
  int frame_length, frame_shift, n_fft, n_mels, n_ceps;
  double pre_emphasis;
  float log_floor;
};

size_t mfcc_smem_bytes(int n_fft, int n_mels);
cudaError_t launch_mfcc(const int16_t* samples, const int64_t* offsets, const int32_t* lengths, const int32_t* seg_start,
                        const int32_t* seg_end, int64_t n_utts, const MfccArgs& a, const float* window, const float2* twiddle,
                        const float* filterbank, const int2* fb_range, const float* dct, const int64_t* mfcc_offsets, float* out,
                        int32_t* n_frames_out, unsigned int* work_counter, int sm_count, cudaStream_t st);
int dtw_max_query_frames();
int dtw_max_dim();
size_t dtw_smem_bytes(int max_t_frames, int dim);     // template frames + one boundary row
cudaError_t launch_dtw(const float* qf, const int64_t* qoff, int64_t nq, int max_q_frames, const float* tf, const int64_t* toff,
                       int64_t nt, int max_t_frames, int dim, float* cost, cudaStream_t st);
cudaError_t launch_dtw_topk(const float* cost, int64_t nq, int64_t nt, int k, int64_t index_base, const int32_t* labels,
                            double* nbr_cost, int64_t* nbr_idx, int32_t* nbr_label, cudaStream_t st);

}  // namespace dsp
