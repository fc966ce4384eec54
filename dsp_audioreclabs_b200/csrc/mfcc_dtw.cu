// MFCC + DTW template matching (BASELINE config 5).
//
// NOT present in the reference (no mfcc / dtw / fft anywhere under /root/reference): the definition is the
// textbook one written down in oracle/mfcc_dtw_oracle.py ("self-oracle, parity unpinned"), built on the reference's
// own pre-processing (src/audio_processing.py:49-90), endpoints (:135-275) and framing rule (:299-333).
//
//   mfcc_kernel   one CTA per utterance (persistent over a work counter): DC / peak of the utterance from exact integer
//                 sums, then per frame of the trimmed segment: pre-emphasis + window -> radix-2 FFT in shared memory ->
//                 power spectrum -> mel filterbank (one warp per filter over its non-zero bins) -> log -> DCT-II.
//                 Filterbank, DCT matrix, twiddles and window are caller-provided tables: the kernel fixes no constants.
//   dtw_kernel    one 32-thread CTA per (query, template) pair, anti-diagonal wavefront: lane L owns a strip of R
//                 consecutive query frames and runs one column behind lane L-1; the boundary value D[i-1][j] travels
//                 by one shuffle per step, template frames are read from shared memory (stride 13 floats: conflict
//                 free), query frames live in registers.  Cost matrix [queries x templates] in fp32.
//   dtw_topk      one warp per query: the k smallest costs (ties to the lower template index) -> candidates in the
//                 layout knn_merge_vote consumes, so the row-sharded multi-GPU exchange is the KNN one.
#include <algorithm>
#include <cfloat>

#include <climits>
#include "kernels.cuh"
#include "knn.cuh"
#include "mfcc_dtw.cuh"

namespace dsp {

namespace {

constexpr int kMfccThreads = 256;

__device__ __forceinline__ long long block_sum_ll(long long v, long long* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) sh[w] = v;
  __syncthreads();
  long long t = 0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
  __syncthreads();
  return t;
}
__device__ __forceinline__ int block_min_i(int v, int* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = __reduce_min_sync(0xffffffffu, v);
  if (lane == 0) sh[w] = v;
  __syncthreads();
  int t = sh[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) t = min(t, sh[i]);
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(kMfccThreads)
mfcc_kernel(const int16_t* __restrict__ samples, const int64_t* __restrict__ offsets, const int32_t* __restrict__ lengths,
            const int32_t* __restrict__ seg_start, const int32_t* __restrict__ seg_end, int64_t n_utts, MfccArgs a,
            const float* __restrict__ window, const float2* __restrict__ twiddle, const float* __restrict__ filterbank,
            const int2* __restrict__ fb_range, const float* __restrict__ dct, const int64_t* __restrict__ mfcc_offsets,
            float* __restrict__ out, int32_t* __restrict__ n_frames_out, unsigned int* __restrict__ work_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* buf = reinterpret_cast<float2*>(smem_raw);                       // n_fft complex values
  float* power = reinterpret_cast<float*>(buf + a.n_fft);                  // n_fft / 2 + 1
  float* logmel = power + (a.n_fft / 2 + 1);                               // n_mels
  __shared__ long long sh_ll[kMfccThreads / 32];
  __shared__ int sh_i[kMfccThreads / 32];
  __shared__ unsigned int sh_u;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int n_fft = a.n_fft, half_fft = n_fft >> 1, n_bins = half_fft + 1;
  const int log2n = 31 - __clz(n_fft);
  for (;;) {
    if (tid == 0) sh_u = atomicAdd(work_counter, 1u);
    __syncthreads();
    const int64_t u = sh_u;
    __syncthreads();
    if (u >= n_utts) break;
    const int16_t* x = samples + offsets[u];
    const int n = lengths ? lengths[u] : (int)(offsets[u + 1] - offsets[u]);
    // ---- remove_dc + normalize_audio constants (src/audio_processing.py:49-75) from exact integer sums
    long long s = 0; int mn = 32767, mx = -32768;
    for (int i = tid; i < n; i += kMfccThreads) { const int k = x[i]; s += k; mn = min(mn, k); mx = max(mx, k); }
    const long long S = block_sum_ll(s, sh_ll);
    mn = block_min_i(mn, sh_i);
    mx = -block_min_i(-mx, sh_i);
    const long long N = n > 0 ? n : 1;
    const long long M = max(N * mx - S, S - N * mn);                      // N * max|k - mean|
    const double mu = (double)S / (double)N;
    const double scale = M > 0 ? (double)N / (double)M : 1.0 / 32768.0;   // all samples equal: x - mean = 0 anyway
    const int start = seg_start[u], end = min(seg_end[u], n);
    const int seg = end - start;
    const int nf = (int)frame_count_host_device(seg, a.frame_length, a.frame_shift);
    if (tid == 0 && n_frames_out) n_frames_out[u] = nf;
    const int64_t obase = mfcc_offsets[u];
    for (int f = 0; f < nf; ++f) {
      const int p = start + f * a.frame_shift;
      const int valid = min(a.frame_length, end - p);
      // ---- pre-emphasis + window, stored in bit-reversed order for the in-place decimation-in-time FFT
      for (int j = tid; j < n_fft; j += kMfccThreads) {
        float v = 0.f;
        if (j < valid) {
          const double cur = ((double)x[p + j] - mu) * scale;
          const double prv = (p + j > start) ? ((double)x[p + j - 1] - mu) * scale : 0.0;
          v = (float)(cur - a.pre_emphasis * prv) * window[j];
        }
        buf[__brev((unsigned)j) >> (32 - log2n)] = make_float2(v, 0.f);
      }
      __syncthreads();
      for (int st = 1; st <= log2n; ++st) {
        const int half = 1 << (st - 1);
        const int tw_stride = half_fft >> (st - 1);
        for (int b = tid; b < half_fft; b += kMfccThreads) {
          const int pos = b & (half - 1);
          const int i0 = ((b >> (st - 1)) << st) + pos, i1 = i0 + half;
          const float2 w = __ldg(twiddle + pos * tw_stride);
          const float2 lo = buf[i0], hi = buf[i1];
          const float tr = hi.x * w.x - hi.y * w.y, ti = hi.x * w.y + hi.y * w.x;
          buf[i0] = make_float2(lo.x + tr, lo.y + ti);
          buf[i1] = make_float2(lo.x - tr, lo.y - ti);
        }
        __syncthreads();
      }
      const float inv_nfft = 1.0f / (float)n_fft;
      for (int b = tid; b < n_bins; b += kMfccThreads) { const float2 c = buf[b]; power[b] = (c.x * c.x + c.y * c.y) * inv_nfft; }
      __syncthreads();
      // ---- mel filterbank: one warp per filter over its non-zero bins
      for (int m = wid; m < a.n_mels; m += kMfccThreads / 32) {
        const int2 r = fb_range[m];
        float acc = 0.f;
        for (int b = r.x + lane; b < r.y; b += 32) acc = fmaf(__ldg(filterbank + (size_t)m * n_bins + b), power[b], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) logmel[m] = logf(fmaxf(acc, a.log_floor));
      }
      __syncthreads();
      // ---- DCT-II
      for (int c = tid; c < a.n_ceps; c += kMfccThreads) {
        float acc = 0.f;
        for (int m = 0; m < a.n_mels; ++m) acc = fmaf(__ldg(dct + c * a.n_mels + m), logmel[m], acc);
        out[(obase + f) * a.n_ceps + c] = acc;
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------
// DTW: one warp-sized CTA per (query, template) pair
// ---------------------------------------------------------------------------------------
// Lane L owns R consecutive query frames of a STRIP of 32 R rows (in registers) and runs one column behind lane L - 1:
// D[i-1][j] crosses the lane boundary by one shfl_up per step.  Queries longer than a strip are processed strip by
// strip: the strip's last row is kept in shared memory (one float per template frame, written in place -- lane 0 has
// read column j long before the last lane overwrites it) and enters the next strip as its top boundary.  DMAX bounds
// the feature dimension held in registers (R * DMAX <= 128 floats per lane): 16 for the 13 MFCCs, 32 / 64 / 128 for
// stacked delta features or other embeddings.
constexpr int kDtwMaxDim = 128;

template <int R, int DMAX>
__global__ void __launch_bounds__(32)
dtw_kernel(const float* __restrict__ qf, const int64_t* __restrict__ qoff, int64_t nq, const float* __restrict__ tf,
           const int64_t* __restrict__ toff, int64_t nt, int dim, int max_t_frames, float* __restrict__ cost) {
  extern __shared__ float st[];                 // template frames [m][dim], then the boundary row [max_t_frames]
  float* brow = st + (size_t)max_t_frames * dim;
  const int lane = threadIdx.x;
  const int64_t ti = blockIdx.x, qi = blockIdx.y;
  const int n = (int)(qoff[qi + 1] - qoff[qi]), m = (int)(toff[ti + 1] - toff[ti]);
  float* dst = cost + qi * nt + ti;
  if (n == 0 || m == 0) { if (lane == 0) *dst = INFINITY; return; }
  const float* tp = tf + toff[ti] * dim;
  for (int i = lane; i < m * dim; i += 32) st[i] = tp[i];
  for (int j = lane; j < m; j += 32) brow[j] = INFINITY;      // "row -1"
  __syncwarp();
  const float* qp = qf + qoff[qi] * dim;
  float result = INFINITY;
  constexpr int kStrip = 32 * R;
#pragma unroll 1
  for (int i0 = 0; i0 < n; i0 += kStrip) {
    const int ns = min(kStrip, n - i0);         // rows of this strip
    // this lane's rows of the strip in registers
    float q[R][DMAX];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = i0 + lane * R + r;
#pragma unroll
      for (int c = 0; c < DMAX; ++c) q[r][c] = (i < n && c < dim) ? qp[(size_t)i * dim + c] : 0.f;
    }
    float prev[R];                              // D[i][j-1] of the lane's rows
#pragma unroll
    for (int r = 0; r < R; ++r) prev[r] = INFINITY;
    float top_cur = INFINITY, top_prev = INFINITY;   // D[i-1][j], D[i-1][j-1] from the lane above (lane 0: the boundary row)
    float bottom = INFINITY;                    // D[last row of the lane][j] of the previous step
    const int last_lane = (ns - 1) / R;
    const int steps = m + last_lane;
#pragma unroll 1
    for (int s = 0; s < steps; ++s) {
      // the lane above finished column j = s - lane in the previous step
      const float from_above = __shfl_up_sync(0xffffffffu, bottom, 1);
      top_prev = top_cur;
      top_cur = lane == 0 ? (s < m ? brow[s] : INFINITY) : from_above;
      const int j = s - lane;
      if (j >= 0 && j < m && lane <= last_lane) {
        float t[DMAX];
#pragma unroll
        for (int c = 0; c < DMAX; ++c) t[c] = c < dim ? st[j * dim + c] : 0.f;
        float up = top_cur, diag = top_prev;
        float cur_r = INFINITY;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i = i0 + lane * R + r;
          float d2 = 0.f;
#pragma unroll
          for (int c = 0; c < DMAX; ++c) { const float e = q[r][c] - t[c]; d2 = fmaf(e, e, d2); }
          const float d = sqrtf(d2);
          float best = fminf(fminf(up, prev[r]), diag);
          if (i == 0 && j == 0) best = 0.f;
          const float v = i < n ? d + best : INFINITY;
          diag = prev[r];                       // D[i][j-1] is the diagonal of the row below
          up = v;
          prev[r] = v;
          if (i == i0 + ns - 1) cur_r = v;
        }
        bottom = prev[R - 1];
        if (lane == last_lane) {
          brow[j] = cur_r;                      // the strip's last row: top boundary of the next strip
          if (i0 + ns == n && j == m - 1) result = cur_r;
        }
      }
    }
    __syncwarp();
  }
  result = __shfl_sync(0xffffffffu, result, ((n - 1) % kStrip) / R);
  if (lane == 0) *dst = result;
}

// k smallest costs per query (ties to the lower template index): one warp per query
__global__ void dtw_topk_kernel(const float* __restrict__ cost, int64_t nq, int64_t nt, int k, int64_t index_base,
                                const int32_t* __restrict__ labels, double* __restrict__ nbr_cost,
                                int64_t* __restrict__ nbr_idx, int32_t* __restrict__ nbr_label) {
  const int lane = threadIdx.x & 31;
  const int64_t qi = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (qi >= nq) return;
  const float* row = cost + qi * nt;
  float bd[kKnnMaxK]; int64_t bi[kKnnMaxK];
  for (int c = 0; c < k; ++c) { bd[c] = INFINITY; bi[c] = INT64_MAX; }
  for (int64_t i = lane; i < nt; i += 32) {
    const float d = row[i];
    if (d < bd[k - 1] || (d == bd[k - 1] && i < bi[k - 1])) {
      int s = k - 1;
      while (s > 0 && (d < bd[s - 1] || (d == bd[s - 1] && i < bi[s - 1]))) { bd[s] = bd[s - 1]; bi[s] = bi[s - 1]; --s; }
      bd[s] = d; bi[s] = i;
    }
  }
  int head = 0;
  for (int c = 0; c < k; ++c) {
    float md = head < k ? bd[head] : INFINITY;
    int64_t mi = head < k ? bi[head] : INT64_MAX;
    int owner = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, md, o);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, mi, o);
      const int oo = __shfl_xor_sync(0xffffffffu, owner, o);
      if (od < md || (od == md && oi < mi)) { md = od; mi = oi; owner = oo; }
    }
    if (owner == lane && mi != INT64_MAX) ++head;
    if (lane == 0) {
      const bool valid = mi != INT64_MAX && md < INFINITY;
      if (nbr_cost) nbr_cost[qi * k + c] = valid ? (double)md : INFINITY;
      if (nbr_idx) nbr_idx[qi * k + c] = valid ? index_base + mi : -1;
      if (nbr_label) nbr_label[qi * k + c] = valid ? labels[mi] : -1;
    }
  }
}

}  // namespace

size_t mfcc_smem_bytes(int n_fft, int n_mels) { return (size_t)n_fft * 8 + (size_t)(n_fft / 2 + 1) * 4 + (size_t)n_mels * 4 + 16; }

cudaError_t launch_mfcc(const int16_t* samples, const int64_t* offsets, const int32_t* lengths, const int32_t* seg_start,
                        const int32_t* seg_end, int64_t n_utts, const MfccArgs& a, const float* window, const float2* twiddle,
                        const float* filterbank, const int2* fb_range, const float* dct, const int64_t* mfcc_offsets, float* out,
                        int32_t* n_frames_out, unsigned int* work_counter, int sm_count, cudaStream_t st) {
  if (n_utts == 0) return cudaSuccess;
  const size_t smem = mfcc_smem_bytes(a.n_fft, a.n_mels);
  cudaError_t e = cudaFuncSetAttribute(mfcc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaMemsetAsync(work_counter, 0, sizeof(unsigned int), st);
  const int grid = (int)std::min<int64_t>(n_utts, (int64_t)sm_count * 4);
  mfcc_kernel<<<grid, kMfccThreads, smem, st>>>(samples, offsets, lengths, seg_start, seg_end, n_utts, a, window, twiddle,
                                                filterbank, fb_range, dct, mfcc_offsets, out, n_frames_out, work_counter);
  return cudaGetLastError();
}

int dtw_max_query_frames() { return INT32_MAX; }      // any length: strips of 32 R rows
int dtw_max_dim() { return kDtwMaxDim; }
size_t dtw_smem_bytes(int max_t_frames, int dim) { return ((size_t)std::max(max_t_frames, 1) * dim + (size_t)std::max(max_t_frames, 1)) * sizeof(float); }

cudaError_t launch_dtw(const float* qf, const int64_t* qoff, int64_t nq, int max_q_frames, const float* tf, const int64_t* toff,
                       int64_t nt, int max_t_frames, int dim, float* cost, cudaStream_t st) {
  if (nq == 0 || nt == 0) return cudaSuccess;
  if (dim > kDtwMaxDim) return cudaErrorInvalidValue;
  const size_t smem = dtw_smem_bytes(max_t_frames, dim);
  // grid.y is limited to 65535: queries are processed in slabs
  for (int64_t q0 = 0; q0 < nq; q0 += 65535) {
    const int64_t qc = std::min<int64_t>(65535, nq - q0);
    const dim3 grid((unsigned)nt, (unsigned)qc);
    const float* cq = qf; const int64_t* oq = qoff + q0; float* cc = cost + q0 * nt;
    auto run = [&](auto kern) -> cudaError_t {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      kern<<<grid, 32, smem, st>>>(cq, oq, qc, tf, toff, nt, dim, std::max(max_t_frames, 1), cc);
      return cudaGetLastError();
    };
    cudaError_t e;
    if (dim > 64) e = run(dtw_kernel<1, 128>);
    else if (dim > 32) e = run(dtw_kernel<2, 64>);
    else if (dim > 16) e = run(dtw_kernel<4, 32>);
    else if (max_q_frames <= 32) e = run(dtw_kernel<1, 16>);
    else if (max_q_frames <= 64) e = run(dtw_kernel<2, 16>);
    else if (max_q_frames <= 128) e = run(dtw_kernel<4, 16>);
    else e = run(dtw_kernel<8, 16>);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_dtw_topk(const float* cost, int64_t nq, int64_t nt, int k, int64_t index_base, const int32_t* labels,
                            double* nbr_cost, int64_t* nbr_idx, int32_t* nbr_label, cudaStream_t st) {
  if (nq == 0) return cudaSuccess;
  dtw_topk_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, st>>>(cost, nq, nt, k, index_base, labels, nbr_cost, nbr_idx, nbr_label);
  return cudaGetLastError();
}

}  // namespace dsp
