// Float64 replay kernel: the whole front end for one utterance per CTA, with the
// reference's float64 operation order restated on the GPU so that every
// integer-valued result (zero-crossing counts, endpoint indices, frame counts) is
// identical to the NumPy path for ANY input, including float arrays passed to the
// per-call API and utterances the int16 fast kernel flags as too close to a
// threshold to certify (frontend_pcm.cu).  It is the general path, not the fast
// one: samples stay in global memory (L2 resident), one thread per frame.
//
// Reference: src/audio_processing.py:49-90 (pre-processing), :135-275 (endpoint
// detection), :299-333 (framing), src/feature_extraction.py:12-88 (features, statistics).
#include "kernels.cuh"

namespace dsp {

namespace {

struct SampleReader {
  const void* base;
  int dtype;
  int channels;
  int64_t off;  // element offset of the utterance
  __device__ __forceinline__ double raw(int64_t e) const {
    switch (dtype) {
      case DSP_S16: return (double)((const int16_t*)base)[off + e] / 32768.0;
      // the reference subtracts 128 from a uint8 array, which wraps modulo 256 in NumPy
      // (audio_processing.py:33-34): values below 128 come out as (u + 128) / 128
      case DSP_U8: return (double)(uint8_t)(((const uint8_t*)base)[off + e] - 128u) / 128.0;
      case DSP_F32: return (double)((const float*)base)[off + e];
      default: return ((const double*)base)[off + e];
    }
  }
  // load_wav: stereo -> mean over the two channels (audio_processing.py:43-44)
  __device__ __forceinline__ double operator()(int64_t i) const {
    if (channels == 2) return ((0.0 + raw(2 * i)) + raw(2 * i + 1)) / 2.0;
    return raw(i);
  }
};

// Leaves of NumPy's pairwise tree over n terms, enumerated left to right by one thread.
constexpr int kMaxLeaves = 2048;

__device__ int enumerate_leaves(int64_t n, int* leaf_off, int* leaf_n) {
  int64_t st_off[kPairwiseDepth], st_n[kPairwiseDepth];
  int sp = 0, cnt = 0;
  st_off[0] = 0; st_n[0] = n;
  while (sp >= 0) {
    const int64_t o = st_off[sp], m = st_n[sp];
    --sp;
    if (m <= kPairwiseBlock) {
      if (cnt < kMaxLeaves) { leaf_off[cnt] = (int)o; leaf_n[cnt] = (int)m; }
      ++cnt;
    } else {
      int64_t n2 = m / 2; n2 -= n2 % 8;
      st_off[sp + 1] = o + n2; st_n[sp + 1] = m - n2;   // right pushed first, popped second
      st_off[sp + 2] = o; st_n[sp + 2] = n2;
      sp += 2;
    }
  }
  return cnt;
}

// Replays the tree over precomputed leaf sums (one thread).
__device__ double combine_leaves(int64_t n, const double* leaf_sum) {
  int64_t st_n[kPairwiseDepth];
  double st_left[kPairwiseDepth];
  int8_t st_state[kPairwiseDepth];
  int sp = 0, next = 0;
  st_n[0] = n; st_state[0] = 0;
  double ret = 0.0;
  bool have = false;
  while (sp >= 0) {
    if (have) {
      if (st_state[sp] == 1) {
        st_left[sp] = ret; st_state[sp] = 2; have = false;
        int64_t n2 = st_n[sp] / 2; n2 -= n2 % 8;
        st_n[sp + 1] = st_n[sp] - n2; st_state[sp + 1] = 0; ++sp;
      } else { ret = st_left[sp] + ret; --sp; }
    } else {
      const int64_t nn = st_n[sp];
      if (nn <= kPairwiseBlock) { ret = leaf_sum[next++]; have = true; --sp; }
      else {
        st_state[sp] = 1;
        int64_t n2 = nn / 2; n2 -= n2 % 8;
        st_n[sp + 1] = n2; st_state[sp + 1] = 0; ++sp;
      }
    }
  }
  return ret;
}

// np.sum over n float64 terms with NumPy's association, block-parallel over the leaves.
template <class Term>
__device__ double block_np_sum(Term term, int64_t n, int* leaf_off, int* leaf_n, double* leaf_sum,
                               double* bcast) {
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = enumerate_leaves(n, leaf_off, leaf_n);
  __syncthreads();
  const int cnt = s_cnt;
  if (cnt <= kMaxLeaves) {
    for (int l = threadIdx.x; l < cnt; l += blockDim.x)
      leaf_sum[l] = np_pairwise_leaf(term, leaf_off[l], leaf_n[l]);
    __syncthreads();
    if (threadIdx.x == 0) *bcast = combine_leaves(n, leaf_sum);
  } else {
    if (threadIdx.x == 0) *bcast = np_pairwise_sum(term, n);
  }
  __syncthreads();
  const double r = *bcast;
  __syncthreads();
  return r;
}

__device__ __forceinline__ int sign_is_pos(double v) { return v > 0.0; }

// compute_zero_crossing_rate (audio_processing.py:119-132): sign changes, zero counted negative.
template <class Term>
__device__ __forceinline__ int count_crossings(Term term, int n) {
  if (n <= 0) return 0;
  int prev = sign_is_pos(term(0)), c = 0;
  for (int j = 1; j < n; ++j) {
    const int s = sign_is_pos(term(j));
    c += (s != prev);
    prev = s;
  }
  return c;
}

}  // namespace

// compute_statistics (feature_extraction.py:46-62) of seq[0..n) (float64, global or shared
// memory) -> out5 = mean, std (ddof 0), max, min, median.  Block-wide; all threads call.
__device__ void block_sequence_stats(const double* seq, int n, double* out5, int* hist,
                                     unsigned long long* sh) {
  __shared__ double s_mean;
  if (threadIdx.x == 0) {
    auto t = [&](int64_t i) { return seq[i]; };
    const double mean = np_pairwise_sum(t, n) / (double)n;
    s_mean = mean;
    out5[0] = mean;
  }
  __syncthreads();
  const double mean = s_mean;
  if (threadIdx.x == 32 % blockDim.x) {
    auto t = [&](int64_t i) { const double d = seq[i] - mean; return d * d; };
    out5[1] = sqrt(np_pairwise_sum(t, n) / (double)n);
  }
  double mx = -INFINITY, mn = INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { const double v = seq[i]; mx = fmax(mx, v); mn = fmin(mn, v); }
  mx = block_reduce<double>(mx, OpMaxD(), -INFINITY, reinterpret_cast<double*>(sh));
  mn = block_reduce<double>(mn, OpMinD(), INFINITY, reinterpret_cast<double*>(sh));
  uint64_t ka, kb;
  auto get = [&](int i) { return f64_key(seq[i]); };
  block_select_pair(get, n, (n - 1) / 2, hist, sh, &ka, &kb);
  if (threadIdx.x == 0) {
    out5[2] = mx; out5[3] = mn;
    const double a = key_f64(ka), b = key_f64(kb);
    out5[4] = (n & 1) ? a : ((0.0 + a) + b) / 2.0;   // np.median: mean of the two middle values
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kExactThreads)
frontend_exact_kernel(const ExactArgs a) {
  __shared__ int s_leaf_off[kMaxLeaves];
  __shared__ int s_leaf_n[kMaxLeaves];
  __shared__ double s_leaf_sum[kMaxLeaves];
  __shared__ int s_hist[256];
  __shared__ unsigned long long s_sh[64];
  __shared__ double s_b[8];
  __shared__ int s_i[8];

  const int64_t n_items = a.list_count ? (int64_t)*a.list_count : a.n_items;
  double* zbuf_cta = a.zbuf ? a.zbuf + (int64_t)blockIdx.x * a.zbuf_stride : nullptr;
  double* seq_cta = a.seqbuf + (int64_t)blockIdx.x * 5 * a.seq_cap;
  double* e_list = seq_cta;
  double* z_list = seq_cta + a.seq_cap;
  double* f_e = seq_cta + 2 * a.seq_cap;
  double* f_m = seq_cta + 3 * a.seq_cap;
  double* f_z = seq_cta + 4 * a.seq_cap;
  const int fl = a.fl, fs = a.fs;

  for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int64_t b = a.list ? (int64_t)a.list[item] : item;
    const int64_t off = a.offsets[b];
    const int64_t n = (a.lengths ? (int64_t)a.lengths[b] : a.offsets[b + 1] - off) / a.channels;
    SampleReader rd{a.samples, a.dtype, a.channels, off};
    int status = DSP_UTT_EXACT;

    // ---- pre-processing (audio_processing.py:49-90) --------------------------------
    const double* z;
    if (a.pre_mode == 0 && a.dtype == DSP_F64 && a.channels == 1) {
      z = (const double*)a.samples + off;
    } else {
      double* zw = a.pre_out ? a.pre_out : zbuf_cta;
      double mean = 0.0;
      if (n > 0 && (a.pre_mode & 1)) {
        if (a.dtype == DSP_S16 || a.dtype == DSP_U8) {
          // PCM: every partial sum is an exact multiple of 2^-15 (2^-8 for stereo 8-bit), so
          // any association gives NumPy's value.
          double s = 0.0;
          for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += rd(i);
          s = block_reduce<double>(s, OpAddD(), 0.0, reinterpret_cast<double*>(s_sh));
          mean = s / (double)n;
        } else {
          mean = block_np_sum(rd, n, s_leaf_off, s_leaf_n, s_leaf_sum, &s_b[0]) / (double)n;
        }
      }
      double peak = 0.0;
      for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double y = (a.pre_mode & 1) ? rd(i) - mean : rd(i);
        zw[i] = y;
        peak = fmax(peak, fabs(y));
      }
      if (a.pre_mode & 2) {
        peak = block_reduce<double>(peak, OpMaxD(), 0.0, reinterpret_cast<double*>(s_sh));
        if (peak > 0.0)
          for (int64_t i = threadIdx.x; i < n; i += blockDim.x) zw[i] = zw[i] / peak;
      }
      __syncthreads();
      z = zw;
    }
    if (n == 0) status |= DSP_UTT_EMPTY;

    // ---- endpoint detection (audio_processing.py:135-275) --------------------------
    int start = 0, end = (int)n, f1 = 0;
    if (a.do_epd && n >= fl) {
      f1 = (int)((n - fl) / fs) + 1;
      for (int f = threadIdx.x; f < f1; f += blockDim.x) {
        const double* fr = z + (int64_t)f * fs;
        auto sq = [&](int64_t i) { const double v = fr[i]; return v * v; };
        auto id = [&](int j) { return fr[j]; };
        e_list[f] = np_pairwise_sum(sq, fl);
        z_list[f] = (double)count_crossings(id, fl);
      }
      __syncthreads();
      const int nf = min(5, f1 / 10);
      uint64_t ka, kb;
      {
        const double v = (double)(f1 - 1) * (90.0 / 100.0);
        const int j = (int)floor(v);
        auto get = [&](int i) { return f64_key(e_list[i]); };
        if (v >= (double)(f1 - 1)) {           // f1 == 1: both neighbours are the last element
          block_select_pair(get, f1, f1 - 1, s_hist, s_sh, &ka, &kb);
          kb = ka;
        } else {
          block_select_pair(get, f1, j, s_hist, s_sh, &ka, &kb);
        }
        if (threadIdx.x == 0) {
          const double g = v - floor(v);
          const double speech = np_lerp(key_f64(ka), key_f64(kb), g);
          double noise_e, noise_z;
          if (nf > 0) {
            auto te = [&](int64_t i) { return i < nf ? e_list[i] : e_list[f1 - 2 * nf + i]; };
            auto tz = [&](int64_t i) { return i < nf ? z_list[i] : z_list[f1 - 2 * nf + i]; };
            noise_e = np_pairwise_sum(te, 2 * nf) / (double)(2 * nf);
            noise_z = np_pairwise_sum(tz, 2 * nf) / (double)(2 * nf);
          } else {
            noise_e = e_list[0]; noise_z = z_list[0];
            for (int i = 1; i < f1; ++i) { noise_e = fmin(noise_e, e_list[i]); noise_z = fmin(noise_z, z_list[i]); }
          }
          s_b[1] = speech * a.hr;                                   // T1 (:202)
          s_b[2] = noise_e + (speech - noise_e) * a.lr;             // T2 (:217)
          s_b[3] = noise_z * a.zr;                                  // T3 (:249)
        }
        __syncthreads();
      }
      const double t1 = s_b[1], t2 = s_b[2], t3 = s_b[3];
      // N3 / N4: first and last frame above T1 (:205-213)
      int n3 = f1, n4 = -1;
      for (int f = threadIdx.x; f < f1; f += blockDim.x)
        if (e_list[f] > t1) { n3 = min(n3, f); n4 = max(n4, f); }
      n3 = block_reduce<int>(n3, OpMinI(), f1, s_i);
      n4 = block_reduce<int>(n4, OpMaxI(), -1, s_i);
      if (n4 >= 0) {
        // N2 / N5: outward to the first frame at or below T2 (:220-237)
        int n2 = 0, n5 = f1 - 1;
        for (int f = threadIdx.x; f < f1; f += blockDim.x) {
          if (e_list[f] <= t2) {
            if (f < n3) n2 = max(n2, f + 1);
            if (f > n4) n5 = min(n5, f - 1);
          }
        }
        n2 = block_reduce<int>(n2, OpMaxI(), 0, s_i);
        n5 = block_reduce<int>(n5, OpMinI(), f1 - 1, s_i);
        // N1 / N6: outward to the first frame at or below T3 in zero-crossing rate (:252-269)
        int n1 = 0, n6 = f1 - 1;
        for (int f = threadIdx.x; f < f1; f += blockDim.x) {
          if (z_list[f] <= t3) {
            if (f < n2) n1 = max(n1, f + 1);
            if (f > n5) n6 = min(n6, f - 1);
          }
        }
        n1 = block_reduce<int>(n1, OpMaxI(), 0, s_i);
        n6 = block_reduce<int>(n6, OpMinI(), f1 - 1, s_i);
        start = n1 * fs;
        end = (int)min((int64_t)n6 * fs + fl, n);
      }
      if (a.out.epd_energy || a.out.epd_zcr || a.epd_zcr_f64) {
        const int64_t eo = a.epd_offsets ? a.epd_offsets[b] : 0;
        for (int f = threadIdx.x; f < f1; f += blockDim.x) {
          if (a.out.epd_energy) a.out.epd_energy[eo + f] = e_list[f];
          if (a.out.epd_zcr) a.out.epd_zcr[eo + f] = (float)z_list[f];
          if (a.epd_zcr_f64) a.epd_zcr_f64[eo + f] = z_list[f];
        }
      }
    }
    const int seg = end - start;
    if (seg <= 0) status |= DSP_UTT_EMPTY;

    // ---- framing + per-frame features (audio_processing.py:299-333, feature_extraction.py:12-43)
    int f2 = 0;
    if (a.do_features && seg > 0) {
      f2 = (int)frame_count_host_device(seg, fl, fs);
      const double* w = a.win;
      const int64_t fo = a.feat_offsets ? a.feat_offsets[b] : 0;
      for (int t = threadIdx.x; t < f2; t += blockDim.x) {
        const int64_t p = (int64_t)start + (int64_t)t * fs;
        const int valid = (int)min((int64_t)fl, (int64_t)end - p);
        const double* fr = z + p;
        auto val = [&](int j) { return j < valid ? fr[j] * w[j] : 0.0 * w[j]; };
        auto sq = [&](int64_t j) { const double v = val((int)j); return v * v; };
        auto ab = [&](int64_t j) { return fabs(val((int)j)); };
        const double e = np_pairwise_sum(sq, fl);
        const double m = np_pairwise_sum(ab, fl);
        const double zc = (double)count_crossings(val, fl);
        f_e[t] = e; f_m[t] = m; f_z[t] = zc;
        if (a.out.energy) a.out.energy[fo + t] = (float)e;
        if (a.out.magnitude) a.out.magnitude[fo + t] = (float)m;
        if (a.out.zcr) a.out.zcr[fo + t] = (float)zc;
        if (a.feat_f64[0]) { a.feat_f64[0][fo + t] = e; a.feat_f64[1][fo + t] = m; a.feat_f64[2][fo + t] = zc; }
        if (a.frames_out)
          for (int j = 0; j < fl; ++j) a.frames_out[(fo + t) * fl + j] = val(j);
      }
      __syncthreads();
      if (f2 == 0) status |= DSP_UTT_NO_FRAMES;
      if (f2 > 0 && (a.out.stats || a.stats_f64)) {
        double* st = &s_leaf_sum[0];  // 15 doubles of scratch (leaf sums are dead here)
        block_sequence_stats(f_e, f2, st + 0, s_hist, s_sh);
        block_sequence_stats(f_m, f2, st + 5, s_hist, s_sh);
        block_sequence_stats(f_z, f2, st + 10, s_hist, s_sh);
        if (threadIdx.x < kStats) {
          if (a.out.stats) a.out.stats[b * kStats + threadIdx.x] = (float)st[threadIdx.x];
          if (a.stats_f64) a.stats_f64[b * kStats + threadIdx.x] = st[threadIdx.x];
        }
      }
    }
    if (threadIdx.x == 0) {
      if (a.out.start) a.out.start[b] = start;
      if (a.out.end) a.out.end[b] = end;
      if (a.out.n_epd_frames) a.out.n_epd_frames[b] = f1;
      if (a.out.n_frames) a.out.n_frames[b] = f2;
      if (a.out.status) a.out.status[b] = status;
    }
    __syncthreads();
  }
}

// extract_frame_features on an arbitrary [n_frames, fl] float64 matrix (feature_extraction.py:12-43).
__global__ void frame_features_kernel(const double* frames, int64_t n_frames, int fl, double* energy,
                                      double* magnitude, double* zcr) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_frames) return;
  const double* fr = frames + t * fl;
  auto sq = [&](int64_t j) { const double v = fr[j]; return v * v; };
  auto ab = [&](int64_t j) { return fabs(fr[j]); };
  auto id = [&](int j) { return fr[j]; };
  energy[t] = np_pairwise_sum(sq, fl);
  magnitude[t] = np_pairwise_sum(ab, fl);
  zcr[t] = (double)count_crossings(id, fl);
}

// compute_statistics for up to three sequences (one CTA): out[5*s .. 5*s+5).
__global__ void __launch_bounds__(kExactThreads)
sequence_stats_kernel(const double* s0, const double* s1, const double* s2, int n, double* out) {
  __shared__ int s_hist[256];
  __shared__ unsigned long long s_sh[64];
  __shared__ double st[kStats];
  const double* seqs[3] = {s0, s1, s2};
  for (int s = 0; s < 3; ++s) {
    if (!seqs[s]) continue;
    block_sequence_stats(seqs[s], n, st + 5 * s, s_hist, s_sh);
    if (threadIdx.x < 5) out[5 * s + threadIdx.x] = st[5 * s + threadIdx.x];
    __syncthreads();
  }
}

}  // namespace dsp
