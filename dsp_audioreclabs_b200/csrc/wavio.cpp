// Batched WAV ingest: the step right before the fused front end.
//
// The reference decodes one file at a time with Python's `wave` module and turns the PCM into a
// float64 array (src/audio_processing.py:9-46).  Here the headers of a whole file list are parsed
// natively by a small thread pool (dsp_wav_scan) and the PCM payloads are read straight into ONE
// caller-provided buffer -- normally pinned memory from dsp_host_alloc -- at 16-byte aligned
// offsets (dsp_wav_read), which is exactly the packed layout dsp_frontend_batch_host uploads.
// Samples stay in their stored encoding (int16 / uint8, interleaved stereo): the conversion and
// the down-mix happen on the device.
//
// The chunk walk restates what `wave.Wave_read.initfp` accepts and rejects (CPython 3.12
// Lib/wave.py): RIFF/WAVE magic, little-endian chunks padded to even sizes, 'fmt ' before 'data',
// format tag PCM (1) or EXTENSIBLE (0xFFFE), sample width (bits + 7) / 8, n_frames =
// data_bytes / (channels * width); the walk stops at the first 'data' chunk.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/dspfront.h"

namespace {

uint32_t le32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint32_t le16(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

bool read_at(int fd, int64_t off, void* dst, size_t n) {
  unsigned char* d = static_cast<unsigned char*>(dst);
  while (n) {
    const ssize_t r = pread(fd, d, n, (off_t)off);
    if (r <= 0) return false;
    d += r; off += r; n -= (size_t)r;
  }
  return true;
}

void scan_one(const char* path, dsp_wav_info* w) {
  std::memset(w, 0, sizeof *w);
  const int fd = open(path, O_RDONLY | O_CLOEXEC);
  if (fd < 0) { w->status = DSP_WAV_ERR_OPEN; return; }
  struct stat st;
  if (fstat(fd, &st) != 0) { close(fd); w->status = DSP_WAV_ERR_OPEN; return; }
  const int64_t fsize = (int64_t)st.st_size;
  unsigned char h[24];
  int status = DSP_WAV_OK;
  do {
    if (fsize < 12 || !read_at(fd, 0, h, 12)) { status = DSP_WAV_ERR_NOT_RIFF; break; }
    if (std::memcmp(h, "RIFF", 4) != 0) { status = DSP_WAV_ERR_NOT_RIFF; break; }      // wave.Error: file does not start with RIFF id
    if (std::memcmp(h + 8, "WAVE", 4) != 0) { status = DSP_WAV_ERR_NOT_WAVE; break; }  // wave.Error: not a WAVE file
    // the RIFF chunk's own size bounds the walk, like Chunk.read() does
    int64_t riff_end = 8 + (int64_t)le32(h + 4);
    if (riff_end > fsize) riff_end = fsize;
    int64_t pos = 12;
    bool have_fmt = false, have_data = false;
    while (pos + 8 <= riff_end) {
      if (!read_at(fd, pos, h, 8)) break;
      const int64_t csize = (int64_t)le32(h + 4);
      const int64_t body = pos + 8;
      if (std::memcmp(h, "fmt ", 4) == 0) {
        if (csize < 14 || body + 14 > fsize || !read_at(fd, body, h + 8, 14)) { status = DSP_WAV_ERR_TRUNCATED; break; }   // EOFError in wave
        const uint32_t tag = le16(h + 8);
        w->channels = (int32_t)le16(h + 10);
        w->sample_rate = (int32_t)le32(h + 12);
        if (tag != 1u && tag != 0xFFFEu) { status = DSP_WAV_ERR_FORMAT; break; }         // wave.Error: unknown format
        unsigned char b[2];
        if (csize < 16 || body + 16 > fsize || !read_at(fd, body + 14, b, 2)) { status = DSP_WAV_ERR_TRUNCATED; break; }
        if (tag == 0xFFFEu) {                                                             // wave.Error: unknown extended format
          static const unsigned char kPcmGuid[16] = {0x01, 0x00, 0x00, 0x00, 0x00, 0x00, 0x10, 0x00, 0x80, 0x00, 0x00, 0xaa, 0x00, 0x38, 0x9b, 0x71};
          unsigned char g[16];
          if (csize < 40 || body + 40 > fsize || !read_at(fd, body + 24, g, 16)) { status = DSP_WAV_ERR_TRUNCATED; break; }
          if (std::memcmp(g, kPcmGuid, 16) != 0) { status = DSP_WAV_ERR_FORMAT; break; }
        }
        w->sample_width = (int32_t)((le16(b) + 7) / 8);
        if (w->sample_width == 0) { status = DSP_WAV_ERR_WIDTH; break; }                 // wave.Error: bad sample width
        if (w->channels == 0) { status = DSP_WAV_ERR_CHANNELS; break; }                  // wave.Error: bad # of channels
        have_fmt = true;
      } else if (std::memcmp(h, "data", 4) == 0) {
        if (!have_fmt) { status = DSP_WAV_ERR_ORDER; break; }                            // wave.Error: data chunk before fmt chunk
        const int64_t fsz = (int64_t)w->channels * w->sample_width;
        w->n_frames = csize / fsz;
        w->data_offset = body;
        // readframes(n_frames) returns what the file really holds (Chunk.read stops at the end of the file)
        int64_t avail = riff_end - body;
        if (avail < 0) avail = 0;
        int64_t want = w->n_frames * fsz;
        w->data_bytes = want < avail ? want : avail;
        have_data = true;
        break;
      }
      pos = body + csize + (csize & 1);      // chunks are padded to even sizes
    }
    if (status == DSP_WAV_OK && !(have_fmt && have_data)) status = DSP_WAV_ERR_MISSING;  // wave.Error: fmt chunk and/or data chunk missing
  } while (false);
  close(fd);
  w->status = status;
}

template <class F>
void parallel_for(int64_t n, int threads, F f) {
  if (threads < 1) threads = 1;
  if (threads > n) threads = (int)n;
  if (threads <= 1) { for (int64_t i = 0; i < n; ++i) f(i); return; }
  std::atomic<int64_t> next{0};
  std::vector<std::thread> pool;
  pool.reserve(threads);
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&]() { for (int64_t i = next.fetch_add(1); i < n; i = next.fetch_add(1)) f(i); });
  for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

int dsp_wav_scan(const char* const* paths, int64_t n_files, int32_t threads, dsp_wav_info* info) {
  if (n_files < 0 || (n_files > 0 && (!paths || !info))) return DSP_ERR_INVALID;
  parallel_for(n_files, threads, [&](int64_t i) { scan_one(paths[i], &info[i]); });
  return DSP_OK;
}

int dsp_wav_read(const char* const* paths, int64_t n_files, int32_t threads, dsp_wav_info* info,
                 const int64_t* dst_byte_offsets, void* dst, int64_t dst_bytes) {
  if (n_files < 0 || (n_files > 0 && (!paths || !info || !dst_byte_offsets || !dst))) return DSP_ERR_INVALID;
  for (int64_t i = 0; i < n_files; ++i)
    if (info[i].status == DSP_WAV_OK && (dst_byte_offsets[i] < 0 || dst_byte_offsets[i] + info[i].data_bytes > dst_bytes)) return DSP_ERR_INVALID;
  parallel_for(n_files, threads, [&](int64_t i) {
    dsp_wav_info& w = info[i];
    if (w.status != DSP_WAV_OK || w.data_bytes == 0) return;
    const int fd = open(paths[i], O_RDONLY | O_CLOEXEC);
    if (fd < 0) { w.status = DSP_WAV_ERR_OPEN; return; }
    if (!read_at(fd, w.data_offset, static_cast<unsigned char*>(dst) + dst_byte_offsets[i], (size_t)w.data_bytes)) w.status = DSP_WAV_ERR_TRUNCATED;
    close(fd);
  });
  return DSP_OK;
}

int dsp_host_alloc(int64_t bytes, void** out) {
  if (!out || bytes < 0) return DSP_ERR_INVALID;
  *out = nullptr;
  if (bytes == 0) return DSP_OK;
  return cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault) == cudaSuccess ? DSP_OK : DSP_ERR_NOMEM;
}

int dsp_host_free(void* p) {
  if (!p) return DSP_OK;
  return cudaFreeHost(p) == cudaSuccess ? DSP_OK : DSP_ERR_CUDA;
}

}  // extern "C"
