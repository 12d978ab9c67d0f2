// Reader for the feature files of the reference's pipeline: NumPy .npy (format 1.0 - 3.0), little-endian
// float32, C order, shape [T, C] or [1, T, C] (prepare_data.py:87 saves last_hidden_state).  Used by
// wavthruvec_pytorch_b200/files.py from reader threads: a ctypes call releases the GIL, so the file
// reads of a batch really run in parallel and land directly in the caller's (pinned) buffer.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.h"

using namespace v2w;

namespace {

struct NpyInfo {
  int64_t rows = 0, cols = 0;
  long data_offset = 0;
};

// Parses the header; returns 0 or a V2W_ERR_* code with the error text set.
int parse_header(FILE* f, const char* path, NpyInfo* out) {
  unsigned char pre[12];
  if (fread(pre, 1, 10, f) != 10 || memcmp(pre, "\x93NUMPY", 6) != 0) {
    set_error("%s: not a .npy file", path);
    return V2W_ERR_INVALID;
  }
  const int major = pre[6];
  size_t hlen = 0, pos = 10;
  if (major == 1) {
    hlen = pre[8] | (pre[9] << 8);
  } else if (major == 2 || major == 3) {
    if (fread(pre + 10, 1, 2, f) != 2) {
      set_error("%s: truncated header", path);
      return V2W_ERR_INVALID;
    }
    hlen = (size_t)pre[8] | ((size_t)pre[9] << 8) | ((size_t)pre[10] << 16) | ((size_t)pre[11] << 24);
    pos = 12;
  } else {
    set_error("%s: unsupported .npy version %d", path, major);
    return V2W_ERR_INVALID;
  }
  if (hlen == 0 || hlen > (1u << 20)) {
    set_error("%s: implausible header length %zu", path, hlen);
    return V2W_ERR_INVALID;
  }
  std::string h(hlen, '\0');
  if (fread(&h[0], 1, hlen, f) != hlen) {
    set_error("%s: truncated header", path);
    return V2W_ERR_INVALID;
  }
  if (h.find("'descr': '<f4'") == std::string::npos && h.find("\"descr\": \"<f4\"") == std::string::npos) {
    set_error("%s: features must be little-endian float32 ('<f4')", path);
    return V2W_ERR_INVALID;
  }
  if (h.find("'fortran_order': False") == std::string::npos) {
    set_error("%s: Fortran-ordered arrays are not supported", path);
    return V2W_ERR_INVALID;
  }
  const size_t sp = h.find("'shape':");
  const size_t lp = sp == std::string::npos ? sp : h.find('(', sp);
  const size_t rp = lp == std::string::npos ? lp : h.find(')', lp);
  if (rp == std::string::npos) {
    set_error("%s: no shape in the header", path);
    return V2W_ERR_INVALID;
  }
  int64_t dims[4];
  int nd = 0;
  const char* c = h.c_str() + lp + 1;
  const char* end = h.c_str() + rp;
  while (c < end && nd < 4) {
    while (c < end && (*c == ' ' || *c == ',')) ++c;
    if (c >= end) break;
    char* nx = nullptr;
    const long long v = strtoll(c, &nx, 10);
    if (nx == c) break;
    dims[nd++] = v;
    c = nx;
  }
  if (nd == 3 && dims[0] == 1) {
    dims[0] = dims[1]; dims[1] = dims[2]; nd = 2;
  }
  if (nd != 2 || dims[0] < 1 || dims[1] < 1) {
    set_error("%s: expected shape [T, C] or [1, T, C]", path);
    return V2W_ERR_INVALID;
  }
  out->rows = dims[0];
  out->cols = dims[1];
  out->data_offset = (long)(pos + hlen);
  return V2W_OK;
}

}  // namespace

extern "C" {

int32_t v2w_npy_shape(const char* path, int64_t* rows, int64_t* cols) {
  if (!path || !rows || !cols) {
    set_error("v2w_npy_shape: null argument");
    return V2W_ERR_INVALID;
  }
  FILE* f = fopen(path, "rb");
  if (!f) {
    set_error("%s: cannot open", path);
    return V2W_ERR_INVALID;
  }
  NpyInfo info;
  const int rc = parse_header(f, path, &info);
  fclose(f);
  if (rc) return rc;
  *rows = info.rows;
  *cols = info.cols;
  return V2W_OK;
}

int32_t v2w_npy_read_f32(const char* path, float* dst, int64_t max_rows, int64_t cols, int64_t* rows) {
  if (!path || !dst || !rows) {
    set_error("v2w_npy_read_f32: null argument");
    return V2W_ERR_INVALID;
  }
  FILE* f = fopen(path, "rb");
  if (!f) {
    set_error("%s: cannot open", path);
    return V2W_ERR_INVALID;
  }
  NpyInfo info;
  int rc = parse_header(f, path, &info);
  if (!rc && info.cols != cols) {
    set_error("%s: %lld feature channels, expected %lld", path, (long long)info.cols, (long long)cols);
    rc = V2W_ERR_INVALID;
  }
  if (!rc && info.rows > max_rows) {
    set_error("%s: %lld frames do not fit the %lld-frame buffer", path, (long long)info.rows, (long long)max_rows);
    rc = V2W_ERR_INVALID;
  }
  if (!rc) {
    const size_t n = (size_t)info.rows * (size_t)info.cols;
    if (fseek(f, info.data_offset, SEEK_SET) != 0 || fread(dst, 4, n, f) != n) {
      set_error("%s: truncated data", path);
      rc = V2W_ERR_INVALID;
    }
  }
  fclose(f);
  if (rc) return rc;
  *rows = info.rows;
  return V2W_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Batch forms: one call per BATCH of utterances instead of one per file, worker threads inside the
// library.  The Python side of files.py makes three ctypes calls per batch (no per-file Python work,
// no intermediate copy: every payload lands at its final offset of the pinned row arena).
// ---------------------------------------------------------------------------------------------
namespace {

// Runs fn(i) for i in [0, n) on `threads` workers; returns the first failing item's code, with its
// error text re-raised on the calling thread (set_error is thread-local).
template <typename Fn>
int for_each_parallel(int32_t n, int32_t threads, Fn fn) {
  std::atomic<int32_t> next(0);
  std::atomic<int> first_rc(0);
  std::mutex mu;
  std::string first_msg;
  auto work = [&]() {
    for (;;) {
      const int32_t i = next.fetch_add(1);
      if (i >= n || first_rc.load() != 0) return;
      const int rc = fn(i);
      if (rc) {
        std::lock_guard<std::mutex> lock(mu);
        if (first_rc.load() == 0) {
          first_rc.store(rc);
          first_msg = v2w_last_error();
        }
      }
    }
  };
  const int32_t nt = std::max(1, std::min(threads, n));
  std::vector<std::thread> pool;
  for (int32_t t = 1; t < nt; ++t) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  if (first_rc.load()) set_error("%s", first_msg.c_str());
  return first_rc.load();
}

}  // namespace

extern "C" {

int32_t v2w_npy_shapes(const char* const* paths, int32_t n, int64_t* rows, int64_t* cols, int32_t threads) {
  if (!paths || !rows || !cols || n < 0) {
    set_error("v2w_npy_shapes: bad argument");
    return V2W_ERR_INVALID;
  }
  return for_each_parallel(n, threads, [&](int32_t i) { return (int)v2w_npy_shape(paths[i], rows + i, cols + i); });
}

int32_t v2w_npy_read_batch(const char* const* paths, int32_t n, float* dst, const int64_t* row_off,
                           const int64_t* rows_expected, int64_t cols, int32_t threads) {
  if (!paths || !dst || !row_off || !rows_expected || n < 0 || cols < 1) {
    set_error("v2w_npy_read_batch: bad argument");
    return V2W_ERR_INVALID;
  }
  return for_each_parallel(n, threads, [&](int32_t i) {
    int64_t got = 0;
    int rc = v2w_npy_read_f32(paths[i], dst + row_off[i] * cols, rows_expected[i], cols, &got);
    if (!rc && got != rows_expected[i]) {
      set_error("%s: %lld frames, the batch layout expected %lld (file changed since its header was read)",
                paths[i], (long long)got, (long long)rows_expected[i]);
      rc = V2W_ERR_INVALID;
    }
    return rc;
  });
}

int32_t v2w_wav_write_batch(const char* const* paths, int32_t n, const int16_t* pcm, const int64_t* sample_off,
                            const int64_t* samples, int32_t sampling_rate, int32_t threads) {
  if (!paths || !pcm || !sample_off || !samples || n < 0 || sampling_rate < 1) {
    set_error("v2w_wav_write_batch: bad argument");
    return V2W_ERR_INVALID;
  }
  return for_each_parallel(n, threads, [&](int32_t i) {
    const uint64_t bytes = (uint64_t)samples[i] * 2;
    if (samples[i] < 0 || bytes > 0xffffffffull - 36) {
      set_error("%s: %lld samples do not fit a RIFF file", paths[i], (long long)samples[i]);
      return (int)V2W_ERR_INVALID;
    }
    FILE* f = fopen(paths[i], "wb");
    if (!f) {
      set_error("%s: cannot create", paths[i]);
      return (int)V2W_ERR_INVALID;
    }
    // 44-byte canonical header: PCM, mono, 16 bit (what scipy.io.wavfile.write emits for int16)
    unsigned char h[44];
    auto u32 = [&](int at, uint32_t v) { h[at] = v & 255; h[at + 1] = (v >> 8) & 255; h[at + 2] = (v >> 16) & 255; h[at + 3] = (v >> 24) & 255; };
    auto u16 = [&](int at, uint32_t v) { h[at] = v & 255; h[at + 1] = (v >> 8) & 255; };
    memcpy(h, "RIFF", 4); u32(4, (uint32_t)(36 + bytes)); memcpy(h + 8, "WAVEfmt ", 8);
    u32(16, 16); u16(20, 1); u16(22, 1); u32(24, (uint32_t)sampling_rate); u32(28, (uint32_t)sampling_rate * 2);
    u16(32, 2); u16(34, 16); memcpy(h + 36, "data", 4); u32(40, (uint32_t)bytes);
    const bool ok = fwrite(h, 1, 44, f) == 44 && fwrite(pcm + sample_off[i], 2, (size_t)samples[i], f) == (size_t)samples[i];
    if (fclose(f) != 0 || !ok) {
      set_error("%s: short write", paths[i]);
      return (int)V2W_ERR_INVALID;
    }
    return (int)V2W_OK;
  });
}

}  // extern "C"
