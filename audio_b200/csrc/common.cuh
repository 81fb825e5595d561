// Shared declarations for libb200audio (sm_100a).  No torch headers anywhere in csrc/.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/b200audio.h"

namespace b200a {

constexpr uint32_t kWsMagic = 0xB200A0D1u;
constexpr int kMaxStages = 16;
constexpr int kMaxFft = 8192;

// Device-side header at the start of a front-end workspace.
struct WsHeader {
  uint32_t magic;
  int32_t n_fft;
  int32_t n_bins;
  int32_t n_mels;
  int32_t n_mfcc;
  float scale;  // frame_length / window normalisation folded into one factor
  int32_t reserved[10];
};
static_assert(sizeof(WsHeader) == 64, "header is 64 bytes");

// Byte offsets of the tables inside a front-end workspace (host + device agree via this struct).
struct WsLayout {
  size_t header, window, twiddle, bands, fb, dct, total;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline WsLayout ws_layout(const b200a_frontend_desc& d) {
  WsLayout l{};
  const size_t n_bins = d.onesided ? d.n_fft / 2 + 1 : d.n_fft;
  size_t off = 0;
  l.header = off;
  off = align_up(off + sizeof(WsHeader), 256);
  l.window = off;
  off = align_up(off + sizeof(float) * d.n_fft, 256);
  l.twiddle = off;
  off = align_up(off + sizeof(float2) * d.n_fft, 256);
  l.bands = off;
  off = align_up(off + sizeof(int2) * (d.n_mels > 0 ? d.n_mels : 1), 256);
  l.fb = off;
  off = align_up(off + sizeof(float) * n_bins * (d.n_mels > 0 ? d.n_mels : 0), 256);
  l.dct = off;
  off = align_up(off + sizeof(float) * (size_t)(d.n_mels > 0 ? d.n_mels : 0) * (d.n_mfcc > 0 ? d.n_mfcc : 0), 256);
  l.total = off;
  return l;
}

inline int launch_status() { return cudaGetLastError() == cudaSuccess ? B200A_OK : B200A_ECUDA; }

// ---- device helpers -----------------------------------------------------------------------
// Index into the raw waveform row for sample i of the (constant `pad`-extended, then centre
// padded) signal; returns -1 for a zero sample.  Mirrors b200a_pad_index on the host.
__device__ __forceinline__ int64_t source_index(int64_t i, int64_t length, int pad, int half, int pad_mode) {
  const int64_t ext = length + 2 * (int64_t)pad;  // length after the constant pre-padding
  int64_t j = i - half;                           // index into the pre-padded signal
  if (j < 0 || j >= ext) {
    if (pad_mode == B200A_PAD_CONSTANT) return -1;
    if (pad_mode == B200A_PAD_REFLECT) {
      j = j < 0 ? -j : 2 * (ext - 1) - j;
    } else if (pad_mode == B200A_PAD_REPLICATE) {
      j = j < 0 ? 0 : ext - 1;
    } else {
      j %= ext;
      if (j < 0) j += ext;
    }
  }
  const int64_t s = j - pad;
  return (s >= 0 && s < length) ? s : -1;
}

__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
  // total order trick: non-negative floats compare like ints, negative like reversed uints
  if (v >= 0.f) {
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  } else {
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
  }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace b200a
