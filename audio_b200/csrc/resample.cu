// Polyphase windowed-sinc resampler.
//
// The reference evaluates  y[r][f*new' + j] = sum_i k[j][i] * xpad[r][f*orig' + i]  as a dense
// conv1d with a (new', 1, 2*width+orig') filter and stride orig'
// (src/torchaudio/functional/functional.py:1405-1432).  Almost all of every filter row is
// (numerically) zero: row j only has a contiguous run of ~2*lowpass_width*orig'/min(orig',new')
// taps around the position of output phase j.  `resample_prepare` finds that run per phase once,
// the kernels then touch only those taps -- no padded copy of the input, no (rows, new', frames)
// intermediate, output written already interleaved and truncated.
#include "common.cuh"

namespace b200a {

struct RsHeader {
  uint32_t magic;
  int32_t orig_r, new_r, width, taps, max_support;
  int32_t reserved[10];
};
static_assert(sizeof(RsHeader) == 64, "header is 64 bytes");

struct RsLayout {
  size_t header, support, total;
};

inline RsLayout rs_layout(int new_r, int taps) {
  (void)taps;
  RsLayout l{};
  size_t off = 0;
  l.header = off;
  off = align_up(off + sizeof(RsHeader), 256);
  l.support = off;
  off = align_up(off + sizeof(int2) * (size_t)new_r, 256);
  l.total = off;
  return l;
}

// One warp per phase: [first, last] index of taps with |k| > 1e-12 * max|k| of that row.
__global__ void resample_support_kernel(const float* __restrict__ kernel, int new_r, int taps, int orig_r,
                                        int width, RsHeader* hdr, int2* support) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    hdr->magic = kWsMagic;
    hdr->orig_r = orig_r;
    hdr->new_r = new_r;
    hdr->width = width;
    hdr->taps = taps;
  }
  if (j >= new_r) return;
  const float* row = kernel + (size_t)j * taps;
  float mx = 0.f;
  for (int i = lane; i < taps; i += 32) mx = fmaxf(mx, fabsf(row[i]));
  mx = warp_max(mx);
  const float thr = mx * 1e-12f;
  int lo = taps, hi = -1;
  for (int i = lane; i < taps; i += 32) {
    if (fabsf(row[i]) > thr) {
      lo = min(lo, i);
      hi = max(hi, i);
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) {
    if (hi < 0) { lo = 0; hi = -1; }
    support[j] = make_int2(lo, hi - lo + 1);
    atomicMax(&hdr->max_support, hi - lo + 1);
  }
}

// Straightforward one-output-per-thread kernel (any ratio).  Consecutive threads are consecutive
// output samples, i.e. consecutive phases of the same input neighbourhood: input loads hit L1.
__global__ void __launch_bounds__(256)
resample_direct_kernel(const float* __restrict__ wave, int64_t length, int64_t row_stride,
                       const float* __restrict__ kernel, const int2* __restrict__ support, int orig_r,
                       int new_r, int width, int taps, float* __restrict__ out, int64_t out_row_stride,
                       int64_t out_len) {
  const int64_t row = blockIdx.y;
  const float* __restrict__ x = wave + row * row_stride;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < out_len; n += (int64_t)gridDim.x * blockDim.x) {
    const int64_t f = n / new_r;
    const int j = (int)(n - f * new_r);
    const int2 sp = support[j];
    const int64_t base = f * orig_r - width + sp.x;  // input index of the first live tap
    const float* __restrict__ k = kernel + (size_t)j * taps + sp.x;
    float acc = 0.f;
    for (int i = 0; i < sp.y; ++i) {
      const int64_t s = base + i;
      const float v = (s >= 0 && s < length) ? x[s] : 0.f;
      acc = fmaf(k[i], v, acc);
    }
    out[row * out_row_stride + n] = acc;
  }
}

size_t resample_workspace_bytes_impl(int new_r, int taps) { return rs_layout(new_r, taps).total; }

int resample_prepare_impl(const float* kernel, int orig_r, int new_r, int width, void* ws, size_t ws_bytes,
                          cudaStream_t stream) {
  if (kernel == nullptr || ws == nullptr || orig_r < 1 || new_r < 1 || width < 0) return B200A_EINVAL;
  const int taps = 2 * width + orig_r;
  const RsLayout l = rs_layout(new_r, taps);
  if (ws_bytes < l.total) return B200A_EWORKSPACE;
  unsigned char* base = static_cast<unsigned char*>(ws);
  if (cudaMemsetAsync(base + l.header, 0, sizeof(RsHeader), stream) != cudaSuccess) return B200A_ECUDA;
  resample_support_kernel<<<(new_r + 7) / 8, 256, 0, stream>>>(kernel, new_r, taps, orig_r, width,
                                                               reinterpret_cast<RsHeader*>(base + l.header),
                                                               reinterpret_cast<int2*>(base + l.support));
  return launch_status();
}

int resample_run_impl(const void* ws, const float* kernel, int orig_r, int new_r, int width, const float* wave,
                      int64_t rows, int64_t length, int64_t row_stride, float* out, int64_t out_row_stride,
                      int64_t out_len, cudaStream_t stream) {
  if (orig_r < 1 || new_r < 1 || width < 0 || rows < 0 || length < 0 || out_len < 0) return B200A_EINVAL;
  if (rows == 0 || out_len == 0) return B200A_OK;  // empty batch: pointers may be null
  if (ws == nullptr || kernel == nullptr || wave == nullptr || out == nullptr) return B200A_EINVAL;
  if (rows > 65535) return B200A_EUNSUPPORTED;
  const int taps = 2 * width + orig_r;
  const RsLayout l = rs_layout(new_r, taps);
  const unsigned char* base = static_cast<const unsigned char*>(ws);
  unsigned bx = (unsigned)((out_len + 255) / 256);
  if (bx > 4096) bx = 4096;
  dim3 grid(bx, (unsigned)rows);
  resample_direct_kernel<<<grid, 256, 0, stream>>>(wave, length, row_stride, kernel,
                                                  reinterpret_cast<const int2*>(base + l.support), orig_r, new_r,
                                                  width, taps, out, out_row_stride, out_len);
  return launch_status();
}

}  // namespace b200a
