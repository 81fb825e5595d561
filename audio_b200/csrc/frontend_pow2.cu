// Register-FFT fast path of the fused front end (n_fft = 1024, one-sided, real output stages).
//
// One WARP transforms one PAIR of consecutive frames (a, b) of an utterance:
//   z[n] = w[n] * (a[n] + i b[n]),  n = lane + 32 j          -- 32 complex values per lane, coalesced loads
//   pass 1  32-point DFT over j in registers (radix-2 DIT, compile-time twiddles, FMA-form butterflies)
//   twiddle W_1024^(lane * k2) from a shared 32x32 table
//   transpose through a per-warp padded shared tile (the only exchange of the whole FFT)
//   pass 2  32-point DFT over the former lane index in registers  -> Z[lane + 32 k1]
//   un-pack the two real spectra with one shuffle per needed value:  A = (Z[k] + conj Z[N-k]) / 2,
//                                                                    B = (Z[k] - conj Z[N-k]) / 2i
//   |.|^p -> global (Spectrogram) or -> the warp's shared tile -> banded mel (-> dB / log) -> global.
// Nothing but the waveform is read from HBM and nothing but the final features is written.
//
// Reference semantics: src/torchaudio/functional/functional.py:54-145 and
// transforms/_transforms.py:403-415, :701-705 (see frontend_generic.cu for the any-size path).
#include <utility>

#include "common.cuh"

namespace b200a {

namespace {

constexpr int kN = 1024;
constexpr int kBins = kN / 2 + 1;
constexpr int kWarps = 8;
constexpr int kTileLd = 33;                      // float2 row pitch of the transpose tile (bank-conflict free)
constexpr int kTileFloat2 = 32 * kTileLd;        // per warp
constexpr int kPowLd = 520;                      // float pitch of one frame's power row inside the same tile
constexpr int kCsrSmemMax = 6144;                // filterbank weights kept in shared memory up to this many

struct Pow2Extra {  // tables appended to the generic workspace
  size_t tw2d, csr_off, csr_lo, csr_w, total;
};

inline Pow2Extra pow2_layout(const b200a_frontend_desc& d, size_t base) {
  Pow2Extra e{};
  const size_t n_bins = d.n_fft / 2 + 1;
  const size_t n_mels = d.n_mels > 0 ? d.n_mels : 0;
  size_t off = base;
  e.tw2d = off;
  off = align_up(off + sizeof(float2) * 32 * 32, 256);
  e.csr_off = off;
  off = align_up(off + sizeof(int) * (n_mels + 1), 256);
  e.csr_lo = off;
  off = align_up(off + sizeof(int) * (n_mels + 1), 256);
  e.csr_w = off;
  off = align_up(off + sizeof(float) * n_bins * n_mels, 256);
  e.total = off;
  return e;
}

bool pow2_applicable(const b200a_frontend_desc& d) { return d.n_fft == kN && d.onesided != 0; }

// ---- compile-time helpers -----------------------------------------------------------------
template <int... Is, typename F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, F&& f) {
  (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
  static_for_impl(std::make_integer_sequence<int, N>{}, static_cast<F&&>(f));
}

__host__ __device__ constexpr int brev5(int v) {
  return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4);
}

// cos / sin of 2 pi k / 32, k = 0..16
__device__ constexpr float kCos32[17] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                                         0.70710678118654757f, 0.55557023301960229f, 0.38268343236508984f,
                                         0.19509032201612833f, 0.f, -0.19509032201612819f, -0.38268343236508973f,
                                         -0.55557023301960196f, -0.70710678118654746f, -0.83146961230254535f,
                                         -0.92387953251128674f, -0.98078528040323043f, -1.f};
__device__ constexpr float kSin32[17] = {0.f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
                                         0.70710678118654746f, 0.83146961230254524f, 0.92387953251128674f,
                                         0.98078528040323043f, 1.f, 0.98078528040323043f, 0.92387953251128674f,
                                         0.83146961230254546f, 0.70710678118654757f, 0.55557023301960218f,
                                         0.38268343236508989f, 0.19509032201612861f, 0.f};

// DIT butterfly (a, b) -> (a + W b, a - W b), W = exp(-2 pi i E / 32), E in [0, 16).
// General case in FMA form: 3 instructions per real output pair instead of 4.
template <int E>
__device__ __forceinline__ void bfly(float2& a, float2& b) {
  if constexpr (E == 0) {
    const float2 t = b;
    b = make_float2(a.x - t.x, a.y - t.y);
    a = make_float2(a.x + t.x, a.y + t.y);
  } else if constexpr (E == 8) {  // W = -i : W b = (b.y, -b.x)
    const float2 t = b;
    b = make_float2(a.x - t.y, a.y + t.x);
    a = make_float2(a.x + t.y, a.y - t.x);
  } else if constexpr (E == 4) {  // W = (1 - i)/sqrt2 : W b = c ((bx + by), (by - bx))
    constexpr float c = 0.70710678118654752f;
    const float tr = b.x + b.y, ti = b.y - b.x;
    b = make_float2(fmaf(-c, tr, a.x), fmaf(-c, ti, a.y));
    a = make_float2(fmaf(c, tr, a.x), fmaf(c, ti, a.y));
  } else if constexpr (E == 12) {  // W = (-1 - i)/sqrt2 : W b = c ((by - bx), -(bx + by))
    constexpr float c = 0.70710678118654752f;
    const float tr = b.y - b.x, ti = -(b.x + b.y);
    b = make_float2(fmaf(-c, tr, a.x), fmaf(-c, ti, a.y));
    a = make_float2(fmaf(c, tr, a.x), fmaf(c, ti, a.y));
  } else {
    constexpr float wr = kCos32[E], wi = -kSin32[E];
    const float pr = fmaf(wr, b.x, fmaf(-wi, b.y, a.x));
    const float pi = fmaf(wr, b.y, fmaf(wi, b.x, a.y));
    b = make_float2(fmaf(2.f, a.x, -pr), fmaf(2.f, a.y, -pi));
    a = make_float2(pr, pi);
  }
}

// In-register 32-point DFT.  Input a[brev5(j)] = x[j]; output a[k] = X[k] (natural order).
__device__ __forceinline__ void fft32(float2 (&a)[32]) {
  static_for<5>([&](auto si) {
    constexpr int len = 2 << decltype(si)::value;  // 2, 4, 8, 16, 32
    constexpr int half = len / 2;
    static_for<16>([&](auto bi) {
      constexpr int b = decltype(bi)::value;
      constexpr int i = (b / half) * len, j = b % half;
      bfly<j*(32 / len)>(a[i + j], a[i + j + half]);
    });
  });
}

struct Pow2Params {
  const float* wave;
  int64_t length, row_stride, frames, pairs_per_row, total_pairs;
  float* out;
  float* group_max;
  int64_t rows_per_group;
  const float* window;   // [1024] centre padded
  const float2* tw2d;    // [32][32]  W_1024^(k2 * g) at [k2][g]
  const int* csr_off;    // [n_mels + 1]
  const int* csr_lo;     // [n_mels]
  const float* csr_w;    // compacted non-zero runs of the filterbank columns
  const WsHeader* hdr;
  int hop, pad, center, pad_mode, n_mels;
  int stage, log_mels;
  float power, db_mult, db_amin, db_offset;
};

template <int POWER_MODE>  // 2: |.|^2, 1: |.|, 0: general exponent
__device__ __forceinline__ float pow_of(float re, float im, float power) {
  if constexpr (POWER_MODE == 2) return fmaf(re, re, im * im);
  const float mag = hypotf(re, im);
  if constexpr (POWER_MODE == 1) return mag;
  return powf(mag, power);
}

template <int POWER_MODE>
__global__ void __launch_bounds__(kWarps * 32, 1) stft1024_kernel(const Pow2Params p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* s_tw = reinterpret_cast<float2*>(smem_raw);                 // [32][32]
  float2* s_tile_all = s_tw + 32 * 32;                                // [kWarps][32 * 33]
  int* s_off = reinterpret_cast<int*>(s_tile_all + kWarps * kTileFloat2);  // [n_mels + 1]
  int* s_lo = s_off + (p.n_mels + 1);                                 // [n_mels]
  float* s_w = reinterpret_cast<float*>(s_lo + p.n_mels + 1);         // [csr_total] (if it fits)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 32 * 32; i += blockDim.x) s_tw[i] = p.tw2d[i];
  const bool mel_stage = p.stage >= B200A_STAGE_MEL;
  const int csr_total = mel_stage ? p.csr_off[p.n_mels] : 0;
  const bool w_in_smem = mel_stage && csr_total <= kCsrSmemMax;
  if (mel_stage) {
    for (int i = tid; i <= p.n_mels; i += blockDim.x) s_off[i] = p.csr_off[i];
    for (int i = tid; i < p.n_mels; i += blockDim.x) s_lo[i] = p.csr_lo[i];
    if (w_in_smem)
      for (int i = tid; i < csr_total; i += blockDim.x) s_w[i] = p.csr_w[i];
  }
  __syncthreads();
  const float* __restrict__ fbw = w_in_smem ? s_w : p.csr_w;

  float2* tile = s_tile_all + warp * kTileFloat2;
  float* ptile = reinterpret_cast<float*>(tile);  // power rows of frames a, b at 0 and kPowLd

  // window (x 1/2 from the un-packing, x the normalisation scale) for n = lane + 32 j
  float wreg[32];
  {
    const float hs = 0.5f * p.hdr->scale;
#pragma unroll
    for (int j = 0; j < 32; ++j) wreg[j] = p.window[lane + 32 * j] * hs;
  }
  const int half = p.center ? kN / 2 : 0;
  float local_max = -CUDART_INF_F;
  int64_t cur_group = -1;

  for (int64_t u = (int64_t)blockIdx.x * kWarps + warp; u < p.total_pairs; u += (int64_t)gridDim.x * kWarps) {
    const int64_t row = u / p.pairs_per_row;
    const int64_t pr = u - row * p.pairs_per_row;
    const int64_t ta = 2 * pr, tb = ta + 1;
    const bool has_b = tb < p.frames;
    const float* __restrict__ x = p.wave + row * p.row_stride;
    if (p.stage == B200A_STAGE_FEAT && p.group_max != nullptr) {
      const int64_t grp = row / p.rows_per_group;
      if (grp != cur_group) {  // flush the running maximum when the warp moves to another top_db group
        const float mx = warp_max(local_max);
        if (lane == 0 && cur_group >= 0 && mx > -CUDART_INF_F) atomic_max_f32(p.group_max + cur_group, mx);
        local_max = -CUDART_INF_F;
        cur_group = grp;
      }
    }
    const int64_t sa = ta * p.hop - half - p.pad;  // first raw sample of frame a
    const int64_t sb = sa + p.hop;

    float2 a[32];
    const bool interior = sa >= 0 && (has_b ? sb : sa) + kN <= p.length;
    if (interior) {
      static_for<32>([&](auto ji) {
        constexpr int j = decltype(ji)::value;
        const float va = __ldg(x + sa + lane + 32 * j);
        const float vb = has_b ? __ldg(x + sb + lane + 32 * j) : 0.f;
        a[brev5(j)] = make_float2(va * wreg[j], vb * wreg[j]);
      });
    } else {
      // edge pair (padding / reflection / ragged end): gather through the warp's tile with a
      // rolled loop so the index arithmetic is not replicated 64 times in the instruction stream
#pragma unroll 1
      for (int j = 0; j < 32; ++j) {
        const int n = lane + 32 * j;
        const int64_t ia = source_index(ta * p.hop + n, p.length, p.pad, half, p.pad_mode);
        const int64_t ib = has_b ? source_index(tb * p.hop + n, p.length, p.pad, half, p.pad_mode) : -1;
        tile[n] = make_float2(ia >= 0 ? __ldg(x + ia) : 0.f, ib >= 0 ? __ldg(x + ib) : 0.f);
      }
      __syncwarp();
      static_for<32>([&](auto ji) {
        constexpr int j = decltype(ji)::value;
        const float2 v = tile[lane + 32 * j];
        a[brev5(j)] = make_float2(v.x * wreg[j], v.y * wreg[j]);
      });
      __syncwarp();
    }

    fft32(a);  // a[k2] = Y[lane][k2]

    // twiddle + transpose: element (g = lane, k2) -> tile[k2][g]
    tile[lane] = a[0];
    static_for<31>([&](auto ki) {
      constexpr int k2 = decltype(ki)::value + 1;
      const float2 w = s_tw[k2 * 32 + lane];
      const float2 v = a[k2];
      tile[k2 * kTileLd + lane] = make_float2(fmaf(v.x, w.x, -v.y * w.y), fmaf(v.x, w.y, v.y * w.x));
    });
    __syncwarp();
    static_for<32>([&](auto gi) {
      constexpr int g = decltype(gi)::value;
      a[brev5(g)] = tile[lane * kTileLd + g];
    });
    __syncwarp();

    fft32(a);  // a[k1] = Z[lane + 32 k1]

    // ---- un-pack the two spectra: need Z[N - k], k = lane + 32 k1, k1 = 0..15 (+ bin 512 on lane 0)
    const int src = (32 - lane) & 31;
    float pa[17], pb[17];
    static_for<16>([&](auto ki) {
      constexpr int k1 = decltype(ki)::value;
      // lanes >= 1: Z[N-k] = Z[(32-lane) + 32 (31-k1)] sits on lane `src`, slot 31-k1
      float mr = __shfl_sync(0xffffffffu, a[31 - k1].x, src);
      float mi = __shfl_sync(0xffffffffu, a[31 - k1].y, src);
      if (lane == 0) {  // lane 0: Z[N-k] = Z[32 (32-k1)] is its own slot 32-k1 (slot 0 for k1 = 0)
        mr = a[(32 - k1) & 31].x;
        mi = a[(32 - k1) & 31].y;
      }
      const float zr = a[k1].x, zi = a[k1].y;
      pa[k1] = pow_of<POWER_MODE>(zr + mr, zi - mi, p.power);
      pb[k1] = pow_of<POWER_MODE>(zi + mi, mr - zr, p.power);
    });
    // bin 512 (lane 0, slot 16) is its own mirror: A = Re, B = Im  (x2 because wreg carries the 1/2)
    pa[16] = pow_of<POWER_MODE>(2.f * a[16].x, 0.f, p.power);
    pb[16] = pow_of<POWER_MODE>(2.f * a[16].y, 0.f, p.power);

    if (p.stage == B200A_STAGE_POWER) {
      float* oa = p.out + (row * p.frames + ta) * kBins;
      float* ob = oa + kBins;
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) {
        oa[lane + 32 * k1] = pa[k1];
        if (has_b) ob[lane + 32 * k1] = pb[k1];
      }
      if (lane == 0) {
        oa[512] = pa[16];
        if (has_b) ob[512] = pb[16];
      }
      continue;
    }

    // ---- mel projection from the warp's shared power rows -------------------------------------
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      ptile[lane + 32 * k1] = pa[k1];
      ptile[kPowLd + lane + 32 * k1] = pb[k1];
    }
    if (lane == 0) {
      ptile[512] = pa[16];
      ptile[kPowLd + 512] = pb[16];
    }
    __syncwarp();
    float* oa = p.out + (row * p.frames + ta) * p.n_mels;
    for (int m = lane; m < p.n_mels; m += 32) {
      const int off = s_off[m], len = s_off[m + 1] - off, lo = s_lo[m];
      const float* ra = ptile + lo;
      const float* rb = ra + kPowLd;
      const float* w = fbw + off;
      float acc_a = 0.f, acc_b = 0.f;
      for (int i = 0; i < len; ++i) {
        const float wv = w[i];
        acc_a = fmaf(ra[i], wv, acc_a);
        acc_b = fmaf(rb[i], wv, acc_b);
      }
      if (p.stage == B200A_STAGE_FEAT) {
        if (p.log_mels) {
          acc_a = logf(acc_a + 1e-6f);
          acc_b = logf(acc_b + 1e-6f);
        } else {
          acc_a = p.db_mult * log10f(fmaxf(acc_a, p.db_amin)) - p.db_offset;
          acc_b = p.db_mult * log10f(fmaxf(acc_b, p.db_amin)) - p.db_offset;
        }
        local_max = fmaxf(local_max, has_b ? fmaxf(acc_a, acc_b) : acc_a);
      }
      oa[m] = acc_a;
      if (has_b) oa[p.n_mels + m] = acc_b;
    }
    __syncwarp();
  }
  if (p.stage == B200A_STAGE_FEAT && p.group_max != nullptr && cur_group >= 0) {
    const float mx = warp_max(local_max);
    if (lane == 0 && mx > -CUDART_INF_F) atomic_max_f32(p.group_max + cur_group, mx);
  }
}

// ---- table preparation ------------------------------------------------------------------------
__global__ void prepare_tw2d_kernel(float2* tw2d) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // i = k2 * 32 + g
  if (i < 32 * 32) {
    const int k2 = i >> 5, g = i & 31;
    double s, c;
    sincospi(-2.0 * (double)(k2 * g) / (double)kN, &s, &c);
    tw2d[i] = make_float2((float)c, (float)s);
  }
}

// Compacts the non-zero run of every filterbank column (bands found by prepare_fbank_kernel).
__global__ void prepare_csr_kernel(const float* __restrict__ fb, const int2* __restrict__ bands, int n_bins, int n_mels,
                                   int* csr_off, int* csr_lo, float* csr_w) {
  __shared__ int s_total;
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int m = 0; m < n_mels; ++m) {
      csr_off[m] = acc;
      csr_lo[m] = bands[m].x;
      acc += bands[m].y - bands[m].x;
    }
    csr_off[n_mels] = acc;
    s_total = acc;
  }
  __syncthreads();
  (void)s_total;
  for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
    const int2 b = bands[m];
    const int off = csr_off[m];
    for (int k = b.x; k < b.y; ++k) csr_w[off + (k - b.x)] = fb[(size_t)k * n_mels + m];
  }
}

}  // namespace

size_t pow2_workspace_extra(const b200a_frontend_desc* d) {
  if (!pow2_applicable(*d)) return 0;
  const size_t base = ws_layout(*d).total;
  return pow2_layout(*d, base).total - base;
}

int pow2_prepare(const b200a_frontend_desc* d, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (!pow2_applicable(*d)) return B200A_OK;
  const WsLayout l = ws_layout(*d);
  const Pow2Extra e = pow2_layout(*d, l.total);
  if (ws_bytes < e.total) return B200A_EWORKSPACE;
  unsigned char* base = static_cast<unsigned char*>(ws);
  prepare_tw2d_kernel<<<4, 256, 0, stream>>>(reinterpret_cast<float2*>(base + e.tw2d));
  if (d->n_mels > 0) {
    prepare_csr_kernel<<<1, 128, 0, stream>>>(reinterpret_cast<const float*>(base + l.fb),
                                              reinterpret_cast<const int2*>(base + l.bands), d->n_fft / 2 + 1, d->n_mels,
                                              reinterpret_cast<int*>(base + e.csr_off),
                                              reinterpret_cast<int*>(base + e.csr_lo),
                                              reinterpret_cast<float*>(base + e.csr_w));
  }
  return launch_status();
}

template <int POWER_MODE>
static int launch_1024(const Pow2Params& p, size_t smem, int64_t grid, cudaStream_t stream) {
  if (cudaFuncSetAttribute(stft1024_kernel<POWER_MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) !=
      cudaSuccess)
    return B200A_ECUDA;
  stft1024_kernel<POWER_MODE><<<(unsigned)grid, kWarps * 32, smem, stream>>>(p);
  return launch_status();
}

int frontend_run_pow2(const b200a_frontend_desc* d, const void* ws, int stage, const float* wave, int64_t rows,
                      int64_t length, int64_t row_stride, int64_t frames, float* out, float* group_max,
                      int64_t rows_per_group, cudaStream_t stream) {
  if (!pow2_applicable(*d) || stage == B200A_STAGE_COMPLEX) return B200A_EUNSUPPORTED;
  const WsLayout l = ws_layout(*d);
  const Pow2Extra e = pow2_layout(*d, l.total);
  const unsigned char* base = static_cast<const unsigned char*>(ws);
  Pow2Params p{};
  p.wave = wave;
  p.length = length;
  p.row_stride = row_stride;
  p.frames = frames;
  p.pairs_per_row = (frames + 1) / 2;
  p.total_pairs = rows * p.pairs_per_row;
  p.out = out;
  p.group_max = group_max;
  p.rows_per_group = rows_per_group > 0 ? rows_per_group : 1;
  p.window = reinterpret_cast<const float*>(base + l.window);
  p.tw2d = reinterpret_cast<const float2*>(base + e.tw2d);
  p.csr_off = reinterpret_cast<const int*>(base + e.csr_off);
  p.csr_lo = reinterpret_cast<const int*>(base + e.csr_lo);
  p.csr_w = reinterpret_cast<const float*>(base + e.csr_w);
  p.hdr = reinterpret_cast<const WsHeader*>(base + l.header);
  p.hop = d->hop;
  p.pad = d->pad;
  p.center = d->center;
  p.pad_mode = d->pad_mode;
  p.n_mels = d->n_mels;
  p.stage = stage;
  p.log_mels = d->log_mels;
  p.power = d->power;
  p.db_mult = d->db_multiplier;
  p.db_amin = d->db_amin;
  p.db_offset = d->db_offset;
  (void)kBins;
  const size_t smem = sizeof(float2) * (32 * 32 + kWarps * kTileFloat2) + sizeof(int) * (2 * (size_t)d->n_mels + 4) +
                      sizeof(float) * kCsrSmemMax;
  int64_t grid = (p.total_pairs + kWarps - 1) / kWarps;
  if (grid > 148 * 64) grid = 148 * 64;
  if (grid < 1) grid = 1;
  if (d->power == 2.f) return launch_1024<2>(p, smem, grid, stream);
  if (d->power == 1.f) return launch_1024<1>(p, smem, grid, stream);
  return launch_1024<0>(p, smem, grid, stream);
}

}  // namespace b200a
