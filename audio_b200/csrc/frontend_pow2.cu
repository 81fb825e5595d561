// Register-FFT fast path of the fused front end (n_fft = 256 / 512 / 1024 / 2048, one-sided, real output stages).
//
// A group of G = n_fft/32 lanes transforms one PAIR of consecutive frames (a, b); a warp holds 32/G
// such groups (1, 2 or 4 pairs = 2, 4 or 8 consecutive frames of one utterance = one "unit"):
//   input   the unit's n_fft + (frames-1)*hop contiguous samples are staged into shared memory by ONE
//           bulk asynchronous copy (cp.async.bulk + mbarrier, the TMA engine's 1-D mode), issued a
//           unit ahead so its latency hides behind the previous unit's arithmetic
//   z[n] = w[n] * (a[n] + i b[n]),  n = g + G j          -- 32 complex values per lane
//   pass 1  32-point DFT over j in registers (radix-2 DIT, compile-time twiddles, FMA-form butterflies)
//   twiddle W_nfft^(g * k2) from a shared 32 x G table, transpose through a padded per-warp shared tile
//   pass 2  32/G DFTs of G points over the former lane index, in registers  -> Z[l + G q + 32 k1]
//   un-pack the two real spectra with one shuffle per needed value:  A = (Z[k] + conj Z[N-k]) / 2,
//                                                                    B = (Z[k] - conj Z[N-k]) / 2i
//   every lane ends with bins k = l + G m (m < 16) of its two frames; |.|^p -> global (Spectrogram: 16
//   independent warps per CTA), or -> shared memory for the mel contraction.
//   mel     warp specialised, two bodies behind one kernel (picked by the plan prepare_tc_kernel built):
//           tcgen05 (n_fft <= 1024, banded filterbank fits shared memory: every real mel / linear bank):
//             12 (n_fft = 256: 8) transform warps publish fp32 power values in UMMA core-matrix order,
//             4 operand warps convert them in place to bf16 hi / lo planes, one thread issues
//             tcgen05.mma kind::f16 (M = 64, accumulator in TMEM, banded k-steps), the operand warps read
//             the previous tile's accumulator with tcgen05.ld; error-compensated bf16, ~2^-16 relative.
//           mma.sync (any other filterbank up to 512 filters, n_fft = 2048, or B200A_TC=0):
//             8 transform warps publish power rows, 4 contraction warps multiply each finished tile with
//             the filterbank: m16n8k8 TF32 with error-compensated operands (P_hi*F_hi + P_lo*F_hi +
//             P_hi*F_lo, ~2^-21 relative), visiting only the k-steps where a group of 8 filters is non-zero.
//           (-> dB / log) -> global.
// Nothing but the waveform is read from HBM and nothing but the final features is written.
//
// tcgen05 at n_fft = 1024: an M = 64 tile of 513 bins does not fit next to the transform working set, so
// the tile has 24 real rows (what 12 warps finish per iteration) and the other 40 operand rows alias
// whatever follows in shared memory -- each accumulator row depends on its own operand row only, and the
// TMEM lanes of those rows are never read.  See DESIGN.md ("mel projection").
//
// Reference semantics: src/torchaudio/functional/functional.py:54-145 and
// transforms/_transforms.py:403-415, :701-705 (see frontend_generic.cu for the any-size path).
#include <cuda_bf16.h>

#include <cstdlib>

#include <type_traits>
#include <utility>

#include "common.cuh"
#include "f32x2.cuh"
#include "ptx.cuh"

namespace b200a {

namespace {

constexpr float kKaldiEps = 1.1920928955078125e-07f;  // numeric_limits<float>::epsilon(), kaldi.py:21-22
constexpr int kPadSymmetric = 4;  // internal pad mode: x[-1-j] = x[j], x[L+j] = x[L-1-j] (Kaldi snip_edges = false)
constexpr int kWarps = 8;     // transform warps per CTA
constexpr int kMelWarps = 4;  // contraction warps per CTA (mel kernel)
constexpr int kMaxItems = 64;  // filter groups (of 8) per contraction
constexpr int kMaxItemsPerWarp = 32;
constexpr int kFragSmemSteps = 112;  // filterbank fragments kept in shared memory (x 512 B)
constexpr int kMaxSlots = 64;

// Geometry of one transform size: G lanes per frame pair.
template <int G>
struct Geo {
  static constexpr int kNfft = 32 * G;
  static constexpr int kBins = kNfft / 2 + 1;
  static constexpr int kGroups = 32 / G;        // frame pairs per warp
  static constexpr int kFrames = 2 * kGroups;   // frames per warp and iteration ("unit")
  static constexpr int kRowLd = G + 1;          // float2 pitch of one transpose row (bank-conflict free)
  static constexpr int kRegion = 32 * (G + 1) + (G == 8 ? 8 : 0);  // float2 per lane group (skewed for G = 8)
  static constexpr int kTileF2 = kGroups * kRegion;                // float2 per warp
  static constexpr int kStageFloats = 2 * kTileF2;                 // a staged unit must fit the tile
  static constexpr int kSlots = kWarps * kFrames;                  // frames finished per CTA iteration
  static constexpr int kPitch = ((kBins + 7 - 4 + 31) / 32) * 32 + 4;  // floats per power row, == 4 (mod 32)
  static constexpr int kLogG = G == 32 ? 5 : (G == 16 ? 4 : 3);
};

// The mel contraction D[16 frames][n_mels] = P[16][bins] * F[bins][n_mels] is cut into ITEMS =
// groups of 8 filters with the k-steps where the group is non-zero, spread over the contraction warps
// by descending size.
struct MelItem {
  int tile;      // filter group: filters [8 tile, 8 tile + 8)
  int kstart;    // first bin of the first k-step (multiple of 8)
  int nsteps;    // k-steps
  int frag_off;  // index of the first step in the fragment array
};
struct MelPlan {  // built on the device by prepare_mma_kernel
  int n_tiles, n_items, total_steps, pad;
  int warp_cnt[kWarps];
  int warp_items[kWarps][kMaxItemsPerWarp];
  MelItem items[kMaxItems];
};

struct Pow2Extra {  // tables appended to the generic workspace
  size_t tw2d, tw_eo, plan, frags, tc_plan, tc_b, total;
};

inline int mel_tiles(int n_mels) { return (n_mels + 7) / 8; }

inline Pow2Extra pow2_layout(const b200a_frontend_desc& d, size_t base) {
  Pow2Extra e{};
  const size_t n_bins = d.n_fft / 2 + 1;
  const size_t nt = d.n_mels > 0 ? mel_tiles(d.n_mels) : 0;
  size_t off = base;
  e.tw2d = off;
  off = align_up(off + sizeof(float2) * 32 * 32, 256);
  e.tw_eo = off;
  off = align_up(off + sizeof(float2) * 17 * 32, 256);
  e.plan = off;
  off = align_up(off + sizeof(MelPlan), 256);
  e.frags = off;  // worst case: every tile spans every bin
  off = align_up(off + sizeof(float4) * 32 * nt * ((n_bins + 7) / 8 + 1), 256);
  const bool tc = d.n_mels > 0 && d.n_fft <= 1024;  // tcgen05 contraction: step table + banded bf16 B blocks
  e.tc_plan = off;
  off = align_up(off + (tc ? 1024 : 0), 256);
  e.tc_b = off;
  off = align_up(off + (tc ? 96 * 1024 : 0), 256);
  e.total = off;
  return e;
}

bool pow2_applicable(const b200a_frontend_desc& d) {
  return (d.n_fft == 2048 || d.n_fft == 1024 || d.n_fft == 512 || d.n_fft == 256) && d.onesided != 0;
}

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
template <int LOG>
__host__ __device__ constexpr int brev(int v) {  // bit reversal of a LOG-bit index
  return brev5(v) >> (5 - LOG);
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

// DIT butterfly (a, b) -> (a + W b, a - W b), W = exp(-2 pi i E / 32), E in [0, 16), on packed FP32 pairs
// (f32x2.cuh): 2 issue slots for the trivial twiddles, 3 for the others (was 4 / 6 with scalar FADD / FFMA).
template <int E>
__device__ __forceinline__ void bfly(float2& a, float2& b) {
  if constexpr (E == 0) {
    const float2 t = b;
    b = sub2(a, t);
    a = add2(a, t);
  } else if constexpr (E == 8) {  // W = -i : W b = (b.y, -b.x) = -i b
    const float2 t = b;
    b = add_i(a, t);
    a = sub_i(a, t);
  } else if constexpr (E == 4) {  // W = (1 - i)/sqrt2 : W b = c (b - i b)
    constexpr float c = 0.70710678118654752f;
    const float2 t = sub_i(b, b);  // (b.x + b.y, b.y - b.x)
    b = fmas2(-c, t, a);
    a = fmas2(c, t, a);
  } else if constexpr (E == 12) {  // W = (-1 - i)/sqrt2 : W b = -c (b + i b)
    constexpr float c = 0.70710678118654752f;
    const float2 u = add_i(b, b);  // (b.x - b.y, b.y + b.x)
    b = fmas2(c, u, a);
    a = fmas2(-c, u, a);
  } else {
    constexpr float wr = kCos32[E], wi = -kSin32[E];
    const float2 p = cfma2(wr, wi, b, a);               // a + W b
    b = fma2(make_float2(2.f, 2.f), a, make_float2(-p.x, -p.y));  // 2 a - p = a - W b
    a = p;
  }
}

// In-register LEN-point DFT on a[OFF .. OFF+LEN).  Input a[OFF + brev<log2 LEN>(j)] = x[j]; output
// a[OFF + k] = X[k] (natural order).
template <int LEN, int OFF, int S0 = 0, int S1 = 5>
__device__ __forceinline__ void fft_regs(float2 (&a)[32]) {
  constexpr int kStages = LEN == 32 ? 5 : (LEN == 16 ? 4 : 3);
  constexpr int kFirst = S0, kLast = S1 < kStages ? S1 : kStages;  // stages [kFirst, kLast)
  static_for<kLast - kFirst>([&](auto si) {
    constexpr int len = 2 << (decltype(si)::value + kFirst);  // 2, 4, ..., LEN
    constexpr int half = len / 2;
    static_for<LEN / 2>([&](auto bi) {
      constexpr int b = decltype(bi)::value;
      constexpr int i = (b / half) * len, j = b % half;
      bfly<j*(32 / len)>(a[OFF + i + j], a[OFF + i + j + half]);
    });
  });
}

struct Pow2Params {
  const float* wave;
  int64_t length, row_stride, frames, units_per_row, total_units;
  float* out;
  float* group_max;
  int64_t rows_per_group;
  const float* window;   // [n_fft] centre padded
  const float2* tw2d;    // [32][G]  W_nfft^(k2 * g) at [k2][g]
  const MelPlan* plan;
  const float4* frags;   // [steps][32] (b0_hi, b1_hi, b0_lo, b1_lo) in mma B-fragment order
  const WsHeader* hdr;
  const struct TcPlan* tc;     // tcgen05 contraction plan (nullptr: not available for this size)
  const unsigned char* tc_b;   // its banded bf16 B blocks
  int hop, pad, center, pad_mode, n_mels;
  int stage, log_mels, bulk_ok, stage_ok;
  // output row geometry: value m of frame t goes to out[(row * frames + t) * out_width + out_col0 + m]
  int out_width, out_col0, out_vec;  // out_vec: floats every row start is aligned to (1, 2 or 4)
  // Kaldi framing / per-frame conditioning (compliance/kaldi.py:44-83, :153-216); kaldi == 0: torch.stft framing
  int kaldi, k_off, k_win, k_dc, k_energy_mode, k_energy_col, k_log;
  float k_preemph, k_energy_floor;
  float power, db_mult, db_amin, db_offset;
};

// samples before t * hop where frame t starts: n_fft/2 (torch.stft center), 0, or Kaldi's win/2 - shift/2
__device__ __forceinline__ int frame_lead(const Pow2Params& p, int n_fft) {
  return p.kaldi ? p.k_off : (p.center ? n_fft / 2 : 0);
}

constexpr int kComplexOut = 3;  // POWER_MODE of the complex (power = None) Spectrogram kernel

template <int POWER_MODE>  // 2: |.|^2, 0: general exponent (1 handled inside)
__device__ __forceinline__ float pow_of(float re, float im, float power) {
  if constexpr (POWER_MODE == 2) return fmaf(re, re, im * im);
  const float mag = hypotf(re, im);
  return power == 1.f ? mag : powf(mag, power);
}

// Running maximum of the dB features per top_db group, flushed with as few atomics as possible.
struct GroupMax {
  float* dst;
  int64_t group;
  float value;
  // Warp-collective (all 32 lanes call it together; `dst` is warp-uniform).
  __device__ __forceinline__ void flush() {
    if (dst == nullptr) return;
    const int64_t g0 = __shfl_sync(0xffffffffu, group, 0);
    if (__all_sync(0xffffffffu, group == g0)) {
      const float mx = warp_max(value);
      if ((threadIdx.x & 31) == 0 && g0 >= 0 && mx > -CUDART_INF_F) atomic_max_f32(dst + g0, mx);
    } else if (group >= 0 && value > -CUDART_INF_F) {
      atomic_max_f32(dst + group, value);
    }
    value = -CUDART_INF_F;
  }
  // Must be called by all 32 lanes together; `valid == false` lanes contribute nothing.
  __device__ __forceinline__ void add(int64_t g, float v, bool valid) {
    if (dst == nullptr) return;
    if (!valid) {
      g = group;
      v = -CUDART_INF_F;
    }
    // warp-collective flush only when ANY lane changes group (keeps the shuffles converged)
    if (__any_sync(0xffffffffu, g != group && group >= 0)) flush();
    group = g;
    value = fmaxf(value, v);
  }
};

template <int N>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// Per-warp walker over this warp's units: (row, unit-in-row) of the current and the next unit,
// advanced without divisions.
struct UnitCursor {
  int64_t u, stride, step_rows, step_units, upr;
  int64_t row, ub, nrow, nub;
  __device__ __forceinline__ void init(int64_t first, int64_t stride_, int64_t units_per_row) {
    u = first;
    stride = stride_;
    upr = units_per_row;
    step_rows = stride / upr;
    step_units = stride - step_rows * upr;
    row = first / upr;
    ub = first - row * upr;
    nrow = row + step_rows;
    nub = ub + step_units;
    if (nub >= upr) { nub -= upr; ++nrow; }
  }
  __device__ __forceinline__ void advance() {
    u += stride;
    row = nrow;
    ub = nub;
    nrow += step_rows;
    nub += step_units;
    if (nub >= upr) { nub -= upr; ++nrow; }
  }
};

// a unit can be staged by one bulk copy iff all its frames exist and lie inside the row un-padded
template <int G>
__device__ __forceinline__ bool bulk_eligible(const Pow2Params& p, int half, int64_t u, int64_t ub) {
  using Ge = Geo<G>;
  if (!p.bulk_ok || u >= p.total_units) return false;
  const int64_t t0 = ub * Ge::kFrames;
  const int64_t s0 = t0 * p.hop - half - p.pad;
  return t0 + Ge::kFrames <= p.frames && s0 >= 0 && s0 + (int64_t)(Ge::kFrames - 1) * p.hop + Ge::kNfft <= p.length;
}
template <int G>
__device__ __forceinline__ void issue_bulk(const Pow2Params& p, int half, int64_t row, int64_t ub, void* dst,
                                           uint64_t* bar) {
  using Ge = Geo<G>;
  const float* src = p.wave + row * p.row_stride + (ub * Ge::kFrames * p.hop - half - p.pad);
  const uint32_t bytes = (uint32_t)(Ge::kNfft + (Ge::kFrames - 1) * p.hop) * 4u;
  mbar_expect_tx(bar, bytes);
  bulk_g2s(dst, src, bytes, bar);
}

// One warp, one unit (32/G frame pairs): samples -> windowed complex signals -> n_fft-point FFTs -> the
// power spectra.  On return lane (group gi, l) holds bins k = l + G m in pa[m] / pb[m] (m < 16) of frames
// t0 + 2 gi and t0 + 2 gi + 1, and lanes with l == 0 bin n_fft/2 in [16].
//   stage      where a prefetched unit was staged (aliases `tile` when STAGE_IS_TILE)
//   STAGE_IS_TILE  the staging buffer is the transpose tile itself: the NEXT unit's bulk copy is issued
//                  only after pass 2 has read the tile back
template <int POWER_MODE, int G, int HG, bool STAGE_IS_TILE, bool KALDI>
__device__ __forceinline__ void transform_unit(const Pow2Params& p, const float (&wreg)[32], const float2* s_tw,
                                               float2* tile, float* stage, uint64_t* bar, uint32_t& parity,
                                               bool& staged, const UnitCursor& cur, int half, int lane,
                                               float (&pa)[17], float (&pb)[17]) {
  using Ge = Geo<G>;
  const int gi = lane / G, l = lane % G;
  const int64_t row = cur.row, t0 = cur.ub * Ge::kFrames;
  const int64_t ta = t0 + 2 * gi, tb = ta + 1;
  const bool has_a = ta < p.frames, has_b = tb < p.frames;
  const float* __restrict__ x = p.wave + row * p.row_stride;
  const int64_t s0 = t0 * p.hop - half - p.pad;  // first raw sample of the unit
  const int64_t sa = s0 + (int64_t)2 * gi * p.hop, sb = sa + p.hop;
  const bool next_staged = bulk_eligible<G>(p, half, cur.u + cur.stride, cur.nub);
  // every frame of the unit inside the signal: plain loads; otherwise the padding-aware gather
  const int64_t last = p.frames - t0 < Ge::kFrames ? p.frames - t0 : Ge::kFrames;  // frames present
  const bool interior = s0 >= 0 && s0 + (last - 1) * p.hop + Ge::kNfft <= p.length;

  float2 a[32];
  float2* grp_tile = tile + gi * Ge::kRegion;
  bool from_stage = staged;
  if (staged) {
    mbar_wait(bar, parity);
    parity ^= 1;
  } else if ((!interior || KALDI) && p.stage_ok) {
    // edge unit (padding / reflection / ragged end): the lanes gather the unit's whole span into the (idle)
    // staging buffer with 4-byte asynchronous copies -- every sample once, all copies in flight together --
    // and the unit then takes the same register-load path as a bulk-staged one
    const int span = Ge::kNfft + (Ge::kFrames - 1) * p.hop;
    // 32-bit index arithmetic (the launch guarantees length + 2 pad + n_fft < 2^31): j indexes the constant
    // pre-padded signal of `ext` samples, exactly as source_index() does in 64 bits
    const int len = (int)p.length, ext = len + 2 * p.pad, j0 = (int)(t0 * p.hop) - half, mode = p.pad_mode;
#pragma unroll 4
    for (int n = lane; n < span; n += 32) {
      int j = j0 + n;
      if ((unsigned)j >= (unsigned)ext) {
        if (mode == B200A_PAD_REFLECT)
          j = j < 0 ? -j : 2 * (ext - 1) - j;
        else if (mode == B200A_PAD_REPLICATE)
          j = j < 0 ? 0 : ext - 1;
        else if (mode == kPadSymmetric)
          j = j < 0 ? -1 - j : 2 * ext - 1 - j;
        else if (mode == B200A_PAD_CIRCULAR) {
          j %= ext;
          if (j < 0) j += ext;
        } else
          j = -1;
      }
      const int src = j - p.pad;
      if (j >= 0 && (unsigned)src < (unsigned)len)
        cp_async4(stage + n, x + src);
      else
        stage[n] = 0.f;
    }
    cp_async_wait_all();
    __syncwarp();
    from_stage = true;
  }
  if (KALDI && from_stage) {
    // Kaldi conditioning of the two staged frames (sample n = l + G j, n < win): DC removal, [raw log energy],
    // pre-emphasis s[n] - c s[max(n - 1, 0)], window (zero beyond win), [log energy after the window]
    const float* fa = stage + 2 * gi * p.hop;
    const float* fb = fa + p.hop;
    const int win = p.k_win;
    float ma = 0.f, mb = 0.f;
    if (p.k_dc) {
#pragma unroll 4
      for (int n = l; n < win; n += G) {
        ma += fa[n];
        mb += fb[n];
      }
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) {
        ma += __shfl_xor_sync(0xffffffffu, ma, o);
        mb += __shfl_xor_sync(0xffffffffu, mb, o);
      }
      ma /= (float)win;
      mb /= (float)win;
    }
    float ea = 0.f, eb = 0.f;
    if (p.k_energy_mode == 1) {
#pragma unroll 4
      for (int n = l; n < win; n += G) {
        const float da = fa[n] - ma, db = fb[n] - mb;
        ea = fmaf(da, da, ea);
        eb = fmaf(db, db, eb);
      }
    }
    const float c = p.k_preemph;
    static_for<32>([&](auto ji) {
      constexpr int j = decltype(ji)::value;
      const int n = l + G * j;
      float va = 0.f, vb = 0.f;
      if (n < win) {
        const int np = n > 0 ? n - 1 : 0;
        va = ((fa[n] - ma) - c * (fa[np] - ma)) * wreg[j];
        vb = ((fb[n] - mb) - c * (fb[np] - mb)) * wreg[j];
      }
      a[brev5(j)] = make_float2(va, vb);
      if (p.k_energy_mode == 2) {  // wreg carries the un-packing's 1/2
        ea = fmaf(2.f * va, 2.f * va, ea);
        eb = fmaf(2.f * vb, 2.f * vb, eb);
      }
    });
    if (p.k_energy_mode != 0) {
#pragma unroll
      for (int o = G / 2; o > 0; o >>= 1) {
        ea += __shfl_xor_sync(0xffffffffu, ea, o);
        eb += __shfl_xor_sync(0xffffffffu, eb, o);
      }
      if (l == 0) {
        const float fl = p.k_energy_floor > 0.f ? logf(p.k_energy_floor) : -CUDART_INF_F;
        float* e_out = p.out + (row * p.frames + ta) * p.out_width + p.k_energy_col;
        if (has_a) e_out[0] = fmaxf(logf(fmaxf(ea, kKaldiEps)), fl);
        if (has_b) e_out[p.out_width] = fmaxf(logf(fmaxf(eb, kKaldiEps)), fl);
      }
    }
    __syncwarp();  // every lane has consumed the staging buffer
  } else if (from_stage) {
    if constexpr (HG >= 0) {  // G == 32: frame b is frame a shifted by HG lane-rows
      constexpr int kV = 32 + (HG >= 0 ? HG : 0);
      float v[kV];
      static_for<kV>([&](auto ji) {
        constexpr int j = decltype(ji)::value;
        v[j] = stage[lane + 32 * j];
      });
      static_for<32>([&](auto ji) {
        constexpr int j = decltype(ji)::value;
        a[brev5(j)] = make_float2(v[j] * wreg[j], v[j + (HG >= 0 ? HG : 0)] * wreg[j]);
      });
    } else {
      const float* pa_ptr = stage + 2 * gi * p.hop + l;
      const float* pb_ptr = pa_ptr + p.hop;
      static_for<32>([&](auto ji) {
        constexpr int j = decltype(ji)::value;
        a[brev5(j)] = scale2(wreg[j], make_float2(pa_ptr[G * j], pb_ptr[G * j]));
      });
    }
    __syncwarp();  // every lane has consumed the staging buffer
  } else if (interior) {
    static_for<32>([&](auto ji) {
      constexpr int j = decltype(ji)::value;
      const float va = has_a ? __ldg(x + sa + l + G * j) : 0.f;
      const float vb = has_b ? __ldg(x + sb + l + G * j) : 0.f;
      a[brev5(j)] = scale2(wreg[j], make_float2(va, vb));
    });
  } else {
    // edge unit whose span does not fit the staging buffer: gather through the group's tile region with a
    // rolled loop so the index arithmetic is not replicated 64 times in the instruction stream
#pragma unroll 1
    for (int j = 0; j < 32; ++j) {
      const int n = l + G * j;
      const int64_t ia = has_a ? source_index(ta * p.hop + n, p.length, p.pad, half, p.pad_mode) : -1;
      const int64_t ib = has_b ? source_index(tb * p.hop + n, p.length, p.pad, half, p.pad_mode) : -1;
      grp_tile[n] = make_float2(ia >= 0 ? __ldg(x + ia) : 0.f, ib >= 0 ? __ldg(x + ib) : 0.f);
    }
    __syncwarp();
    static_for<32>([&](auto ji) {
      constexpr int j = decltype(ji)::value;
      const float2 v = grp_tile[l + G * j];
      a[brev5(j)] = scale2(wreg[j], v);
    });
    __syncwarp();
  }
  if constexpr (!STAGE_IS_TILE) {  // separate staging buffer: prefetch right away
    staged = next_staged;
    if (staged && lane == 0) issue_bulk<G>(p, half, cur.nrow, cur.nub, stage, bar);
  }

  // the inter-pass twiddles: kTwAhead loads are in flight before the last butterfly stage, and every multiply
  // issues the load kTwAhead positions ahead of it, so no multiply waits for its own load
#ifndef B200A_TW_AHEAD
#define B200A_TW_AHEAD 8
#endif
  constexpr int kTwAhead = B200A_TW_AHEAD;
  fft_regs<32, 0, 0, 4>(a);
  float2 tw[32];
  static_for<kTwAhead>([&](auto ki) {
    constexpr int k2 = decltype(ki)::value + 1;
    tw[k2] = s_tw[k2 * G + l];
  });
  fft_regs<32, 0, 4, 5>(a);  // a[k2] = Y[l][k2]
  grp_tile[l] = a[0];
  static_for<31>([&](auto ki) {
    constexpr int k2 = decltype(ki)::value + 1;
    if constexpr (k2 + kTwAhead < 32) tw[k2 + kTwAhead] = s_tw[(k2 + kTwAhead) * G + l];
    grp_tile[k2 * Ge::kRowLd + l] = cmul2(a[k2], tw[k2]);
  });
  __syncwarp();
  // lane l now owns k2 = l + G q, q < 32/G: slot q*G + brev(g) <- element (g, l + G q)
  static_for<32>([&](auto si) {
    constexpr int s = decltype(si)::value;
    constexpr int q = s / G, g = s % G;
    a[q * G + brev<Ge::kLogG>(g)] = grp_tile[(l + G * q) * Ge::kRowLd + g];
  });
  __syncwarp();
  if constexpr (STAGE_IS_TILE) {  // the tile is free until the next unit's transpose: stage into it
    staged = next_staged;
    if (staged && lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads above -> async write
      issue_bulk<G>(p, half, cur.nrow, cur.nub, stage, bar);
    }
  }

  static_for<Ge::kGroups>([&](auto qi) { fft_regs<G, decltype(qi)::value * G>(a); });
  // a[q*G + k1] = Z[(l + G q) + 32 k1]; bin k = l + G m with m = q + (32/G) k1  <->  slot (m % NG)*G + m / NG

  // ---- un-pack the two spectra: need Z[N - k] for k = l + G m, m = 0..15 (+ bin N/2 on l == 0) ----
  constexpr int NG = Ge::kGroups;
  const int src = (lane & ~(G - 1)) | ((G - l) & (G - 1));
  static_for<16>([&](auto mi) {
    constexpr int m = decltype(mi)::value;
    constexpr int slot = (m % NG) * G + m / NG;
    // l >= 1: N - k = (G - l) + G (31 - m): on lane `src`, slot of m' = 31 - m
    constexpr int mm = 31 - m, mslot = (mm % NG) * G + mm / NG;
    float mr = __shfl_sync(0xffffffffu, a[mslot].x, src);
    float mi_ = __shfl_sync(0xffffffffu, a[mslot].y, src);
    if (l == 0) {  // l == 0: N - k = G (32 - m): own slot of m' = 32 - m (slot 0 for m = 0)
      constexpr int m0 = (32 - m) & 31, slot0 = (m0 % NG) * G + m0 / NG;
      mr = a[slot0].x;
      mi_ = a[slot0].y;
    }
    // A = Z[k] + conj Z[N-k] = (sx.x, sy.x),  B = (Z[k] - conj Z[N-k]) / i = (sy.y, -sx.y): two packed adds
    const float zr = a[slot].x, zi = a[slot].y;
    const float2 sx = add2(make_float2(zr, zr), make_float2(mr, -mr));    // (zr + mr, zr - mr)
    const float2 sy = add2(make_float2(zi, zi), make_float2(-mi_, mi_));  // (zi - mi, zi + mi)
    if constexpr (POWER_MODE == kComplexOut) {  // power = None: the two spectra go straight to out[row][t][bin] (complex64)
      float2* oc = reinterpret_cast<float2*>(p.out) + (row * p.frames + ta) * Ge::kBins + l + G * m;
      if (has_a) oc[0] = make_float2(sx.x, sy.x);
      if (has_b) oc[Ge::kBins] = make_float2(sy.y, -sx.y);
    } else if constexpr (POWER_MODE == 2) {  // (|A|^2, |B|^2) as one packed multiply + one packed FMA
      const float2 pw = fma2(sx, sx, mul2(sy, sy));
      pa[m] = pw.x;
      pb[m] = pw.y;
    } else {
      pa[m] = pow_of<POWER_MODE>(sx.x, sy.x, p.power);
      pb[m] = pow_of<POWER_MODE>(sy.y, sx.y, p.power);
    }
  });
  // bin N/2 (l == 0, m = 16) is its own mirror: A = Re, B = Im  (x2 because wreg carries the 1/2)
  constexpr int slot16 = (16 % NG) * G + 16 / NG;
  if constexpr (POWER_MODE == kComplexOut) {
    if (l == 0) {
      float2* oc = reinterpret_cast<float2*>(p.out) + (row * p.frames + ta) * Ge::kBins + Ge::kNfft / 2;
      if (has_a) oc[0] = make_float2(2.f * a[slot16].x, 0.f);
      if (has_b) oc[Ge::kBins] = make_float2(2.f * a[slot16].y, 0.f);
    }
  } else {
    pa[16] = pow_of<POWER_MODE>(2.f * a[slot16].x, 0.f, p.power);
    pb[16] = pow_of<POWER_MODE>(2.f * a[slot16].y, 0.f, p.power);
  }
}

template <int G>
__device__ __forceinline__ void load_window(const Pow2Params& p, int lane, float (&wreg)[32]) {
  // window (x 1/2 from the un-packing, x the normalisation scale) for n = l + G j
  const float hs = 0.5f * p.hdr->scale;
  const int l = lane % G;
#pragma unroll
  for (int j = 0; j < 32; ++j) wreg[j] = p.window[l + G * j] * hs;
}

// ------------------------------------------------------------------------------------------------
// Spectrogram kernel: 8 independent warps, power spectra straight to global memory.
// ------------------------------------------------------------------------------------------------
template <int POWER_MODE, int G, int HG, int NW, bool STAGE_IS_TILE, bool KALDI>
__global__ void __launch_bounds__(NW * 32, 1) stft_pow2_power_kernel(const Pow2Params p) {
  using Ge = Geo<G>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* s_tw = reinterpret_cast<float2*>(smem_raw);                                    // [32][G]
  float2* s_tile_all = s_tw + 32 * 32;                                                   // [NW][kTileF2]
  float* s_stage_all = reinterpret_cast<float*>(s_tile_all + NW * Ge::kTileF2);      // [NW][kStageFloats]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_stage_all + (STAGE_IS_TILE ? 0 : NW * Ge::kStageFloats));  // [NW]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 32 * G; i += blockDim.x) s_tw[i] = p.tw2d[i];
  if (tid < NW) mbar_init(s_bar + tid, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  float2* tile = s_tile_all + warp * Ge::kTileF2;
  float* stage = STAGE_IS_TILE ? reinterpret_cast<float*>(tile) : s_stage_all + warp * Ge::kStageFloats;
  uint64_t* bar = s_bar + warp;
  float wreg[32];
  load_window<G>(p, lane, wreg);
  const int half = frame_lead(p, Ge::kNfft);
  const int gi = lane / G, l = lane % G;
  uint32_t parity = 0;
  bool staged = false;
  UnitCursor cur;
  cur.init((int64_t)blockIdx.x * NW + warp, (int64_t)gridDim.x * NW, p.units_per_row);
  if (bulk_eligible<G>(p, half, cur.u, cur.ub)) {
    if (lane == 0) issue_bulk<G>(p, half, cur.row, cur.ub, stage, bar);
    staged = true;
  }
  for (; cur.u < p.total_units; cur.advance()) {
    float pa[17], pb[17];
    transform_unit<POWER_MODE, G, HG, STAGE_IS_TILE, KALDI>(p, wreg, s_tw, tile, stage, bar, parity, staged, cur, half, lane, pa,
                                                            pb);
    if constexpr (POWER_MODE == kComplexOut) continue;  // transform_unit has written the complex spectra
    const int64_t ta = cur.ub * Ge::kFrames + 2 * gi;
    const bool has_a = ta < p.frames, has_b = ta + 1 < p.frames;
    float* oa = p.out + (cur.row * p.frames + ta) * p.out_width + p.out_col0;
    float* ob = oa + p.out_width;
    if (KALDI && p.k_log) {  // Kaldi spectrogram: log(max(|X|^2, eps)), kaldi.py:310
#pragma unroll
      for (int m = 0; m < 17; ++m) {
        pa[m] = logf(fmaxf(pa[m], kKaldiEps));
        pb[m] = logf(fmaxf(pb[m], kKaldiEps));
      }
    }
    const int skip = KALDI ? p.k_energy_col - p.out_col0 : -1;  // the bin whose column holds the frame's log energy
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      if (l + G * m == skip) continue;
      if (has_a) oa[l + G * m] = pa[m];
      if (has_b) ob[l + G * m] = pb[m];
    }
    if (l == 0 && Ge::kNfft / 2 != skip) {
      if (has_a) oa[Ge::kNfft / 2] = pa[16];
      if (has_b) ob[Ge::kNfft / 2] = pb[16];
    }
  }
}

// One contraction warp's share of a 16-frame power tile: D[16 x 8] = P[16 x bins] * F[bins x 8] for each of its
// filter groups on the tensor pipe, (dB / log), store.  pw: the tile's 16 power rows (pitch PITCH);
// slot / grp: output offset (or -1) and top_db group of each of the 16 rows.
template <int PITCH>
__device__ __forceinline__ void contract_tile(const Pow2Params& p, const MelPlan* s_plan, const float4* s_frags,
                                              bool frags_in_smem, int mw, int lane, const float* pw,
                                              const int64_t* slot, const int64_t* grp, GroupMax& gmax) {
  const int r = lane >> 2, c = lane & 3;
  const int cnt = s_plan->warp_cnt[mw];
  const int64_t o_lo = slot[r], o_hi = slot[r + 8];
  const int64_t g_lo = grp[r], g_hi = grp[r + 8];
  for (int ii = 0; ii < cnt; ++ii) {
    const MelItem mi = s_plan->items[s_plan->warp_items[mw][ii]];
    const float* a_lo_row = pw + (size_t)r * PITCH + mi.kstart + c;
    const float* a_hi_row = a_lo_row + 8 * PITCH;
    // three independent accumulator chains (hi*hi, lo*hi, hi*lo), summed in a fixed order
    float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
    auto contract = [&](auto in_smem) {
      const float4* fr = (decltype(in_smem)::value ? s_frags : p.frags) + (size_t)mi.frag_off * 32 + lane;
#pragma unroll 4
      for (int s = 0; s < mi.nsteps; ++s) {
        float4 bf;
        if constexpr (decltype(in_smem)::value) bf = fr[(size_t)s * 32];
        else bf = __ldg(fr + (size_t)s * 32);
        const float av[4] = {a_lo_row[8 * s], a_hi_row[8 * s], a_lo_row[8 * s + 4], a_hi_row[8 * s + 4]};
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) split_tf32(av[q], hi[q], lo[q]);
        mma_tf32(d0, hi, __float_as_uint(bf.x), __float_as_uint(bf.y));
        mma_tf32(d1, lo, __float_as_uint(bf.x), __float_as_uint(bf.y));
        mma_tf32(d2, hi, __float_as_uint(bf.z), __float_as_uint(bf.w));
      }
    };
    if (frags_in_smem) contract(std::true_type{});
    else contract(std::false_type{});
    float d[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) d[q] = d0[q] + (d1[q] + d2[q]);
    const int n0 = 8 * mi.tile + 2 * c;
    const bool n0_ok = n0 < p.n_mels, n1_ok = n0 + 1 < p.n_mels;
    if (p.k_log) {  // Kaldi fbank: log(max(mel, FLT_EPSILON)), kaldi.py:629-631
#pragma unroll
      for (int q = 0; q < 4; ++q) d[q] = logf(fmaxf(d[q], kKaldiEps));
    }
    if (p.stage == B200A_STAGE_FEAT) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        d[q] = p.log_mels ? logf(d[q] + 1e-6f) : p.db_mult * log10f(fmaxf(d[q], p.db_amin)) - p.db_offset;
      const float m_lo = fmaxf(n0_ok ? d[0] : -CUDART_INF_F, n1_ok ? d[1] : -CUDART_INF_F);
      const float m_hi = fmaxf(n0_ok ? d[2] : -CUDART_INF_F, n1_ok ? d[3] : -CUDART_INF_F);
      gmax.add(g_lo, m_lo, o_lo >= 0);
      gmax.add(g_hi, m_hi, o_hi >= 0);
    }
    const bool vec = n1_ok && p.out_vec >= 2;  // 8-byte aligned pair
    if (o_lo >= 0) {
      if (vec) *reinterpret_cast<float2*>(p.out + o_lo + n0) = make_float2(d[0], d[1]);
      else {
        if (n0_ok) p.out[o_lo + n0] = d[0];
        if (n1_ok) p.out[o_lo + n0 + 1] = d[1];
      }
    }
    if (o_hi >= 0) {
      if (vec) *reinterpret_cast<float2*>(p.out + o_hi + n0) = make_float2(d[2], d[3]);
      else {
        if (n0_ok) p.out[o_hi + n0] = d[2];
        if (n1_ok) p.out[o_hi + n0 + 1] = d[3];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Mel / MFCC-feature kernel, warp specialised: warps 0-7 transform (one unit each per iteration)
// and publish power rows into a double-buffered shared tile; warps 8-11 contract each finished tile
// (16, 32 or 64 frames) with the filterbank on the tensor pipe and store the features.  The two groups
// only meet at the tile's full/empty mbarriers, so the FFT warps never wait for the contraction.
// ------------------------------------------------------------------------------------------------
constexpr int kFftRegs = 200, kMelRegs = 96;  // 256*200 + 128*96 = 63488 <= 64512 = 384 * 168

template <int POWER_MODE, int G, int HG, int FFT_REGS, bool KALDI>
__device__ __forceinline__ void mel_body_mma(const Pow2Params& p, unsigned char* smem_raw) {
  using Ge = Geo<G>;
  constexpr int kSlots = Ge::kSlots, kPitch = Ge::kPitch;
  float2* s_tw = reinterpret_cast<float2*>(smem_raw);                              // [32][G]
  float2* s_tile_all = s_tw + 32 * 32;                                             // [kWarps][kTileF2] (also staging)
  float* s_pow = reinterpret_cast<float*>(s_tile_all + kWarps * Ge::kTileF2);      // [2][kSlots][kPitch]
  int64_t* s_slot = reinterpret_cast<int64_t*>(s_pow + 2 * kSlots * kPitch);        // [2][kSlots] out offsets
  int64_t* s_grp = s_slot + 2 * kSlots;                                            // [2][kSlots] top_db group
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_grp + 2 * kSlots);               // [kWarps] staging
  uint64_t* s_full = s_bar + kWarps;                                               // [2]
  uint64_t* s_empty = s_full + 2;                                                  // [2]
  MelPlan* s_plan = reinterpret_cast<MelPlan*>(s_empty + 2);
  float4* s_frags = reinterpret_cast<float4*>(s_plan + 1);                         // [<= kFragSmemSteps][32]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 32 * G; i += blockDim.x) s_tw[i] = p.tw2d[i];
  {
    const int* src = reinterpret_cast<const int*>(p.plan);
    int* dst = reinterpret_cast<int*>(s_plan);
    for (int i = tid; i < (int)(sizeof(MelPlan) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
  }
  const int total_steps = p.plan->total_steps;
  const bool frags_in_smem = total_steps <= kFragSmemSteps;
  if (frags_in_smem)
    for (int i = tid; i < total_steps * 32; i += blockDim.x) s_frags[i] = p.frags[i];
  // columns >= n_bins of every power row are read by the last k-step: keep them finite (zero)
  for (int i = tid; i < 2 * kSlots * (kPitch - Ge::kBins); i += blockDim.x) {
    const int r = i / (kPitch - Ge::kBins), c = i - r * (kPitch - Ge::kBins);
    s_pow[r * kPitch + Ge::kBins + c] = 0.f;
  }
  if (tid < kWarps) mbar_init(s_bar + tid, 1);
  if (tid == 0) {
    mbar_init(s_full + 0, kWarps);
    mbar_init(s_full + 1, kWarps);
    mbar_init(s_empty + 0, kMelWarps);
    mbar_init(s_empty + 1, kMelWarps);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const int64_t stride = (int64_t)gridDim.x * kWarps;
  const int64_t u0 = (int64_t)blockIdx.x * kWarps;
  const int width = p.out_width;

  if (warp < kWarps) {
    // =============================== transform warps ===========================================
    reg_alloc<FFT_REGS>();
    float2* tile = s_tile_all + warp * Ge::kTileF2;
    float* stage = reinterpret_cast<float*>(tile);
    uint64_t* bar = s_bar + warp;
    float wreg[32];
    load_window<G>(p, lane, wreg);
    const int half = frame_lead(p, Ge::kNfft);
    const int gi = lane / G, l = lane % G;
    uint32_t parity = 0;
    bool staged = false;
    UnitCursor cur;
    cur.init(u0 + warp, stride, p.units_per_row);
    if (bulk_eligible<G>(p, half, cur.u, cur.ub)) {
      if (lane == 0) issue_bulk<G>(p, half, cur.row, cur.ub, stage, bar);
      staged = true;
    }
    int it = 0;
    for (int64_t base = u0; base < p.total_units; base += stride, ++it, cur.advance()) {
      const bool valid = cur.u < p.total_units;
      float pa[17], pb[17];
      if (valid)
        transform_unit<POWER_MODE, G, HG, true, KALDI>(p, wreg, s_tw, tile, stage, bar, parity, staged, cur, half, lane, pa,
                                                pb);
      const int b = it & 1;
      if (it >= 2) mbar_wait(s_empty + b, ((it >> 1) & 1) ^ 1);  // the mel warps have drained this buffer
      const int slot_a = Ge::kFrames * warp + 2 * gi;
      float* prow_a = s_pow + (size_t)(b * kSlots + slot_a) * kPitch;
      float* prow_b = prow_a + kPitch;
      if (valid) {
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          prow_a[l + G * m] = pa[m];
          prow_b[l + G * m] = pb[m];
        }
        if (l == 0) {
          prow_a[Ge::kNfft / 2] = pa[16];
          prow_b[Ge::kNfft / 2] = pb[16];
        }
      }
      if (l == 0) {
        const int64_t ta = cur.ub * Ge::kFrames + 2 * gi;
        const int64_t oa = (cur.row * p.frames + ta) * (int64_t)width + p.out_col0;
        s_slot[b * kSlots + slot_a] = (valid && ta < p.frames) ? oa : -1;
        s_slot[b * kSlots + slot_a + 1] = (valid && ta + 1 < p.frames) ? oa + width : -1;
        const int64_t g = cur.row / p.rows_per_group;
        s_grp[b * kSlots + slot_a] = g;
        s_grp[b * kSlots + slot_a + 1] = g;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(s_full + b);
    }
  } else if (warp >= kWarps + kMelWarps) {
    reg_dealloc<24>();  // launched for the tcgen05 body's 12 transform warps: hand the registers back and leave
  } else {
    // =============================== contraction warps =========================================
    reg_dealloc<kMelRegs>();
    const int mw = warp - kWarps;
    GroupMax gmax{p.stage == B200A_STAGE_FEAT ? p.group_max : nullptr, -1, -CUDART_INF_F};
    int it = 0;
    for (int64_t base = u0; base < p.total_units; base += stride, ++it) {
      const int b = it & 1;
      mbar_wait(s_full + b, (it >> 1) & 1);
#pragma unroll 1
      for (int mt = 0; mt < kSlots / 16; ++mt)  // 16-frame MMA tiles of this iteration
        contract_tile<kPitch>(p, s_plan, s_frags, frags_in_smem, mw, lane, s_pow + (size_t)(b * kSlots + 16 * mt) * kPitch,
                              s_slot + b * kSlots + 16 * mt, s_grp + b * kSlots + 16 * mt, gmax);
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty + b);
    }
    gmax.flush();
  }
}

// ================================================================================================
// n_fft = 2048: one real frame per 1024-point complex FFT ("even/odd packing").
//   z[m] = w[2m] x[2m] + i w[2m+1] x[2m+1],  Z = FFT_1024(z);  with E = (Z[k] + conj Z[1024-k]) / 2 and
//   O = (Z[k] - conj Z[1024-k]) / 2i:   X[k] = E + W_2048^k O,   X[1024-k] = conj(E - W_2048^k O),
// so each lane turns its 16 (Z[k], Z[1024-k]) pairs into 32 power bins.  The same 32 x 32 register FFT,
// tile transposes, bulk staging and tensor-pipe contraction as above; a warp handles 2 frames per iteration
// one after the other, and the (16 x 1025) power tile is single buffered.
// ================================================================================================
constexpr int kEoN = 2048, kEoBins = 1025, kEoPitch = 1060, kEoSlots = 16;
constexpr int kEoFragSteps = 136;  // fragments kept in shared memory when the plan has at most this many steps

__device__ __forceinline__ bool eo_frame_bulk_ok(const Pow2Params& p, int half, int64_t t) {
  const int64_t sa = t * p.hop - half - p.pad;
  return p.bulk_ok && t < p.frames && sa >= 0 && sa + kEoN <= p.length;
}

// One warp, one frame of 2048 samples.  On return lane l holds bins l + 32 k1 in plo[k1], bins
// 1024 - l - 32 k1 in phi[k1] (k1 < 16), and lane 0 bin 512 in pmid.
template <int POWER_MODE>
__device__ __forceinline__ void transform_frame_eo(const Pow2Params& p, const float2* s_win, const float2* s_tw,
                                                   const float2* s_tw2, float2* tile, uint64_t* bar, uint32_t& parity,
                                                   bool& staged, int64_t row, int64_t t, bool next_ok, int64_t next_row,
                                                   int64_t next_t, int half, int lane, float (&plo)[16],
                                                   float (&phi)[16], float& pmid) {
  const float* __restrict__ x = p.wave + row * p.row_stride;
  const int64_t sa = t * p.hop - half - p.pad;
  float* stage = reinterpret_cast<float*>(tile);
  float2 a[32];
  if (staged) {
    mbar_wait(bar, parity);
    parity ^= 1;
    const float2* st2 = reinterpret_cast<const float2*>(stage);
    static_for<32>([&](auto ji) {
      constexpr int j = decltype(ji)::value;
      const float2 v = st2[lane + 32 * j], w = s_win[lane + 32 * j];
      a[brev5(j)] = make_float2(v.x * w.x, v.y * w.y);
    });
    __syncwarp();
  } else if (sa >= 0 && sa + kEoN <= p.length) {
    static_for<32>([&](auto ji) {
      constexpr int j = decltype(ji)::value;
      const float2 w = s_win[lane + 32 * j];
      const float ve = __ldg(x + sa + 2 * (lane + 32 * j)), vo = __ldg(x + sa + 2 * (lane + 32 * j) + 1);
      a[brev5(j)] = make_float2(ve * w.x, vo * w.y);
    });
  } else {
#pragma unroll 1
    for (int j = 0; j < 32; ++j) {
      const int m = lane + 32 * j;
      const int64_t ie = source_index(t * p.hop + 2 * m, p.length, p.pad, half, p.pad_mode);
      const int64_t io = source_index(t * p.hop + 2 * m + 1, p.length, p.pad, half, p.pad_mode);
      tile[m] = make_float2(ie >= 0 ? __ldg(x + ie) : 0.f, io >= 0 ? __ldg(x + io) : 0.f);
    }
    __syncwarp();
    static_for<32>([&](auto ji) {
      constexpr int j = decltype(ji)::value;
      const float2 v = tile[lane + 32 * j], w = s_win[lane + 32 * j];
      a[brev5(j)] = make_float2(v.x * w.x, v.y * w.y);
    });
    __syncwarp();
  }

  fft_regs<32, 0>(a);
  tile[lane] = a[0];
  static_for<31>([&](auto ki) {
    constexpr int k2 = decltype(ki)::value + 1;
    const float2 w = s_tw[k2 * 32 + lane];
    const float2 v = a[k2];
    tile[k2 * 33 + lane] = cmul2(v, w);
  });
  __syncwarp();
  static_for<32>([&](auto gi) {
    constexpr int g = decltype(gi)::value;
    a[brev5(g)] = tile[lane * 33 + g];
  });
  __syncwarp();
  staged = next_ok && eo_frame_bulk_ok(p, half, next_t);
  if (staged && lane == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const float* src = p.wave + next_row * p.row_stride + (next_t * p.hop - half - p.pad);
    mbar_expect_tx(bar, kEoN * 4u);
    bulk_g2s(stage, src, kEoN * 4u, bar);
  }
  fft_regs<32, 0>(a);  // a[k1] = Z[lane + 32 k1]

  const int src_lane = (32 - lane) & 31;
  static_for<16>([&](auto ki) {
    constexpr int k1 = decltype(ki)::value;
    float mr = __shfl_sync(0xffffffffu, a[31 - k1].x, src_lane);
    float mi = __shfl_sync(0xffffffffu, a[31 - k1].y, src_lane);
    if (lane == 0) {
      mr = a[(32 - k1) & 31].x;
      mi = a[(32 - k1) & 31].y;
    }
    const float zr = a[k1].x, zi = a[k1].y;
    const float er = zr + mr, ei = zi - mi;   // E (x 2, the 1/2 rides on the window)
    const float orr = zi + mi, oi = mr - zr;  // O
    const float2 w = s_tw2[k1 * 32 + lane];   // W_2048^(lane + 32 k1)
    const float tr = fmaf(orr, w.x, -oi * w.y), ti = fmaf(orr, w.y, oi * w.x);
    plo[k1] = pow_of<POWER_MODE>(er + tr, ei + ti, p.power);
    phi[k1] = pow_of<POWER_MODE>(er - tr, ei - ti, p.power);
  });
  // bin 512 = lane 0, slot 16: E = 2 Re Z, O = 2 Im Z, W^512 = -i  ->  X = E - i O
  pmid = pow_of<POWER_MODE>(2.f * a[16].x, -2.f * a[16].y, p.power);
}

__device__ __forceinline__ void eo_store_row(float* row, int lane, const float (&plo)[16], const float (&phi)[16],
                                             float pmid) {
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) {
    row[lane + 32 * k1] = plo[k1];
    row[1024 - lane - 32 * k1] = phi[k1];
  }
  if (lane == 0) row[512] = pmid;
}

__device__ __forceinline__ void eo_load_tables(const Pow2Params& p, const float2* tw_eo, float2* s_win, float2* s_tw,
                                               float2* s_tw2, int tid, int nthreads) {
  const float hs = 0.5f * p.hdr->scale;
  for (int i = tid; i < 1024; i += nthreads) {
    s_tw[i] = p.tw2d[i];
    s_win[i] = make_float2(p.window[2 * i] * hs, p.window[2 * i + 1] * hs);
  }
  for (int i = tid; i < 17 * 32; i += nthreads) s_tw2[i] = tw_eo[i];
}

struct EoNext {  // the frame that follows (row, t) in this warp's walk
  bool ok;
  int64_t row, t;
};
__device__ __forceinline__ EoNext eo_next_frame(const Pow2Params& p, const UnitCursor& cur, int f) {
  if (f == 0) return EoNext{true, cur.row, 2 * cur.ub + 1};
  return EoNext{cur.u + cur.stride < p.total_units, cur.nrow, 2 * cur.nub};
}

template <int POWER_MODE>
__global__ void __launch_bounds__(kWarps * 32, 1) stft2048_power_kernel(const Pow2Params p, const float2* tw_eo) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* s_tw = reinterpret_cast<float2*>(smem_raw);   // [32][32]
  float2* s_win = s_tw + 1024;                          // [1024] (w[2m], w[2m+1]) x scale
  float2* s_tw2 = s_win + 1024;                         // [17][32] W_2048^(l + 32 k1)
  float2* s_tile_all = s_tw2 + 17 * 32;                 // [kWarps][32 * 33]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_tile_all + kWarps * 32 * 33);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  eo_load_tables(p, tw_eo, s_win, s_tw, s_tw2, tid, blockDim.x);
  if (tid < kWarps) mbar_init(s_bar + tid, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  float2* tile = s_tile_all + warp * 32 * 33;
  uint64_t* bar = s_bar + warp;
  const int half = p.center ? kEoN / 2 : 0;
  uint32_t parity = 0;
  bool staged = false;
  UnitCursor cur;
  cur.init((int64_t)blockIdx.x * kWarps + warp, (int64_t)gridDim.x * kWarps, p.units_per_row);
  if (cur.u < p.total_units && eo_frame_bulk_ok(p, half, 2 * cur.ub)) {
    if (lane == 0) {
      mbar_expect_tx(bar, kEoN * 4u);
      bulk_g2s(tile, p.wave + cur.row * p.row_stride + (2 * cur.ub * p.hop - half - p.pad), kEoN * 4u, bar);
    }
    staged = true;
  }
  for (; cur.u < p.total_units; cur.advance()) {
#pragma unroll 1
    for (int f = 0; f < 2; ++f) {
      const int64_t t = 2 * cur.ub + f;
      if (t >= p.frames) { staged = false; break; }  // (never staged: eo_frame_bulk_ok checks t < frames)
      const EoNext nx = eo_next_frame(p, cur, f);
      float plo[16], phi[16], pmid;
      transform_frame_eo<POWER_MODE>(p, s_win, s_tw, s_tw2, tile, bar, parity, staged, cur.row, t, nx.ok, nx.row, nx.t,
                                     half, lane, plo, phi, pmid);
      eo_store_row(p.out + (cur.row * p.frames + t) * kEoBins, lane, plo, phi, pmid);
    }
  }
}

template <int POWER_MODE>
__global__ void __launch_bounds__((kWarps + kMelWarps) * 32, 1) stft2048_mel_kernel(const Pow2Params p,
                                                                                     const float2* tw_eo) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* s_tw = reinterpret_cast<float2*>(smem_raw);
  float2* s_win = s_tw + 1024;
  float2* s_tw2 = s_win + 1024;
  float2* s_tile_all = s_tw2 + 17 * 32;
  float* s_pow = reinterpret_cast<float*>(s_tile_all + kWarps * 32 * 33);  // [kEoSlots][kEoPitch]
  int64_t* s_slot = reinterpret_cast<int64_t*>(s_pow + kEoSlots * kEoPitch);
  int64_t* s_grp = s_slot + kEoSlots;
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_grp + kEoSlots);  // [kWarps] staging, full, empty
  uint64_t* s_full = s_bar + kWarps;
  uint64_t* s_empty = s_full + 1;
  MelPlan* s_plan = reinterpret_cast<MelPlan*>(s_empty + 1);
  float4* s_frags = reinterpret_cast<float4*>(s_plan + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  eo_load_tables(p, tw_eo, s_win, s_tw, s_tw2, tid, blockDim.x);
  {
    const int* src = reinterpret_cast<const int*>(p.plan);
    int* dst = reinterpret_cast<int*>(s_plan);
    for (int i = tid; i < (int)(sizeof(MelPlan) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
  }
  const int total_steps = p.plan->total_steps;
  const bool frags_in_smem = total_steps <= kEoFragSteps;
  if (frags_in_smem)
    for (int i = tid; i < total_steps * 32; i += blockDim.x) s_frags[i] = p.frags[i];
  for (int i = tid; i < kEoSlots * (kEoPitch - kEoBins); i += blockDim.x) {
    const int r = i / (kEoPitch - kEoBins), c = i - r * (kEoPitch - kEoBins);
    s_pow[r * kEoPitch + kEoBins + c] = 0.f;
  }
  if (tid < kWarps) mbar_init(s_bar + tid, 1);
  if (tid == 0) {
    mbar_init(s_full, kWarps);
    mbar_init(s_empty, kMelWarps);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  const int64_t stride = (int64_t)gridDim.x * kWarps;
  const int64_t u0 = (int64_t)blockIdx.x * kWarps;
  if (warp < kWarps) {
    reg_alloc<kFftRegs>();
    float2* tile = s_tile_all + warp * 32 * 33;
    uint64_t* bar = s_bar + warp;
    const int half = p.center ? kEoN / 2 : 0;
    uint32_t parity = 0;
    bool staged = false;
    UnitCursor cur;
    cur.init(u0 + warp, stride, p.units_per_row);
    if (cur.u < p.total_units && eo_frame_bulk_ok(p, half, 2 * cur.ub)) {
      if (lane == 0) {
        mbar_expect_tx(bar, kEoN * 4u);
        bulk_g2s(tile, p.wave + cur.row * p.row_stride + (2 * cur.ub * p.hop - half - p.pad), kEoN * 4u, bar);
      }
      staged = true;
    }
    int it = 0;
    for (int64_t base = u0; base < p.total_units; base += stride, ++it, cur.advance()) {
      const bool valid = cur.u < p.total_units;
#pragma unroll 1
      for (int f = 0; f < 2; ++f) {
        const int64_t t = 2 * cur.ub + f;
        const bool live = valid && t < p.frames;
        float plo[16], phi[16], pmid = 0.f;
        if (live) {
          const EoNext nx = eo_next_frame(p, cur, f);
          transform_frame_eo<POWER_MODE>(p, s_win, s_tw, s_tw2, tile, bar, parity, staged, cur.row, t, nx.ok, nx.row,
                                         nx.t, half, lane, plo, phi, pmid);
        } else {
          staged = false;
        }
        if (f == 0 && it >= 1) mbar_wait(s_empty, (it - 1) & 1);  // the contraction warps have drained the tile
        if (live) eo_store_row(s_pow + (size_t)(2 * warp + f) * kEoPitch, lane, plo, phi, pmid);
        if (lane == 0) {
          s_slot[2 * warp + f] = live ? (cur.row * p.frames + t) * (int64_t)p.n_mels : -1;
          s_grp[2 * warp + f] = cur.row / p.rows_per_group;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(s_full);
    }
  } else {
    reg_dealloc<kMelRegs>();
    const int mw = warp - kWarps;
    GroupMax gmax{p.stage == B200A_STAGE_FEAT ? p.group_max : nullptr, -1, -CUDART_INF_F};
    int it = 0;
    for (int64_t base = u0; base < p.total_units; base += stride, ++it) {
      mbar_wait(s_full, it & 1);
      contract_tile<kEoPitch>(p, s_plan, s_frags, frags_in_smem, mw, lane, s_pow, s_slot, s_grp, gmax);
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty);
    }
    gmax.flush();
  }
}

__global__ void prepare_tw_eo_kernel(float2* tw_eo) {  // [17][32]: W_2048^(l + 32 k1)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 17 * 32) {
    const int k1 = i >> 5, l = i & 31;
    double s, c;
    sincospi(-2.0 * (double)(l + 32 * k1) / 2048.0, &s, &c);
    tw_eo[i] = make_float2((float)c, (float)s);
  }
}

// ================================================================================================
// Mel contraction on tcgen05 (5th-generation tensor cores, accumulator in tensor memory).
//
// Moving the contraction to the tensor core frees the registers and issue slots of four mma.sync warps, so
// this body runs NW = 12 transform warps (n_fft >= 512; the transform is latency bound and scales with
// resident warps) next to four light "operand" warps.  One CTA iteration finishes R = 2 NW (32 / G) frames.
//   transform warps     publish fp32 power values into the OPERAND BUFFER, already in the tensor core's
//                       K-major core-matrix order: chunk = 8 bins; per chunk, per group of 8 frames, a
//                       256-byte block [8 rows x 8 floats]; arrive on `full`
//   operand warps       convert each block IN PLACE to [8 rows x 8 bf16 hi | 8 rows x 8 bf16 lo] (read 32 B,
//                       __syncwarp, write 16 + 16 B), fence to the async proxy, arrive on `ready`
//   warp NW + 3         issues, per k-step of 16 bins,
//                           D[:, 2 n0 : 2 n0 + 2 N] += P_hi [F_hi | F_lo] + P_lo [F_hi | F_lo]   (tcgen05.mma kind::f16,
//                       M = 64, two instructions) and commits to the buffer's `mma` barrier, which is also what
//                       the transform warps wait on before they publish into that buffer again.  The filterbank
//                       sits in shared memory as BANDED UMMA B blocks: for each k-step only the filters that
//                       are non-zero there (groups of 8: 8 rows F_hi, 8 rows F_lo), so filter f owns accumulator
//                       columns 16 (f / 8) + f % 8 and + 8.  Tile rows >= R of the M = 64 instruction alias
//                       whatever follows in shared memory and land in TMEM lanes nobody reads (an accumulator
//                       row depends on its own operand row only).
//   epilogue            ceil(R / 16) of the operand warps (one TMEM lane quadrant each) tcgen05.ld the PREVIOUS
//                       tile's accumulator (two accumulators alternate) while the tensor core works on the
//                       current one: hi + lo columns, dB / log, top_db maximum, store.
// Error-compensated bf16 carries ~2^-16 relative error per product (the 1e-4 bar; the TF32x3 mma.sync body
// ~2^-21).
// ================================================================================================
constexpr int kTcMaxSteps = 34;       // k-steps of 16 bins (n_fft = 1024: 33)
constexpr int kTcMaxN = 128;          // filters: 2 accumulator columns each, two accumulators in 512 TMEM columns
constexpr int kTcWsBBytes = 96 * 1024;  // workspace reserved for the banded B blocks

struct TcStep {
  uint32_t b_off;  // byte offset of the step's block in the B region (64 n bytes: 2 k-chunks x 2 n operand rows)
  uint32_t n;      // filters covered (multiple of 8)
  uint32_t col;    // first filter
  uint32_t kstep;  // which 16 bins: [16 kstep, 16 kstep + 16)
};
struct TcPlan {  // built on the device by prepare_tc_kernel
  int ok, steps, n_pad, b_bytes;
  TcStep step[kTcMaxSteps];
};
struct __align__(16) TcIssue {  // ready-to-issue descriptors of one k-step (operand buffer 0), built per CTA
  uint64_t a_hi, a_lo, b;
  uint32_t idesc, col;
};

template <int G>
struct TcGeo {
  using Ge = Geo<G>;
  static constexpr int NW = G == 8 ? 8 : 12;               // transform warps
  static constexpr int NB = G == 8 ? 2 : 1;                // operand buffers
  static constexpr int kThreads = (NW + kMelWarps) * 32;
  static constexpr int kFftRegs = NW == 12 ? 144 : 216;    // 384*144 + 128*80 = 65536 = 512*128; 256*216 + 128*72 = 64512
  static constexpr int kOpRegs = NW == 12 ? 80 : 72;
  static constexpr int kRows = NW * Ge::kFrames;           // frames per tile: 24 / 48 / 64
  static constexpr int kChunks = 2 * G + 2;                // ceil((16 G + 1) / 8) rounded up to even
  static constexpr int kSteps = kChunks / 2;
  static constexpr int kChunkStride = kRows * 32 + 32;     // bytes; == 32 (mod 128): conflict-free publishing
  static constexpr int kOperand = kChunks * kChunkStride;  // bytes of one operand buffer
  static constexpr int kEpi = (kRows + 15) / 16;           // epilogue warps == TMEM lane quadrants in use
  static constexpr int kFixed = NB * kOperand + 8 * (NW * Ge::kTileF2 + 32 * G) + 16 * NB * kRows +
                                (int)sizeof(TcIssue) * kTcMaxSteps + 8 * (NW + 4 * NB + 4) + 16;
  static constexpr int kBBudget = ((227 * 1024 - kFixed) / 128) * 128;
  static_assert(kSteps <= kTcMaxSteps && kRows <= 64 && kRows % 8 == 0, "one M = 64 tile per iteration");
};
struct Tc2;
int tc_b_budget(int n_fft);  // defined after Tc2

// (x, y) -> packed bf16 pair (x in the low half) and the packed pair of the residuals
__device__ __forceinline__ void split_bf16x2(float x, float y, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(y), "f"(x));
  const float rx = x - __uint_as_float(hi << 16), ry = y - __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(ry), "f"(rx));
}

template <int POWER_MODE, int G, int HG, bool KALDI>
__device__ __forceinline__ void mel_body_tc(const Pow2Params& p, unsigned char* smem_raw) {
  using Ge = Geo<G>;
  using Tc = TcGeo<G>;
  constexpr int NW = Tc::NW, NB = Tc::NB, kRows = Tc::kRows, kEpi = Tc::kEpi, kStride = Tc::kChunkStride;
  unsigned char* s_a = smem_raw;                                                   // [NB] operand buffers
  unsigned char* s_b = s_a + NB * Tc::kOperand;                                    // banded B blocks
  float2* s_tile_all = reinterpret_cast<float2*>(s_b + Tc::kBBudget);              // [NW][kTileF2]
  float2* s_tw = s_tile_all + NW * Ge::kTileF2;                                    // [32][G]
  int64_t* s_slot = reinterpret_cast<int64_t*>(s_tw + 32 * G);                     // [NB][kRows]
  int64_t* s_grp = s_slot + NB * kRows;                                            // [NB][kRows]
  TcIssue* s_issue = reinterpret_cast<TcIssue*>(s_grp + NB * kRows);               // [kTcMaxSteps]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_issue + kTcMaxSteps);            // [NW] staging
  uint64_t* s_full = s_bar + NW;                                                   // [NB] fp32 values published
  uint64_t* s_ready = s_full + NB;                                                 // [NB] converted to bf16 planes
  uint64_t* s_mma = s_ready + NB;                                                  // [NB] MMAs of the buffer complete
  uint64_t* s_tfree = s_mma + NB;                                                  // [2] accumulator read out
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_tfree + 2);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const int n_steps = p.tc->steps, n_pad = p.tc->n_pad;
  for (int i = tid; i < 32 * G; i += blockDim.x) s_tw[i] = p.tw2d[i];
  {  // the pad chunk of every operand buffer stays zero; B blocks come prepared
    uint4* a4 = reinterpret_cast<uint4*>(s_a);
    for (int i = tid; i < NB * Tc::kOperand / 16; i += blockDim.x) a4[i] = make_uint4(0, 0, 0, 0);
    const uint4* src = reinterpret_cast<const uint4*>(p.tc_b);
    uint4* b4 = reinterpret_cast<uint4*>(s_b);
    const int n16 = p.tc->b_bytes / 16;
    for (int i = tid; i < n16; i += blockDim.x) b4[i] = src[i];
  }
  if (tid < n_steps) {
    const TcStep st = p.tc->step[tid];
    const uint32_t a_addr = smem_u32(s_a) + st.kstep * 2 * kStride;
    TcIssue o;
    o.a_hi = umma_smem_desc(a_addr, kStride, 256);        // 8-row groups 256 B apart: [hi 128 B | lo 128 B]
    o.a_lo = umma_smem_desc(a_addr + 128, kStride, 256);
    o.b = umma_smem_desc(smem_u32(s_b) + st.b_off, st.n * 32, 128);  // 2 n rows: per 8 filters, 8 hi rows then 8 lo rows
    o.idesc = umma_idesc_bf16(64, 2 * (int)st.n);
    o.col = 2 * st.col;
    s_issue[tid] = o;
  }
  if (tid < NW) mbar_init(s_bar + tid, 1);
  if (tid == 0) {
    for (int i = 0; i < NB; ++i) {
      mbar_init(s_full + i, NW);
      mbar_init(s_ready + i, kMelWarps);
      mbar_init(s_mma + i, 1);
    }
    mbar_init(s_tfree + 0, kEpi);
    mbar_init(s_tfree + 1, kEpi);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // B blocks, zeroed buffers -> the tensor core
  __syncthreads();

  const int64_t stride = (int64_t)gridDim.x * NW;
  const int64_t u0 = (int64_t)blockIdx.x * NW;
  const int width = p.out_width;

  if (warp < NW) {
    // =============================== transform warps ===========================================
    reg_alloc<Tc::kFftRegs>();
    float2* tile = s_tile_all + warp * Ge::kTileF2;
    float* stage = reinterpret_cast<float*>(tile);
    uint64_t* bar = s_bar + warp;
    float wreg[32];
    load_window<G>(p, lane, wreg);
    const int half = frame_lead(p, Ge::kNfft);
    const int gi = lane / G, l = lane % G;
    uint32_t parity = 0;
    bool staged = false;
    UnitCursor cur;
    cur.init(u0 + warp, stride, p.units_per_row);
    if (bulk_eligible<G>(p, half, cur.u, cur.ub)) {
      if (lane == 0) issue_bulk<G>(p, half, cur.row, cur.ub, stage, bar);
      staged = true;
    }
    // value (row, bin k) lives at chunk (k / 8), 8-row group (row / 8): [row % 8][k % 8] floats
    const int row_a = Ge::kFrames * warp + 2 * gi;  // even: rows a and a + 1 share the 8-row group
    const int lane_off = (l >> 3) * kStride + (row_a >> 3) * 256 + (row_a & 7) * 32 + (l & 7) * 4;
    constexpr int kStepM = (G / 8) * kStride;
    int it = 0;
    for (int64_t base = u0; base < p.total_units; base += stride, ++it, cur.advance()) {
      const bool valid = cur.u < p.total_units;
      float pa[17], pb[17];
      if (valid)
        transform_unit<POWER_MODE, G, HG, true, KALDI>(p, wreg, s_tw, tile, stage, bar, parity, staged, cur, half, lane, pa,
                                                pb);
      const int b = it % NB;
      if (it >= NB) mbar_wait(s_mma + b, ((it / NB) & 1) ^ 1);  // the tensor core has consumed this buffer
      unsigned char* dst = s_a + (size_t)b * Tc::kOperand + lane_off;
      if (valid) {
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          *reinterpret_cast<float*>(dst + m * kStepM) = pa[m];
          *reinterpret_cast<float*>(dst + m * kStepM + 32) = pb[m];
        }
        if (l == 0) {  // bin n_fft/2 opens chunk 2 G: write the whole 8-float row, the other 7 are zeros
          float4* ra = reinterpret_cast<float4*>(dst + 16 * kStepM);
          ra[0] = make_float4(pa[16], 0.f, 0.f, 0.f);
          ra[1] = make_float4(0.f, 0.f, 0.f, 0.f);
          ra[2] = make_float4(pb[16], 0.f, 0.f, 0.f);
          ra[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      if (l == 0) {
        const int64_t ta = cur.ub * Ge::kFrames + 2 * gi;
        const int64_t oa = (cur.row * p.frames + ta) * (int64_t)width + p.out_col0;
        s_slot[b * kRows + row_a] = (valid && ta < p.frames) ? oa : -1;
        s_slot[b * kRows + row_a + 1] = (valid && ta + 1 < p.frames) ? oa + width : -1;
        const int64_t g = cur.row / p.rows_per_group;
        s_grp[b * kRows + row_a] = g;
        s_grp[b * kRows + row_a + 1] = g;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(s_full + b);
    }
  } else {
    // ============ operand warps (all four), MMA issue (the last), epilogue (the first kEpi of them) ============
    reg_dealloc<Tc::kOpRegs>();  // one setmaxnreg for the whole warpgroup
    const int cw = warp - NW;    // == TMEM lane quadrant (NW % 4 == 0)
    static_assert(NW % 4 == 0, "operand warp i must own TMEM lane quadrant i");
    const uint32_t acc_cols = 2 * n_pad;  // one accumulator; two of them alternate
    const uint32_t tmem_cols =
        acc_cols <= 16 ? 32u : (acc_cols <= 32 ? 64u : (acc_cols <= 64 ? 128u : (acc_cols <= 128 ? 256u : 512u)));
    if (cw == 0) tmem_alloc(s_tmem, tmem_cols);
    tc_fence_before();
    asm volatile("bar.sync 1, %0;" ::"n"(kMelWarps * 32) : "memory");
    tc_fence_after();
    const uint32_t tmem_d = *reinterpret_cast<volatile uint32_t*>(s_tmem);
    GroupMax gmax{p.stage == B200A_STAGE_FEAT ? p.group_max : nullptr, -1, -CUDART_INF_F};
    // conversion items: (chunk, row) with the row fastest (8 consecutive lanes = one 256-byte block)
    constexpr int kItems = (Tc::kChunks - 1) * kRows, kRounds = (kItems + 127) / 128;
    const int erow = 16 * cw + (lane & 15);  // epilogue: thread i < 16 owns tile row 16 cw + i == TMEM lane 32 cw + i

    // accumulator of tile `t` -> dB / log -> global; o / g: output offset and top_db group of this thread's row
    auto epilogue = [&](int t, int64_t o, int64_t g) {
      const uint32_t acc = tmem_d + ((uint32_t)(32 * cw) << 16) + (uint32_t)(t & 1) * acc_cols;
      const bool row_ok = lane < 16 && o >= 0;
#pragma unroll 1
      for (int f0 = 0; f0 < n_pad; f0 += 16) {
        float u[16], w[16], v[16];
        tmem_ld16(acc + 2 * f0, u);
        tmem_ld16(acc + 2 * f0 + 16, w);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          v[q] = u[q] + u[q + 8];
          v[q + 8] = w[q] + w[q + 8];
        }
        if (p.k_log) {  // Kaldi fbank: log(max(mel, FLT_EPSILON)), kaldi.py:629-631
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = 0.69314718055994531f * __log2f(fmaxf(v[q], kKaldiEps));
        }
        if (p.stage == B200A_STAGE_FEAT) {
          // one thread owns a whole row here, so the logarithms are the epilogue's critical path: MUFU.LG2
          // (2^-22 absolute on the log2, i.e. < 1e-5 dB) instead of the ~30-instruction log10f
          float mx = -CUDART_INF_F;
          const float scale = p.log_mels ? 0.69314718055994531f : p.db_mult * 0.30102999566398120f;
          const float offs = p.log_mels ? 0.f : p.db_offset;
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float arg = p.log_mels ? v[q] + 1e-6f : fmaxf(v[q], p.db_amin);
            v[q] = fmaf(scale, __log2f(arg), -offs);
            if (f0 + q < p.n_mels) mx = fmaxf(mx, v[q]);
          }
          gmax.add(g, mx, row_ok);
        }
        if (row_ok) {
          float* dst = p.out + o + f0;
          if (p.out_vec >= 4 && f0 + 16 <= p.n_mels) {
#pragma unroll
            for (int q = 0; q < 16; q += 4)
              *reinterpret_cast<float4*>(dst + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
          } else {
#pragma unroll
            for (int q = 0; q < 16; ++q)
              if (f0 + q < p.n_mels) dst[q] = v[q];
          }
        }
      }
      tc_fence_before();  // my tcgen05.ld are done before this accumulator is handed back
      __syncwarp();
      if (lane == 0) mbar_arrive(s_tfree + (t & 1));
    };

    int it = 0;
    int64_t o_prev = -1, g_prev = -1;
    for (int64_t base = u0; base < p.total_units; base += stride, ++it) {
      const int b = it % NB;
      const uint32_t use = (uint32_t)(it / NB) & 1;  // parity of this use of buffer b
      mbar_wait(s_full + b, use);
      int64_t o_cur = -1, g_cur = -1;
      if (cw < kEpi && erow < kRows) {
        o_cur = s_slot[b * kRows + erow];
        g_cur = s_grp[b * kRows + erow];
      }
      {
        unsigned char* buf = s_a + (size_t)b * Tc::kOperand;
#pragma unroll 2
        for (int rd = 0; rd < kRounds; ++rd) {
          const int item = rd * 128 + cw * 32 + lane;
          const bool live = item < kItems;
          const int chunk = item / kRows, row = item % kRows;
          unsigned char* blk = buf + chunk * kStride + (row >> 3) * 256;
          float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
          if (live) {
            v0 = *reinterpret_cast<const float4*>(blk + (row & 7) * 32);
            v1 = *reinterpret_cast<const float4*>(blk + (row & 7) * 32 + 16);
          }
          uint4 hi, lo;
          split_bf16x2(v0.x, v0.y, hi.x, lo.x);
          split_bf16x2(v0.z, v0.w, hi.y, lo.y);
          split_bf16x2(v1.x, v1.y, hi.z, lo.z);
          split_bf16x2(v1.z, v1.w, hi.w, lo.w);
          __syncwarp();  // the 8 rows of a block are 8 lanes of this warp: all reads before any write
          if (live) {
            *reinterpret_cast<uint4*>(blk + (row & 7) * 16) = hi;
            *reinterpret_cast<uint4*>(blk + 128 + (row & 7) * 16) = lo;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my plane writes -> async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(s_ready + b);
      if (cw == kMelWarps - 1) {
        // ---- issue: D[it & 1][:, 2 n0 : 2 n0 + 2 N] += P_hi [F_hi | F_lo] + P_lo [F_hi | F_lo] per k-step ----
        mbar_wait(s_ready + b, use);
        if (it >= 2) mbar_wait(s_tfree + (it & 1), ((it >> 1) & 1) ^ 1);  // tile it - 2 has left this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_d + (uint32_t)(it & 1) * acc_cols;
        const uint64_t a_off = (uint64_t)((uint32_t)b * (Tc::kOperand >> 4));
#pragma unroll 3
        for (int s = 0; s < n_steps; ++s) {
          const TcIssue e = s_issue[s];
          if (elect_one()) {
            umma_bf16(acc + e.col, e.a_hi + a_off, e.b, e.idesc, s > 0 ? 1u : 0u);  // step 0 spans every column
            umma_bf16(acc + e.col, e.a_lo + a_off, e.b, e.idesc, 1u);
          }
        }
        if (elect_one()) umma_commit(s_mma + b);
        __syncwarp();
      }
      if (cw < kEpi && it > 0) {  // the previous tile's accumulator, while the tensor core works on this one
        // NB == 1: tile it could only be published (`full`, waited above) after the MMAs of tile it - 1 completed,
        // and a second look at that barrier could race with the completion of tile it
        if (NB > 1) mbar_wait(s_mma + (it - 1) % NB, (uint32_t)((it - 1) / NB) & 1);
        tc_fence_after();
        epilogue(it - 1, o_prev, g_prev);
      }
      o_prev = o_cur;
      g_prev = g_cur;
    }
    if (cw < kEpi && it > 0) {
      mbar_wait(s_mma + (it - 1) % NB, (uint32_t)((it - 1) / NB) & 1);
      tc_fence_after();
      epilogue(it - 1, o_prev, g_prev);
    }
    gmax.flush();
    tc_fence_before();
    asm volatile("bar.sync 1, %0;" ::"n"(kMelWarps * 32) : "memory");  // every tcgen05.ld is done
    if (cw == 0) tmem_dealloc(tmem_d, tmem_cols);
  }
}

// Position kp of the contraction's (permuted) K axis -> spectrum bin, or -1 for a padding position.
//   perm_g == 0: identity (operand buffer in bin order; n_fft 256 / 512 body).
//   perm_g == G: the n_fft = 32 G register FFT leaves lane l with bins l + G m; the transform warps publish the PAIR
//                (m, m + 1) of a lane as one packed bf16x2 word, so K position 2 G (m >> 1) + 2 l + (m & 1) holds
//                bin l + G m, and the Nyquist bin sits at position n_fft / 2.
__host__ __device__ __forceinline__ int tc_bin_of(int kp, int perm_g, int n_bins) {
  if (perm_g == 0) return kp < n_bins ? kp : -1;
  const int half = 16 * perm_g;  // n_fft / 2
  if (kp == half) return half;
  if (kp > half) return -1;
  const int j = kp / (2 * perm_g), rem = kp % (2 * perm_g);
  return (rem >> 1) + perm_g * (2 * j + (rem & 1));
}

// Banded bf16 hi / lo UMMA B blocks and their step table.  One block.
__global__ void prepare_tc_kernel(const float* __restrict__ fb, int n_bins, int n_mels, int n_fft, int budget,
                                  int perm_g, TcPlan* plan, unsigned char* blocks) {
  __shared__ int s_lo[kTcMaxSteps], s_hi[kTcMaxSteps];
  __shared__ TcPlan s_plan;
  const int k_steps = (n_fft / 16 + 2) / 2;  // (2 G + 2) / 2
  const int n_pad = (n_mels + 15) / 16 * 16;
  if ((int)threadIdx.x < k_steps) {
    int lo = n_mels, hi = -1;
    for (int kp = 16 * threadIdx.x; kp < 16 * (int)threadIdx.x + 16; ++kp) {
      const int k = tc_bin_of(kp, perm_g, n_bins);
      if (k < 0) continue;
      for (int n = 0; n < n_mels; ++n)
        if (fb[(size_t)k * n_mels + n] != 0.f) {
          lo = min(lo, n);
          hi = max(hi, n);
        }
    }
    s_lo[threadIdx.x] = lo;
    s_hi[threadIdx.x] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int steps = 0, off = 0;
    for (int s = 0; s < k_steps; ++s) {
      int n0, n;
      if (s == 0) {  // the first MMA clears the accumulator: every column
        n0 = 0;
        n = n_pad;
      } else if (s_hi[s] < 0) {
        continue;
      } else {
        n0 = s_lo[s] / 8 * 8;
        n = (s_hi[s] + 1 - n0 + 7) / 8 * 8;
      }
      s_plan.step[steps] = TcStep{(uint32_t)off, (uint32_t)n, (uint32_t)n0, (uint32_t)s};
      off += n * 64;
      ++steps;
    }
    s_plan.steps = steps;
    s_plan.n_pad = n_pad;
    s_plan.b_bytes = off;
    s_plan.ok = (off <= budget && off <= kTcWsBBytes && n_pad <= kTcMaxN) ? 1 : 0;
  }
  __syncthreads();
  if (s_plan.ok) {
    for (int s = 0; s < s_plan.steps; ++s) {
      const TcStep st = s_plan.step[s];
      for (int i = threadIdx.x; i < (int)st.n * 16; i += blockDim.x) {
        const int nl = i >> 4, kk = i & 15;
        const int n = (int)st.col + nl, k = tc_bin_of(16 * (int)st.kstep + kk, perm_g, n_bins);
        const float v = (n < n_mels && k >= 0) ? fb[(size_t)k * n_mels + n] : 0.f;
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(h));
        // operand row of filter nl: 16 (nl / 8) + nl % 8 for F_hi, + 8 for F_lo; 2 n rows x 16 B per k chunk
        const size_t row = (size_t)(nl >> 3) * 16 + (nl & 7);
        const size_t o = st.b_off + (size_t)(kk >> 3) * st.n * 32 + row * 16 + (size_t)(kk & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(blocks + o) = h;
        *reinterpret_cast<__nv_bfloat16*>(blocks + o + 128) = lo;
      }
    }
  }
  for (int i = threadIdx.x; i < (int)(sizeof(TcPlan) / sizeof(int)); i += blockDim.x)
    reinterpret_cast<int*>(plan)[i] = reinterpret_cast<const int*>(&s_plan)[i];
}

// ================================================================================================
// n_fft = 1024 contraction, second generation (mel_body_tc2).  What changed against mel_body_tc:
//   * the transform warps publish bf16 hi / lo words DIRECTLY (no fp32 staging, no conversion pass by other warps,
//     one barrier hop less): a lane packs the powers of its bins (l + 32 m, l + 32 (m + 1)) into one bf16x2 word,
//     which makes the contraction's K axis a permutation of the spectrum (tc_bin_of) -- the banded filterbank
//     blocks are built in the same order, so nothing else notices;
//   * the hi and lo planes of a frame are two ROWS of the same M = 64 operand tile (row group 2 (r / 8) = hi,
//     + 1 = lo), so ONE tcgen05.mma per k-step produces P_hi F and P_lo F side by side in tensor memory lanes
//     i and i + 8 of the frame group's lane quadrant, and the epilogue adds them with one shuffle.  Against two
//     M = 64 instructions per k-step (each reading 64 operand rows of which 24 were real) this halves the
//     tensor core's shared-memory reads: 48 of the 64 rows are real now.
//   * the four service warps only issue MMAs (one of them) and run the epilogue (three: one per 8-frame group).
// ================================================================================================
struct Tc2 {
  static constexpr int NW = 12;                       // transform warps
  static constexpr int kThreads = (NW + kMelWarps) * 32;
  static constexpr int kFftRegs = 144, kSvcRegs = 80;  // 384*144 + 128*80 = 65536
  static constexpr int kRows = 2 * NW;                // frames per tile: 24
  static constexpr int kQuads = (kRows + 7) / 8;      // 8-frame groups == TMEM lane quadrants in use == epilogue warps
  static constexpr int kChunks = 66;                  // 8-position K chunks: 512 + Nyquist chunk + one of padding
  static constexpr int kSteps = kChunks / 2;
  static constexpr int kChunkStride = kQuads * 256 + 16;  // bytes: [hi 128 B | lo 128 B] per frame group; / 16 odd
  static constexpr int kOperand = ((kChunks * kChunkStride + 1024 + 127) / 128) * 128;  // + the M = 64 over-read
  static constexpr int kFixed = kOperand + 8 * (NW * Geo<32>::kTileF2 + 32 * 32) + 16 * kRows + 32 * kTcMaxSteps +
                                8 * (NW + 8) + 16;
  static constexpr int kBBudget = ((227 * 1024 - kFixed) / 128) * 128;
  static_assert(kQuads <= 3 && (kChunkStride / 16) % 2 == 1 && kSteps <= kTcMaxSteps, "layout");
  static_assert(kBBudget >= 40 * 1024, "banded filterbank blocks need room");
};
struct __align__(16) TcIssue2 {  // ready-to-issue descriptors of one k-step
  uint64_t a, b;
  uint32_t idesc, col, pad0, pad1;
};

template <int POWER_MODE, int HG, bool KALDI>
__device__ __forceinline__ void mel_body_tc2(const Pow2Params& p, unsigned char* smem_raw) {
  constexpr int G = 32;
  using Ge = Geo<G>;
  constexpr int NW = Tc2::NW, kRows = Tc2::kRows, kQuads = Tc2::kQuads, kStride = Tc2::kChunkStride;
  unsigned char* s_a = smem_raw;                                                   // operand buffer
  unsigned char* s_b = s_a + Tc2::kOperand;                                        // banded B blocks
  float2* s_tile_all = reinterpret_cast<float2*>(s_b + Tc2::kBBudget);             // [NW][kTileF2]
  float2* s_tw = s_tile_all + NW * Ge::kTileF2;                                    // [32][G]
  int64_t* s_slot = reinterpret_cast<int64_t*>(s_tw + 32 * G);                     // [kRows] output offsets
  int64_t* s_grp = s_slot + kRows;                                                 // [kRows] top_db groups
  TcIssue2* s_issue = reinterpret_cast<TcIssue2*>(s_grp + kRows);                  // [kTcMaxSteps]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_issue + kTcMaxSteps);            // [NW] staging
  uint64_t* s_full = s_bar + NW;                                                   // operand tile published
  uint64_t* s_mma = s_full + 1;                                                    // its MMAs complete
  uint64_t* s_tfree = s_mma + 1;                                                   // [2] accumulator read out
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_tfree + 2);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // provably warp-uniform
  const int n_steps = p.tc->steps, n_pad = p.tc->n_pad;
  for (int i = tid; i < 32 * G; i += blockDim.x) s_tw[i] = p.tw2d[i];
  {  // positions nobody publishes (the Nyquist chunk's tail, the padding chunk) stay zero; B blocks come prepared
    uint4* a4 = reinterpret_cast<uint4*>(s_a);
    for (int i = tid; i < Tc2::kOperand / 16; i += blockDim.x) a4[i] = make_uint4(0, 0, 0, 0);
    const uint4* src = reinterpret_cast<const uint4*>(p.tc_b);
    uint4* b4 = reinterpret_cast<uint4*>(s_b);
    const int n16 = p.tc->b_bytes / 16;
    for (int i = tid; i < n16; i += blockDim.x) b4[i] = src[i];
  }
  if (tid < n_steps) {
    const TcStep st = p.tc->step[tid];
    TcIssue2 o;
    // A: 64 rows = 8 row groups 128 B apart (hi / lo of four 8-frame groups), the two K chunks kStride apart
    o.a = umma_smem_desc(smem_u32(s_a) + st.kstep * 2 * kStride, kStride, 128);
    o.b = umma_smem_desc(smem_u32(s_b) + st.b_off, st.n * 32, 128);  // 2 n rows: per 8 filters, 8 hi rows then 8 lo rows
    o.idesc = umma_idesc_bf16(64, 2 * (int)st.n);
    o.col = 2 * st.col;
    o.pad0 = o.pad1 = 0;
    s_issue[tid] = o;
  }
  if (tid < NW) mbar_init(s_bar + tid, 1);
  if (tid == 0) {
    mbar_init(s_full, NW);
    mbar_init(s_mma, 1);
    mbar_init(s_tfree + 0, kQuads);
    mbar_init(s_tfree + 1, kQuads);
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // B blocks, zeroed buffer -> the tensor core
  __syncthreads();

  const int64_t stride = (int64_t)gridDim.x * NW;
  const int64_t u0 = (int64_t)blockIdx.x * NW;
  const int width = p.out_width;

  if (warp < NW) {
    // =============================== transform warps ===========================================
    reg_alloc<Tc2::kFftRegs>();
    float2* tile = s_tile_all + warp * Ge::kTileF2;
    float* stage = reinterpret_cast<float*>(tile);
    uint64_t* bar = s_bar + warp;
    float wreg[32];
    load_window<G>(p, lane, wreg);
    const int half = frame_lead(p, Ge::kNfft);
    const int l = lane;
    uint32_t parity = 0;
    bool staged = false;
    UnitCursor cur;
    cur.init(u0 + warp, stride, p.units_per_row);
    if (bulk_eligible<G>(p, half, cur.u, cur.ub)) {
      if (lane == 0) issue_bulk<G>(p, half, cur.row, cur.ub, stage, bar);
      staged = true;
    }
    // frames 2 warp (a) and 2 warp + 1 (b) are rows (r & 7) of frame group r >> 3; K position 64 j + 2 l (+1):
    // chunk 8 j + (l >> 2), byte (l & 3) * 4 of the 16-byte row.  32 lanes -> 8 chunks x 4 words: conflict free.
    const int row_a = 2 * warp;
    unsigned char* dst = s_a + (l >> 2) * kStride + (row_a >> 3) * 256 + (row_a & 7) * 16 + (l & 3) * 4;
    int it = 0;
    for (int64_t base = u0; base < p.total_units; base += stride, ++it, cur.advance()) {
      const bool valid = cur.u < p.total_units;
      float pa[17], pb[17];
      if (valid)
        transform_unit<POWER_MODE, G, HG, true, KALDI>(p, wreg, s_tw, tile, stage, bar, parity, staged, cur, half, lane, pa,
                                                       pb);
      if (it >= 1) mbar_wait(s_mma, (uint32_t)(it & 1) ^ 1u);  // the tensor core has consumed the previous tile
      if (valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint32_t ha, la, hb, lb;
          split_bf16x2(pa[2 * j], pa[2 * j + 1], ha, la);
          split_bf16x2(pb[2 * j], pb[2 * j + 1], hb, lb);
          unsigned char* d = dst + j * (8 * kStride);
          *reinterpret_cast<uint32_t*>(d) = ha;
          *reinterpret_cast<uint32_t*>(d + 16) = hb;
          *reinterpret_cast<uint32_t*>(d + 128) = la;
          *reinterpret_cast<uint32_t*>(d + 144) = lb;
        }
        if (l == 0) {  // bin n_fft / 2 opens chunk 64; its partner position is padding
          uint32_t ha, la, hb, lb;
          split_bf16x2(pa[16], 0.f, ha, la);
          split_bf16x2(pb[16], 0.f, hb, lb);
          unsigned char* d = dst + 8 * (8 * kStride);
          *reinterpret_cast<uint32_t*>(d) = ha;
          *reinterpret_cast<uint32_t*>(d + 16) = hb;
          *reinterpret_cast<uint32_t*>(d + 128) = la;
          *reinterpret_cast<uint32_t*>(d + 144) = lb;
        }
      }
      if (l == 0) {
        const int64_t ta = cur.ub * Ge::kFrames;
        const int64_t oa = (cur.row * p.frames + ta) * (int64_t)width + p.out_col0;
        s_slot[row_a] = (valid && ta < p.frames) ? oa : -1;
        s_slot[row_a + 1] = (valid && ta + 1 < p.frames) ? oa + width : -1;
        const int64_t g = cur.row / p.rows_per_group;
        s_grp[row_a] = g;
        s_grp[row_a + 1] = g;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my operand words -> the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(s_full);
    }
  } else {
    // ============ service warps: epilogue (the first kQuads of them), MMA issue (the last) ============
    reg_dealloc<Tc2::kSvcRegs>();  // one setmaxnreg for the whole warpgroup
    const int cw = warp - NW;      // == TMEM lane quadrant (NW % 4 == 0)
    static_assert(NW % 4 == 0, "service warp i must own TMEM lane quadrant i");
    const uint32_t acc_cols = 2 * n_pad;  // one accumulator; two of them alternate
    const uint32_t tmem_cols =
        acc_cols <= 16 ? 32u : (acc_cols <= 32 ? 64u : (acc_cols <= 64 ? 128u : (acc_cols <= 128 ? 256u : 512u)));
    if (cw == 0) tmem_alloc(s_tmem, tmem_cols);
    tc_fence_before();
    asm volatile("bar.sync 1, %0;" ::"n"(kMelWarps * 32) : "memory");
    tc_fence_after();
    const uint32_t tmem_d = *reinterpret_cast<volatile uint32_t*>(s_tmem);
    GroupMax gmax{p.stage == B200A_STAGE_FEAT ? p.group_max : nullptr, -1, -CUDART_INF_F};
    // epilogue thread i < 8 owns frame 8 cw + i: its hi-plane row is TMEM lane 32 cw + i, its lo-plane row lane + 8
    const int erow = 8 * cw + (lane & 7);

    auto epilogue = [&](int t, int64_t o, int64_t g) {
      const uint32_t acc = tmem_d + ((uint32_t)(32 * cw) << 16) + (uint32_t)(t & 1) * acc_cols;
      const bool row_ok = lane < 8 && o >= 0;
#pragma unroll 1
      for (int f0 = 0; f0 < n_pad; f0 += 16) {
        float u[16], w[16], v[16];
        tmem_ld16(acc + 2 * f0, u);
        tmem_ld16(acc + 2 * f0 + 16, w);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          v[q] = u[q] + u[q + 8];          // x F_hi + x F_lo
          v[q + 8] = w[q] + w[q + 8];
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] += __shfl_down_sync(0xffffffffu, v[q], 8);  // P_hi row + P_lo row
        if (p.k_log) {  // Kaldi fbank: log(max(mel, FLT_EPSILON)), kaldi.py:629-631
#pragma unroll
          for (int q = 0; q < 16; ++q) v[q] = 0.69314718055994531f * __log2f(fmaxf(v[q], kKaldiEps));
        }
        if (p.stage == B200A_STAGE_FEAT) {
          float mx = -CUDART_INF_F;
          const float scale = p.log_mels ? 0.69314718055994531f : p.db_mult * 0.30102999566398120f;
          const float offs = p.log_mels ? 0.f : p.db_offset;
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float arg = p.log_mels ? v[q] + 1e-6f : fmaxf(v[q], p.db_amin);
            v[q] = fmaf(scale, __log2f(arg), -offs);
            if (f0 + q < p.n_mels) mx = fmaxf(mx, v[q]);
          }
          gmax.add(g, mx, row_ok);
        }
        if (row_ok) {
          float* dsto = p.out + o + f0;
          if (p.out_vec >= 4 && f0 + 16 <= p.n_mels) {
#pragma unroll
            for (int q = 0; q < 16; q += 4)
              *reinterpret_cast<float4*>(dsto + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
          } else {
#pragma unroll
            for (int q = 0; q < 16; ++q)
              if (f0 + q < p.n_mels) dsto[q] = v[q];
          }
        }
      }
      tc_fence_before();  // my tcgen05.ld are done before this accumulator is handed back
      __syncwarp();
      if (lane == 0) mbar_arrive(s_tfree + (t & 1));
    };

    int it = 0;
    int64_t o_prev = -1, g_prev = -1;
    for (int64_t base = u0; base < p.total_units; base += stride, ++it) {
      mbar_wait(s_full, (uint32_t)it & 1u);
      if (cw == kMelWarps - 1) {
        // ---- issue: D[it & 1][:, 2 n0 : 2 n0 + 2 N] += [P_hi ; P_lo] [F_hi | F_lo], ONE instruction per k-step ----
        if (it >= 2) mbar_wait(s_tfree + (it & 1), ((uint32_t)(it >> 1) & 1u) ^ 1u);  // tile it - 2 has left this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_d + (uint32_t)(it & 1) * acc_cols;
#pragma unroll 3
        for (int s = 0; s < n_steps; ++s) {
          const TcIssue2 e = s_issue[s];
          if (elect_one()) umma_bf16(acc + e.col, e.a, e.b, e.idesc, s > 0 ? 1u : 0u);  // step 0 spans every column
        }
        if (elect_one()) umma_commit(s_mma);
        __syncwarp();
      } else if (cw < kQuads) {
        // the slots of tile `it` must be read before its MMAs complete (then the transform warps overwrite them)
        int64_t o_cur = -1, g_cur = -1;
        if (erow < kRows) {
          o_cur = s_slot[erow];
          g_cur = s_grp[erow];
        }
        if (it > 0) {  // tile it could only be published after the MMAs of tile it - 1 completed (single operand buffer)
          tc_fence_after();
          epilogue(it - 1, o_prev, g_prev);
        }
        o_prev = o_cur;
        g_prev = g_cur;
      }
    }
    if (cw < kQuads && cw != kMelWarps - 1 && it > 0) {
      mbar_wait(s_mma, (uint32_t)(it - 1) & 1u);
      tc_fence_after();
      epilogue(it - 1, o_prev, g_prev);
    }
    gmax.flush();
    tc_fence_before();
    asm volatile("bar.sync 1, %0;" ::"n"(kMelWarps * 32) : "memory");  // every tcgen05.ld is done
    if (cw == 0) tmem_dealloc(tmem_d, tmem_cols);
  }
}

int tc_b_budget(int n_fft) {
  return n_fft == 1024 ? Tc2::kBBudget : (n_fft == 512 ? TcGeo<16>::kBBudget : TcGeo<8>::kBBudget);
}

// The mel / MFCC-feature kernel: tcgen05 contraction when the prepared plan says the banded filterbank fits
// shared memory (every real mel / linear filterbank does), mma.sync contraction otherwise.
template <int POWER_MODE, int G, int HG, bool KALDI>
__global__ void __launch_bounds__(TcGeo<G>::kThreads, 1) stft_pow2_mel_kernel(const Pow2Params p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // 512 threads start with 128 registers each: 8 x 32 x 192 + 4 x 32 x 96 + 4 x 32 x 24 <= 65536
  constexpr int kMmaFftRegs = TcGeo<G>::kThreads == 512 ? 192 : kFftRegs;
  if (p.tc != nullptr && p.tc->ok) {
    if constexpr (G == 32) mel_body_tc2<POWER_MODE, HG, KALDI>(p, smem_raw);
    else mel_body_tc<POWER_MODE, G, HG, KALDI>(p, smem_raw);
  } else {
    mel_body_mma<POWER_MODE, G, HG, kMmaFftRegs, KALDI>(p, smem_raw);
  }
}

// ================================================================================================
// Inverse STFT frames on the register FFT (n_fft = 256 / 512 / 1024): the first half of b200a_istft_run.
// A lane group rebuilds the PAIR of frames (a, b) from their two Hermitian spectra with ONE complex transform:
//   Z[k] = A[k] + i B[k] (k <= N/2),  Z[N-k] = conj(A[k]) + i conj(B[k]);  z = IFFT(Z) = a + i b
// computed as conj(FFT(conj Z)) / N with the forward passes of transform_unit (32-point register DFT, twiddle,
// transpose through the padded tile, G-point register DFTs).  Lane l loads bins n = l + G j and ends with time
// samples n = l + G m, which it multiplies by window / (N * forward normalisation) and stores to the frame buffer.
// C2R semantics: the imaginary parts of bins 0 and N/2 are ignored.
// ================================================================================================
struct IstftPow2Params {
  const float2* spec;  // logical [rows][bins][frames], element strides below
  int64_t stride_row, stride_bin, stride_frame;
  int64_t frames, units_per_row, total_units;
  float* frame_buf;  // [rows][frames][n_fft]
  const float* window;
  const float2* tw2d;
  const WsHeader* hdr;
};

constexpr int kIsWarps = 16;

template <int G>
__global__ void __launch_bounds__(kIsWarps * 32, 1) istft_pow2_kernel(const IstftPow2Params p) {
  using Ge = Geo<G>;
  constexpr int N = Ge::kNfft, NG = Ge::kGroups;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float2* s_tw = reinterpret_cast<float2*>(smem_raw);  // [32][G]
  float2* s_tile_all = s_tw + 32 * G;                  // [kIsWarps][kTileF2]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 32 * G; i += blockDim.x) s_tw[i] = p.tw2d[i];
  __syncthreads();
  float2* grp_tile = s_tile_all + warp * Ge::kTileF2 + (lane / G) * Ge::kRegion;
  const int gi = lane / G, l = lane % G;
  float wreg[32];
  {
    const float gain = 1.f / ((float)N * p.hdr->scale);
#pragma unroll
    for (int m = 0; m < 32; ++m) wreg[m] = p.window[l + G * m] * gain;
  }
  UnitCursor cur;
  cur.init((int64_t)blockIdx.x * kIsWarps + warp, (int64_t)gridDim.x * kIsWarps, p.units_per_row);
  for (; cur.u < p.total_units; cur.advance()) {
    const int64_t ta = cur.ub * Ge::kFrames + 2 * gi, tb = ta + 1;
    const bool has_a = ta < p.frames, has_b = tb < p.frames;
    const float2* __restrict__ sp = p.spec + cur.row * p.stride_row;
    float2 a[32];
    static_for<32>([&](auto ji) {
      constexpr int j = decltype(ji)::value;
      const int n = l + G * j;
      const int kk = n <= N / 2 ? n : N - n;
      float2 za = make_float2(0.f, 0.f), zb = za;
      if (has_a) za = __ldg(sp + kk * p.stride_bin + ta * p.stride_frame);
      if (has_b) zb = __ldg(sp + kk * p.stride_bin + tb * p.stride_frame);
      if (kk == 0 || kk == N / 2) za.y = zb.y = 0.f;
      if (n > N / 2) {
        za.y = -za.y;
        zb.y = -zb.y;
      }
      a[brev5(j)] = make_float2(za.x - zb.y, -(za.y + zb.x));  // conj(Z[n])
    });
    fft_regs<32, 0>(a);
    grp_tile[l] = a[0];
    static_for<31>([&](auto ki) {
      constexpr int k2 = decltype(ki)::value + 1;
      const float2 w = s_tw[k2 * G + l];
      const float2 v = a[k2];
      grp_tile[k2 * Ge::kRowLd + l] = cmul2(v, w);
    });
    __syncwarp();
    static_for<32>([&](auto si) {
      constexpr int s = decltype(si)::value;
      constexpr int q = s / G, g = s % G;
      a[q * G + brev<Ge::kLogG>(g)] = grp_tile[(l + G * q) * Ge::kRowLd + g];
    });
    __syncwarp();
    static_for<NG>([&](auto qi) { fft_regs<G, decltype(qi)::value * G>(a); });
    // a[(m % NG) G + m / NG] = FFT(conj Z)[l + G m] = N (a[n] - i b[n])
    float* fa = p.frame_buf + (cur.row * p.frames + ta) * N + l;
    static_for<32>([&](auto mi) {
      constexpr int m = decltype(mi)::value;
      constexpr int slot = (m % NG) * G + m / NG;
      if (has_a) fa[G * m] = a[slot].x * wreg[m];
      if (has_b) fa[N + G * m] = -a[slot].y * wreg[m];
    });
  }
}

// ---- table preparation ------------------------------------------------------------------------
__global__ void prepare_tw2d_kernel(float2* tw2d, int G) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // i = k2 * G + g
  if (i < 32 * G) {
    const int k2 = i / G, g = i - k2 * G;
    double s, c;
    sincospi(-2.0 * (double)(k2 * g) / (double)(32 * G), &s, &c);
    tw2d[i] = make_float2((float)c, (float)s);
  }
}

// Builds the mel contraction plan: per group of 8 filters the 8-bin k-steps its non-zero bins span,
// groups spread over the contraction warps by descending size; and the filterbank values split
// into TF32 hi/lo parts in mma.m16n8k8 B-fragment order.
__global__ void prepare_mma_kernel(const float* __restrict__ fb, const int2* __restrict__ bands, int n_bins, int n_mels,
                                   int n_tiles, MelPlan* plan, float4* frags) {
  __shared__ int t_kstart[kMaxItems], t_steps[kMaxItems];
  if (threadIdx.x == 0) {
    int total = 0;
    for (int t = 0; t < n_tiles; ++t) {
      int lo = n_bins, hi = 0;
      for (int m = 8 * t; m < min(8 * t + 8, n_mels); ++m) {
        const int2 b = bands[m];
        if (b.y > b.x) { lo = min(lo, b.x); hi = max(hi, b.y); }
      }
      t_kstart[t] = hi > lo ? (lo & ~7) : 0;
      t_steps[t] = hi > lo ? (hi - t_kstart[t] + 7) / 8 : 0;
      plan->items[t] = MelItem{t, t_kstart[t], t_steps[t], total};
      total += t_steps[t];
    }
    plan->n_tiles = n_tiles;
    plan->n_items = n_tiles;
    plan->total_steps = total;
    plan->pad = 0;
    // longest-processing-time-first assignment of the groups to the contraction warps
    int load[kWarps];
    bool used[kMaxItems];
    for (int w = 0; w < kWarps; ++w) { load[w] = 0; plan->warp_cnt[w] = 0; }
    for (int i = 0; i < n_tiles; ++i) used[i] = false;
    for (int k = 0; k < n_tiles; ++k) {
      int best = -1;
      for (int i = 0; i < n_tiles; ++i)
        if (!used[i] && (best < 0 || t_steps[i] > t_steps[best])) best = i;
      used[best] = true;
      int w = 0;
      for (int q = 1; q < kMelWarps; ++q)
        if (load[q] < load[w] || (load[q] == load[w] && plan->warp_cnt[q] < plan->warp_cnt[w])) w = q;
      plan->warp_items[w][plan->warp_cnt[w]++] = best;
      load[w] += t_steps[best] + 2;  // + epilogue cost
    }
  }
  __syncthreads();
  int off = 0;
  for (int t = 0; t < n_tiles; ++t) {
    for (int i = threadIdx.x; i < t_steps[t] * 32; i += blockDim.x) {
      const int s = i >> 5, lane = i & 31;
      const int n = 8 * t + (lane >> 2);
      const int k0 = t_kstart[t] + 8 * s + (lane & 3), k1 = k0 + 4;
      const float b0 = (n < n_mels && k0 < n_bins) ? fb[(size_t)k0 * n_mels + n] : 0.f;
      const float b1 = (n < n_mels && k1 < n_bins) ? fb[(size_t)k1 * n_mels + n] : 0.f;
      const float b0h = __uint_as_float(__float_as_uint(b0) & 0xffffe000u);
      const float b1h = __uint_as_float(__float_as_uint(b1) & 0xffffe000u);
      frags[(size_t)(off + s) * 32 + lane] = make_float4(b0h, b1h, b0 - b0h, b1 - b1h);
    }
    off += t_steps[t];
  }
}

static_assert(kMaxItemsPerWarp * kMelWarps >= kMaxItems,
              "every filter group must find a place in a contraction warp's list");

}  // namespace

// mel stages of n_fft <= 1024 run their contraction on tcgen05 unless B200A_TC=0 asks for the mma.sync path
static bool tc_enabled(const b200a_frontend_desc& d) {
  static const bool enabled = [] {
    const char* e = std::getenv("B200A_TC");
    return !(e && e[0] == '0');
  }();
  return enabled && d.n_fft <= 1024 && d.onesided && d.n_mels > 0 && d.n_mels <= kTcMaxN;
}

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
  const int G = d->n_fft == 2048 ? 32 : d->n_fft / 32;  // 2048 runs on the 1024-point complex core
  prepare_tw2d_kernel<<<(32 * G + 255) / 256, 256, 0, stream>>>(reinterpret_cast<float2*>(base + e.tw2d), G);
  if (d->n_fft == 2048) prepare_tw_eo_kernel<<<3, 256, 0, stream>>>(reinterpret_cast<float2*>(base + e.tw_eo));
  if (d->n_mels > 0 && mel_tiles(d->n_mels) <= kMaxItems) {
    prepare_mma_kernel<<<1, 256, 0, stream>>>(reinterpret_cast<const float*>(base + l.fb),
                                              reinterpret_cast<const int2*>(base + l.bands), d->n_fft / 2 + 1, d->n_mels,
                                              mel_tiles(d->n_mels), reinterpret_cast<MelPlan*>(base + e.plan),
                                              reinterpret_cast<float4*>(base + e.frags));
  }
  if (tc_enabled(*d)) {
    prepare_tc_kernel<<<1, 256, 0, stream>>>(reinterpret_cast<const float*>(base + l.fb), d->n_fft / 2 + 1, d->n_mels,
                                             d->n_fft, tc_b_budget(d->n_fft), d->n_fft == 1024 ? 32 : 0,
                                             reinterpret_cast<TcPlan*>(base + e.tc_plan), base + e.tc_b);
  }
  return launch_status();
}

static int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      return -1;
    cached = n > 0 ? n : 148;
  }
  return cached;
}

// persistent: one resident CTA per SM, units dealt round-robin (every CTA gets the same count +-1)
static int64_t persistent_grid(const Pow2Params& p, int warps = kWarps) {
  const int sms = num_sms();
  if (sms < 0) return -1;
  const int64_t iters = (p.total_units + warps - 1) / warps;
  const int64_t grid = iters < sms ? iters : sms;
  return grid < 1 ? 1 : grid;
}

template <int POWER_MODE, int G, int HG>
static int launch_power(const Pow2Params& p, cudaStream_t stream) {
  using Ge = Geo<G>;
  // the transform is latency bound: as many warps as shared memory (tile + staging buffer each) and the
  // register file (168 registers at 12 warps, no spills) allow
  constexpr int NW = 16;
  constexpr bool kShare = true;  // the staged input lives in the warp's transpose tile
  const size_t smem = sizeof(float2) * (32 * 32 + NW * Ge::kTileF2) + sizeof(uint64_t) * NW;
  auto kern = stft_pow2_power_kernel<POWER_MODE, G, HG, NW, kShare, false>;
  if constexpr (POWER_MODE != kComplexOut)
    if (p.kaldi) kern = stft_pow2_power_kernel<POWER_MODE, G, -1, NW, kShare, true>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
    return B200A_ECUDA;
  const int64_t grid = persistent_grid(p, NW);
  if (grid < 0) return B200A_ECUDA;
  kern<<<(unsigned)grid, NW * 32, smem, stream>>>(p);
  return launch_status();
}

template <int POWER_MODE, int G, int HG>
static int launch_mel(const Pow2Params& p, cudaStream_t stream) {
  using Ge = Geo<G>;
  static_assert(sizeof(MelPlan) % 16 == 0, "fragment array must stay 16-byte aligned");
  static_assert(Ge::kSlots <= kMaxSlots, "slot tables");
  const size_t smem = sizeof(float2) * (32 * 32 + kWarps * Ge::kTileF2) + sizeof(float) * 2 * Ge::kSlots * Ge::kPitch +
                      sizeof(int64_t) * 4 * Ge::kSlots + sizeof(uint64_t) * (kWarps + 4) + sizeof(MelPlan) +
                      sizeof(float4) * 32 * kFragSmemSteps;
  if (smem > 227 * 1024) return B200A_EUNSUPPORTED;
  static_assert(TcGeo<G>::kFixed + TcGeo<G>::kBBudget <= 227 * 1024 && TcGeo<G>::kBBudget >= 24 * 1024, "tcgen05 layout");
  constexpr int kThreads = TcGeo<G>::kThreads;
  auto kern = p.kaldi ? stft_pow2_mel_kernel<POWER_MODE, G, -1, true> : stft_pow2_mel_kernel<POWER_MODE, G, HG, false>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
    return B200A_ECUDA;
  const int64_t grid = persistent_grid(p, p.tc != nullptr ? TcGeo<G>::NW : kWarps);
  if (grid < 0) return B200A_ECUDA;
  kern<<<(unsigned)grid, kThreads, p.tc != nullptr ? (size_t)227 * 1024 : smem, stream>>>(p);
  return launch_status();
}

template <int POWER_MODE, int G>
static int launch_g(const Pow2Params& p, bool mel, cudaStream_t stream) {
  if constexpr (POWER_MODE == kComplexOut) {  // complex spectra: only the Spectrogram kernel
    if constexpr (G == 32)
      if (p.bulk_ok && p.hop == 256) return launch_power<POWER_MODE, 32, 8>(p, stream);
    return launch_power<POWER_MODE, G, -1>(p, stream);
  } else {
    if constexpr (G == 32) {
      if (p.bulk_ok && p.hop == 256 && !p.kaldi)  // frame b = frame a shifted by 8 lane-rows: shared register loads
        return mel ? launch_mel<POWER_MODE, 32, 8>(p, stream) : launch_power<POWER_MODE, 32, 8>(p, stream);
    }
    return mel ? launch_mel<POWER_MODE, G, -1>(p, stream) : launch_power<POWER_MODE, G, -1>(p, stream);
  }
}

template <int POWER_MODE>
static int launch_any(Pow2Params& p, int n_fft, bool mel, cudaStream_t stream) {
  if (n_fft == 1024) return launch_g<POWER_MODE, 32>(p, mel, stream);
  if (n_fft == 512) return launch_g<POWER_MODE, 16>(p, mel, stream);
  return launch_g<POWER_MODE, 8>(p, mel, stream);
}

template <int POWER_MODE>
static int launch_eo(const Pow2Params& p, const float2* tw_eo, bool mel, cudaStream_t stream) {
  const int64_t grid = persistent_grid(p);
  if (grid < 0) return B200A_ECUDA;
  const size_t tables = sizeof(float2) * (1024 + 1024 + 17 * 32 + kWarps * 32 * 33);
  if (!mel) {
    auto kern = stft2048_power_kernel<POWER_MODE>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
      return B200A_ECUDA;
    kern<<<(unsigned)grid, kWarps * 32, tables + sizeof(uint64_t) * kWarps, stream>>>(p, tw_eo);
    return launch_status();
  }
  const size_t smem = tables + sizeof(float) * kEoSlots * kEoPitch + sizeof(int64_t) * 2 * kEoSlots +
                      sizeof(uint64_t) * (kWarps + 2) + sizeof(MelPlan) + sizeof(float4) * 32 * kEoFragSteps;
  if (smem > 227 * 1024) return B200A_EUNSUPPORTED;
  auto kern = stft2048_mel_kernel<POWER_MODE>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
    return B200A_ECUDA;
  kern<<<(unsigned)grid, (kWarps + kMelWarps) * 32, smem, stream>>>(p, tw_eo);
  return launch_status();
}

int frontend_run_pow2(const b200a_frontend_desc* d, const void* ws, int stage, const float* wave, int64_t rows,
                      int64_t length, int64_t row_stride, int64_t frames, float* out, float* group_max,
                      int64_t rows_per_group, cudaStream_t stream, const b200a_kaldi_desc* kd) {
  if (!pow2_applicable(*d)) return B200A_EUNSUPPORTED;
  if (stage == B200A_STAGE_COMPLEX && (d->n_fft > 1024 || kd != nullptr)) return B200A_EUNSUPPORTED;
  // Kaldi features with a 256 / 512 / 1024-point FFT; every other size takes the generic kernel
  if (kd != nullptr && d->n_fft > 1024) return B200A_EUNSUPPORTED;
  if (stage >= B200A_STAGE_MEL && mel_tiles(d->n_mels) > kMaxItems) return B200A_EUNSUPPORTED;  // > 512 filters
  const WsLayout l = ws_layout(*d);
  const Pow2Extra e = pow2_layout(*d, l.total);
  const unsigned char* base = static_cast<const unsigned char*>(ws);
  const bool eo = d->n_fft == 2048;
  const int G = eo ? 32 : d->n_fft / 32;
  const int frames_per_unit = 2 * (32 / G);
  Pow2Params p{};
  p.wave = wave;
  p.length = length;
  p.row_stride = row_stride;
  p.frames = frames;
  p.units_per_row = (frames + frames_per_unit - 1) / frames_per_unit;
  p.total_units = rows * p.units_per_row;
  p.out = out;
  p.group_max = group_max;
  p.rows_per_group = rows_per_group > 0 ? rows_per_group : 1;
  p.window = reinterpret_cast<const float*>(base + l.window);
  p.tw2d = reinterpret_cast<const float2*>(base + e.tw2d);
  p.plan = reinterpret_cast<const MelPlan*>(base + e.plan);
  p.frags = reinterpret_cast<const float4*>(base + e.frags);
  p.hdr = reinterpret_cast<const WsHeader*>(base + l.header);
  p.tc = tc_enabled(*d) ? reinterpret_cast<const TcPlan*>(base + e.tc_plan) : nullptr;
  p.tc_b = base + e.tc_b;
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
  // bulk staging needs 16-byte aligned sources and sizes (every unit starts at a multiple of
  // frames_per_unit*hop, minus half + pad) and the unit's span must fit the staging buffer
  const int half = d->center ? d->n_fft / 2 : 0;
  const int stage_floats = 2 * (32 / G) * (32 * (G + 1) + (G == 8 ? 8 : 0));
  p.bulk_ok = d->hop % 4 == 0 && (half + d->pad) % 4 == 0 && row_stride % 4 == 0 &&
              (reinterpret_cast<uintptr_t>(wave) & 15) == 0 &&
              d->n_fft + (frames_per_unit - 1) * (int64_t)d->hop <= stage_floats;
  p.stage_ok = d->n_fft + (frames_per_unit - 1) * (int64_t)d->hop <= stage_floats &&  // edge units gather into it
               length + 2 * (int64_t)d->pad + d->n_fft < (int64_t)1 << 31;           // with 32-bit indices
  p.out_width = stage >= B200A_STAGE_MEL ? d->n_mels : d->n_fft / 2 + 1;
  p.out_col0 = 0;
  p.k_energy_col = -1;
  int lead = half;
  if (kd != nullptr) {
    p.kaldi = 1;
    p.k_off = kd->snip_edges ? 0 : kd->window_size / 2 - kd->window_shift / 2;
    p.k_win = kd->window_size;
    p.k_dc = kd->remove_dc_offset;
    p.k_preemph = kd->preemphasis;
    p.k_energy_mode = kd->energy_col >= 0 ? kd->energy_mode : 0;
    p.k_energy_floor = kd->energy_floor;
    p.k_energy_col = kd->energy_col;
    p.k_log = kd->use_log;
    p.out_width = kd->out_width;
    p.out_col0 = kd->out_col0;
    p.pad_mode = kPadSymmetric;  // only reached when snip_edges == 0 (frames never leave the signal otherwise)
    lead = p.k_off;
    if (!p.stage_ok) return B200A_EUNSUPPORTED;  // the conditioning reads the staged span
    p.bulk_ok = d->hop % 4 == 0 && lead % 4 == 0 && row_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(wave) & 15) == 0;
  }
  p.out_vec = (p.out_width % 4 == 0 && p.out_col0 % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)   ? 4
              : (p.out_width % 2 == 0 && p.out_col0 % 2 == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0) ? 2
                                                                                                              : 1;
  const bool mel = stage >= B200A_STAGE_MEL;
  if (eo) {
    p.bulk_ok = d->hop % 4 == 0 && (half + d->pad) % 4 == 0 && row_stride % 4 == 0 &&
                (reinterpret_cast<uintptr_t>(wave) & 15) == 0;
    const float2* tw_eo = reinterpret_cast<const float2*>(base + e.tw_eo);
    return d->power == 2.f ? launch_eo<2>(p, tw_eo, mel, stream) : launch_eo<0>(p, tw_eo, mel, stream);
  }
  if (stage == B200A_STAGE_COMPLEX) return launch_any<kComplexOut>(p, d->n_fft, false, stream);
  return d->power == 2.f ? launch_any<2>(p, d->n_fft, mel, stream) : launch_any<0>(p, d->n_fft, mel, stream);
}

// First half of b200a_istft_run for n_fft = 256 / 512 / 1024: windowed time frames into frame_buf.
int istft_frames_pow2(const b200a_frontend_desc* d, const void* ws, const float* spec, int64_t rows, int64_t frames,
                      int64_t stride_row, int64_t stride_bin, int64_t stride_frame, float* frame_buf, cudaStream_t stream) {
  if (!pow2_applicable(*d) || d->n_fft > 1024) return B200A_EUNSUPPORTED;
  const WsLayout l = ws_layout(*d);
  const Pow2Extra e = pow2_layout(*d, l.total);
  const unsigned char* base = static_cast<const unsigned char*>(ws);
  const int G = d->n_fft / 32;
  const int frames_per_unit = 2 * (32 / G);
  IstftPow2Params p{};
  p.spec = reinterpret_cast<const float2*>(spec);
  p.stride_row = stride_row;
  p.stride_bin = stride_bin;
  p.stride_frame = stride_frame;
  p.frames = frames;
  p.units_per_row = (frames + frames_per_unit - 1) / frames_per_unit;
  p.total_units = rows * p.units_per_row;
  p.frame_buf = frame_buf;
  p.window = reinterpret_cast<const float*>(base + l.window);
  p.tw2d = reinterpret_cast<const float2*>(base + e.tw2d);
  p.hdr = reinterpret_cast<const WsHeader*>(base + l.header);
  const int sms = num_sms();
  if (sms < 0) return B200A_ECUDA;
  const int64_t iters = (p.total_units + kIsWarps - 1) / kIsWarps;
  const unsigned grid = (unsigned)(iters < sms ? (iters < 1 ? 1 : iters) : sms);
  auto launch = [&](auto kern, size_t smem) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return (int)B200A_ECUDA;
    kern<<<grid, kIsWarps * 32, smem, stream>>>(p);
    return launch_status();
  };
  if (G == 32) return launch(istft_pow2_kernel<32>, sizeof(float2) * (32 * 32 + kIsWarps * Geo<32>::kTileF2));
  if (G == 16) return launch(istft_pow2_kernel<16>, sizeof(float2) * (32 * 16 + kIsWarps * Geo<16>::kTileF2));
  return launch(istft_pow2_kernel<8>, sizeof(float2) * (32 * 8 + kIsWarps * Geo<8>::kTileF2));
}

}  // namespace b200a
