// Pippenger bucket-method kernels for one B200 (sm_100a), generic over the curve.
//
// What this replaces on the reference's hot path:
//   a10  signed Booth window extraction      reference constantine/math/arithmetic/bigints.nim:360-379, 806-861
//   a7   bucket accumulation + reduction     reference constantine/math/elliptic/ec_multi_scalar_mul.nim:177-235
//   a5/a6 one-task-per-window decomposition  reference constantine/math/elliptic/ec_multi_scalar_mul_parallel.nim:148-208, 316-431
//   a2   Fr (Montgomery) -> canonical scalars reference constantine/math/elliptic/ec_multi_scalar_mul_parallel.nim:611-628
//
// B200 mapping (DESIGN.md has the full data-flow):
//   1. k_digits        one thread per scalar: all W signed digits -> (key = window*B + bucket, val = point index | sign<<31)
//   2. radix sort       (key,val) pairs by key; zero digits carry the key W*B and sort to the tail
//   3. k_accumulate    the sorted list is cut into fixed slices of K entries, one thread per slice; a thread sums the
//                      points of each run of equal keys in registers (XYZZ mixed adds). Runs that start inside the slice
//                      own their bucket and store it; a run that continues from the previous slice is emitted as a
//                      "partial" -- perfectly load-balanced whatever the digit distribution is.
//   4. k_fixup         the partials form a (much shorter) sorted list themselves: same slicing, XYZZ+XYZZ adds, added
//                      into the owning bucket; repeated until a level has a single slice.
//   5. k_bucket_reduce per window, T threads x L consecutive buckets: running sum + (t*L)*run offset, then k_sum_groups
//                      tree-sums the T partial results -> one point per window.
//   6. host tail       Horner over the <= W window sums (host_field.hpp).
#pragma once
#include "ec.cuh"

namespace b200 {

// ------------------------------------------------------------------------------------------- digits
struct DigitPlan {
  int bits;        // declared scalar width b (254 / 255)
  int c;           // window size
  int num_full;    // b / c
  int num_windows; // num_full + 1   (reference parallel.nim:157-158: the recoding must see one more window)
  int excess;      // b % c
  int top;         // b - excess
  uint32_t buckets_per_window;  // 2^(c-1)
  uint32_t total_buckets;       // num_windows * 2^(c-1); also the key of "no bucket" (zero digit)
  int win_begin, win_end;       // windows handled by this device (multi-GPU window sharding), [begin, end)
};

__device__ __forceinline__ uint32_t scalar_window(const uint32_t* k, int bit_index, int nbits) {
  // bits [bit_index, bit_index + nbits) of the 256-bit little-endian scalar; bits >= 256 read as zero
  // (reference bigints.nim:360-379 getWindowAt: the next limb is only read if it exists). nbits <= 24.
  int word = bit_index >> 5, pos = bit_index & 31;
  uint64_t lo = (word < 8) ? k[word] : 0u;
  uint64_t hi = (word + 1 < 8) ? k[word + 1] : 0u;
  uint64_t v = (lo | (hi << 32)) >> pos;
  return (uint32_t)v & ((1u << nbits) - 1u);
}

// Booth recoding of a (bitsize+1)-bit digit (reference bigints.nim:806-832 signedWindowEncoding)
__device__ __forceinline__ void signed_encode(uint32_t digit, int bitsize, uint32_t& val, uint32_t& neg) {
  neg = digit >> bitsize;
  uint32_t mask = 0u - neg;
  uint32_t enc = (digit + 1u) >> 1;
  val = ((enc + mask) ^ mask) & ((1u << bitsize) - 1u);
}

// digit of window w under plan p (reference bigints.nim:834-861; window kinds per parallel.nim:171-196)
__device__ __forceinline__ void window_digit(const uint32_t* k, const DigitPlan& p, int w, uint32_t& val, uint32_t& neg) {
  if (w == p.num_full) {
    if (p.top == 0) signed_encode(scalar_window(k, 0, p.c) << 1, p.c, val, neg);
    else if (p.excess == 0) signed_encode(scalar_window(k, p.top - 1, p.c + 1), p.c, val, neg);
    else signed_encode(scalar_window(k, p.top - 1, p.excess + 1), p.excess + 1, val, neg);
  } else if (w == 0) {
    signed_encode(scalar_window(k, 0, p.c) << 1, p.c, val, neg);
  } else {
    signed_encode(scalar_window(k, w * p.c - 1, p.c + 1), p.c, val, neg);
  }
}

// One thread per scalar. FR_MONT: scalars arrive as Fr Montgomery residues and are first converted to canonical
// integers by one Montgomery reduction (multiplication by the integer 1), like the reference's fromField pass.
// key = m * msm_key_stride + w_local * key_stride + bucket, val = (w_local * val_stride + point index) | sign << 31.
//   single MSM:  m = 0; key_stride = B; val_stride = 0 (every window references point i)
//   table mode:  key_stride = 0 (all windows share one bucket set because the points come from a table of precomputed
//                window multiples), val_stride = table row length
//   batch:       scalar i belongs to MSM m = i / per; msm_key_stride = bucket sets' size per MSM; the point index is
//                i (every MSM has its own bases) or i - m * per (`shared`: all MSMs use the same bases)
template <class FrParams, bool FR_MONT>
__global__ void __launch_bounds__(256) k_digits(const uint32_t* __restrict__ scalars, uint32_t n, DigitPlan plan,
                                                uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                                uint32_t key_stride, uint32_t no_key, uint32_t val_stride,
                                                uint32_t per, uint32_t msm_key_stride, uint32_t shared) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k[8];
  const uint4* src = reinterpret_cast<const uint4*>(scalars) + (size_t)i * 2;
  uint4 a = __ldg(src), b = __ldg(src + 1);
  k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w; k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
  if (FR_MONT) {
    static_assert(FrParams::N == 8, "256-bit scalar fields");
    uint32_t one[8] = {1, 0, 0, 0, 0, 0, 0, 0};
    fe_mul<FrParams>(k, k, one);
  }
  const uint32_t m = i / per;
  const uint32_t ref = shared ? i - m * per : i;
  const uint32_t key_base = m * msm_key_stride;
  for (int w = plan.win_begin; w < plan.win_end; w++) {
    uint32_t val, neg;
    window_digit(k, plan, w, val, neg);
    const uint32_t wl = (uint32_t)(w - plan.win_begin);
    size_t slot = (size_t)wl * n + i;
    keys[slot] = val ? key_base + wl * key_stride + (val - 1u) : no_key;
    vals[slot] = (wl * val_stride + ref) | (neg << 31);
  }
}

// ------------------------------------------------------------------------------------------- bucket accumulation
constexpr uint32_t KEY_NONE = 0xFFFFFFFFu;

// Slice kernel over the sorted (key, point-ref) list.  `no_key` = first key value that means "no bucket".
//
// A run of equal keys that starts inside a slice is owned by that slice's thread, which stores the bucket. A run that
// continues from the previous slice ("head") cannot be stored by this thread. Heads are handed to the previous lane of
// the same warp through shared memory and added into that lane's last run before it is stored (one extra XYZZ add per
// warp), so only heads whose predecessor lives in another warp -- or is itself a single-run continuation slice --
// are spilled to the global partial list that k_fixup resolves.
// `into` != 0: the buckets already hold sums (earlier input chunks of the same MSM, msm_engine.cuh) and this launch adds to them.
#ifndef B200_ACC_MIN_BLOCKS
#define B200_ACC_MIN_BLOCKS 2
#endif
#ifndef B200_ACC_THREADS
#define B200_ACC_THREADS 128
#endif
// Slice length that makes ceil(n / K) slices fill m = ceil(n / (fit_threads * k_max)) resident waves exactly: K = ceil(n / (m * fit_threads)),
// at least B200_ACC_MIN_SLICE (tiny inputs: one partial wave). Then ceil(n / K) <= m * fit_threads for every n, and m is monotone
// in n, so a grid of m(n_upper_bound) * fit_threads threads covers any actual n <= n_upper_bound.
#ifndef B200_ACC_MIN_SLICE
#define B200_ACC_MIN_SLICE 16
#endif
__host__ __device__ inline size_t accumulate_waves(size_t n, uint32_t fit_threads, int k_max) {
  const size_t per_wave = (size_t)fit_threads * (size_t)k_max;
  const size_t m = (n + per_wave - 1) / per_wave;
  return m < 1 ? 1 : m;
}
__host__ __device__ inline int accumulate_slice_len(size_t n, uint32_t fit_threads, int k_max) {
  const size_t per = accumulate_waves(n, fit_threads, k_max) * (size_t)fit_threads;
  const size_t k = (n + per - 1) / per;
  return k < B200_ACC_MIN_SLICE ? B200_ACC_MIN_SLICE : (int)k;
}

template <class T>
__global__ void __launch_bounds__(B200_ACC_THREADS, (T::WORDS <= 12) ? B200_ACC_MIN_BLOCKS : 1)
k_accumulate(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, const unsigned long long* __restrict__ bounds,
             int w0, int w1, uint32_t no_key, const uint32_t* __restrict__ points, uint32_t* buckets, uint32_t* part_pts,
             uint32_t* part_keys, size_t max_slices, int K, int into, uint32_t fit_threads) {
  // entries of windows [w0, w1) occupy sorted positions [begin, total); slices are counted from `begin`
  const size_t begin = (size_t)bounds[w0], total = (size_t)bounds[w1];
  // `fit_threads` != 0 (the number of threads of this kernel the device keeps resident): K is the upper limit and the slice length is
  // chosen HERE, from the actual entry count, so that the slices fill a whole number m of resident waves -- every thread of a wave
  // walks the same number of entries, so a last wave that is 10% full costs as much as a full one (accumulate_slice_len, the host
  // twin that sizes the grid, has the arithmetic)
  if (fit_threads) K = accumulate_slice_len(total - begin, fit_threads, K);
  const size_t num_slices = (total - begin + (size_t)K - 1) / (size_t)K;
  __shared__ uint32_t head_smem[B200_ACC_THREADS * 4 * T::WORDS];
  const unsigned lane = threadIdx.x & 31u;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = t < num_slices;
  size_t base = begin + t * (size_t)K;
  uint32_t prev_key = (live && t > 0) ? keys[base - 1] : KEY_NONE;
  uint32_t cur_key = KEY_NONE;
  bool head_valid = false;     // first run continues the previous slice (its sum is parked in head_smem[threadIdx.x])
  bool first_run = true;
  Xyzz<T> acc = Xyzz<T>::inf();
  if (live) {
    // software pipeline: the (key, ref, point) of entry j+1 is fetched before the point addition of entry j is issued,
    // so the dependent gather keys -> vals -> points overlaps ~10 field multiplications instead of stalling the warp.
    // Single-field coordinates hold the next point in registers. For Fp2 (a point is 48 words, the accumulator 96) the live set of
    // a mixed addition exceeds the 255-register file either way (ptxas spills ~0.6 KB to local memory, L1-resident), so the next
    // point is only pulled into L2 and loaded at the top of its own iteration (-DB200_FP2_REG_PREFETCH=1 restores the register form).
#ifndef B200_FP2_REG_PREFETCH
#define B200_FP2_REG_PREFETCH 0
#endif
    constexpr bool REG_PREFETCH = T::WORDS <= 12 || B200_FP2_REG_PREFETCH;
    uint32_t key_n = (base < total) ? keys[base] : no_key;
    uint32_t v_n = 0;
    Aff<T> p_n;
    if constexpr (REG_PREFETCH) { p_n.x = T::zero(); p_n.y = T::zero(); }
    if (key_n < no_key) {
      v_n = vals[base];
      if constexpr (REG_PREFETCH) p_n = load_affine<T>(points, v_n & 0x7FFFFFFFu);
    }
#pragma unroll 1
    for (int j = 0; j < K; j++) {
      const uint32_t key = key_n;
      if (key >= no_key) break;  // zero digits are sorted to the tail (or end of list): nothing left in this slice
      const uint32_t v = v_n;
      Aff<T> p;
      if constexpr (REG_PREFETCH) p = p_n;
      const size_t nidx = base + j + 1;
      key_n = (j + 1 < K && nidx < total) ? keys[nidx] : no_key;
      if (key_n < no_key) {
        v_n = vals[nidx];
        if constexpr (REG_PREFETCH) p_n = load_affine<T>(points, v_n & 0x7FFFFFFFu);
        else {
          const uint32_t* nx = points + (size_t)(v_n & 0x7FFFFFFFu) * (2 * T::WORDS);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 32));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + 2 * T::WORDS - 1));
        }
      }
      if constexpr (!REG_PREFETCH) p = load_affine<T>(points, v & 0x7FFFFFFFu);
      if (!p.is_inf()) p.y.cneg((v >> 31) != 0);
      if (key != cur_key) {
        if (cur_key != KEY_NONE) {
          if (first_run && cur_key == prev_key) { store_xyzz(head_smem, threadIdx.x, acc); head_valid = true; }
          else store_xyzz(buckets, (size_t)cur_key, acc);
          first_run = false;
        }
        // a run this thread owns starts from the bucket's current value when earlier input chunks already accumulated into
        // the buckets (`into`), from infinity otherwise; a run continuing the previous slice ("head") always starts from infinity
        const bool owned = !(first_run && key == prev_key);
        cur_key = key;
        if (into && owned) acc = load_xyzz<T>(buckets, (size_t)key);
        else acc = Xyzz<T>::inf();
      }
      xyzz_madd(acc, p);   // onto infinity: a copy
    }
  }
  // the last run of the slice is still in `acc`
  const bool has_run = cur_key != KEY_NONE;
  const bool last_is_head = has_run && first_run && cur_key == prev_key;  // single-run slice continuing its predecessor
  if (last_is_head) { store_xyzz(head_smem, threadIdx.x, acc); head_valid = true; }
  const bool absorber = has_run && !last_is_head;                        // owns the bucket of its last run
  __syncwarp();
  // A lane that owns the bucket of its last run takes over the head of the next lane (same key by construction);
  // heads whose predecessor lane cannot do that (lane 0, or a predecessor that is itself a single-run continuation
  // slice) go to the global partial list. Longer chains are cheaper to resolve there (warp butterflies in k_fixup)
  // than by serial additions here.
  const bool prev_absorbs = __shfl_up_sync(0xFFFFFFFFu, absorber ? 1 : 0, 1) != 0 && lane > 0;
  const bool next_has_head = __shfl_down_sync(0xFFFFFFFFu, head_valid ? 1 : 0, 1) != 0 && lane < 31;
  const bool head_absorbed = head_valid && prev_absorbs;
  if (absorber && next_has_head) {
    Xyzz<T> h = load_xyzz<T>(head_smem, threadIdx.x + 1);
    xyzz_add(acc, h);
  }
  if (absorber) store_xyzz(buckets, (size_t)cur_key, acc);
  if (t < max_slices) {   // every slot of the group's partial list is written (slots past the last slice are holes)
    if (live && head_valid && !head_absorbed) {
      Xyzz<T> h = load_xyzz<T>(head_smem, threadIdx.x);
      store_xyzz(part_pts, t, h);
      part_keys[t] = prev_key;
    } else {
      part_keys[t] = KEY_NONE;
    }
  }
}

// bounds[w] = first sorted position whose key is >= w * B  (w = 0..num_windows; keys of zero digits sort last)
static __global__ void k_window_bounds(const uint32_t* __restrict__ keys, size_t total, uint32_t B, int num_windows, unsigned long long* bounds) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w > num_windows) return;
  const uint64_t target = (uint64_t)w * B;
  size_t lo = 0, hi = total;
  while (lo < hi) {
    size_t mid = lo + (hi - lo) / 2;
    if ((uint64_t)keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  bounds[w] = (unsigned long long)lo;
}

// Fix-up level: input = list of (key, XYZZ) partials in slice order (KEY_NONE entries are holes); equal keys are
// contiguous. One thread per entry, one warp per chunk of 32 entries. Runs of equal keys inside a chunk are summed with a
// segmented butterfly over shuffles (5 dependent additions for a chunk-long run instead of 31; steps in which no lane has
// a partner are skipped, so sparse lists cost nothing). The sum of a run that starts at the chunk boundary and continues
// the previous chunk is forwarded to the next level (slot = chunk index); every other run is added into its bucket --
// exactly one thread per key and level does so.
template <class T>
__global__ void __launch_bounds__(128) k_fixup(const uint32_t* __restrict__ in_keys, const uint32_t* in_pts, size_t count,
                                               uint32_t* buckets, uint32_t* out_pts, uint32_t* out_keys) {
  const unsigned lane = threadIdx.x & 31u;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool in_range = i < count;
  const uint32_t key = in_range ? in_keys[i] : KEY_NONE;
  const uint32_t prev = (in_range && i > 0) ? in_keys[i - 1] : KEY_NONE;
  const bool valid = key != KEY_NONE;
  const bool continues = valid && prev == key;                 // same run as the previous entry
  Xyzz<T> acc = valid ? load_xyzz<T>(in_pts, i) : Xyzz<T>::inf();
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t other_key = __shfl_down_sync(0xFFFFFFFFu, key, d);
    const bool take = valid && (lane + d < 32u) && other_key == key;
    if (__any_sync(0xFFFFFFFFu, take)) {
      Xyzz<T> other;
#pragma unroll
      for (int k = 0; k < T::WORDS; k++) {
        other.x.set_word(k, __shfl_down_sync(0xFFFFFFFFu, acc.x.word(k), d));
        other.y.set_word(k, __shfl_down_sync(0xFFFFFFFFu, acc.y.word(k), d));
        other.zz.set_word(k, __shfl_down_sync(0xFFFFFFFFu, acc.zz.word(k), d));
        other.zzz.set_word(k, __shfl_down_sync(0xFFFFFFFFu, acc.zzz.word(k), d));
      }
      if (take) xyzz_add_ni(acc, other);
    }
  }
  if (!in_range) return;
  if (lane == 0 && !continues) out_keys[i / 32] = KEY_NONE;    // this chunk forwards nothing
  if (!valid) return;
  if (continues && lane != 0) return;                           // interior of a run: its first lane holds the sum
  if (continues) {                                              // lane 0 continuing the previous chunk's run
    store_xyzz(out_pts, i / 32, acc);
    out_keys[i / 32] = key;
  } else {
    Xyzz<T> b = load_xyzz<T>(buckets, (size_t)key);
    xyzz_add_ni(b, acc);
    store_xyzz(buckets, (size_t)key, b);
  }
}

// ------------------------------------------------------------------------------------------- bucket reduction
// Window sum  S = sum_{j=0}^{B-1} (j+1) * bucket[j]   (reference ec_multi_scalar_mul.nim:186-197 bucketReduce computes
// the same value with one serial running sum).  Here B buckets are cut into chunks of L consecutive buckets:
//   chunk t:  run = sum_i b[tL+i],  acc = sum_i (i+1) b[tL+i]  (running sum, 2 adds per bucket)
//   S = sum_t ( acc_t + (t*L) * run_t )
// (t*L)*run_t is a double-and-add over `nbits` bits (uniform loop bound). One point per chunk comes out;
// k_row_sum_warp adds them up with warp butterflies.
template <class T, bool INL>
B200_DEV void padd(Xyzz<T>& a, const Xyzz<T>& b) {
  if constexpr (INL) xyzz_add(a, b); else xyzz_add_ni(a, b);
}
template <class T, bool INL>
B200_DEV void pdbl(Xyzz<T>& a) {
  if constexpr (INL) a = xyzz_dbl(a); else xyzz_dbl_ni(a);
}

template <class T, bool INL>
__global__ void __launch_bounds__(64) k_bucket_reduce(const uint32_t* buckets, uint32_t buckets_per_window, uint32_t L,
                                                      uint32_t chunks_per_window, uint32_t num_windows, uint32_t* out_acc,
                                                      uint32_t* out_run) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (size_t)chunks_per_window * num_windows) return;
  uint32_t w = (uint32_t)(g / chunks_per_window), t = (uint32_t)(g % chunks_per_window);
  size_t first = (size_t)w * buckets_per_window + (size_t)t * L;
  uint32_t cnt = min(L, buckets_per_window - t * L);
  Xyzz<T> run = Xyzz<T>::inf(), acc = Xyzz<T>::inf();
#pragma unroll 1
  for (int i = (int)cnt - 1; i >= 0; i--) {
    Xyzz<T> b = load_xyzz<T>(buckets, first + i);
    padd<T, INL>(run, b);
    padd<T, INL>(acc, run);
  }
  store_xyzz(out_acc, g, acc);
  store_xyzz(out_run, g, run);
}

// acc[g] += (t*L) * run[g]   -- 2-bit fixed-window scalar multiplication over `nbits` bits (uniform trip count):
// table {run, 2 run, 3 run}, then per digit two doublings and one table addition (skipped by lanes whose digit is 0).
template <class T, bool INL>
__global__ void __launch_bounds__(64) k_chunk_offset(uint32_t* acc_io, const uint32_t* run_in, uint32_t L, uint32_t chunks_per_window,
                                                     uint32_t num_windows, int nbits) {
  size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (size_t)chunks_per_window * num_windows) return;
  const uint32_t off = (uint32_t)(g % chunks_per_window) * L;
  if (off == 0) return;
  Xyzz<T> run = load_xyzz<T>(run_in, g);
  if (run.is_inf()) return;
  Xyzz<T> run2 = run;
  pdbl<T, INL>(run2);
  Xyzz<T> run3 = run2;
  padd<T, INL>(run3, run);
  Xyzz<T> m = Xyzz<T>::inf();
  const int digits = (nbits + 1) / 2;
#pragma unroll 1
  for (int d = digits - 1; d >= 0; d--) {
    pdbl<T, INL>(m);
    pdbl<T, INL>(m);
    const uint32_t w = (off >> (2 * d)) & 3u;
    if (w) {
      Xyzz<T> sel = (w == 1u) ? run : ((w == 2u) ? run2 : run3);
      padd<T, INL>(m, sel);
    }
  }
  Xyzz<T> acc = load_xyzz<T>(acc_io, g);
  padd<T, false>(acc, m);
  store_xyzz(acc_io, g, acc);
}

// One warp per group of 32 consecutive entries of a row: out[row][g] = sum_{i<32} in[row][32 g + i] (butterfly over
// shuffles: 5 dependent additions instead of 31).
template <class T, bool INL>
__global__ void __launch_bounds__(128) k_row_sum_warp(const uint32_t* in, uint32_t row_len, uint32_t out_row_len, uint32_t num_rows,
                                                      uint32_t* out) {
  const unsigned lane = threadIdx.x & 31u;
  size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= (size_t)out_row_len * num_rows) return;
  uint32_t row = (uint32_t)(warp / out_row_len), o = (uint32_t)(warp % out_row_len);
  uint32_t src = o * 32u + lane;
  Xyzz<T> acc = (src < row_len) ? load_xyzz<T>(in, (size_t)row * row_len + src) : Xyzz<T>::inf();
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) {
    Xyzz<T> other;
#pragma unroll
    for (int k = 0; k < T::WORDS; k++) {
      other.x.set_word(k, __shfl_xor_sync(0xFFFFFFFFu, acc.x.word(k), d));
      other.y.set_word(k, __shfl_xor_sync(0xFFFFFFFFu, acc.y.word(k), d));
      other.zz.set_word(k, __shfl_xor_sync(0xFFFFFFFFu, acc.zz.word(k), d));
      other.zzz.set_word(k, __shfl_xor_sync(0xFFFFFFFFu, acc.zzz.word(k), d));
    }
    padd<T, INL>(acc, other);
  }
  if (lane == 0) store_xyzz(out, warp, acc);
}

// ------------------------------------------------------------------------------------------- bucket reduction, bit-plane form
// Same window sum  S = sum_j (j+1) * bucket[j]  (reference ec_multi_scalar_mul.nim:186-197), reorganised so that no thread
// walks a long chain of dependent point additions (the running-sum form above needs ~2L + log2(B) of them):
//   view the B = 2^(c-1) buckets of a window as a matrix of R = 2^rbits rows and C = 2^a columns, j = h*C + l, so that
//   j + 1 = h*C + (l + 1)  and   S = C * sum_h h*H_h + sum_l (l+1)*L_l   with the row sums H_h and the column sums L_l;
//   a sum  sum_i i*X_i  over 2^k points is  sum_b 2^b * (sum of the X_i whose index has bit b set): "bit planes".
//   k_rowcol_sums: every row sum and every column sum, `lanes` lanes per sum (strided serial part + xor butterfly);
//   k_plane_sums : one warp per (window, bit position q of the weight j+1), q = 0 .. c-2;
//   k_plane_combine: radix-16 digits D_g = sum_{k<4} 2^k P_{4g+k} of the c-1 bit-position sums.
// A window leaves the device as ceil((c-1)/4) partial points with  S = sum_g 16^g D_g; the host tail folds the 16^g into the
// Horner evaluation over the windows it runs anyway (one doubling per bit position, one addition per partial point).
// Work: 2 additions per bucket as before, but the longest dependent chain is C/lanes + R/lanes + 2 log2(lanes) + ~13
// additions instead of ~66, and every stage is a plain sum (any number of lanes per sum).
template <class T, bool INL>
B200_DEV void group_butterfly(Xyzz<T>& acc, int lanes) {
#pragma unroll 1
  for (int d = lanes >> 1; d >= 1; d >>= 1) {
    Xyzz<T> other;
#pragma unroll
    for (int k = 0; k < T::WORDS; k++) {
      other.x.set_word(k, __shfl_xor_sync(0xFFFFFFFFu, acc.x.word(k), d));
      other.y.set_word(k, __shfl_xor_sync(0xFFFFFFFFu, acc.y.word(k), d));
      other.zz.set_word(k, __shfl_xor_sync(0xFFFFFFFFu, acc.zz.word(k), d));
      other.zzz.set_word(k, __shfl_xor_sync(0xFFFFFFFFu, acc.zzz.word(k), d));
    }
    padd<T, INL>(acc, other);
  }
}

// Groups wider than a warp (64 or 128 lanes of a 128-thread block; groups are aligned, so warp `wid` belongs to the group that starts
// at warp wid - wid % wg): after the in-warp butterfly every warp parks its sum in shared memory and the first warp of the group adds
// the wg sums with a wg-lane butterfly -- log2(wg) more dependent additions instead of a longer serial part. Every thread of the
// block must call it (one __syncthreads). The total ends up in lane 0 of the group's first warp.
template <class T, bool INL>
B200_DEV void block_group_finish(Xyzz<T>& acc, int lanes, uint32_t* smem) {
  const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  const unsigned wg = (unsigned)lanes >> 5;
  if (lane == 0) store_xyzz(smem, wid, acc);
  __syncthreads();
  if (wid % wg == 0) {
    Xyzz<T> v = Xyzz<T>::inf();
    if (lane < wg) v = load_xyzz<T>(smem, wid + lane);
    group_butterfly<T, INL>(v, (int)wg);
    acc = v;
  }
}

// Blocks [0, row_blocks): out_rows[w*R + h] = sum_l bucket[w][h*C + l];  the others: out_cols[w*C + l] = sum_h bucket[w][h*C + l].
// `lanes` (power of two <= 128) lanes share one sum; a block is a whole number of groups and every lane reaches the shuffles.
template <class T, bool INL>
__global__ void __launch_bounds__(128) k_rowcol_sums(const uint32_t* buckets, uint32_t buckets_per_window, int a, uint32_t num_windows,
                                                     int lanes_r, int lanes_c, unsigned row_blocks, uint32_t* out_rows, uint32_t* out_cols) {
  const bool cols = blockIdx.x >= row_blocks;               // uniform per block
  const uint32_t C = 1u << a, R = buckets_per_window >> a;
  const uint32_t per_window = cols ? C : R;                 // sums per window
  const uint32_t len = cols ? R : C;                        // terms per sum
  const int lanes = cols ? lanes_c : lanes_r;
  const size_t g = (size_t)(blockIdx.x - (cols ? row_blocks : 0u)) * blockDim.x + threadIdx.x;
  const size_t sum_id = g / (unsigned)lanes;
  const uint32_t t = (uint32_t)(g % (unsigned)lanes);
  const bool live = sum_id < (size_t)per_window * num_windows;
  Xyzz<T> acc = Xyzz<T>::inf();
  if (live) {
    const uint32_t w = (uint32_t)(sum_id / per_window), i = (uint32_t)(sum_id % per_window);
    const size_t base = (size_t)w * buckets_per_window;
    // row sum: elements i*C + k (stride 1);  column sum: elements k*C + i (stride C)
    const size_t first = cols ? (size_t)i : (size_t)i * C;
    const size_t stride = cols ? (size_t)C : 1;
#pragma unroll 1
    for (uint32_t k = t; k < len; k += (uint32_t)lanes) {
      Xyzz<T> b = load_xyzz<T>(buckets, base + first + (size_t)k * stride);
      padd<T, INL>(acc, b);
    }
  }
  __shared__ uint32_t warp_sums[4 * 4 * T::WORDS];
  group_butterfly<T, INL>(acc, lanes < 32 ? lanes : 32);
  if (lanes > 32) block_group_finish<T, INL>(acc, lanes, warp_sums);   // uniform per block
  if (live && t == 0) store_xyzz(cols ? out_cols : out_rows, sum_id, acc);
}

// One warp per (window w, bit position q), q = 0 .. c-2:  P_q = sum of everything that carries weight bit q of j + 1 = h*C + (l+1):
//   q <  a: the column sums L_l with bit q of l + 1 set (l + 1 = C has no bit below a);
//   q >= a: the row sums H_h with bit q - a of h set, and for q == a also the last column L_{C-1} (its weight l + 1 = C = 2^a).
// Then S_w = sum_q 2^q P_q.  out[w * (c - 1) + q].
template <class T, bool INL>
__global__ void __launch_bounds__(128) k_plane_sums(const uint32_t* row_sums, const uint32_t* col_sums, int a, int rbits,
                                                    uint32_t num_windows, int lanes, uint32_t* out) {
  // `lanes` = 32 (one warp per sum) or 128 (one block per sum: 4 more dependent additions saved on sums of 256 terms)
  const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t sum_id = g / (unsigned)lanes;
  const uint32_t t = (uint32_t)(g % (unsigned)lanes);
  const uint32_t planes = (uint32_t)(rbits + a);
  const bool live = sum_id < (size_t)planes * num_windows;
  Xyzz<T> acc = Xyzz<T>::inf();
  if (live) {
    const uint32_t w = (uint32_t)(sum_id / planes), q = (uint32_t)(sum_id % planes);
    const bool is_h = q >= (uint32_t)a;
    const uint32_t bit = is_h ? q - (uint32_t)a : q;
    const uint32_t len = is_h ? (1u << rbits) : (1u << a);
    const uint32_t* src = is_h ? row_sums : col_sums;
    const size_t base = (size_t)w * len;
#pragma unroll 1
    for (uint32_t i = t; i < len; i += (uint32_t)lanes) {
      const uint32_t weight = is_h ? i : i + 1u;
      if ((weight >> bit) & 1u) {
        Xyzz<T> b = load_xyzz<T>(src, base + i);
        padd<T, INL>(acc, b);
      }
    }
    if (q == (uint32_t)a && t == 0) {
      Xyzz<T> b = load_xyzz<T>(col_sums, (size_t)w * (1u << a) + ((1u << a) - 1u));
      padd<T, INL>(acc, b);
    }
  }
  __shared__ uint32_t warp_sums[4 * 4 * T::WORDS];
  group_butterfly<T, INL>(acc, 32);
  if (lanes > 32) block_group_finish<T, INL>(acc, lanes, warp_sums);
  if (live && t == 0) store_xyzz(out, sum_id, acc);
}

// Radix-16 digits of the window sums: out[w * groups + g] = sum_{k<4} 2^k P_{4g+k}, so that the host tail adds one point per FOUR
// bit positions instead of one per position. Four lanes per digit: lane k doubles P_{4g+k} k times, a 4-lane butterfly adds them
// (at most 3 + 2 dependent operations; a single thread per digit needs 8).
template <class T, bool INL>
__global__ void __launch_bounds__(64) k_plane_combine(const uint32_t* planes_in, uint32_t planes, uint32_t groups, uint32_t num_windows,
                                                      uint32_t* out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t g = t >> 2;
  const uint32_t k = (uint32_t)(t & 3u);
  const bool live = g < (size_t)groups * num_windows;
  Xyzz<T> r = Xyzz<T>::inf();
  if (live) {
    const uint32_t w = (uint32_t)(g / groups), grp = (uint32_t)(g % groups);
    const uint32_t q = 4u * grp + k;
    if (q < planes) {
      r = load_xyzz<T>(planes_in, (size_t)w * planes + q);
#pragma unroll 1
      for (uint32_t i = 0; i < k; i++) pdbl<T, INL>(r);
    }
  }
  group_butterfly<T, INL>(r, 4);
  if (live && k == 0) store_xyzz(out, g, r);
}

// Batch tail: one thread per MSM of a batch. parts[(m * nwd + w) * row + i] are the <= 4 partial sums of window w of
// MSM m; the thread adds them up and runs Horner over the windows (c doublings + one addition per window, reference
// ec_multi_scalar_mul_parallel.nim:198-203). nwd = 1 (table mode) leaves just the partial sums.
template <class T>
__global__ void __launch_bounds__(64) k_batch_tail(const uint32_t* parts, uint32_t row, int nwd, int c, uint32_t batch, uint32_t* out) {
  constexpr bool FAST = T::WORDS <= 12;   // single-field coordinates: unrolled multipliers inline (latency); Fp2: out of line
  uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= batch) return;
  Xyzz<T> r = Xyzz<T>::inf();
#pragma unroll 1
  for (int w = nwd - 1; w >= 0; w--) {
    if (w != nwd - 1) {
#pragma unroll 1
      for (int i = 0; i < c; i++) {
        if constexpr (FAST) r = xyzz_dbl_u(r); else xyzz_dbl_ni(r);
      }
    }
#pragma unroll 1
    for (uint32_t i = 0; i < row; i++) {
      Xyzz<T> p = load_xyzz<T>(parts, ((size_t)m * nwd + w) * row + i);
      if constexpr (FAST) xyzz_add_u(r, p); else xyzz_add_ni(r, p);
    }
  }
  store_xyzz(out, m, r);
}

// Sum of affine points (reference sum_reduce_vartime, ec_shortweierstrass_batch_ops.nim:649-664): thread t adds the
// points t, t + T, t + 2T, ... (neighbouring threads read neighbouring points); k_row_sum_warp finishes.
template <class T>
__global__ void __launch_bounds__(128) k_sum_strided(const uint32_t* __restrict__ points, size_t n, uint32_t* out) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  Xyzz<T> acc = Xyzz<T>::inf();
#pragma unroll 1
  for (size_t i = t; i < n; i += stride) {
    Aff<T> p = load_affine<T>(points, (uint32_t)i);
    xyzz_madd_ni(acc, p);
  }
  store_xyzz(out, t, acc);
}

}  // namespace b200
