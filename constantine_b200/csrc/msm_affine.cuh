// Batched-affine bucket accumulation for one B200 (sm_100a).
//
// Replaces on the reference's hot path:
//   Scheduler / schedule / sparseVectorAddition   reference constantine/math/elliptic/ec_multi_scalar_mul_scheduler.nim:234-553
//   affineAdd, lambdaAdd / lambdaDouble            reference constantine/math/elliptic/ec_shortweierstrass_batch_ops.nim:424-455
//   inv_vartime                                    reference constantine/math/arithmetic/finite_fields.nim:386-396 (field_inv.cuh here)
// The reference queues bucket updates until a collision-free batch can share one inversion (6 multiplications per affine
// addition instead of 11 for a Jacobian mixed addition). The same arithmetic is mapped onto the GPU very differently:
//
//   * after the radix sort a bucket is a RUN of equal keys; its sum is a balanced binary tree over the run. Level r of ALL
//     trees is one dense array of independent pair additions, because every bucket's slots at every level are laid out by
//     prefix sums of ceil(n_b / 2^r) (k_level_blocksums / k_level_scan / k_level_offsets): slot j of bucket b at level r+1 is the
//     sum of slots 2j and 2j+1 of level r. No collisions exist by construction, nothing is queued or rescheduled.
//   * k_affine_plan (one thread per sorted entry) writes, for every level, the source of each slot -- so the arithmetic kernel
//     k_affine_pairs is a plain list processor: lane = slot, perfectly regular, independent of the digit distribution.
//   * the shared inversion is PER THREAD: a thread walks M slots (M = level size / resident threads, ~150 at N = 2^20),
//     multiplies their denominators into a running product (prefix products parked in a coalesced global scratch), inverts
//     once (safegcd, field_inv.cuh) and unwinds: 1 + 5 multiplications per addition + inversion / M. Lanes never wait for
//     each other: no block-wide scan, no barrier, every lane of a warp inverts at the same time.
//   * after L levels a bucket has ceil(n_b / 2^L) survivors (one, for the typical run); they go through the generic XYZZ
//     slice kernel (k_accumulate), which also absorbs any adversarial distribution (all scalars equal: one run of N entries).
// Special cases inside a batch (infinity operand, P + P, P - P) contribute nothing to the product and are resolved outside the
// shared inversion (copy / affine doubling with its own denominator 2y / infinity), the cases the reference's
// scheduler handles at ec_multi_scalar_mul_scheduler.nim:465-479,513-516.
#pragma once
#include "ec.cuh"
#include "field_inv.cuh"

namespace b200 {

constexpr int AFF_MAX_LEVELS = 8;
constexpr uint32_t AFF_NONE = 0xFFFFFFFFu;

// ------------------------------------------------------------------------------------------------ run bounds
// head[b] = first sorted position of key b, tail[b] = one past its last (both pre-zeroed: empty buckets have length 0)
static __global__ void k_bucket_bounds(const uint32_t* __restrict__ keys, size_t n, uint32_t no_key, uint32_t* head, uint32_t* tail) {
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const uint32_t key = keys[q];
  if (key >= no_key) return;
  if (q == 0 || keys[q - 1] != key) head[key] = (uint32_t)q;
  if (q + 1 == n || keys[q + 1] != key) tail[key] = (uint32_t)(q + 1);
}

// ------------------------------------------------------------------------------------------------ level offsets
// off[r * (nb + 1) + b] = sum_{b' < b} ceil(n_b' / 2^r) for r = 0..L; off[r * (nb + 1) + nb] = size of level r.
// Three small kernels (block sums, scan of the block sums, offsets); a block covers SCAN_ITEMS buckets.
constexpr int SCAN_THREADS = 256, SCAN_PER_THREAD = 4, SCAN_ITEMS = SCAN_THREADS * SCAN_PER_THREAD;

// exclusive prefix of v over the block's threads; total = sum over the block (all threads get it)
B200_DEV uint32_t block_exclusive_scan(uint32_t v, uint32_t& total, uint32_t* smem /* SCAN_THREADS / 32 + 1 words */) {
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= (unsigned)d) inc += t;
  }
  __syncthreads();   // smem may still be read by the previous call
  if (lane == 31u) smem[warp] = inc;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; w++) {
    const uint32_t s = smem[w];
    if ((unsigned)w < warp) base += s;
    tot += s;
  }
  total = tot;
  return base + inc - v;
}

B200_DEV uint32_t level_count(uint32_t n, int r) { return (n + (1u << r) - 1u) >> r; }

static __global__ void __launch_bounds__(SCAN_THREADS) k_level_blocksums(const uint32_t* __restrict__ head, const uint32_t* __restrict__ tail,
                                                                          uint32_t nb, int L, uint32_t nblk, uint32_t* blocksum /* [(L+1)][nblk] */) {
  __shared__ uint32_t sm[SCAN_THREADS / 32 + 1];
  const uint32_t b0 = blockIdx.x * SCAN_ITEMS + threadIdx.x * SCAN_PER_THREAD;
  uint32_t cnt[SCAN_PER_THREAD];
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; k++) cnt[k] = (b0 + k < nb) ? tail[b0 + k] - head[b0 + k] : 0u;
  for (int r = 0; r <= L; r++) {
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) s += level_count(cnt[k], r);
    uint32_t tot;
    block_exclusive_scan(s, tot, sm);
    if (threadIdx.x == 0) blocksum[(size_t)r * nblk + blockIdx.x] = tot;
  }
}

// one block: exclusive scan of the block sums of every level; also stores the level sizes at off[r][nb]
static __global__ void __launch_bounds__(SCAN_THREADS) k_level_scan(uint32_t* blocksum, uint32_t nblk, int L, uint32_t nb, uint32_t* off) {
  __shared__ uint32_t sm[SCAN_THREADS / 32 + 1];
  for (int r = 0; r <= L; r++) {
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nblk; base += SCAN_THREADS) {
      const uint32_t i = base + threadIdx.x;
      const uint32_t v = (i < nblk) ? blocksum[(size_t)r * nblk + i] : 0u;
      uint32_t tot;
      const uint32_t ex = block_exclusive_scan(v, tot, sm);
      if (i < nblk) blocksum[(size_t)r * nblk + i] = carry + ex;
      carry += tot;
    }
    if (threadIdx.x == 0) off[(size_t)r * (nb + 1) + nb] = carry;
  }
}

static __global__ void __launch_bounds__(SCAN_THREADS) k_level_offsets(const uint32_t* __restrict__ head, const uint32_t* __restrict__ tail,
                                                                        uint32_t nb, int L, uint32_t nblk, const uint32_t* __restrict__ blocksum,
                                                                        uint32_t* off) {
  __shared__ uint32_t sm[SCAN_THREADS / 32 + 1];
  const uint32_t b0 = blockIdx.x * SCAN_ITEMS + threadIdx.x * SCAN_PER_THREAD;
  uint32_t cnt[SCAN_PER_THREAD];
#pragma unroll
  for (int k = 0; k < SCAN_PER_THREAD; k++) cnt[k] = (b0 + k < nb) ? tail[b0 + k] - head[b0 + k] : 0u;
  for (int r = 0; r <= L; r++) {
    uint32_t c[SCAN_PER_THREAD], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) { c[k] = level_count(cnt[k], r); s += c[k]; }
    uint32_t tot;
    uint32_t ex = block_exclusive_scan(s, tot, sm) + blocksum[(size_t)r * nblk + blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; k++) {
      if (b0 + k < nb) off[(size_t)r * (nb + 1) + b0 + k] = ex;
      ex += c[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------ plan
struct AffinePlan {
  uint2* plan0;                        // level 0 -> 1: (point ref | sign << 31, partner ref or AFF_NONE)
  uint32_t* plan[AFF_MAX_LEVELS];      // level r -> r+1 (r >= 1): index of the left operand in level r | has-partner << 31
  uint32_t* surv_keys;                 // survivors of level L: bucket key per slot (sorted), and the identity map as "point refs"
  uint32_t* surv_vals;
};

// One thread per sorted entry q (bucket b, offset i in its run): entry q is the leftmost leaf of the level-(r+1) slot
// i >> (r+1) of its bucket iff 2^(r+1) divides i.
static __global__ void k_affine_plan(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, size_t n, uint32_t no_key,
                                     const uint32_t* __restrict__ head, const uint32_t* __restrict__ tail,
                                     const uint32_t* __restrict__ off, uint32_t nb, int L, AffinePlan P) {
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const uint32_t b = keys[q];
  if (b >= no_key) return;
  const uint32_t h = head[b], i = (uint32_t)q - h, cnt = tail[b] - h;
  const size_t stride = (size_t)nb + 1;
  for (int r = 0; r < L; r++) {
    if (i & ((2u << r) - 1u)) break;
    const uint32_t p = off[(size_t)(r + 1) * stride + b] + (i >> (r + 1));
    const uint32_t s = i >> r;
    const bool has2 = s + 1u < level_count(cnt, r);
    if (r == 0) P.plan0[p] = make_uint2(vals[q], has2 ? vals[q + 1] : AFF_NONE);
    else P.plan[r][p] = (off[(size_t)r * stride + b] + s) | (has2 ? 0x80000000u : 0u);
  }
  if ((i & ((1u << L) - 1u)) == 0u) {
    const uint32_t ps = off[(size_t)L * stride + b] + (i >> L);
    P.surv_keys[ps] = b;
    P.surv_vals[ps] = ps;
  }
}

// ------------------------------------------------------------------------------------------------ pair additions
template <class T>
B200_DEV Aff<T> load_affine_rw(const uint32_t* base, size_t idx) {
  const uint32_t* p = base + idx * (2 * T::WORDS);
  Aff<T> a;
  load_words_rw(a.x, p);
  load_words_rw(a.y, p + T::WORDS);
  return a;
}
template <class T>
B200_DEV void store_affine(uint32_t* base, size_t idx, const Aff<T>& a) {
  uint32_t* p = base + idx * (2 * T::WORDS);
  store_words(p, a.x);
  store_words(p + T::WORDS, a.y);
}

// per-thread prefix products: element j of thread t at rows (j * V + v) of a [rows][threads] uint4 matrix (V vectors per element)
template <class T>
B200_DEV void scratch_store(uint4* scratch, size_t threads, size_t tid, uint32_t j, const T& v) {
  constexpr int V = T::WORDS / 4;
#pragma unroll
  for (int k = 0; k < V; k++) {
    uint4 w;
    w.x = v.word(4 * k + 0); w.y = v.word(4 * k + 1); w.z = v.word(4 * k + 2); w.w = v.word(4 * k + 3);
    scratch[((size_t)j * V + k) * threads + tid] = w;
  }
}
template <class T>
B200_DEV T scratch_load(const uint4* scratch, size_t threads, size_t tid, uint32_t j) {
  constexpr int V = T::WORDS / 4;
  T v;
#pragma unroll
  for (int k = 0; k < V; k++) {
    const uint4 w = scratch[((size_t)j * V + k) * threads + tid];
    v.set_word(4 * k + 0, w.x); v.set_word(4 * k + 1, w.y); v.set_word(4 * k + 2, w.z); v.set_word(4 * k + 3, w.w);
  }
  return v;
}

// ------------------------------------------------------------------------------------------------ level 0 by arrival of the points
// A host call moves its points over PCIe in P chunks. A level-0 pair only needs the chunks of its own operands, so the pair list is
// partitioned (stable counting sort, P <= 8 classes) by the LAST chunk a pair touches: launch q of the pair kernel runs as soon as
// chunk q has landed, and only the last launch waits for the whole transfer.
constexpr int PART_TILE = 1024, PART_THREADS = 256, PART_MAX = 8;

B200_DEV uint32_t pair_chunk(const uint2 task, uint32_t n_points, uint32_t P) {
  uint32_t idx = task.x & 0x7FFFFFFFu;
  if (task.y != AFF_NONE) { const uint32_t j = task.y & 0x7FFFFFFFu; if (j > idx) idx = j; }
  uint32_t q = (uint32_t)(((unsigned long long)idx * P) / n_points);
  return q < P ? q : P - 1u;
}

// counts[q * nblk + blk] = slots of tile blk whose pair belongs to class q
static __global__ void __launch_bounds__(PART_THREADS) k_part_count(const uint2* __restrict__ plan0, const uint32_t* __restrict__ total_ptr,
                                                                    uint32_t n_points, uint32_t P, uint32_t nblk, uint32_t* counts) {
  __shared__ uint32_t sm[PART_MAX];
  if (threadIdx.x < PART_MAX) sm[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t total = *total_ptr;
  const uint32_t base = blockIdx.x * PART_TILE;
  uint32_t local[PART_MAX];
#pragma unroll
  for (int q = 0; q < PART_MAX; q++) local[q] = 0;
  for (uint32_t i = threadIdx.x; i < PART_TILE; i += PART_THREADS) {
    const uint32_t p = base + i;
    if (p < total) {
      const uint32_t q = pair_chunk(plan0[p], n_points, P);
#pragma unroll
      for (int k = 0; k < PART_MAX; k++) local[k] += (q == (uint32_t)k) ? 1u : 0u;
    }
  }
#pragma unroll
  for (int q = 0; q < PART_MAX; q++) {
    uint32_t v = local[q];
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    if ((threadIdx.x & 31u) == 0 && v) atomicAdd(&sm[q], v);
  }
  __syncthreads();
  if (threadIdx.x < P) counts[threadIdx.x * nblk + blockIdx.x] = sm[threadIdx.x];
}

// one block: exclusive scan of counts in (class, tile) order, in place; starts[q] = first position of class q, starts[P] = total
static __global__ void __launch_bounds__(SCAN_THREADS) k_part_scan(uint32_t* counts, uint32_t P, uint32_t nblk, uint32_t* starts) {
  __shared__ uint32_t sm[SCAN_THREADS / 32 + 1];
  uint32_t carry = 0;
  const uint32_t len = P * nblk;
  for (uint32_t base = 0; base < len; base += SCAN_THREADS) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = (i < len) ? counts[i] : 0u;
    uint32_t tot;
    const uint32_t ex = block_exclusive_scan(v, tot, sm);
    if (i < len) {
      counts[i] = carry + ex;
      if (i % nblk == 0) starts[i / nblk] = carry + ex;
    }
    carry += tot;
  }
  if (threadIdx.x == 0) starts[P] = carry;
}

// perm[offset(class, tile) + rank within (class, tile)] = slot, ranks in slot order (stable)
static __global__ void __launch_bounds__(PART_THREADS) k_part_scatter(const uint2* __restrict__ plan0, const uint32_t* __restrict__ total_ptr,
                                                                      uint32_t n_points, uint32_t P, uint32_t nblk,
                                                                      const uint32_t* __restrict__ offsets, uint32_t* perm) {
  __shared__ uint32_t warp_cnt[PART_THREADS / 32][PART_MAX];
  __shared__ uint32_t run_base[PART_MAX];
  const uint32_t total = *total_ptr;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x < PART_MAX) run_base[threadIdx.x] = (threadIdx.x < P) ? offsets[threadIdx.x * nblk + blockIdx.x] : 0u;
  __syncthreads();
  // the tile is walked in rounds of PART_THREADS consecutive slots so that ranks follow slot order
  for (uint32_t round = 0; round < PART_TILE / PART_THREADS; round++) {
    const uint32_t p = blockIdx.x * PART_TILE + round * PART_THREADS + threadIdx.x;
    const bool live = p < total;
    const uint32_t q = live ? pair_chunk(plan0[p], n_points, P) : 0xFFFFFFFFu;
    uint32_t my_rank = 0;
    for (uint32_t k = 0; k < P; k++) {
      const unsigned m = __ballot_sync(0xFFFFFFFFu, q == k);
      if (q == k) my_rank = __popc(m & ((1u << lane) - 1u));
      if (lane == 0) warp_cnt[warp][k] = __popc(m);
    }
    __syncthreads();
    if (live) {
      uint32_t before = 0;
      for (unsigned w = 0; w < warp; w++) before += warp_cnt[w][q];
      perm[run_base[q] + before + my_rank] = p;
    }
    __syncthreads();
    if (threadIdx.x < P) {
      uint32_t tot = 0;
      for (unsigned w = 0; w < PART_THREADS / 32; w++) tot += warp_cnt[w][threadIdx.x];
      run_base[threadIdx.x] += tot;
    }
    __syncthreads();
  }
}

enum PairKind { PAIR_COPY1 = 0, PAIR_COPY2 = 1, PAIR_INF = 2, PAIR_ADD = 3, PAIR_DBL = 4 };

// one slot's operands. FIRST: references into the caller's point array (sign in bit 31); else slots of the previous level.
struct PairTask {
  uint32_t a, b;     // b == AFF_NONE: single operand
};
template <bool FIRST>
B200_DEV PairTask load_task(const void* plan, size_t p) {
  PairTask t;
  if constexpr (FIRST) {
    const uint2 v = reinterpret_cast<const uint2*>(plan)[p];
    t.a = v.x; t.b = v.y;
  } else {
    const uint32_t v = reinterpret_cast<const uint32_t*>(plan)[p];
    t.a = v & 0x7FFFFFFFu;
    t.b = (v >> 31) ? t.a + 1u : AFF_NONE;
  }
  return t;
}
template <class T, bool FIRST>
B200_DEV T load_x(const uint32_t* src, uint32_t ref) {
  T x;
  if constexpr (FIRST) load_words(x, src + (size_t)(ref & 0x7FFFFFFFu) * (2 * T::WORDS));
  else load_words_rw(x, src + (size_t)ref * (2 * T::WORDS));
  return x;
}
template <class T, bool FIRST>
B200_DEV T load_y(const uint32_t* src, uint32_t ref, bool& is_inf_if_x_zero) {
  T y;
  if constexpr (FIRST) {
    load_words(y, src + (size_t)(ref & 0x7FFFFFFFu) * (2 * T::WORDS) + T::WORDS);
    is_inf_if_x_zero = y.is_zero();
    y.cneg((ref >> 31) != 0);
  } else {
    load_words_rw(y, src + (size_t)ref * (2 * T::WORDS) + T::WORDS);
    is_inf_if_x_zero = y.is_zero();
  }
  return y;
}

// Classification shared by both passes. den is the factor this pair contributes to the batch product (ADD: x2 - x1, DBL: 2 y1).
template <class T>
B200_DEV int classify_pair(bool single, const T& x1, const T& y1, bool inf1, const T& x2, const T& y2, bool inf2, T& den) {
  if (single) return PAIR_COPY1;
  if (inf1) return PAIR_COPY2;
  if (inf2) return PAIR_COPY1;
  den = x2 - x1;
  if (!den.is_zero()) return PAIR_ADD;
  if (!(y1 == y2) || y1.is_zero()) return PAIR_INF;     // P + (-P), or doubling a point of order two
  den = y1.dbl();
  return PAIR_DBL;
}

#ifndef B200_AFF_THREADS
#define B200_AFF_THREADS 128
#endif
#ifndef B200_AFF_MIN_BLOCKS
#define B200_AFF_MIN_BLOCKS 4      // measured at N = 2^20: 4 blocks (128 registers) 7.50 ms, 3 blocks (141 registers) 7.65 ms per MSM
#endif

B200_DEV void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// both 128-byte lines an operand of `bytes` bytes at p can touch
B200_DEV void prefetch_span(const void* p, int bytes) {
  prefetch_l2(p);
  prefetch_l2((const char*)p + bytes - 1);
}
template <class T, bool FIRST>
B200_DEV void prefetch_point(const uint32_t* src, uint32_t ref) {
  const uint32_t idx = FIRST ? (ref & 0x7FFFFFFFu) : ref;
  prefetch_span(src + (size_t)idx * (2 * T::WORDS), 2 * T::WORDS * 4);
}

// dst[p] = src[a_p] (+ src[b_p]) for p < *total_ptr. Persistent: the grid's threads split the slots evenly; a warp owns a
// contiguous range of 32 M slots and lane l takes slots l, l + 32, ... of it (coalesced plans, outputs and level >= 1 operands).
template <class T, bool FIRST>
__global__ void __launch_bounds__(B200_AFF_THREADS, (T::WORDS <= 12) ? B200_AFF_MIN_BLOCKS : 1)
k_affine_pairs(const void* __restrict__ plan, const uint32_t* __restrict__ total_ptr, const uint32_t* src, uint32_t* dst, uint4* scratch,
               const uint32_t* __restrict__ perm = nullptr, const uint32_t* __restrict__ range = nullptr) {
  // perm / range (level 0 of a host call whose points arrive in chunks): this launch handles the slots perm[range[0] .. range[1]),
  // i.e. the pairs whose operands all lie in the chunks that have arrived; otherwise slots 0 .. *total_ptr in order.
  const size_t threads = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t range_begin = range ? range[0] : 0u;
  const uint32_t total = range ? range[1] - range[0] : *total_ptr;
  if (perm) perm += range_begin;
  const uint32_t M = (uint32_t)(((size_t)total + threads - 1) / threads);
  const unsigned lane = threadIdx.x & 31u;
  const size_t warp_base = (tid - lane) * (size_t)M;          // first slot of this warp
  if (warp_base >= total) return;
  // slots of this lane: warp_base + lane + 32 j, j < cnt
  uint32_t cnt = 0;
  {
    const size_t first = warp_base + lane;
    if (first < total) {
      const size_t left = (size_t)total - first;
      cnt = (uint32_t)((left + 31) / 32);
      if (cnt > M) cnt = M;
    }
  }
  // ---- pass 1: running product of the denominators; prefix products to the scratch
  T run = T::one();
  {
    auto slot_of = [&](size_t t) -> size_t { return perm ? (size_t)perm[t] : t; };
    PairTask t_next = cnt ? load_task<FIRST>(plan, slot_of(warp_base + lane)) : PairTask{0u, AFF_NONE};
    T x1n = T::zero(), x2n = T::zero();
    if (cnt) {
      x1n = load_x<T, FIRST>(src, t_next.a);
      if (t_next.b != AFF_NONE) x2n = load_x<T, FIRST>(src, t_next.b);
    }
#pragma unroll 1
    for (uint32_t j = 0; j < cnt; j++) {
      const PairTask t = t_next;
      const T x1 = x1n, x2 = x2n;
      if (j + 1 < cnt) {
        t_next = load_task<FIRST>(plan, slot_of(warp_base + lane + 32u * (size_t)(j + 1)));
        x1n = load_x<T, FIRST>(src, t_next.a);
        if (t_next.b != AFF_NONE) x2n = load_x<T, FIRST>(src, t_next.b);
      }
      const bool single = t.b == AFF_NONE;
      T den = x2 - x1;
      bool contributes = !single;
      if (!single && (den.is_zero() || x1.is_zero() || x2.is_zero())) {
        // rare: equal abscissae (doubling or cancellation) or a possible infinity operand -- needs the ordinates
        bool z1, z2;
        const T y1 = load_y<T, FIRST>(src, t.a, z1), y2 = load_y<T, FIRST>(src, t.b, z2);
        const int kind = classify_pair(false, x1, y1, z1 && x1.is_zero(), x2, y2, z2 && x2.is_zero(), den);
        contributes = kind >= PAIR_ADD;
      }
      if (contributes) run = run.mul_u(den);
      scratch_store(scratch, threads, tid, j, run);
    }
  }
  // ---- the shared inversion of this thread's batch
  T inv = fe_inverse(run);
  // ---- pass 2: unwind, last slot first. The operands of slot j - 1 (found through its plan entry) and the prefix product it
  // will need are pulled into L2 while slot j is being computed: the dependent plan -> point gather then costs an L2 hit.
  auto slot_of2 = [&](size_t t) -> size_t { return perm ? (size_t)perm[t] : t; };
  size_t p_prev = cnt ? slot_of2(warp_base + lane + 32u * (size_t)(cnt - 1)) : 0;
  PairTask t_prev = cnt ? load_task<FIRST>(plan, p_prev) : PairTask{0u, AFF_NONE};
#pragma unroll 1
  for (uint32_t jj = cnt; jj > 0; jj--) {
    const uint32_t j = jj - 1;
    const size_t p = p_prev;
    const PairTask t = t_prev;
    if (j > 0) {
      p_prev = slot_of2(warp_base + lane + 32u * (size_t)(j - 1));
      t_prev = load_task<FIRST>(plan, p_prev);
      prefetch_point<T, FIRST>(src, t_prev.a);
      if (t_prev.b != AFF_NONE) prefetch_point<T, FIRST>(src, t_prev.b);
      if (j > 1) {
        constexpr int V = T::WORDS / 4;
#pragma unroll
        for (int k = 0; k < V; k++) prefetch_l2(scratch + ((size_t)(j - 2) * V + k) * threads + tid);
      }
    }
    const bool single = t.b == AFF_NONE;
    Aff<T> P1, P2;
    bool z1 = false, z2 = false;
    P1.x = load_x<T, FIRST>(src, t.a);
    P1.y = load_y<T, FIRST>(src, t.a, z1);
    if (!single) {
      P2.x = load_x<T, FIRST>(src, t.b);
      P2.y = load_y<T, FIRST>(src, t.b, z2);
    } else {
      P2.x = T::zero(); P2.y = T::zero();
    }
    T den;
    const int kind = classify_pair(single, P1.x, P1.y, z1 && P1.x.is_zero(), P2.x, P2.y, z2 && P2.x.is_zero(), den);
    Aff<T> R;
    if (kind < PAIR_ADD) {
      if (kind == PAIR_COPY1) R = P1;
      else if (kind == PAIR_COPY2) R = P2;
      else { R.x = T::zero(); R.y = T::zero(); }
    } else {
      T inv_den = inv;
      if (j > 0) inv_den = inv.mul_u(scratch_load<T>(scratch, threads, tid, j - 1));
      inv = inv.mul_u(den);
      T num;
      if (kind == PAIR_ADD) num = P2.y - P1.y;
      else { const T xx = P1.x.sqr(); num = xx.dbl() + xx; }     // 3 x^2 (a = 0); rare: rolled multiplier
      // unrolled multipliers: the rolled form spends a quarter of its issue slots rotating registers (ncu, profiles/)
      const T lam = num.mul_u(inv_den);
      R.x = lam.sqr_u() - P1.x - P2.x;
      R.y = lam.mul_u(P1.x - R.x) - P1.y;
    }
    store_affine(dst, p, R);
  }
}

}  // namespace b200
