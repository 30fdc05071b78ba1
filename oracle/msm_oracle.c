/* ORACLE -- test infrastructure, NOT product code.
 *
 * CPU restatement (plain C, gcc, pthreads) of the reference's multi-scalar-multiplication hot path:
 *   reference constantine/math/elliptic/ec_multi_scalar_mul_parallel.nim:148-208, 519-553, 588-628
 *   reference constantine/math/elliptic/ec_multi_scalar_mul.nim:40-95, 177-296
 *   reference constantine/math/elliptic/ec_multi_scalar_mul_scheduler.nim:172-223
 *   reference constantine/math/arithmetic/bigints.nim:360-379, 806-861
 * plus the field / point arithmetic in msm_oracle_impl.h.
 *
 * The reference itself (Nim) cannot be built in this image (no nim/nimble), so there is
 * no oracle/_ref.  Parity status of THIS restatement: PINNED -- tests/test_oracle_golden.py checks it against
 * the reference's EIP-2537 G1/G2 MSM vectors and Sage scalar-mul vectors (tests/golden/), and
 * tests/test_oracle_vs_exact.py against the exact big-integer tier (oracle/pyref.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this library.
 *
 * Build: see oracle/Makefile  (gcc -O3 -march=x86-64-v3 -shared -fPIC -pthread).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

typedef unsigned __int128 u128;

/* MULX / ADCX / ADOX multiplication for the TIMED CPU baseline (impl 3): compiled in on x86-64, switched on at run time only when
 * the CPU has BMI2 + ADX (the library is built in the dev container and travels to the GPU box) and only for impl 3, so that the
 * checker paths stay on the portable multiplication. */
#if defined(__x86_64__) && defined(__GNUC__)
#define ORACLE_HAVE_MULX_ADX 1
#endif
static int g_fast_mul = 0;

#define SCALAR_LIMBS 4 /* every scalar field on this path is 254/255 bits = 4 x u64 (reference curves/bigints.h:18-21) */

typedef struct {
  int nl;            /* 64-bit limbs */
  int bits;
  uint64_t p[6];
  uint64_t one[6];   /* R mod p */
  uint64_t r2[6];    /* R^2 mod p */
  uint64_t m0ninv;   /* -p^-1 mod 2^64 */
} field_t;

typedef struct {
  field_t fp;        /* coordinate prime field */
  field_t fr;        /* scalar field (for fr_coefs entry points) */
  int ext;           /* 1: G1 over Fp, 2: G2 over Fp2 */
  int scalar_bits;   /* declared BigInt width: 254 or 255 */
} curve_t;

/* ------------------------------------------------------------------ scalar windows */

/* reference bigints.nim:360-379 (getWindowAt): `window_size` bits starting at `bit_index`; reads the next
 * limb only if it exists, i.e. bits beyond the top limb are zero. */
static inline uint64_t get_window_at(const uint64_t* a, int bit_index, int window_size) {
  int slot = bit_index >> 6, pos = bit_index & 63;
  uint64_t mask = ((uint64_t)1 << window_size) - 1;
  if (slot >= SCALAR_LIMBS) return 0;
  uint64_t word = a[slot];
  if (pos + window_size > 64 && slot + 1 < SCALAR_LIMBS)
    return ((word >> pos) | (a[slot + 1] << (64 - pos))) & mask;
  return (word >> pos) & mask;
}

/* reference bigints.nim:806-832 (signedWindowEncoding, Booth recoding of a (bitsize+1)-bit digit) */
static inline void signed_window_encoding(uint64_t digit, int bitsize, uint64_t* val, int* neg) {
  uint64_t n = digit >> bitsize;
  uint64_t neg_mask = (uint64_t)0 - n;
  uint64_t val_mask = ((uint64_t)1 << bitsize) - 1;
  uint64_t encode = (digit + 1) >> 1;
  *val = ((encode + neg_mask) ^ neg_mask) & val_mask;
  *neg = (int)n;
}
/* reference bigints.nim:834-843 (getSignedFullWindowAt) */
static inline void signed_full_window(const uint64_t* a, int bit_index, int c, uint64_t* val, int* neg) {
  signed_window_encoding(get_window_at(a, bit_index - 1, c + 1), c, val, neg);
}
/* reference bigints.nim:845-851 (getSignedBottomWindow) */
static inline void signed_bottom_window(const uint64_t* a, int c, uint64_t* val, int* neg) {
  signed_window_encoding(get_window_at(a, 0, c) << 1, c, val, neg);
}
/* reference bigints.nim:853-861 (getSignedTopWindow) */
static inline void signed_top_window(const uint64_t* a, int top_index, int excess, uint64_t* val, int* neg) {
  signed_window_encoding(get_window_at(a, top_index - 1, excess + 1), excess + 1, val, neg);
}

/* reference ec_multi_scalar_mul_scheduler.nim:172-223 (bestBucketBitSize), float32 arithmetic as there */
int oracle_best_bucket_bit_size(long input_size, int scalar_bitwidth, int use_signed, int use_manual_tuning) {
  const float A = 10.0f, D = 6.0f;
  int s = use_signed ? 1 : 0;
  float b = (float)scalar_bitwidth;
  float min_cost = __builtin_inff();
  int result = 0;
  for (int c = 2; c <= 20; c++) {
    float b_over_c = b / (float)c;
    float bucket_accumulate_reduce = b_over_c * (float)(input_size + ((long)1 << (c - s)) - 2) * A;
    float final_reduction = (b_over_c - 1.0f) * ((float)c * D + A);
    float cost = bucket_accumulate_reduce + final_reduction;
    if (cost < min_cost) { min_cost = cost; result = c; }
  }
  if (use_manual_tuning) {
    if (14 <= result) result -= 1;
    if (15 <= result) result -= 1;
    if (16 <= result) result -= 1;
  }
  return result;
}

/* the window size the reference's parallel dispatch actually runs with
 * (reference ec_multi_scalar_mul_parallel.nim:519-553: c in 2..10 as is, 11..17 -> c-1) */
int oracle_parallel_dispatch_c(long n, int bits) {
  int c = oracle_best_bucket_bit_size(n, bits, 1, 1);
  if (c >= 11) c -= 1;
  if (c > 16) c = 16;
  return c;
}

/* ------------------------------------------------------------------ tiny task runner (one task per window) */
typedef struct {
  void (*fn)(void*); char* tasks; size_t stride; int count; int next; pthread_mutex_t mu;
} pool_t;

static void* pool_worker(void* arg) {
  pool_t* p = (pool_t*)arg;
  for (;;) {
    pthread_mutex_lock(&p->mu);
    int i = p->next++;
    pthread_mutex_unlock(&p->mu);
    if (i >= p->count) break;
    p->fn(p->tasks + (size_t)i * p->stride);
  }
  return NULL;
}

static void run_tasks(void (*fn)(void*), void* tasks, size_t stride, int count, int nthreads) {
  pool_t p;
  p.fn = fn; p.tasks = (char*)tasks; p.stride = stride; p.count = count; p.next = 0;
  pthread_mutex_init(&p.mu, NULL);
  if (nthreads > count) nthreads = count;
  if (nthreads <= 1) { pool_worker(&p); pthread_mutex_destroy(&p.mu); return; }
  pthread_t* th = (pthread_t*)malloc((size_t)nthreads * sizeof(pthread_t));
  for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, pool_worker, &p);
  for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
  free(th);
  pthread_mutex_destroy(&p.mu);
}

/* ------------------------------------------------------------------ instantiations */
#define NL 4
#define EXT 1
#include "msm_oracle_impl.h"
#undef EXT
#define EXT 2
#include "msm_oracle_impl.h"
#undef EXT
#undef NL
#define NL 6
#define EXT 1
#include "msm_oracle_impl.h"
#undef EXT
#define EXT 2
#include "msm_oracle_impl.h"
#undef EXT
#undef NL

/* Fr Montgomery residue -> canonical BigInt: one Montgomery reduction = mulMont(a, 1)
 * (reference finite_fields.nim:70-75 fromField -> limbs_montgomery.nim:577-603 fromMont) */
static void fr_from_mont(uint64_t* out, const uint64_t* in, const field_t* fr) {
  fp_4_1 a, one, r;
  memcpy(a.l, in, 32);
  memset(&one, 0, sizeof(one));
  one.l[0] = 1;
  fp_mul_4_1(&r, &a, &one, fr);
  memcpy(out, r.l, 32);
}

/* ------------------------------------------------------------------ exported API (ctypes) */

/* impl: 0 = naive double-and-add sum, 1 = unsigned-window reference bucket method, 2 = signed-window
 *       one-task-per-window method (the parallel hot path, Jacobian buckets: the checker), 3 = the same with batched-affine
 *       bucket sums (the reference's arithmetic for c >= 9; the timed CPU baseline).
 * c:    window size; <= 0 selects the reference's own choice (bestBucketBitSize / parallel dispatch).
 * coefs: n x 4 x u64; canonical BigInt when fr_mont == 0, Fr Montgomery residues when fr_mont != 0
 *        (reference bindings/c_curve_decls_parallel.nim:31-45: big_coefs vs fr_coefs entry points).
 * out:  Jacobian point, 3 coordinates of nl*ext limbs (Montgomery residues). Returns the c used, <0 on error. */
int oracle_msm(const curve_t* cv, void* out, const uint64_t* coefs, const void* points, size_t n,
               int fr_mont, int impl, int c, int nthreads) {
  const field_t* f = &cv->fp;
  int bits = cv->scalar_bits;
  uint64_t* big = NULL;
  if (fr_mont) {
    big = (uint64_t*)malloc((n ? n : 1) * 32);
    for (size_t i = 0; i < n; i++) fr_from_mont(big + 4 * i, coefs + 4 * i, &cv->fr);
    coefs = big;
  }
  if (c <= 0) c = (impl >= 2) ? oracle_parallel_dispatch_c((long)n, bits) : oracle_best_bucket_bit_size((long)n, bits, impl >= 2, 1);
  if (c < 2) c = 2;
  if (c > 20) c = 20;
  int rc = c;
#if defined(ORACLE_HAVE_MULX_ADX)
  g_fast_mul = (impl == 3) && __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("adx");
#endif
#define DISPATCH(NLV, EXTV)                                                                              \
  do {                                                                                                   \
    jac_##NLV##_##EXTV* r = (jac_##NLV##_##EXTV*)out;                                                    \
    const aff_##NLV##_##EXTV* pts = (const aff_##NLV##_##EXTV*)points;                                   \
    if (n == 0) jac_set_inf_##NLV##_##EXTV(r, f);                                                        \
    else if (impl == 0) msm_naive_##NLV##_##EXTV(r, coefs, pts, n, bits, f);                             \
    else if (impl == 1) msm_reference_##NLV##_##EXTV(r, coefs, pts, n, c, bits, f);                      \
    else msm_signed_##NLV##_##EXTV(r, coefs, pts, n, c, bits, f, nthreads, impl == 3);                   \
  } while (0)
  if (f->nl == 4 && cv->ext == 1) DISPATCH(4, 1);
  else if (f->nl == 4 && cv->ext == 2) DISPATCH(4, 2);
  else if (f->nl == 6 && cv->ext == 1) DISPATCH(6, 1);
  else if (f->nl == 6 && cv->ext == 2) DISPATCH(6, 2);
  else rc = -1;
#undef DISPATCH
  g_fast_mul = 0;
  free(big);
  return rc;
}

/* field / point primitives for unit tests: op 0 mul, 1 add, 2 sub, 3 neg(a), 4 div2(a); elements are nl limbs */
int oracle_fp_op(const field_t* f, int op, uint64_t* r, const uint64_t* a, const uint64_t* b, size_t count) {
  for (size_t i = 0; i < count; i++) {
    if (f->nl == 4) {
      fp_4_1 x, y, z; memcpy(x.l, a + 4 * i, 32); memcpy(y.l, b + 4 * i, 32);
      if (op == 0) fp_mul_4_1(&z, &x, &y, f); else if (op == 1) fp_add_4_1(&z, &x, &y, f);
      else if (op == 2) fp_sub_4_1(&z, &x, &y, f); else if (op == 3) fp_neg_4_1(&z, &x, f); else fp_div2_4_1(&z, &x, f);
      memcpy(r + 4 * i, z.l, 32);
    } else if (f->nl == 6) {
      fp_6_1 x, y, z; memcpy(x.l, a + 6 * i, 48); memcpy(y.l, b + 6 * i, 48);
      if (op == 0) fp_mul_6_1(&z, &x, &y, f); else if (op == 1) fp_add_6_1(&z, &x, &y, f);
      else if (op == 2) fp_sub_6_1(&z, &x, &y, f); else if (op == 3) fp_neg_6_1(&z, &x, f); else fp_div2_6_1(&z, &x, f);
      memcpy(r + 6 * i, z.l, 48);
    } else return -1;
  }
  return 0;
}

/* point ops on Jacobian structs: op 0: r = p + q (sum_vartime), 1: r = 2p, 2: r = p + affine q (mixedSum_vartime) */
int oracle_ec_op(const curve_t* cv, int op, void* r, const void* p, const void* q) {
  const field_t* f = &cv->fp;
#define EC(NLV, EXTV)                                                                                     \
  do {                                                                                                    \
    if (op == 0) jac_add_##NLV##_##EXTV((jac_##NLV##_##EXTV*)r, (const jac_##NLV##_##EXTV*)p, (const jac_##NLV##_##EXTV*)q, f); \
    else if (op == 1) jac_dbl_##NLV##_##EXTV((jac_##NLV##_##EXTV*)r, (const jac_##NLV##_##EXTV*)p, f);    \
    else jac_madd_##NLV##_##EXTV((jac_##NLV##_##EXTV*)r, (const jac_##NLV##_##EXTV*)p, (const aff_##NLV##_##EXTV*)q, f); \
  } while (0)
  if (f->nl == 4 && cv->ext == 1) EC(4, 1);
  else if (f->nl == 4 && cv->ext == 2) EC(4, 2);
  else if (f->nl == 6 && cv->ext == 1) EC(6, 1);
  else if (f->nl == 6 && cv->ext == 2) EC(6, 2);
  else return -1;
#undef EC
  return 0;
}

/* signed digit of scalar `k` for window w of size c (tests of the recoding: sum_w d_w 2^(wc) == k) */
int oracle_signed_digit(const uint64_t* k, int bits, int c, int w, uint64_t* val, int* neg) {
  int num_full = bits / c, excess = bits % c, top = bits - excess;
  if (w < 0 || w > num_full) return -1;
  if (w == num_full) {
    if (top == 0) signed_bottom_window(k, c, val, neg);
    else if (excess == 0) signed_full_window(k, top, c, val, neg);
    else signed_top_window(k, top, excess, val, neg);
  } else if (w == 0) signed_bottom_window(k, c, val, neg);
  else signed_full_window(k, w * c, c, val, neg);
  return 0;
}
