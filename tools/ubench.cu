// Micro-benchmarks + self-check for the field / point primitives on one B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/ubench tools/ubench.cu
// Prints one JSON object per line.  Used to fix the integer-pipe roofline denominator
// ("measure with a register-only IMAD microbenchmark before quoting a roofline").
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../constantine_b200/csrc/ec.cuh"
#include "../constantine_b200/csrc/host_field.hpp"
#include "field_rr.cuh"
#include "field_two_pipe.cuh"

using namespace b200;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// ---------------------------------------------------------------- raw pipe rates
template <int MODE>
__global__ void k_pipe(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed * 3 + 1;
  uint32_t x0 = 1, x1 = 2, x2 = 3, x3 = 4, x4 = 5, x5 = 6, x6 = 7, x7 = 8;
  uint64_t w0 = 1, w1 = 2, w2 = 3, w3 = 4, w4 = 5, w5 = 6, w6 = 7, w7 = 8;
  double d0 = 1.0, d1 = 1.1, d2 = 1.2, d3 = 1.3, d4 = 1.4, d5 = 1.5, d6 = 1.6, d7 = 1.7, da = 1.0000001, db = 1e-9;
  for (int i = 0; i < iters; i++) {
    if (MODE == 0) {  // mad.lo.u32
#define OP(x) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x) : "r"(a), "r"(b));
      OP(x0) OP(x1) OP(x2) OP(x3) OP(x4) OP(x5) OP(x6) OP(x7)
#undef OP
    } else if (MODE == 1) {  // mad.hi.u32
#define OP(x) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x) : "r"(a), "r"(b));
      OP(x0) OP(x1) OP(x2) OP(x3) OP(x4) OP(x5) OP(x6) OP(x7)
#undef OP
    } else if (MODE == 2) {  // mad.wide.u32 with a 64-bit addend; operands differ per chain so ptxas cannot share one product
#define OP(x, k) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x) : "r"(a + k), "r"((uint32_t)x));
      OP(w0, 0) OP(w1, 1) OP(w2, 2) OP(w3, 3) OP(w4, 4) OP(w5, 5) OP(w6, 6) OP(w7, 7)
#undef OP
    } else if (MODE == 8) {  // mul.wide.u32 (no addend), results folded with xor on the ALU pipe
#define OP(x, k) { uint64_t t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"((uint32_t)x + k), "r"(b)); x ^= t; }
      OP(w0, 0) OP(w1, 1) OP(w2, 2) OP(w3, 3) OP(w4, 4) OP(w5, 5) OP(w6, 6) OP(w7, 7)
#undef OP
    } else if (MODE == 3) {  // carry chains: mad.lo.cc / madc.hi.cc pairs (4 pairs per chain, 2 chains)
      asm volatile(
          "mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %8, %9, %2; madc.hi.cc.u32 %3, %8, %9, %3;"
          "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5; madc.lo.cc.u32 %6, %8, %9, %6; madc.hi.u32 %7, %8, %9, %7;"
          : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(a), "r"(b));
    } else if (MODE == 4) {  // fma.rn.f64
#define OP(x) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(da), "d"(db));
      OP(d0) OP(d1) OP(d2) OP(d3) OP(d4) OP(d5) OP(d6) OP(d7)
#undef OP
    } else if (MODE == 5) {  // add.cc chains (IADD3 path)
      asm volatile(
          "add.cc.u32 %0, %0, %8; addc.cc.u32 %1, %1, %9; addc.cc.u32 %2, %2, %8; addc.cc.u32 %3, %3, %9;"
          "addc.cc.u32 %4, %4, %8; addc.cc.u32 %5, %5, %9; addc.cc.u32 %6, %6, %8; addc.u32 %7, %7, %9;"
          : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(a), "r"(b));
    } else if (MODE == 6) {  // 4 mad.lo + 4 fma.f64 interleaved (do the pipes overlap?)
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x0) : "r"(a), "r"(b));
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d0) : "d"(da), "d"(db));
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x1) : "r"(a), "r"(b));
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d1) : "d"(da), "d"(db));
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x2) : "r"(a), "r"(b));
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d2) : "d"(da), "d"(db));
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x3) : "r"(a), "r"(b));
      asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d3) : "d"(da), "d"(db));
    } else if (MODE == 7) {  // 4 mad.lo + 4 add.cc (FMA pipe + ALU pipe overlap?)
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x0) : "r"(a), "r"(b));
      asm volatile("add.u32 %0, %0, %1;" : "+r"(x4) : "r"(a));
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x1) : "r"(a), "r"(b));
      asm volatile("add.u32 %0, %0, %1;" : "+r"(x5) : "r"(a));
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x2) : "r"(a), "r"(b));
      asm volatile("add.u32 %0, %0, %1;" : "+r"(x6) : "r"(a));
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x3) : "r"(a), "r"(b));
      asm volatile("add.u32 %0, %0, %1;" : "+r"(x7) : "r"(a));
    }
  }
  uint32_t r = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7 ^ (uint32_t)(w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7) ^
               (uint32_t)((w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7) >> 32) ^ (uint32_t)(d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7);
  if (r == 0x12345678u) out[0] = r;
}

template <int MODE>
void run_pipe(const char* name, int sms, double clock_ghz) {
  uint32_t* d; CK(cudaMalloc(&d, 4));
  const int iters = 20000, threads = 256, blocks = sms * 8;
  k_pipe<MODE><<<blocks, threads>>>(d, 100, 1);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_pipe<MODE><<<blocks, threads>>>(d, iters, 7);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * threads * iters * 8.0;
  double per_s = ops / (ms * 1e-3);
  printf("{\"bench\":\"pipe\",\"op\":\"%s\",\"ms\":%.4f,\"Gops_per_s\":%.1f,\"ops_per_clk_per_sm_at_%.3fGHz\":%.2f}\n", name, ms,
         per_s * 1e-9, clock_ghz, per_s / (sms * clock_ghz * 1e9));
  cudaFree(d);
}

// ---------------------------------------------------------------- field / point throughput + correctness
template <class T>
__global__ void k_mul_chain(const uint32_t* in, uint32_t* out, int n, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T x, y;
  load_words(x, in + (size_t)(2 * i) * T::WORDS);
  load_words(y, in + (size_t)(2 * i + 1) * T::WORDS);
  for (int k = 0; k < iters; k++) { x = x * y; y = y * x; }
  store_words(out + (size_t)i * T::WORDS, x + y);
}

template <class T>
__global__ void __launch_bounds__(128) k_madd_chain(const uint32_t* pts, uint32_t* out, int n, int npts, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Xyzz<T> acc = Xyzz<T>::inf();
  uint32_t idx = (uint32_t)i % npts;
  for (int k = 0; k < iters; k++) {
    Aff<T> p = load_affine<T>(pts, idx);
    xyzz_madd(acc, p);
    idx = (idx * 1664525u + 1013904223u) % npts;
  }
  store_xyzz(out, (size_t)i, acc);
}

// two-pipe multiplier variants (field.cuh fe_mul_v): MASK selects which product rows go through IMAD.WIDE (no addend) + ALU adds
template <class F, int MASK>
__global__ void __launch_bounds__(128) k_mul_chain_v(const uint32_t* in, uint32_t* out, int n, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  typedef Fp<F> T;
  T x, y;
  load_words(x, in + (size_t)(2 * i) * T::WORDS);
  load_words(y, in + (size_t)(2 * i + 1) * T::WORDS);
#pragma unroll 1
  for (int k = 0; k < iters; k++) {
    if constexpr (MASK < 0) { fe_mul<F>(x.l, x.l, y.l); fe_mul<F>(y.l, y.l, x.l); }
    else { fe_mul_v<F, MASK>(x.l, x.l, y.l); fe_mul_v<F, MASK>(y.l, y.l, x.l); }
  }
  store_words(out + (size_t)i * T::WORDS, x + y);
}

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd64() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

template <class F>
host::HFp<F> rand_fe() {
  host::HFp<F> r;
  for (int i = 0; i < F::N64; i++) r.l[i] = rnd64();
  int top_bits = F::BITS - 64 * (F::N64 - 1);
  r.l[F::N64 - 1] &= (top_bits >= 64) ? ~0ull : ((1ull << top_bits) - 1);
  while (host::HFp<F>::geq_p(r.l)) host::HFp<F>::sub_p(r.l);
  return r;
}

template <class F>
void bench_field(const char* name, int sms, double clock_ghz) {
  typedef Fp<F> T;
  typedef host::HFp<F> H;
  const int n = sms * 2048, iters = 64;
  std::vector<H> in(2 * n);
  for (auto& v : in) v = rand_fe<F>();
  uint32_t *d_in, *d_out;
  CK(cudaMalloc(&d_in, sizeof(H) * 2 * n));
  CK(cudaMalloc(&d_out, sizeof(H) * n));
  CK(cudaMemcpy(d_in, in.data(), sizeof(H) * 2 * n, cudaMemcpyHostToDevice));
  int threads = 128;
  k_mul_chain<T><<<(n + threads - 1) / threads, threads>>>(d_in, d_out, n, 2);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_mul_chain<T><<<(n + threads - 1) / threads, threads>>>(d_in, d_out, n, iters);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<H> out(n);
  CK(cudaMemcpy(out.data(), d_out, sizeof(H) * n, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < 512; i++) {
    H x = in[2 * i], y = in[2 * i + 1];
    for (int k = 0; k < iters; k++) { x = x * y; y = y * x; }
    H r = x + y;
    if (!(r == out[i])) bad++;
  }
  double muls = (double)n * iters * 2;
  double per_s = muls / (ms * 1e-3);
  int L = F::N;
  printf("{\"bench\":\"fe_mul\",\"field\":\"%s\",\"mismatch_of_512\":%d,\"ms\":%.4f,\"Gmul_per_s\":%.2f,\"clk_per_mul_per_sm\":%.1f,"
         "\"int_macs_per_mul\":%d,\"Tmac_per_s\":%.2f}\n",
         name, bad, ms, per_s * 1e-9, sms * clock_ghz * 1e9 / per_s, 2 * L * L + L, per_s * (2 * L * L + L) * 1e-12);
  cudaFree(d_in); cudaFree(d_out);
}

template <class F, int MASK>
void bench_field_v(const char* name, int sms, double clock_ghz) {
  typedef Fp<F> T;
  typedef host::HFp<F> H;
  const int n = sms * 2048, iters = 64;
  std::vector<H> in(2 * n);
  for (auto& v : in) v = rand_fe<F>();
  uint32_t *d_in, *d_out;
  CK(cudaMalloc(&d_in, sizeof(H) * 2 * n));
  CK(cudaMalloc(&d_out, sizeof(H) * n));
  CK(cudaMemcpy(d_in, in.data(), sizeof(H) * 2 * n, cudaMemcpyHostToDevice));
  int threads = 128;
  k_mul_chain_v<F, MASK><<<(n + threads - 1) / threads, threads>>>(d_in, d_out, n, 2);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    k_mul_chain_v<F, MASK><<<(n + threads - 1) / threads, threads>>>(d_in, d_out, n, iters);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  std::vector<H> out(n);
  CK(cudaMemcpy(out.data(), d_out, sizeof(H) * n, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < 512; i++) {
    H x = in[2 * i], y = in[2 * i + 1];
    for (int k = 0; k < iters; k++) { x = x * y; y = y * x; }
    H r = x + y;
    if (!(r == out[i])) bad++;
  }
  int regs = 0;
  { cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_mul_chain_v<F, MASK>); regs = fa.numRegs; }
  double muls = (double)n * iters * 2;
  double per_s = muls / (best * 1e-3);
  int L = F::N;
  printf("{\"bench\":\"fe_mul_two_pipe\",\"field\":\"%s\",\"alu_row_mask\":%d,\"regs\":%d,\"mismatch_of_512\":%d,\"ms\":%.4f,\"Gmul_per_s\":%.2f,"
         "\"clk_per_mul_per_sm\":%.2f,\"Tmac_per_s\":%.2f}\n",
         name, MASK, regs, bad, best, per_s * 1e-9, sms * clock_ghz * 1e9 / per_s, per_s * (2 * L * L + L) * 1e-12);
  fflush(stdout);
  cudaFree(d_in); cudaFree(d_out);
}

template <class F>
void bench_two_pipe(const char* name, int sms, double clock_ghz) {
  bench_field_v<F, -1>(name, sms, clock_ghz);   // fe_mul as shipped (all rows on the multiplier pipe)
  bench_field_v<F, 0>(name, sms, clock_ghz);
  bench_field_v<F, 1>(name, sms, clock_ghz);
  bench_field_v<F, 4>(name, sms, clock_ghz);
  bench_field_v<F, 5>(name, sms, clock_ghz);
  bench_field_v<F, 10>(name, sms, clock_ghz);
  bench_field_v<F, 3>(name, sms, clock_ghz);
  bench_field_v<F, 12>(name, sms, clock_ghz);
  bench_field_v<F, 7>(name, sms, clock_ghz);
  bench_field_v<F, 15>(name, sms, clock_ghz);
}

// ---------------------------------------------------------------- reduced-radix (carry-free) multiplier prototype
template <class R, class F>
__global__ void k_rr_mul_chain(const uint32_t* in, uint32_t* out, int n, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a[F::N], b[F::N];
  for (int k = 0; k < F::N; k++) { a[k] = in[(size_t)(2 * i) * F::N + k]; b[k] = in[(size_t)(2 * i + 1) * F::N + k]; }
  R x = R::from_abi(a), y = R::from_abi(b);
  for (int k = 0; k < iters; k++) { x = x * y; y = y * x; }
  for (int k = 0; k < R::PR::NL; k++) out[(size_t)i * R::PR::NL + k] = x.l[k];
}

template <class R, class F>
__global__ void k_rr_mulps_chain(const uint32_t* in, uint32_t* out, int n, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t a[F::N], b[F::N];
  for (int k = 0; k < F::N; k++) { a[k] = in[(size_t)(2 * i) * F::N + k]; b[k] = in[(size_t)(2 * i + 1) * F::N + k]; }
  R x = R::from_abi(a), y = R::from_abi(b);
  for (int k = 0; k < iters; k++) { x = x.mul_ps(y); y = y.mul_ps(x); }
  for (int k = 0; k < R::PR::NL; k++) out[(size_t)i * R::PR::NL + k] = x.l[k];
}

template <class F, int W, int NL>
void bench_rr(const char* name, int sms, double clock_ghz) {
  typedef FpRR<F, W, NL> R;
  typedef host::HFp<F> H;
  const int n = sms * 2048, iters = 64;
  std::vector<H> in(2 * n);
  for (auto& v : in) v = rand_fe<F>();
  uint32_t *d_in, *d_out;
  CK(cudaMalloc(&d_in, sizeof(H) * 2 * n));
  CK(cudaMalloc(&d_out, 4 * NL * (size_t)n));
  CK(cudaMemcpy(d_in, in.data(), sizeof(H) * 2 * n, cudaMemcpyHostToDevice));
  int threads = 128;
  k_rr_mul_chain<R, F><<<(n + threads - 1) / threads, threads>>>(d_in, d_out, n, 2);
  CK(cudaDeviceSynchronize());
  {
    cudaEvent_t f0, f1; cudaEventCreate(&f0); cudaEventCreate(&f1);
    k_rr_mulps_chain<R, F><<<(n + threads - 1) / threads, threads>>>(d_in, d_out, n, 2);
    cudaEventRecord(f0);
    k_rr_mulps_chain<R, F><<<(n + threads - 1) / threads, threads>>>(d_in, d_out, n, iters);
    cudaEventRecord(f1);
    CK(cudaDeviceSynchronize());
    float ms2; cudaEventElapsedTime(&ms2, f0, f1);
    std::vector<uint32_t> o2((size_t)NL * n);
    CK(cudaMemcpy(o2.data(), d_out, 4 * NL * (size_t)n, cudaMemcpyDeviceToHost));
    k_rr_mul_chain<R, F><<<(n + threads - 1) / threads, threads>>>(d_in, d_out, n, iters);
    CK(cudaDeviceSynchronize());
    std::vector<uint32_t> o1((size_t)NL * n);
    CK(cudaMemcpy(o1.data(), d_out, 4 * NL * (size_t)n, cudaMemcpyDeviceToHost));
    size_t diff = 0; for (size_t q = 0; q < o1.size(); q++) diff += (o1[q] != o2[q]);
    double per_s2 = (double)n * iters * 2 / (ms2 * 1e-3);
    printf("{\"bench\":\"fe_mul_reduced_radix_product_scanning\",\"field\":\"%s\",\"W\":%d,\"limbs\":%d,\"limb_diffs_vs_operand_scanning\":%zu,\"ms\":%.4f,\"Gmul_per_s\":%.2f,\"clk_per_mul_per_sm\":%.1f}\n",
           name, W, NL, diff, ms2, per_s2 * 1e-9, sms * clock_ghz * 1e9 / per_s2);
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_rr_mul_chain<R, F><<<(n + threads - 1) / threads, threads>>>(d_in, d_out, n, iters);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<uint32_t> out((size_t)NL * n);
  CK(cudaMemcpy(out.data(), d_out, 4 * NL * (size_t)n, cudaMemcpyDeviceToHost));
  // expected: host chain in the ABI Montgomery domain, then times 2^SHIFT (change of radix R -> R')
  H two_shift = H::zero(); two_shift.l[0] = 1ull << R::PR::SHIFT;
  H r2; for (int i = 0; i < F::N64; i++) r2.l[i] = F::R264(i);
  H c = two_shift * r2;  // Montgomery form of 2^SHIFT
  int bad = 0;
  for (int i = 0; i < 512; i++) {
    H x = in[2 * i], y = in[2 * i + 1];
    for (int k = 0; k < iters; k++) { x = x * y; y = y * x; }
    H want = x * c;
    // device limbs (radix 2^W, value < 2p) -> integer -> canonical
    unsigned __int128 acc = 0; int accbits = 0; uint64_t limbs[F::N64 + 1]; int li = 0;
    for (int k = 0; k < F::N64 + 1; k++) limbs[k] = 0;
    for (int k = 0; k < NL; k++) {
      acc += (unsigned __int128)out[(size_t)i * NL + k] << accbits; accbits += W;
      while (accbits >= 64 && li < F::N64 + 1) { limbs[li++] = (uint64_t)acc; acc >>= 64; accbits -= 64; }
    }
    if (li < F::N64 + 1) limbs[li++] = (uint64_t)acc;
    H got; for (int k = 0; k < F::N64; k++) got.l[k] = limbs[k];
    bool ok = limbs[F::N64] == 0;
    for (int t = 0; t < 3 && H::geq_p(got.l); t++) H::sub_p(got.l);
    if (!ok || !(got == want)) bad++;
  }
  double muls = (double)n * iters * 2;
  double per_s = muls / (ms * 1e-3);
  printf("{\"bench\":\"fe_mul_reduced_radix\",\"field\":\"%s\",\"W\":%d,\"limbs\":%d,\"mismatch_of_512\":%d,\"ms\":%.4f,\"Gmul_per_s\":%.2f,\"clk_per_mul_per_sm\":%.1f}\n",
         name, W, NL, bad, ms, per_s * 1e-9, sms * clock_ghz * 1e9 / per_s);
  cudaFree(d_in); cudaFree(d_out);
}

template <class F>
void bench_madd(const char* name, int sms, double clock_ghz, const uint64_t* gx, const uint64_t* gy) {
  typedef Fp<F> T;
  typedef host::HFp<F> H;
  typedef host::HXyzz<H> HP;
  const int npts = 4096;
  // points k*G, k = 1..npts, normalised to affine on the host
  std::vector<H> aff(2 * npts);
  HP g; for (int i = 0; i < F::N64; i++) { g.x.l[i] = gx[i]; g.y.l[i] = gy[i]; }
  H r2; for (int i = 0; i < F::N64; i++) r2.l[i] = F::R264(i);
  g.x = g.x * r2; g.y = g.y * r2; g.zz = H::one(); g.zzz = H::one();
  HP cur = g;
  for (int k = 0; k < npts; k++) {
    H izz = cur.zz.inv(), izzz = cur.zzz.inv();
    aff[2 * k] = cur.x * izz; aff[2 * k + 1] = cur.y * izzz;
    cur = host::xyzz_add(cur, g);
  }
  const int n = sms * 512, iters = 32;
  uint32_t *d_pts, *d_out;
  CK(cudaMalloc(&d_pts, sizeof(H) * 2 * npts));
  CK(cudaMalloc(&d_out, sizeof(H) * 4 * n));
  CK(cudaMemcpy(d_pts, aff.data(), sizeof(H) * 2 * npts, cudaMemcpyHostToDevice));
  int threads = 128;
  k_madd_chain<T><<<(n + threads - 1) / threads, threads>>>(d_pts, d_out, n, npts, 2);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_madd_chain<T><<<(n + threads - 1) / threads, threads>>>(d_pts, d_out, n, npts, iters);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<H> out(4 * (size_t)n);
  CK(cudaMemcpy(out.data(), d_out, sizeof(H) * 4 * n, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < 256; i++) {
    // thread i sums points idx_0 = i % npts, idx_{k+1} = (idx_k*1664525+1013904223) % npts  (k*G has index k-1)
    HP acc = HP::inf();
    uint32_t idx = (uint32_t)i % npts;
    for (int k = 0; k < iters; k++) {
      HP q; q.x = aff[2 * idx]; q.y = aff[2 * idx + 1]; q.zz = H::one(); q.zzz = H::one();
      acc = host::xyzz_add(acc, q);
      idx = (idx * 1664525u + 1013904223u) % npts;
    }
    HP o; o.x = out[4 * i]; o.y = out[4 * i + 1]; o.zz = out[4 * i + 2]; o.zzz = out[4 * i + 3];
    bool ok = (acc.x * o.zz == o.x * acc.zz) && (acc.y * o.zzz == o.y * acc.zzz) && (acc.is_inf() == o.is_inf());
    if (!ok) bad++;
  }
  double adds = (double)n * iters;
  printf("{\"bench\":\"xyzz_madd\",\"curve\":\"%s\",\"mismatch_of_256\":%d,\"ms\":%.4f,\"Gadd_per_s\":%.3f,\"clk_per_add_per_sm\":%.0f}\n", name, bad,
         ms, adds / (ms * 1e-3) * 1e-9, sms * clock_ghz * 1e9 / (adds / (ms * 1e-3)));
  cudaFree(d_pts); cudaFree(d_out);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double clock_ghz = khz * 1e-6;
  printf("{\"device\":\"%s\",\"sms\":%d,\"clock_ghz_max\":%.3f}\n", prop.name, sms, clock_ghz);
  run_pipe<0>("mad.lo.u32", sms, clock_ghz);
  run_pipe<1>("mad.hi.u32", sms, clock_ghz);
  run_pipe<2>("mad.wide.u32 (64-bit addend, distinct operands)", sms, clock_ghz);
  run_pipe<8>("mul.wide.u32 + xor", sms, clock_ghz);
  run_pipe<3>("mad.lo.cc+madc.hi.cc chain", sms, clock_ghz);
  run_pipe<4>("fma.rn.f64", sms, clock_ghz);
  run_pipe<5>("add.cc chain", sms, clock_ghz);
  run_pipe<6>("mad.lo.u32 + fma.f64 interleaved", sms, clock_ghz);
  run_pipe<7>("mad.lo.u32 + add.u32 interleaved", sms, clock_ghz);
  bench_two_pipe<Bls12381Fp>("bls12_381_fp", sms, clock_ghz);
  bench_two_pipe<Bn254SnarksFp>("bn254_snarks_fp", sms, clock_ghz);
  if (getenv("UBENCH_TWO_PIPE_ONLY")) return 0;
  bench_field<Bls12381Fp>("bls12_381_fp", sms, clock_ghz);
  bench_field<Bn254SnarksFp>("bn254_snarks_fp", sms, clock_ghz);
  bench_field<PallasFp>("pallas_fp", sms, clock_ghz);
  bench_field<Bls12381Fr>("bls12_381_fr", sms, clock_ghz);
  bench_rr<Bls12381Fp, 28, 14>("bls12_381_fp", sms, clock_ghz);
  bench_rr<Bls12381Fp, 29, 14>("bls12_381_fp", sms, clock_ghz);
  bench_rr<Bn254SnarksFp, 29, 9>("bn254_snarks_fp", sms, clock_ghz);
  bench_rr<PallasFp, 29, 9>("pallas_fp", sms, clock_ghz);
  bench_rr<PallasFp, 26, 10>("pallas_fp", sms, clock_ghz);
  {
    const uint64_t gx[6] = {0xfb3af00adb22c6bbull, 0x6c55e83ff97a1aefull, 0xa14e3a3f171bac58ull, 0xc3688c4f9774b905ull, 0x2695638c4fa9ac0full, 0x17f1d3a73197d794ull};
    const uint64_t gy[6] = {0x0caa232946c5e7e1ull, 0xd03cc744a2888ae4ull, 0x00db18cb2c04b3edull, 0xfcf5e095d5d00af6ull, 0xa09e30ed741d8ae4ull, 0x08b3f481e3aaa0f1ull};
    bench_madd<Bls12381Fp>("bls12_381_g1", sms, clock_ghz, gx, gy);
    const uint64_t bx[4] = {1, 0, 0, 0}, by[4] = {2, 0, 0, 0};
    bench_madd<Bn254SnarksFp>("bn254_snarks_g1", sms, clock_ghz, bx, by);
  }
  return 0;
}
