// SASS-level issue rates of the integer instructions a multi-limb multiplier is made of, measured in SM cycles
// (clock64 inside the kernel, so no clock-frequency assumption enters the figure).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/bin/ubench_pipes tools/ubench_pipes.cu
//   cuobjdump -sass tools/bin/ubench_pipes   (profiles/ubench_pipes_r2.sass holds the loop bodies this file compiled to)
// Every variant runs 8 independent dependency chains per thread; one "op" = one PTX statement group as labelled.
// Prints one JSON object per line: ops per SM cycle at 64 and at 8 resident warps per SM, and the SM clock seen.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

enum Mode {
  M_IMAD_LO = 0,        // mad.lo.u32 d, a, b, d                         -> IMAD
  M_IMAD_HI,            // mad.hi.u32 d, a, b, d                         -> IMAD.HI.U32
  M_MULWIDE,            // mul.wide.u32 t, lo(t), b                      -> IMAD.WIDE.U32 ..., RZ
  M_MADWIDE_ADD64,      // mad.lo.cc lo,a,b,lo ; madc.hi hi,a,b,hi       -> IMAD.WIDE.U32 with a 64-bit addend (no carry in/out)?
  M_MADWIDE_CHAIN4,     // 4 pairs mad.lo.cc/madc.hi.cc, one carry chain -> IMAD.WIDE.U32 + 3 IMAD.WIDE.U32.X
  M_MADWIDE_CHAIN12,    // 12 pairs, one carry chain (a BLS12-381 row)
  M_MULWIDE_IADD,       // mul.wide + 64-bit add (what ptxas makes of mad.wide.u32 with a 64-bit addend)
  M_IADD3,              // add.u32 x, x, a ; add.u32 x, x, b  (ptxas fuses to one 3-input IADD3?)
  M_IADD_CHAIN8,        // add.cc chain of 8
  M_MULWIDE_IADD_CC,    // mul.wide products accumulated with add.cc/addc.cc pairs (carry-deferred column sums)
  M_LOP3,               // xor/and mixes -> LOP3.LUT
  M_SHF,                // funnel shifts -> SHF
  M_DFMA,               // fma.rn.f64 on 8 chains (all warps)
  M_WIDE_HALF,          // even warps: chains of 4 mad.lo.cc/madc.hi.cc pairs; odd warps idle
  M_DFMA_HALF,          // odd warps: fma.rn.f64; even warps idle
  M_MIX_WIDE_DFMA,      // even warps IMAD.WIDE chains, odd warps DFMA: do the two pipes run side by side?
  M_COUNT
};

template <int MODE>
__global__ void k_pipe(uint32_t* out, unsigned long long* cycles, int iters, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed * 3 + 1;
  uint32_t x0 = 1, x1 = 2, x2 = 3, x3 = 4, x4 = 5, x5 = 6, x6 = 7, x7 = 8;
  uint32_t y0 = 9, y1 = 10, y2 = 11, y3 = 12, y4 = 13, y5 = 14, y6 = 15, y7 = 16;
  uint32_t z0 = 1, z1 = 1, z2 = 1, z3 = 1, z4 = 1, z5 = 1, z6 = 1, z7 = 1;
  uint64_t w0 = 1, w1 = 2, w2 = 3, w3 = 4, w4 = 5, w5 = 6, w6 = 7, w7 = 8;
  w0 += a; w1 += a; w2 += a; w3 += a; w4 += a; w5 += a; w6 += a; w7 += a;   // per-thread values: keep the chains off the uniform datapath
  double f0 = 1.0 + a, f1 = 2.0 + a, f2 = 3.0, f3 = 4.0, f4 = 5.0, f5 = 6.0, f6 = 7.0, f7 = 8.0;
  const double fa = 1.0000001 + 1e-9 * threadIdx.x, fb = 1e-7;
  const bool odd_warp = (threadIdx.x >> 5) & 1;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
    if (MODE == M_IMAD_LO) {
#define OP(x) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x) : "r"(a), "r"(b));
      OP(x0) OP(x1) OP(x2) OP(x3) OP(x4) OP(x5) OP(x6) OP(x7)
#undef OP
    } else if (MODE == M_IMAD_HI) {
#define OP(x) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(x) : "r"(a), "r"(b));
      OP(x0) OP(x1) OP(x2) OP(x3) OP(x4) OP(x5) OP(x6) OP(x7)
#undef OP
    } else if (MODE == M_MULWIDE) {
#define OP(x) asm volatile("{ .reg .u32 lo; cvt.u32.u64 lo, %0; mul.wide.u32 %0, lo, %1; }" : "+l"(x) : "r"(a));
      OP(w0) OP(w1) OP(w2) OP(w3) OP(w4) OP(w5) OP(w6) OP(w7)
#undef OP
    } else if (MODE == M_MADWIDE_ADD64) {
#define OP(lo, hi) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
      OP(x0, y0) OP(x1, y1) OP(x2, y2) OP(x3, y3) OP(x4, y4) OP(x5, y5) OP(x6, y6) OP(x7, y7)
#undef OP
    } else if (MODE == M_MADWIDE_CHAIN4) {
      // two chains of 4 pairs = 8 ops (one op = one lo/hi pair = one 32x32+64 MAC)
      asm volatile(
          "mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %8, %9, %2; madc.hi.cc.u32 %3, %8, %9, %3;"
          "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5; madc.lo.cc.u32 %6, %8, %9, %6; madc.hi.u32 %7, %8, %9, %7;"
          : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(a), "r"(b));
      asm volatile(
          "mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %8, %9, %2; madc.hi.cc.u32 %3, %8, %9, %3;"
          "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5; madc.lo.cc.u32 %6, %8, %9, %6; madc.hi.u32 %7, %8, %9, %7;"
          : "+r"(y0), "+r"(y1), "+r"(y2), "+r"(y3), "+r"(y4), "+r"(y5), "+r"(y6), "+r"(y7) : "r"(b), "r"(a));
    } else if (MODE == M_MADWIDE_CHAIN12) {
      // one chain of 8 pairs over x0..x7,y0..y7 (16 limbs): 8 ops
      asm volatile(
          "mad.lo.cc.u32 %0, %16, %17, %0; madc.hi.cc.u32 %1, %16, %17, %1; madc.lo.cc.u32 %2, %16, %17, %2; madc.hi.cc.u32 %3, %16, %17, %3;"
          "madc.lo.cc.u32 %4, %16, %17, %4; madc.hi.cc.u32 %5, %16, %17, %5; madc.lo.cc.u32 %6, %16, %17, %6; madc.hi.cc.u32 %7, %16, %17, %7;"
          "madc.lo.cc.u32 %8, %16, %17, %8; madc.hi.cc.u32 %9, %16, %17, %9; madc.lo.cc.u32 %10, %16, %17, %10; madc.hi.cc.u32 %11, %16, %17, %11;"
          "madc.lo.cc.u32 %12, %16, %17, %12; madc.hi.cc.u32 %13, %16, %17, %13; madc.lo.cc.u32 %14, %16, %17, %14; madc.hi.u32 %15, %16, %17, %15;"
          : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7),
            "+r"(y0), "+r"(y1), "+r"(y2), "+r"(y3), "+r"(y4), "+r"(y5), "+r"(y6), "+r"(y7) : "r"(a), "r"(b));
    } else if (MODE == M_MULWIDE_IADD) {
#define OP(x, k) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x) : "r"(a + k), "r"((uint32_t)x));
      OP(w0, 0) OP(w1, 1) OP(w2, 2) OP(w3, 3) OP(w4, 4) OP(w5, 5) OP(w6, 6) OP(w7, 7)
#undef OP
    } else if (MODE == M_IADD3) {
#define OP(x) asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x) : "r"(a), "r"(b));
      OP(x0) OP(x1) OP(x2) OP(x3) OP(x4) OP(x5) OP(x6) OP(x7)
#undef OP
    } else if (MODE == M_IADD_CHAIN8) {
      asm volatile(
          "add.cc.u32 %0, %0, %8; addc.cc.u32 %1, %1, %9; addc.cc.u32 %2, %2, %8; addc.cc.u32 %3, %3, %9;"
          "addc.cc.u32 %4, %4, %8; addc.cc.u32 %5, %5, %9; addc.cc.u32 %6, %6, %8; addc.u32 %7, %7, %9;"
          : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(a), "r"(b));
    } else if (MODE == M_MULWIDE_IADD_CC) {
      // 8 products (a+k)*lo(acc_k); each is added into a 3-word column accumulator (x,y,z) with add.cc/addc.cc/addc:
      // the shape of a carry-deferred product-scanning multiplier (FMA pipe: 1 IMAD.WIDE per op; ALU pipe: 3 adds per op)
#define OP(x, y, z, k) asm volatile("{ .reg .u64 t; .reg .u32 tl, th; mul.wide.u32 t, %3, %0; mov.b64 {tl, th}, t;" \
                                    " add.cc.u32 %0, %0, tl; addc.cc.u32 %1, %1, th; addc.u32 %2, %2, 0; }" \
                                    : "+r"(x), "+r"(y), "+r"(z) : "r"(a + k));
      OP(x0, y0, z0, 0) OP(x1, y1, z1, 1) OP(x2, y2, z2, 2) OP(x3, y3, z3, 3) OP(x4, y4, z4, 4) OP(x5, y5, z5, 5) OP(x6, y6, z6, 6) OP(x7, y7, z7, 7)
#undef OP
    } else if (MODE == M_DFMA || MODE == M_WIDE_HALF || MODE == M_DFMA_HALF || MODE == M_MIX_WIDE_DFMA) {
      const bool do_wide = (MODE == M_WIDE_HALF || MODE == M_MIX_WIDE_DFMA) && !odd_warp;
      const bool do_dfma = MODE == M_DFMA || ((MODE == M_DFMA_HALF || MODE == M_MIX_WIDE_DFMA) && odd_warp);
      if (do_wide) {
        asm volatile(
            "mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %8, %9, %2; madc.hi.cc.u32 %3, %8, %9, %3;"
            "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5; madc.lo.cc.u32 %6, %8, %9, %6; madc.hi.u32 %7, %8, %9, %7;"
            : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4), "+r"(x5), "+r"(x6), "+r"(x7) : "r"(a), "r"(b));
        asm volatile(
            "mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %8, %9, %2; madc.hi.cc.u32 %3, %8, %9, %3;"
            "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5; madc.lo.cc.u32 %6, %8, %9, %6; madc.hi.u32 %7, %8, %9, %7;"
            : "+r"(y0), "+r"(y1), "+r"(y2), "+r"(y3), "+r"(y4), "+r"(y5), "+r"(y6), "+r"(y7) : "r"(b), "r"(a));
      }
      if (do_dfma) {
#define OP(x) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(fa), "d"(fb));
        OP(f0) OP(f1) OP(f2) OP(f3) OP(f4) OP(f5) OP(f6) OP(f7)
#undef OP
      }
    } else if (MODE == M_LOP3) {
#define OP(x) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(a), "r"(b));
      OP(x0) OP(x1) OP(x2) OP(x3) OP(x4) OP(x5) OP(x6) OP(x7)
#undef OP
    } else if (MODE == M_SHF) {
#define OP(x) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
      OP(x0) OP(x1) OP(x2) OP(x3) OP(x4) OP(x5) OP(x6) OP(x7)
#undef OP
    }
  }
  const long long t1 = clock64();
  uint32_t r = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7 ^ y0 ^ y1 ^ y2 ^ y3 ^ y4 ^ y5 ^ y6 ^ y7 ^ z0 ^ z1 ^ z2 ^ z3 ^ z4 ^ z5 ^ z6 ^ z7 ^
               (uint32_t)(w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7) ^ (uint32_t)((w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7) >> 32);
  if (r == 0x12345678u || f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7 == 0.123) out[0] = r;
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

static const char* kNames[M_COUNT] = {
    "IMAD (mad.lo.u32)", "IMAD.HI.U32 (mad.hi.u32)", "IMAD.WIDE.U32 RZ addend (mul.wide.u32)",
    "mad.lo.cc+madc.hi single pair (64-bit addend, no carry)", "mad.lo.cc/madc.hi.cc chains of 4 pairs", "chain of 8 pairs",
    "mad.wide.u32 + 64-bit addend as ptxas splits it (IMAD.WIDE RZ + IADD3)", "two dependent add.u32 (IADD3 fusion?) -- op = 2 adds",
    "add.cc chain of 8 (IADD3.X) -- op = 1 add", "mul.wide + 3-word add.cc accumulate -- op = 1 product",
    "LOP3.LUT", "SHF.R.W", "DFMA (fma.rn.f64), all warps", "IMAD.WIDE chains on even warps only (odd idle) -- op = 1 MAC per even-warp thread-op",
    "DFMA on odd warps only (even idle)", "even warps IMAD.WIDE chains + odd warps DFMA side by side"};

template <int MODE>
void run(int sms, int blocks_per_sm, int threads, int iters = 20000) {
  const int blocks = sms * blocks_per_sm;
  uint32_t* d; unsigned long long* cyc;
  CK(cudaMalloc(&d, 4)); CK(cudaMalloc(&cyc, 8 * blocks));
  k_pipe<MODE><<<blocks, threads>>>(d, cyc, 100, 1);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_pipe<MODE><<<blocks, threads>>>(d, cyc, iters, 7);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long* h = (unsigned long long*)malloc(8 * blocks);
  CK(cudaMemcpy(h, cyc, 8 * blocks, cudaMemcpyDeviceToHost));
  double mean = 0; for (int i = 0; i < blocks; i++) mean += (double)h[i]; mean /= blocks;
  const double ops_per_sm = (double)blocks_per_sm * threads * iters * 8.0;
  // ops_per_clk_per_sm / sm_clock_ghz_seen are only meaningful when every block is resident at once (<= 32 registers at 64 warps/SM);
  // Tops_per_s is time-based and always valid
  printf("{\"bench\":\"pipe2\",\"op\":\"%s\",\"warps_per_sm\":%d,\"iters\":%d,\"ms\":%.4f,\"Tops_per_s\":%.3f,\"ops_per_clk_per_sm_at_1965MHz\":%.2f,\"sm_cycles\":%.0f,\"ops_per_clk_per_sm\":%.2f,\"sm_clock_ghz_seen\":%.3f}\n",
         kNames[MODE], blocks_per_sm * threads / 32, iters, ms, ops_per_sm * sms / (ms * 1e-3) * 1e-12, ops_per_sm / (ms * 1e-3 * 1.965e9), mean, ops_per_sm / mean,
         mean / (ms * 1e-3) * 1e-9);
  fflush(stdout);
  free(h); cudaFree(d); cudaFree(cyc);
}

template <int MODE>
void both(int sms) { run<MODE>(sms, 8, 256); run<MODE>(sms, 2, 128); }

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("{\"device\":\"%s\",\"sms\":%d}\n", prop.name, sms);
  both<M_IMAD_LO>(sms); both<M_IMAD_HI>(sms); both<M_MULWIDE>(sms); both<M_MADWIDE_ADD64>(sms); both<M_MADWIDE_CHAIN4>(sms);
  both<M_MADWIDE_CHAIN12>(sms); both<M_MULWIDE_IADD>(sms); both<M_IADD3>(sms); both<M_IADD_CHAIN8>(sms); both<M_MULWIDE_IADD_CC>(sms);
  both<M_LOP3>(sms); both<M_SHF>(sms);
  both<M_DFMA>(sms); both<M_WIDE_HALF>(sms); both<M_DFMA_HALF>(sms); both<M_MIX_WIDE_DFMA>(sms);
  // occupancy sweep of the multiplier-row shape (chains of 4 pairs), long enough (~0.1-0.3 s each) for nvidia-smi to sample clocks and
  // power beside it (tools/gpu_r2_d.sh): is the sustained rate set by the pipe or by the power cap?
  run<M_MADWIDE_CHAIN4>(sms, 1, 128, 400000); run<M_MADWIDE_CHAIN4>(sms, 2, 128, 400000); run<M_MADWIDE_CHAIN4>(sms, 4, 128, 400000);
  run<M_MADWIDE_CHAIN4>(sms, 4, 256, 400000); run<M_MADWIDE_CHAIN4>(sms, 6, 256, 400000);
  run<M_IMAD_LO>(sms, 2, 128, 800000); run<M_IMAD_LO>(sms, 4, 256, 800000); run<M_IMAD_LO>(sms, 8, 256, 800000);
  run<M_DFMA>(sms, 2, 128, 400000); run<M_DFMA>(sms, 4, 256, 400000);
  run<M_MIX_WIDE_DFMA>(sms, 2, 128, 400000); run<M_MIX_WIDE_DFMA>(sms, 4, 256, 400000);
  return 0;
}
