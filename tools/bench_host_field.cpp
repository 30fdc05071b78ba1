// Host tail arithmetic (constantine_b200/csrc/host_field.hpp): self-check of the MULX/ADX multiplication against the portable form
// (random and edge operands, every coordinate field) and ns per multiplication of both.
//   g++ -O3 -D__host__= -D__device__= -I constantine_b200/csrc tools/bench_host_field.cpp -o tools/bin/bench_host_field
// Exit code 0 = every product agreed (or the CPU has no BMI2 + ADX: nothing to compare). tests/test_host_logic.py runs it.
#include <chrono>
#include <cstdio>
#include "host_field.hpp"
using namespace b200;
using namespace b200::host;

static uint64_t rng_state = 0x243F6A8885A308D3ull;
static uint64_t next64() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

template <class F> static HFp<F> canonical(int kind) {
  typedef HFp<F> Fp;
  Fp a;
  const int N = F::N64;
  switch (kind) {
    case 0: return Fp::zero();
    case 1: return Fp::one();
    case 2: for (int i = 0; i < N; i++) a.l[i] = F::P64(i); a.l[0] -= 1; return a;                 // p - 1
    case 3: a = Fp::zero(); a.l[0] = 1; return a;                                                    // raw 1
    case 4: for (int i = 0; i < N; i++) a.l[i] = F::P64(i); a.l[0] -= 2; return a;                 // p - 2
    case 5: a = Fp::zero(); a.l[N - 1] = F::P64(N - 1) - 1; for (int i = 0; i < N - 1; i++) a.l[i] = ~0ull; return a;   // all-ones low limbs
    default:
      for (int i = 0; i < N; i++) a.l[i] = next64();
      if ((kind & 3) == 0) a.l[next64() % N] = ~0ull;
      if ((kind & 3) == 1) a.l[next64() % N] = 0;
      a.l[N - 1] %= F::P64(N - 1);                                                                   // top limb below the modulus' top limb: canonical
      return a;
  }
}

template <class F> static int run(const char* name) {
  typedef HFp<F> Fp;
  int bad = 0;
#if defined(B200_HOST_MULX_ADX)
  if (cpu_has_mulx_adx()) {
    const int CASES = 400000;
    for (int i = 0; i < CASES && bad < 5; i++) {
      Fp a = canonical<F>(i % 23), b = canonical<F>((i / 23) % 29);
      Fp x = a.mul_mulx_adx(b), y = a.mul_portable(b);
      if (!(x == y)) { bad++; printf("%s: MISMATCH at case %d\n", name, i); }
      Fp s1 = a.mul_mulx_adx(a), s2 = a.mul_portable(a);
      if (!(s1 == s2)) { bad++; printf("%s: MISMATCH (square) at case %d\n", name, i); }
    }
    // a long dependent chain: values that are products of products
    Fp a = canonical<F>(100), b = canonical<F>(101), c = a, d = b;
    for (int i = 0; i < 200000; i++) { a = a.mul_mulx_adx(b); b = b.mul_mulx_adx(a); c = c.mul_portable(d); d = d.mul_portable(c); }
    if (!(a == c) || !(b == d)) { bad++; printf("%s: MISMATCH in the chained products\n", name); }
  } else {
    printf("%s: no BMI2 + ADX on this CPU, portable multiplication only\n", name);
  }
#endif
  const Fp a0 = canonical<F>(7), b0 = canonical<F>(8);
  Fp a = a0, b = b0;
  const int R = 1000000;
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < R; i++) { a = a * b; b = b * a; }
  auto t1 = std::chrono::steady_clock::now();
  Fp c = a0, d = b0;
  for (int i = 0; i < R; i++) { c = c.mul_portable(d); d = d.mul_portable(c); }
  auto t2 = std::chrono::steady_clock::now();
  if (!(a == c) || !(b == d)) { bad++; printf("%s: MISMATCH between operator* and the portable chain\n", name); }
  printf("%s: operator* %.1f ns, portable %.1f ns per multiplication, %s\n", name, std::chrono::duration<double, std::nano>(t1 - t0).count() / (2.0 * R),
         std::chrono::duration<double, std::nano>(t2 - t1).count() / (2.0 * R), bad ? "FAILED" : "ok");
  return bad;
}

int main() {
  int bad = 0;
  bad += run<Bls12381Fp>("bls12_381 fp (6 limbs)");
  bad += run<Bn254SnarksFp>("bn254_snarks fp (4 limbs)");
  bad += run<PallasFp>("pallas fp (4 limbs)");
  bad += run<VestaFp>("vesta fp (4 limbs)");
  return bad ? 1 : 0;
}
