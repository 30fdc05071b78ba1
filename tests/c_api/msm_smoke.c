/* C caller of the drop-in boundary: includes include/ctt_b200_msm.h exactly as a Constantine C user would include
 * constantine/curves/bls12_381_parallel.h, and calls the reference-named symbol.
 * Input: EIP-2537 vector "bls_g1multiexp_(g1+g1=2*g1)" (reference tests/protocol_ethereum_evm_precompiles/eip-2537/
 * multiexp_G1_bls.json) = generator with scalar 2; the program prints the Jacobian result limbs, the pytest wrapper
 * normalises and compares with the vector's expected point.
 * Build: gcc -std=c99 -I include tests/c_api/msm_smoke.c -L constantine_b200/lib -lctt_b200_msm -Wl,-rpath,... */
#include <stdio.h>
#include <string.h>
#include "ctt_b200_msm.h"

/* BLS12-381 G1 generator, Montgomery residues (R = 2^384), little-endian 64-bit limbs */
static const bls12_381_g1_aff G1_GEN = {
  {{0x5cb38790fd530c16ull, 0x7817fc679976fff5ull, 0x154f95c7143ba1c1ull, 0xf0ae6acdf3d0e747ull, 0xedce6ecc21dbf440ull, 0x120177419e0bfb75ull}},
  {{0xbaac93d50ce72271ull, 0x8c22631a7918fd8eull, 0xdd595f13570725ceull, 0x51ac582950405194ull, 0x0e1c8c3fad0059c0ull, 0x0bbc3efc5008a26aull}}};

int main(int argc, char** argv) {
  int symbols_only = argc > 1 && strcmp(argv[1], "--link-only") == 0;
  ctt_threadpool* tp = ctt_threadpool_new(2);
  if (!tp) return 2;
  if (symbols_only) {   /* no GPU on this machine: prove that the program links and the handle API works */
    printf("linked ok, %d host threads\n", ctt_cpu_get_num_threads_os());
    ctt_threadpool_shutdown(tp);
    return 0;
  }
  big255 coefs[1];
  memset(coefs, 0, sizeof(coefs));
  coefs[0].limbs[0] = 2;
  bls12_381_g1_aff points[1];
  points[0] = G1_GEN;
  bls12_381_g1_jac r;
  ctt_bls12_381_g1_jac_multi_scalar_mul_big_coefs_vartime_parallel(tp, &r, coefs, points, 1);
  const secret_word* w = (const secret_word*)&r;
  for (int i = 0; i < 18; i++) printf("%016llx%c", (unsigned long long)w[i], (i % 6 == 5) ? '\n' : ' ');
  ctt_threadpool_shutdown(tp);
  return 0;
}
