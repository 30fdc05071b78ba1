/* C caller of the UNCHANGED reference symbol with several GPUs behind it (ctt_b200_set_devices / CTT_B200_DEVICES): N copies of
 * the generator with scalars 1..N, so the result is [N (N + 1) / 2] G -- the program prints the Jacobian result limbs once for
 * the single-device call and once with the device list ("all" devices, or the one device listed twice when there is only one),
 * the pytest wrapper normalises both and compares them with the closed form.
 * Build: gcc -std=c99 -I include tests/c_api/msm_multi_gpu.c -L constantine_b200/lib -lctt_b200_msm -Wl,-rpath,... */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ctt_b200_msm.h"

static const bls12_381_g1_aff G1_GEN = {
  {{0x5cb38790fd530c16ull, 0x7817fc679976fff5ull, 0x154f95c7143ba1c1ull, 0xf0ae6acdf3d0e747ull, 0xedce6ecc21dbf440ull, 0x120177419e0bfb75ull}},
  {{0xbaac93d50ce72271ull, 0x8c22631a7918fd8eull, 0xdd595f13570725ceull, 0x51ac582950405194ull, 0x0e1c8c3fad0059c0ull, 0x0bbc3efc5008a26aull}}};

static void print_jac(const bls12_381_g1_jac* r) {
  const secret_word* w = (const secret_word*)r;
  for (int i = 0; i < 18; i++) printf("%016llx%c", (unsigned long long)w[i], (i % 6 == 5) ? '\n' : ' ');
}

int main(int argc, char** argv) {
  const size_t n = argc > 1 ? (size_t)strtoull(argv[1], NULL, 10) : 100000;
  ctt_threadpool* tp = ctt_threadpool_new(4);
  big255* coefs = (big255*)calloc(n, sizeof(big255));          /* ordinary heap memory, as a Constantine caller passes */
  bls12_381_g1_aff* points = (bls12_381_g1_aff*)malloc(n * sizeof(bls12_381_g1_aff));
  if (!tp || !coefs || !points) return 2;
  for (size_t i = 0; i < n; i++) { coefs[i].limbs[0] = i + 1; points[i] = G1_GEN; }
  bls12_381_g1_jac r;
  ctt_bls12_381_g1_jac_multi_scalar_mul_big_coefs_vartime_parallel(tp, &r, coefs, points, n);
  print_jac(&r);
  int ids[8], count = ctt_b200_device_count();
  if (count > 8) count = 8;
  for (int i = 0; i < count; i++) ids[i] = i;
  if (count == 1) { ids[1] = 0; count = 2; }
  if (ctt_b200_set_devices(ids, count) != 0) return 3;
  ctt_bls12_381_g1_jac_multi_scalar_mul_big_coefs_vartime_parallel(tp, &r, coefs, points, n);
  print_jac(&r);
  printf("devices %d\n", count);
  ctt_threadpool_shutdown(tp);
  free(coefs); free(points);
  return 0;
}
