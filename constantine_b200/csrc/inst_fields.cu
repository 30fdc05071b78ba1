// Explicit instantiation of the element-wise field test hook for the eight prime fields.
#include "msm_hooks.cuh"
namespace b200 {
B200_INSTANTIATE_FIELD(Bls12381Fp) B200_INSTANTIATE_FIELD(Bn254SnarksFp) B200_INSTANTIATE_FIELD(PallasFp) B200_INSTANTIATE_FIELD(VestaFp)
B200_INSTANTIATE_FIELD(Bls12381Fr) B200_INSTANTIATE_FIELD(Bn254SnarksFr) B200_INSTANTIATE_FIELD(PallasFr) B200_INSTANTIATE_FIELD(VestaFr)
}
