// Explicit instantiation of the engine for one curve (own translation unit so the curves compile in parallel).
#include "msm_hooks.cuh"
namespace b200 { B200_INSTANTIATE_CURVE(PallasEc) }
