// Windowed one-CTA y scan (mp_sweep_wpc.cuh), linear geometry; separate TU so the geometries compile in parallel.
#include "../../include/libmidaspom_cuda.h"
#define MP_WPC_GEOM MP_GEOM_LINEAR
#include "mp_sweep_wpc.cuh"
int mp_launch_sweep_wpc_linear(mp_engine *h, int window, int nl_max, int nclusters, const void *btasks) { return mp::launch_wpc_any(h, window, nl_max, nclusters, btasks); }
