// Windowed one-CTA y scan (mp_sweep_wpc.cuh), coords geometry; separate TU so the geometries compile in parallel.
#include "../../include/libmidaspom_cuda.h"
#define MP_WPC_GEOM MP_GEOM_COORDS
#include "mp_sweep_wpc.cuh"
int mp_launch_sweep_wpc_coords(mp_engine *h, int window, int nl_max, int nclusters, const void *btasks) { return mp::launch_wpc_any(h, window, nl_max, nclusters, btasks); }
size_t mp_wpc_smem_bytes(int nl_max) { return mp::wpc_smem_bytes(nl_max); }
