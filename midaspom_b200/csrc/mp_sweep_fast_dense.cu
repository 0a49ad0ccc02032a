// FP32 fast y sweep, dense geometry (see mp_sweep_fast.cuh); separate TU so the geometries compile in parallel.
#include "../../include/libmidaspom_cuda.h"
#define MP_FAST_GEOM MP_GEOM_DENSE
#include "mp_sweep_fast.cuh"
int mp_launch_sweep_fast_dense(mp_engine *h, int cs, int tpt) { return mp::launch_fast_any(h, cs, tpt); }
