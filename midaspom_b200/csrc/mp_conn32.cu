// mp_conn32.cu -- k_conn with the year contraction on the FP32 pipe (A32 variant of mp_conn.cuh): the connectivity kernel of
// the FP32 engines.  A separate translation unit so that its instantiations compile beside those of mp_engine.cu.
#include "mp_host.h"
#include "mp_conn.cuh"

using namespace mp;

template <int GEOM>
static int launch_conn32_g(mp_engine *h, const ConnArgs<float> &a, int ny, int shape, bool cull, dim3 grid)
{
    constexpr bool CAN_CULL = GEOM != MP_GEOM_DENSE;
#define MP_CONN_K(NYB, NT, CULLED) do { auto kern = k_conn<float, GEOM, NYB, CULLED, 2, NT, true>;                                  \
                                        const size_t smem = conn_a32_smem(NYB, NT);                                              \
                                        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
                                        kern<<<grid, NT, smem, h->stream>>>(a); } while (0)
#define MP_CONN_T(NYB, NT) do { if (cull && CAN_CULL) MP_CONN_K(NYB, NT, CAN_CULL); else MP_CONN_K(NYB, NT, false); } while (0)
#define MP_CONN(NYB) do { if (shape == 1) MP_CONN_T(NYB, 128); else if (shape == 2) MP_CONN_T(NYB, 64); else MP_CONN_T(NYB, 32); } while (0)
    if (ny <= 4) MP_CONN(4); else if (ny <= 8) MP_CONN(8); else if (ny <= 12) MP_CONN(12); else if (ny <= 16) MP_CONN(16);
    else if (ny <= 20) MP_CONN(20); else if (ny <= 24) MP_CONN(24); else if (ny <= 28) MP_CONN(28); else MP_CONN(32);
#undef MP_CONN
#undef MP_CONN_T
#undef MP_CONN_K
    CK(cudaGetLastError());
    return MP_OK;
}

// args: the ConnArgs<float> of this launch (built by launch_conn_g in mp_engine.cu, which also chose the CTA shape and the grid)
int mp_launch_conn32(mp_engine *h, const void *args, int geom, int ny, int shape, int cull, unsigned gx, unsigned gy, unsigned gz)
{
    const ConnArgs<float> &a = *(const ConnArgs<float> *)args;
    const dim3 grid(gx, gy, gz);
    switch (geom) {
    case MP_GEOM_LINEAR: return launch_conn32_g<MP_GEOM_LINEAR>(h, a, ny, shape, cull != 0, grid);
    case MP_GEOM_COORDS: return launch_conn32_g<MP_GEOM_COORDS>(h, a, ny, shape, cull != 0, grid);
    default: return launch_conn32_g<MP_GEOM_DENSE>(h, a, ny, shape, cull != 0, grid);
    }
}
