// mp_conn_gemm.cu -- tensor-core connectivity for a batch of chains that share (alpha, b).
//
// When every chain of the engine has the same dispersal parameters the connectivity really is a dense contraction
// (north star (1); the matrix form `c*M%*%pti` of Rscript/simuls_traj.R:16,203,214 == main_MIDASPOM.c:350-358):
//
//     S[col][k] = sum_l W[k][l] * Y[col][l],    W[k][l] = A_l^b exp(-alpha d_kl) (0 on the diagonal),  col = (chain, year)
//
// with ONE kernel matrix W for all columns.  W is never stored: each CTA owns 128 target patches and a range of
// sources; its 8 producer warps evaluate 128 x 64 weights per pipeline stage with exactly the FP32 expression of
// every other kernel (weight_of), split each weight into three BF16 parts w = w1 + w2 + w3 (8 + 8 + 8 mantissa bits:
// the split is exact) and store them as three K-major UMMA operand tiles in shared memory.  The 0/1 occupancy columns
// are BF16 tiles fetched by TMA (cp.async.bulk.tensor, 128-byte swizzle).  One elected thread issues tcgen05.mma
// (M = 128, N = the column block, K = 16) with FP32 accumulators in tensor memory.  Three accumulators keep the
// truncation of the FP32 accumulation away from the result: the leading parts w1 alternate between two of them by
// stage, the two trailing parts (<= 2^-8 of w) share the third.  The epilogue reads them with tcgen05.ld, adds them in
// FP64 and accumulates into S (FP64 atomics: the source range of a target tile is split over several CTAs so that the
// 148 SMs are busy).
//
// Accuracy: products are exact (BF16 x {0,1}); what remains is the FP32 accumulation of the leading parts, ~1e-6
// relative (measured against the oracle in tests/test_gpu_gemm.py at 1e-5).  This is the batch-evaluation path behind
// mp_connectivity / mp_loglik; the sampler keeps k_conn, whose FP64 accumulation lets rank-1 removals cancel exactly.
#include <cuda.h>

#include "mp_host.h"

namespace mp {

constexpr int GM_BLOCK_M = 128, GM_BLOCK_K = 64, GM_STAGES = 3, GM_UMMA_K = 16;
constexpr int GM_PRODUCERS = 256;                  // 8 warps: weights in, accumulators out
constexpr int GM_THREADS = GM_PRODUCERS + 32;      // + one warp: TMA of the occupancy tiles, MMA issue, TMEM allocation
constexpr int GM_A_SPLIT_BYTES = GM_BLOCK_M * GM_BLOCK_K * 2;      // one BF16 part of a weight tile: 16 KB
constexpr int GM_A_STAGE_BYTES = 3 * GM_A_SPLIT_BYTES;

struct GemmArgs {
    const float4 *src;      // [Kpad] {x, y, log2 A^b, -} of every source patch (patch order; zero records beyond n)
    double *S;              // [ncols][n], accumulated into (zeroed by the caller)
    int n, ncols, nstage_total, stages_per_cta;
    float apre;             // -alpha log2(e)
    float spacing;          // linear landscapes
    int linear;
    unsigned long long *stats;
};

// ---- PTX wrappers (sm_100a)
__device__ __forceinline__ uint32_t gm_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gm_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void gm_mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void gm_mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// A wait that cannot hang the GPU: a phase that has not completed after ~4 s of polling is a protocol bug -> trap.
__device__ __forceinline__ void gm_mbar_wait(uint32_t bar, uint32_t parity)
{
    const long long t0 = clock64();
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity), "r"(10000u) : "memory");
        if (done) return;
        if (clock64() - t0 > 8000000000LL) __trap();
    }
}
__device__ __forceinline__ void gm_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void gm_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void gm_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void gm_tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void gm_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void gm_commit(uint32_t bar)
{   // arrives on the mbarrier when every tcgen05.mma issued so far by this thread has completed (implies fence::before_thread_sync)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void gm_tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void gm_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, leading / stride byte offsets in
// 16-byte units, descriptor version 1 (Blackwell), layout type (0 = no swizzle, 2 = 128-byte swizzle)
__device__ __forceinline__ uint64_t gm_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout)
{
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46) | ((uint64_t)layout << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = BF16, both K-major, dense, M = 128, N = NPAD
__host__ __device__ constexpr uint32_t gm_idesc(int npad)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(npad >> 3) << 17) | ((uint32_t)(GM_BLOCK_M >> 4) << 24);
}

// Dynamic shared memory of one CTA (1024-byte aligned: the 128-byte swizzle of the TMA tiles needs it)
template <int NPAD> struct GemmSmem {
    static constexpr int B_STAGE_BYTES = NPAD * GM_BLOCK_K * 2;      // NPAD rows of 128 bytes
    static constexpr int A_OFF = 0;
    static constexpr int B_OFF = GM_STAGES * GM_A_STAGE_BYTES;
    static constexpr int BAR_OFF = B_OFF + GM_STAGES * B_STAGE_BYTES;
    static constexpr int BYTES = BAR_OFF + 256;
};

// TMEM columns of the three accumulators, rounded to the power of two tcgen05.alloc wants
__host__ __device__ constexpr int gm_tmem_cols(int npad) { return 3 * npad <= 128 ? 128 : 3 * npad <= 256 ? 256 : 512; }

template <int NPAD>
__global__ void __launch_bounds__(GM_THREADS, 1)
k_conn_gemm(const __grid_constant__ CUtensorMap ymap, GemmArgs a)
{
    using SM = GemmSmem<NPAD>;
    static_assert(NPAD % 32 == 0 && NPAD >= 32 && 3 * NPAD <= 512, "column block: a multiple of 32, three accumulators in 512 TMEM columns");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (gm_smem(smem_raw) & 1023u)) & 1023u);   // the launch adds 1 KB of slack for this
    const uint32_t sbase = gm_smem(smem);
    const uint32_t bars = sbase + SM::BAR_OFF;                       // full_a[3], full_b[3], empty[3], accum, then the TMEM address word
    auto full_a = [&](int s) { return bars + 8u * s; };
    auto full_b = [&](int s) { return bars + 8u * (GM_STAGES + s); };
    auto empty = [&](int s) { return bars + 8u * (2 * GM_STAGES + s); };
    const uint32_t accum_bar = bars + 8u * (3 * GM_STAGES);
    volatile uint32_t *tmem_word = reinterpret_cast<volatile uint32_t *>(smem + SM::BAR_OFF + 8 * (3 * GM_STAGES + 1));
    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * GM_BLOCK_M;
    const int st0 = blockIdx.y * a.stages_per_cta, st1 = min(a.nstage_total, st0 + a.stages_per_cta);
    const int col0 = blockIdx.z * NPAD;
    const int nst = st1 - st0;
    if (nst <= 0) return;

    if (tid == 0) {
        for (int s = 0; s < GM_STAGES; s++) { gm_mbar_init(full_a(s), GM_PRODUCERS); gm_mbar_init(full_b(s), 1); gm_mbar_init(empty(s), 1); }
        gm_mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == GM_PRODUCERS / 32) {                                  // the MMA warp owns the tensor memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(gm_smem((const void *)tmem_word)), "r"((uint32_t)gm_tmem_cols(NPAD)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    gm_tc_fence_before();
    __syncthreads();
    gm_tc_fence_after();
    const uint32_t tmem = *tmem_word;

    if (wid < GM_PRODUCERS / 32) {
        // ===== producers: 128 x 64 weights per stage, split into three BF16 operand tiles
        const int m = tid & (GM_BLOCK_M - 1), khalf = tid >> 7;       // row of the tile, which 32 of the 64 sources
        const int k = min(m0 + m, a.n - 1);                           // target patch (rows beyond n repeat the last one; masked in the epilogue)
        const float4 tk = a.src[k];
        for (int i = 0; i < nst; i++) {
            const int s = i % GM_STAGES, use = i / GM_STAGES;
            if (use > 0) gm_mbar_wait(empty(s), (uint32_t)(use - 1) & 1u);
            const int l0 = (st0 + i) * GM_BLOCK_K + khalf * 32;
            const uint32_t abase = sbase + SM::A_OFF + s * GM_A_STAGE_BYTES + (uint32_t)(m >> 3) * 128u + (uint32_t)(m & 7) * 16u;
#pragma unroll
            for (int c8 = 0; c8 < 4; c8++) {                          // 8 sources = one 16-byte K chunk of every part
                uint32_t p1[4], p2[4], p3[4];
#pragma unroll
                for (int u = 0; u < 8; u += 2) {
                    float w[2];
#pragma unroll
                    for (int v = 0; v < 2; v++) {
                        const int l = l0 + c8 * 8 + u + v;
                        const float4 sl = __ldg(&a.src[l]);
                        float d;
                        if (a.linear) d = (float)(k > l ? k - l : l - k) * a.spacing;
                        else { const float dx = tk.x - sl.x, dy = tk.y - sl.y; d = Num<float>::sqrtv(fmaf(dx, dx, dy * dy)); }
                        w[v] = l == k ? 0.f : weight_of(a.apre, sl.z, d);          // the expression every FP32 kernel uses
                    }
                    // exact three-way split: w = h1 + h2 + h3, each with 8 significant bits (the top 16 bits of an FP32 word are a BF16)
                    uint32_t q1[2], q2[2], q3[2];
#pragma unroll
                    for (int v = 0; v < 2; v++) {
                        const uint32_t b1 = __float_as_uint(w[v]) & 0xFFFF0000u;
                        const float r1 = w[v] - __uint_as_float(b1);
                        const uint32_t b2 = __float_as_uint(r1) & 0xFFFF0000u;
                        const float r2 = r1 - __uint_as_float(b2);
                        q1[v] = b1; q2[v] = b2; q3[v] = __float_as_uint(r2) & 0xFFFF0000u;
                    }
                    p1[u >> 1] = (q1[0] >> 16) | q1[1]; p2[u >> 1] = (q2[0] >> 16) | q2[1]; p3[u >> 1] = (q3[0] >> 16) | q3[1];
                }
                const uint32_t off = (uint32_t)(khalf * 4 + c8) * 2048u;    // K chunk: 16 row groups x 128 bytes apart
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(abase + off), "r"(p1[0]), "r"(p1[1]), "r"(p1[2]), "r"(p1[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(abase + off + GM_A_SPLIT_BYTES), "r"(p2[0]), "r"(p2[1]), "r"(p2[2]), "r"(p2[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(abase + off + 2 * GM_A_SPLIT_BYTES), "r"(p3[0]), "r"(p3[1]), "r"(p3[2]), "r"(p3[3]) : "memory");
            }
            gm_fence_async_smem();                                    // generic-proxy stores -> visible to the tensor core's async proxy
            gm_mbar_arrive(full_a(s));
        }
        // ===== epilogue: the three accumulators -> FP64 -> S
        gm_mbar_wait(accum_bar, 0u);
        gm_tc_fence_after();
        const int q = wid & 3, half = wid >> 2;                       // TMEM lane quarter of this warp, which half of the columns
        const int row = q * 32 + lane, kt = m0 + row;
        constexpr int HALF = NPAD / 2;
#pragma unroll 1
        for (int cb = 0; cb < HALF; cb += 16) {
            const int c = half * HALF + cb;
            uint32_t v0[16], v1[16], v2[16];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
            gm_tmem_ld16(taddr, v0); gm_tmem_ld16(taddr + NPAD, v1); gm_tmem_ld16(taddr + 2 * NPAD, v2);
            gm_tmem_ld_wait();
            if (kt < a.n) {
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const int col = col0 + c + j;
                    if (col < a.ncols) {
                        const double lead2 = nst >= 2 ? (double)__uint_as_float(v1[j]) : 0.0;     // one stage only: the second accumulator was never written
                        const double v = ((double)__uint_as_float(v0[j]) + lead2) + (double)__uint_as_float(v2[j]);
                        atomicAdd(&a.S[(size_t)col * a.n + kt], v);
                    }
                }
            }
        }
        gm_tc_fence_before();
    } else if (lane == 0) {
        // ===== one thread: TMA of the occupancy tiles, MMA issue
        constexpr uint32_t IDESC = gm_idesc(NPAD);
        // the TMA runs GM_STAGES - 1 tiles ahead of the tensor core
        int loaded = 0;
        auto load_b = [&](int i) {
            const int s = i % GM_STAGES, use = i / GM_STAGES;
            if (use > 0) gm_mbar_wait(empty(s), (uint32_t)(use - 1) & 1u);
            gm_mbar_expect_tx(full_b(s), (uint32_t)SM::B_STAGE_BYTES);
            gm_tma_load_2d(sbase + SM::B_OFF + s * SM::B_STAGE_BYTES, &ymap, (st0 + i) * GM_BLOCK_K, col0, full_b(s));
        };
        for (; loaded < min(nst, GM_STAGES - 1); loaded++) load_b(loaded);
        for (int i = 0; i < nst; i++) {
            if (loaded < nst) load_b(loaded++);
            const int s = i % GM_STAGES, use = i / GM_STAGES;
            gm_mbar_wait(full_a(s), (uint32_t)use & 1u);
            gm_mbar_wait(full_b(s), (uint32_t)use & 1u);
            gm_tc_fence_after();
            const uint32_t a_st = sbase + SM::A_OFF + s * GM_A_STAGE_BYTES, b_st = sbase + SM::B_OFF + s * SM::B_STAGE_BYTES;
            const uint32_t d_lead = tmem + (uint32_t)((i & 1) * NPAD), d_trail = tmem + 2u * NPAD;
#pragma unroll
            for (int kk = 0; kk < GM_BLOCK_K / GM_UMMA_K; kk++) {
                // A: K-major, no swizzle: 8 x 16-byte core matrices, 128 B between row groups (SBO), 2048 B between K chunks (LBO)
                // B: K-major, 128-byte swizzle: rows of 128 B, 1024 B between 8-row groups (SBO); a K step is 32 B inside the row
                const uint64_t bdesc = gm_desc(b_st + kk * 32, 16, 1024, 2);
                const uint64_t a1 = gm_desc(a_st + kk * 2 * 2048, 2048, 128, 0);
                const uint64_t a2 = gm_desc(a_st + GM_A_SPLIT_BYTES + kk * 2 * 2048, 2048, 128, 0);
                const uint64_t a3 = gm_desc(a_st + 2 * GM_A_SPLIT_BYTES + kk * 2 * 2048, 2048, 128, 0);
                gm_mma(d_lead, a1, bdesc, IDESC, (i >= 2 || kk > 0) ? 1u : 0u);           // leading parts: accumulator i mod 2
                gm_mma(d_trail, a2, bdesc, IDESC, (i > 0 || kk > 0) ? 1u : 0u);           // trailing parts share the third
                gm_mma(d_trail, a3, bdesc, IDESC, 1u);
            }
            gm_commit(empty(s));                                      // stage s may be refilled when these MMAs have read it
        }
        gm_commit(accum_bar);
    }
    __syncthreads();
    if (wid == GM_PRODUCERS / 32) {
        gm_tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)gm_tmem_cols(NPAD)) : "memory");
    }
    if (a.stats && tid == 0) atomicAdd(&a.stats[MP_CNT_GEMM_TILES], (unsigned long long)nst);
}

// BF16 occupancy columns [ncols_pad][Kpad] and the source records of the shared parameter set
static __global__ void k_gemm_pack(const uint8_t *__restrict__ y, int n, int kpad, int ncols, unsigned short *__restrict__ ybf)
{
    const int col = blockIdx.y;
    for (int l = blockIdx.x * blockDim.x + threadIdx.x; l < kpad; l += gridDim.x * blockDim.x)
        ybf[(size_t)col * kpad + l] = (col < ncols && l < n && y[(size_t)col * n + l]) ? 0x3F80 : 0;      // BF16 1.0 / 0.0
}
static __global__ void k_gemm_sources(const float *__restrict__ px, const float *__restrict__ py, const float *__restrict__ aw0, int n,
                                      int kpad, float4 *__restrict__ src)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= kpad) return;
    src[l] = l < n ? make_float4(px ? px[l] : 0.f, py ? py[l] : 0.f, aw0[l], 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace mp

using namespace mp;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int NPAD> static int launch_gemm_npad(mp_engine *h, const CUtensorMap &map, GemmArgs a, dim3 grid)
{
    auto kern = k_conn_gemm<NPAD>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<NPAD>::BYTES + 1024));
    kern<<<grid, GM_THREADS, GemmSmem<NPAD>::BYTES + 1024, h->stream>>>(map, a);
    CK(cudaGetLastError());
    return MP_OK;
}

// S (set 0) of every chain from the resident y with the parameters of chain 0 (the caller has checked that all chains
// share alpha and b, that the engine is FP32 and the landscape has positions or is linear).
int mp_launch_conn_gemm(mp_engine *h, double alpha)
{
    const int n = h->cfg.n_patches, ncols = h->cfg.n_chains * (h->cfg.n_years - 1);
    const int kpad = (n + GM_BLOCK_K - 1) / GM_BLOCK_K * GM_BLOCK_K;
    const int npad = ncols <= 32 ? 32 : ncols <= 64 ? 64 : ncols <= 96 ? 96 : ncols <= 128 ? 128 : 160;
    const int ncolblk = (ncols + npad - 1) / npad;
    const size_t ybytes = (size_t)ncolblk * npad * kpad * 2, sbytes = (size_t)kpad * sizeof(float4);
    if (h->gemm_bytes < ybytes + sbytes) {
        if (h->d_gemm) CK(cudaFree(h->d_gemm));
        h->d_gemm = nullptr; h->gemm_bytes = 0;
        CK(cudaMalloc(&h->d_gemm, ybytes + sbytes));
        h->gemm_bytes = ybytes + sbytes;
    }
    unsigned short *ybf = (unsigned short *)h->d_gemm;
    float4 *src = (float4 *)((char *)h->d_gemm + ybytes);
    static PFN_encodeTiled encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        REQUIRE(fn && qres == cudaDriverEntryPointSuccess, MP_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        encode = (PFN_encodeTiled)fn;
    }
    CUtensorMap map;
    const cuuint64_t dims[2] = { (cuuint64_t)kpad, (cuuint64_t)ncolblk * npad };
    const cuuint64_t strides[1] = { (cuuint64_t)kpad * 2 };
    const cuuint32_t box[2] = { (cuuint32_t)GM_BLOCK_K, (cuuint32_t)npad }, estr[2] = { 1, 1 };
    const CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ybf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    REQUIRE(cr == CUDA_SUCCESS, MP_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    {
        Timed tm(h, MP_K_SMALL);
        k_gemm_pack<<<dim3((kpad + 255) / 256, ncolblk * npad), 256, 0, h->stream>>>(h->d_y, n, kpad, ncols, ybf);
        CK(cudaGetLastError());
        k_gemm_sources<<<(kpad + 255) / 256, 256, 0, h->stream>>>(h->geom == MP_GEOM_COORDS ? (const float *)h->d_px : nullptr,
                                                                  h->geom == MP_GEOM_COORDS ? (const float *)h->d_py : nullptr,
                                                                  (const float *)h->d_aw[0], n, kpad, src);
        CK(cudaGetLastError());
        CK(cudaMemsetAsync(h->d_S[0], 0, (size_t)ncols * n * 8, h->stream));
    }
    Timed tm(h, MP_K_CONN);
    GemmArgs a;
    a.src = src; a.S = h->d_S[0]; a.n = n; a.ncols = ncols; a.nstage_total = kpad / GM_BLOCK_K;
    a.apre = -1.4426950408889634f * (float)alpha; a.spacing = (float)h->spacing; a.linear = h->geom == MP_GEOM_LINEAR;
    a.stats = h->d_work;
    // split the sources of a target tile over several CTAs until the grid covers the SMs about twice
    const int mtiles = (n + GM_BLOCK_M - 1) / GM_BLOCK_M;
    int ksplit = std::max(1, std::min(a.nstage_total, (2 * h->sm_count + mtiles * ncolblk - 1) / (mtiles * ncolblk)));
    a.stages_per_cta = (a.nstage_total + ksplit - 1) / ksplit;
    ksplit = (a.nstage_total + a.stages_per_cta - 1) / a.stages_per_cta;
    const dim3 grid(mtiles, ksplit, ncolblk);
    switch (npad) {
    case 32: return launch_gemm_npad<32>(h, map, a, grid);
    case 64: return launch_gemm_npad<64>(h, map, a, grid);
    case 96: return launch_gemm_npad<96>(h, map, a, grid);
    case 128: return launch_gemm_npad<128>(h, map, a, grid);
    default: return launch_gemm_npad<160>(h, map, a, grid);
    }
}
