// mp_exact.cu -- the reference's OWN algorithm (exact likelihood by state enumeration on a
// parameter grid) on the GPU, FP64.  Rows a5-a7 of SURVEY section 8 and "next" row f3.
//
// Reference (paths under /root/reference/sources/):
//   state tables   main_MIDASPOM.c:198-290   piall, npstates, pstates, simppstates, priorst, all2short
//                                            (bit order: MSB = first variable patch, :206-207,244,247)
//   S and pC       main_MIDASPOM.c:350-358   (independent of (e, c): computed ONCE here, the reference
//                                            recomputes it for each of the 10,201 grid points)
//   compPePc       main_MIDASPOM.c:18-50     Pe[i][j] = E^s1 (1-E)^s2, Pc[j][i] = prod_k (...)
//   Pe.Pc          main_MIDASPOM.c:363       cblas_dgemm, nextid x nstates x nextid
//   forward pass   main_MIDASPOM.c:368-392   product over years of the gathered sub-blocks of P
//   dieoff / loss  main_MIDASPOM_dieoff.c:307-351, main_MIDASPOM_loss.c:345-386
//
// Layout: the integer tables are built on the host exactly as the reference builds them (they are
// tiny); one CTA per grid point evaluates P = Pe.Pc with the enumerated states streamed in chunks
// through shared memory, then one thread runs the forward recursion.  All sums in FP64, states in
// ascending order (the order of the reference's dgemm), so results agree to ~1e-15 relative.
//
// One deliberate difference: with -1 cells in the FIRST row the reference multiplies by a partly
// uninitialised Pold (main_MIDASPOM.c:368-369,379).  Here Pold starts as the
// identity -- the evident intent, and what the data-augmented likelihood sums to.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include "../../include/libmidaspom_cuda.h"

namespace {

thread_local std::string g_exact_error;

#define XCK(call)                                                                       \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) {                              \
        g_exact_error = std::string(#call) + ": " + cudaGetErrorString(e_); rc = MP_ERR_CUDA; goto done; } } while (0)

struct Tables {
    int n = 0, T = 0, nvar = 0, nextid = 0;
    uint32_t nstates = 0;
    std::vector<int> var;                 // n
    std::vector<uint32_t> zmask_all;      // nstates: bit k (patch k) of enumerated state
    std::vector<int> npstates;            // T
    std::vector<std::vector<uint32_t>> pstates, simpp;   // per year
    std::vector<float> priorst;           // npstates[0]   (float like the reference, :218)
    std::vector<uint32_t> all2short;      // nextid -> enumerated state id
};

// patch-bit mask (bit k = patch k occupied) of enumerated state id: piall[i][j] = i / 2^(nvar-jt-1) % 2  (:203-209)
static uint32_t state_mask(const Tables &tb, uint32_t id)
{
    uint32_t m = 0; int jt = 0;
    for (int j = 0; j < tb.n; j++)
        if (tb.var[j]) { if ((id >> (tb.nvar - jt - 1)) & 1u) m |= 1u << j; jt++; }
    return m;
}

static int build_tables(const int8_t *obs, int T, int n, float prioroc, Tables &tb)
{
    tb.n = n; tb.T = T;
    tb.var.assign(n, 0);
    for (int i = 0; i < T; i++) for (int j = 0; j < n; j++) if (obs[(size_t)i * n + j] != 0) tb.var[j] = 1;   // :162-164
    tb.nvar = 0; for (int j = 0; j < n; j++) tb.nvar += tb.var[j];
    if (tb.nvar > 24 || n > 31) { g_exact_error = "exact engine: more than 24 variable patches (2^nvar states)"; return MP_ERR_UNSUPPORTED; }
    tb.nstates = 1u << tb.nvar;
    tb.zmask_all.resize(tb.nstates);
    for (uint32_t i = 0; i < tb.nstates; i++) tb.zmask_all[i] = state_mask(tb, i);
    tb.npstates.assign(T, 0); tb.pstates.assign(T, {}); tb.simpp.assign(T, {});
    tb.nextid = 0;
    for (int i = 0; i < T; i++) {                                                  // :222-279
        int s1 = 0;
        for (int j = 0; j < n; j++) if (obs[(size_t)i * n + j] == -1) s1++;
        if (s1 > 12) { g_exact_error = "exact engine: more than 12 missing cells in one year (4096 completions)"; return MP_ERR_UNSUPPORTED; }
        const uint32_t np = 1u << s1;
        tb.npstates[i] = (int)np;
        if (i == 0) tb.priorst.assign(np, 1.0f);
        tb.pstates[i].assign(np, 0); tb.simpp[i].assign(np, 0);
        s1 = 0; int jt = 0;
        for (int j = 0; j < n; j++) {
            const int o = obs[(size_t)i * n + j];
            if (o == -1) s1++;
            for (uint32_t k = 0; k < np; k++) {
                if (o > -1) { if (o == 1) tb.pstates[i][k] += 1u << (tb.nvar - jt - 1); }      // :244 (o == 1 implies var[j])
                else {
                    const uint32_t st1 = np >> s1;                                  // npstates / 2^s1
                    const uint32_t bit = (k / st1) % 2;
                    tb.pstates[i][k] += bit << (tb.nvar - jt - 1);
                    if (i == 0) tb.priorst[k] *= (float)bit * prioroc + (float)(1 - bit) * (1 - prioroc);   // :249, float arithmetic
                }
            }
            if (tb.var[j]) jt++;
        }
        for (uint32_t k = 0; k < np; k++) {                                        // :256-278
            if (i == 0) { tb.simpp[0][k] = tb.nextid++; continue; }
            bool present = false;
            for (int j = 0; j < i; j++)
                for (uint32_t l = 0; l < (uint32_t)tb.npstates[j]; l++)
                    if (tb.pstates[i][k] == tb.pstates[j][l]) { tb.simpp[i][k] = tb.simpp[j][l]; present = true; }
            if (!present) tb.simpp[i][k] = tb.nextid++;
        }
    }
    tb.all2short.assign(tb.nextid, 0);
    for (int i = 0; i < T; i++) for (uint32_t k = 0; k < (uint32_t)tb.npstates[i]; k++) tb.all2short[tb.simpp[i][k]] = tb.pstates[i][k];   // :282-287
    return MP_OK;
}

// S[j][k] = sum_{l != k} M[l][k] y_l(j), M[l][k] = exp(-a |l-k| d)   (:180-188, :351-355); l ascending
__global__ void k_exact_S(const uint32_t *__restrict__ masks, uint32_t nstates, int n, double a, double d, double *__restrict__ S)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nstates) return;
    const uint32_t m = masks[j];
    for (int k = 0; k < n; k++) {
        double s = 0.0;
        for (int l = 0; l < n; l++)
            if (l != k && ((m >> l) & 1u)) s += exp(-a * (double)(l > k ? l - k : k - l) * d);
        S[(size_t)j * n + k] = s;
    }
}

constexpr int XCHUNK = 64;      // enumerated states per shared-memory chunk
constexpr int XTILE = 32;       // P = Pe.Pc is formed in XTILE x XTILE tiles of short-list states (any number of them)

// One CTA per (grid point, row tile, column tile): P[i][i'] = sum_j Pe[i][j] Pc[j][i'] over all enumerated states j
// (ascending, the order of the reference's dgemm) for the short-list states i in the row tile and i' in the column tile.
__global__ void __launch_bounds__(256)
k_exact_P(const uint32_t *__restrict__ masks, uint32_t nstates, int n, const double *__restrict__ S,
          const uint32_t *__restrict__ short_masks, int nextid, const double *__restrict__ egrid,
          const double *__restrict__ cgrid, int nstep, int gp0, double *__restrict__ Pout)
{
    __shared__ double sPe[XCHUNK][XTILE], sPc[XCHUNK][XTILE];
    const int gp = gp0 + blockIdx.x, ie = gp / nstep, ic = gp - ie * nstep;
    const int a0 = blockIdx.y * XTILE, b0 = blockIdx.z * XTILE;
    double E = egrid[ie]; if (E > 1.0) E = 1.0;                                     // compPePc:21-22
    const double c = cgrid[ic];
    const int tid = threadIdx.x;
    // each thread owns the entries tid, tid+256, ... of the tile
    constexpr int MAXOWN = (XTILE * XTILE + 255) / 256;
    double acc[MAXOWN];
#pragma unroll
    for (int o = 0; o < MAXOWN; o++) acc[o] = 0.0;
    for (uint32_t j0 = 0; j0 < nstates; j0 += XCHUNK) {
        // factors of the chunk: thread handles (state, short state) pairs -- Pe for the row tile, Pc for the column tile
        for (int p = tid; p < XCHUNK * XTILE * 2; p += 256) {
            const int which = p / (XCHUNK * XTILE), r = p - which * (XCHUNK * XTILE);
            const int js = r / XTILE, il = r - js * XTILE, i = (which ? b0 : a0) + il;
            const uint32_t j = j0 + js;
            double v = 0.0;
            if (j < nstates && i < nextid) {
                const uint32_t y = masks[j], zs = short_masks[i];
                if ((y & ~zs) == 0u) {                                              // compPePc:34-37  (0 -> 1 impossible)
                    if (!which) {
                        const int s1 = __popc(zs & ~y), s2 = __popc(zs & y);        // :38-39
                        v = pow(E, (double)s1) * pow(1.0 - E, (double)s2);          // :43
                    } else {
                        v = 1.0;
                        for (int k = 0; k < n; k++) {                               // :40
                            if ((y >> k) & 1u) continue;
                            double pC = c * S[(size_t)j * n + k];                   // :356-357
                            if (pC > 1.0) pC = 1.0;
                            v *= ((zs >> k) & 1u) ? pC : 1.0 - pC;
                        }
                    }
                }
            }
            (which ? sPc : sPe)[js][il] = v;
        }
        __syncthreads();
#pragma unroll
        for (int o = 0; o < MAXOWN; o++) {
            const int ent = tid + o * 256, al = ent / XTILE, bl = ent - al * XTILE;
            double s = acc[o];
            for (int js = 0; js < XCHUNK; js++) s += sPe[js][al] * sPc[js][bl];
            acc[o] = s;
        }
        __syncthreads();
    }
    double *P = Pout + (size_t)blockIdx.x * nextid * nextid;
#pragma unroll
    for (int o = 0; o < MAXOWN; o++) {
        const int ent = tid + o * 256, ai = a0 + ent / XTILE, bi = b0 + ent % XTILE;
        if (ai < nextid && bi < nextid) P[(size_t)ai * nextid + bi] = acc[o];
    }
}

// One CTA per grid point: the forward recursion of main_MIDASPOM.c:368-392 over the years and Lik = log sum.
// Pold (np0 x np_{t-1}) starts as the identity; the entries of each product are spread over the threads, every entry
// summed in the reference's order.
__global__ void __launch_bounds__(256)
k_exact_forward(const double *__restrict__ Pall, int nextid, int T, const int *__restrict__ npstates, const int *__restrict__ simpp_off,
                const uint32_t *__restrict__ simpp, const float *__restrict__ priorst, int maxnp, int gp0, double *__restrict__ lik,
                double *__restrict__ work)
{
    const int tid = threadIdx.x, np0 = npstates[0];
    const double *P = Pall + (size_t)blockIdx.x * nextid * nextid;
    double *A = work + (size_t)blockIdx.x * 2 * np0 * maxnp, *B = A + (size_t)np0 * maxnp;
    for (int p = tid; p < np0 * np0; p += 256) A[p] = (p / np0) == (p % np0) ? 1.0 : 0.0;
    __syncthreads();
    for (int t = 1; t < T; t++) {
        const int npp = npstates[t - 1], npc = npstates[t];
        const uint32_t *sp = simpp + simpp_off[t - 1], *sc = simpp + simpp_off[t];
        for (int p = tid; p < np0 * npc; p += 256) {
            const int k = p / npc, l = p - k * npc;
            double s = 0.0;
            for (int m = 0; m < npp; m++) s += A[k * npp + m] * P[(size_t)sp[m] * nextid + sc[l]];   // :375,379
            B[k * npc + l] = s;
        }
        __syncthreads();
        double *tmp = A; A = B; B = tmp;
    }
    if (tid == 0) {
        double L = 0.0;
        const int npl = npstates[T - 1];
        for (int k = 0; k < np0; k++) for (int l = 0; l < npl; l++) L += A[k * npl + l] * (double)priorst[k];   // :386-390
        lik[gp0 + blockIdx.x] = log(L);                                             // :392
    }
}


// ---- dieoff / loss: likelihood of the FIRST survey after ts pre-event and tdis post-event unobserved years
// (dieoff.c:307-351, loss.c:345-386).  The reference forms P = Pe.Pc, matrix powers PK^ts and P^tdis
// and sums every row of their product; summing the rows first turns that into ts + tdis
// vector-matrix products  v <- (v.Pe).Pc  starting from v = (1,...,1).  One CTA per grid value.
__global__ void __launch_bounds__(256)
k_exact_variant(int n, const double *__restrict__ S, double a, double eB, double cB, const double *__restrict__ Kgrid,
                const double *__restrict__ dgrid, int nd, int variant, int ts, int tdis, const uint32_t *__restrict__ pst,
                const float *__restrict__ prior, int npst, double *__restrict__ lik,
                double *__restrict__ gscratch /* 14..16 patches: the two state vectors of every CTA in global memory; else nullptr */)
{
    extern __shared__ double xs[];
    const uint32_t nstates = 1u << n;
    double *v = gscratch ? gscratch + (size_t)blockIdx.x * 2 * nstates : xs, *w = v + nstates;
    double *powE = gscratch ? xs : w + nstates, *pow1 = powE + 32, *g = pow1 + 32;
    const int iK = blockIdx.x / nd, id = blockIdx.x - iK * nd, tid = threadIdx.x;
    const double Kval = Kgrid[iK], dL = dgrid ? dgrid[id] : 0.0;
    for (uint32_t i = tid; i < nstates; i += 256) v[i] = 1.0;
    if (tid < n) g[tid] = exp(-a * (double)(tid + 1) * dL);            // loss.c:365  M[n][j] = exp(-a (j+1) d_L)
    __syncthreads();
    for (int step = 0; step < ts + tdis; step++) {
        const bool pre = step < ts;
        double E = (pre && variant == 1) ? eB / Kval : eB;              // dieoff.c:56-57 ; loss.c:57
        if (E > 1.0) E = 1.0;
        const double Kt = (pre && variant == 1) ? Kval : 1.0, Ks = (pre && variant == 2) ? Kval : 0.0;
        if (tid <= n) { powE[tid] = pow(E, (double)tid); pow1[tid] = pow(1.0 - E, (double)tid); }
        __syncthreads();
        for (uint32_t j = tid; j < nstates; j += 256) {                 // w = v . Pe   (pije)
            double s = 0.0;
            for (uint32_t i = 0; i < nstates; i++)
                if ((j & ~i) == 0u) s += v[i] * (powE[__popc(i & ~j)] * pow1[__popc(i & j)]);
            w[j] = s;
        }
        __syncthreads();
        for (uint32_t t = tid; t < nstates; t += 256) {                 // v = w . Pc   (pijc / pijcsource)
            double s = 0.0;
            for (uint32_t j = 0; j < nstates; j++) {
                if ((j & ~t) != 0u) continue;
                double pc = 1.0;
                for (int k = 0; k < n; k++) {
                    if ((j >> k) & 1u) continue;
                    double C = variant == 1 ? cB * S[(size_t)j * n + k] * Kt : cB * (S[(size_t)j * n + k] + g[k] * Ks);   // dieoff.c:78 ; loss.c:97-98
                    if (C > 1.0) C = 1.0;
                    pc *= ((t >> k) & 1u) ? C : 1.0 - C;
                }
                s += w[j] * pc;
            }
            v[t] = s;
        }
        __syncthreads();
    }
    if (tid == 0) {
        double L = 0.0;
        for (int j = 0; j < npst; j++) L += v[pst[j]] * (double)prior[j];     // dieoff.c:345-350
        lik[blockIdx.x] = L;
    }
}

}  // namespace

extern "C" {

const char *mp_exact_last_error(void) { return g_exact_error.c_str(); }

// Exact log-likelihood of the observations on an nstep x nstep grid of (e, c) -- the table
// MIDASPOM.out computes (main_MIDASPOM.c:341-395).  loglik_out[ie*nstep+ic] = log L(e_ie, c_ic);
// ltot_out (nullable) = 2 log(win) + log sum coef exp(Lik) (:414-424).  state_info (nullable, 4 ints):
// n variable patches, enumerated states, short-list states, max states per year.
int mp_exact_posterior(int device, const int8_t *obs, int n_years, int n_patches, double a, double d, double prior_occ,
                       int nstep, double ecmin, double ecmax, double *loglik_out, double *ltot_out, int *state_info)
{
    if (!obs || !loglik_out || n_years < 2 || n_patches < 1 || nstep < 2) { g_exact_error = "mp_exact_posterior: bad argument"; return MP_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_exact_error = "mp_exact_posterior: no CUDA device; there is no CPU fallback"; return MP_ERR_CUDA; }
    if (device < 0 || device >= ndev) { g_exact_error = "mp_exact_posterior: bad device ordinal"; return MP_ERR_ARG; }
    Tables tb;
    int rc = build_tables(obs, n_years, n_patches, (float)prior_occ, tb);
    if (rc != MP_OK) return rc;
    if (tb.nextid > 4096) { g_exact_error = "exact engine: more than 4096 distinct observation-compatible states"; return MP_ERR_UNSUPPORTED; }
    const int n = n_patches, T = n_years, ng = nstep * nstep;
    // grid axes (:312-319): i*win + ecmin, last = ecmax
    const double win = (ecmax - ecmin) / (nstep - 1);
    std::vector<double> axis(nstep);
    for (int i = 0; i < nstep - 1; i++) axis[i] = (double)i * win + ecmin;
    axis[nstep - 1] = ecmax;
    std::vector<uint32_t> short_masks(tb.nextid);
    for (int i = 0; i < tb.nextid; i++) short_masks[i] = tb.zmask_all[tb.all2short[i]];
    std::vector<int> simpp_off(T + 1, 0);
    std::vector<uint32_t> simpp_flat;
    int maxnp = 1;
    for (int t = 0; t < T; t++) {
        simpp_off[t] = (int)simpp_flat.size();
        simpp_flat.insert(simpp_flat.end(), tb.simpp[t].begin(), tb.simpp[t].end());
        if (tb.npstates[t] > maxnp) maxnp = tb.npstates[t];
    }
    if (state_info) { state_info[0] = tb.nvar; state_info[1] = (int)tb.nstates; state_info[2] = tb.nextid; state_info[3] = maxnp; }

    // one arena from the stream-ordered pool (kept by the driver between calls: no cudaMalloc / cudaFree per grid),
    // carved into the 11 device buffers at 256-byte boundaries.  P and the recursion's work space are held for a batch
    // of grid points (at most ~2 GB), the grid is walked batch by batch.
    uint32_t *d_masks = nullptr, *d_short = nullptr, *d_simpp = nullptr;
    double *d_S = nullptr, *d_axis = nullptr, *d_P = nullptr, *d_lik = nullptr, *d_work = nullptr;
    int *d_np = nullptr, *d_off = nullptr;
    float *d_prior = nullptr;
    unsigned char *arena = nullptr;
    const size_t per_point = ((size_t)tb.nextid * tb.nextid + (size_t)2 * tb.npstates[0] * maxnp) * 8;
    const int batch = (int)std::max<size_t>(1, std::min<size_t>((size_t)ng, ((size_t)2 << 30) / per_point));
    const int ntile = (tb.nextid + XTILE - 1) / XTILE;
    XCK(cudaSetDevice(device));
    {
        size_t off = 0;
        auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
        const size_t o_masks = carve((size_t)tb.nstates * 4), o_short = carve((size_t)tb.nextid * 4), o_simpp = carve(simpp_flat.size() * 4),
                     o_S = carve((size_t)tb.nstates * n * 8), o_axis = carve((size_t)nstep * 8),
                     o_P = carve((size_t)batch * tb.nextid * tb.nextid * 8), o_lik = carve((size_t)ng * 8),
                     o_work = carve((size_t)batch * 2 * tb.npstates[0] * maxnp * 8), o_np = carve((size_t)T * 4),
                     o_off = carve((size_t)(T + 1) * 4), o_prior = carve(tb.priorst.size() * 4);
        static bool pool_ready = false;
        if (!pool_ready) {                                // keep freed memory in the pool instead of returning it to the OS
            cudaMemPool_t pool; unsigned long long keep = ~0ull;
            if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            pool_ready = true;
        }
        XCK(cudaMallocAsync((void **)&arena, off, 0));
        d_masks = (uint32_t *)(arena + o_masks); d_short = (uint32_t *)(arena + o_short); d_simpp = (uint32_t *)(arena + o_simpp);
        d_S = (double *)(arena + o_S); d_axis = (double *)(arena + o_axis); d_P = (double *)(arena + o_P); d_lik = (double *)(arena + o_lik);
        d_work = (double *)(arena + o_work); d_np = (int *)(arena + o_np); d_off = (int *)(arena + o_off); d_prior = (float *)(arena + o_prior);
    }
    XCK(cudaMemcpyAsync(d_masks, tb.zmask_all.data(), (size_t)tb.nstates * 4, cudaMemcpyHostToDevice, 0));
    XCK(cudaMemcpyAsync(d_short, short_masks.data(), (size_t)tb.nextid * 4, cudaMemcpyHostToDevice, 0));
    XCK(cudaMemcpyAsync(d_simpp, simpp_flat.data(), simpp_flat.size() * 4, cudaMemcpyHostToDevice, 0));
    XCK(cudaMemcpyAsync(d_axis, axis.data(), (size_t)nstep * 8, cudaMemcpyHostToDevice, 0));
    XCK(cudaMemcpyAsync(d_np, tb.npstates.data(), (size_t)T * 4, cudaMemcpyHostToDevice, 0));
    XCK(cudaMemcpyAsync(d_off, simpp_off.data(), (size_t)(T + 1) * 4, cudaMemcpyHostToDevice, 0));
    XCK(cudaMemcpyAsync(d_prior, tb.priorst.data(), tb.priorst.size() * 4, cudaMemcpyHostToDevice, 0));
    k_exact_S<<<(tb.nstates + 127) / 128, 128>>>(d_masks, tb.nstates, n, a, d, d_S);
    XCK(cudaGetLastError());
    for (int gp0 = 0; gp0 < ng; gp0 += batch) {
        const int nb = std::min(batch, ng - gp0);
        k_exact_P<<<dim3(nb, ntile, ntile), 256>>>(d_masks, tb.nstates, n, d_S, d_short, tb.nextid, d_axis, d_axis, nstep, gp0, d_P);
        XCK(cudaGetLastError());
        k_exact_forward<<<nb, 256>>>(d_P, tb.nextid, T, d_np, d_off, d_simpp, d_prior, maxnp, gp0, d_lik, d_work);
        XCK(cudaGetLastError());
    }
    XCK(cudaMemcpy(loglik_out, d_lik, (size_t)ng * 8, cudaMemcpyDeviceToHost));
    if (ltot_out) {                                                                 // :414-424 (host, nstep^2 terms)
        double Ltot = 0.0;
        for (int k = 0; k < nstep; k++)
            for (int l = 0; l < nstep; l++) {
                double coef = 1.0;
                if (k == 0 || k == nstep - 1) coef *= 0.5;
                if (l == 0 || l == nstep - 1) coef *= 0.5;
                Ltot += exp(loglik_out[(size_t)k * nstep + l]) * coef;
            }
        *ltot_out = 2.0 * log(win) + log(Ltot);
    }
done:
    if (arena) cudaFreeAsync(arena, 0);
    return rc;
}


// dieoff (variant 1: MIDASPOM_dieoff.out) / loss (variant 2: MIDASPOM_loss.out): likelihood of the first
// survey row on the log-spaced K grid (dieoff.c:283-286) x, for loss, the d_L grid (loss.c:319-322).
// lik_out: nstepK (dieoff) or nstepK*nstepd (loss) raw likelihoods, as the reference writes them.
int mp_exact_variant(int device, int variant, const int8_t *first_row, int n_patches, double a, double d, double prior_occ,
                     double eB, double cB, int ts, int tdis, int nstepK, double Kmin, double Kmax, int nstepd, double dmin,
                     double dmax, double *lik_out)
{
    if (!first_row || !lik_out || n_patches < 1 || n_patches > 16 || nstepK < 2 || ts < 0 || tdis < 0 || (variant != 1 && variant != 2) ||
        (variant == 2 && nstepd < 2)) { g_exact_error = "mp_exact_variant: bad argument (at most 16 patches)"; return MP_ERR_ARG; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_exact_error = "mp_exact_variant: no CUDA device; there is no CPU fallback"; return MP_ERR_CUDA; }
    const int n = n_patches, nd = variant == 2 ? nstepd : 1;
    const uint32_t nstates = 1u << n;
    // completions of the -1 cells of the survey row and their prior (dieoff.c:204-233): all 2^n states, bit j of the
    // id = patch j counted from the MSB (pow(2, n-j-1))
    int s1 = 0;
    for (int j = 0; j < n; j++) s1 += first_row[j] == -1;
    const uint32_t npst = 1u << s1;
    std::vector<uint32_t> pst(npst, 0);
    std::vector<float> prior(npst, 1.0f);
    const float pr = (float)prior_occ;
    s1 = 0;
    for (int j = 0; j < n; j++) {
        if (first_row[j] == -1) s1++;
        for (uint32_t k = 0; k < npst; k++) {
            uint32_t bit;
            if (first_row[j] > -1) bit = first_row[j] == 1;
            else { const uint32_t st1 = npst >> s1; bit = (k / st1) % 2; prior[k] *= (float)bit * pr + (float)(1 - bit) * (1 - pr); }
            pst[k] |= bit << j;                                        // our masks use bit j = patch j (see k_exact_S)
        }
    }
    std::vector<uint32_t> masks(nstates);
    for (uint32_t i = 0; i < nstates; i++) masks[i] = i;               // state id == patch mask (order of the sums is immaterial here)
    std::vector<double> Kg(nstepK), dg(nd, 0.0);
    for (int i = 0; i < nstepK; i++) Kg[i] = pow(10.0, ((double)i) / (nstepK - 1) * (log10(Kmax) - log10(Kmin)) + log10(Kmin));
    if (variant == 2) for (int i = 0; i < nd; i++) dg[i] = i * (dmax - dmin) / (nd - 1) + dmin;
    int rc = MP_OK;
    uint32_t *d_masks = nullptr, *d_pst = nullptr;
    double *d_S = nullptr, *d_K = nullptr, *d_d = nullptr, *d_lik = nullptr, *d_scr = nullptr;
    float *d_prior = nullptr;
    // the two state vectors of a CTA fit its shared memory up to 13 patches (2 x 8,192 doubles); beyond that they live in
    // global memory (MP_EXACT_GLOBAL=1 forces that path at any size)
    const char *force = getenv("MP_EXACT_GLOBAL");
    const bool global_vec = n > 13 || (force && atoi(force) != 0);
    const size_t smem = ((global_vec ? 0 : (size_t)2 * nstates) + 64 + 32) * sizeof(double);
    XCK(cudaSetDevice(device));
    XCK(cudaMalloc(&d_masks, (size_t)nstates * 4)); XCK(cudaMalloc(&d_pst, (size_t)npst * 4)); XCK(cudaMalloc(&d_prior, (size_t)npst * 4));
    XCK(cudaMalloc(&d_S, (size_t)nstates * n * 8)); XCK(cudaMalloc(&d_K, (size_t)nstepK * 8)); XCK(cudaMalloc(&d_d, (size_t)nd * 8));
    XCK(cudaMalloc(&d_lik, (size_t)nstepK * nd * 8));
    XCK(cudaMemcpy(d_masks, masks.data(), (size_t)nstates * 4, cudaMemcpyHostToDevice));
    XCK(cudaMemcpy(d_pst, pst.data(), (size_t)npst * 4, cudaMemcpyHostToDevice));
    XCK(cudaMemcpy(d_prior, prior.data(), (size_t)npst * 4, cudaMemcpyHostToDevice));
    XCK(cudaMemcpy(d_K, Kg.data(), (size_t)nstepK * 8, cudaMemcpyHostToDevice));
    XCK(cudaMemcpy(d_d, dg.data(), (size_t)nd * 8, cudaMemcpyHostToDevice));
    k_exact_S<<<(nstates + 127) / 128, 128>>>(d_masks, nstates, n, a, d, d_S);
    XCK(cudaGetLastError());
    XCK(cudaFuncSetAttribute(k_exact_variant, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (global_vec) XCK(cudaMalloc(&d_scr, (size_t)nstepK * nd * 2 * nstates * 8));
    k_exact_variant<<<nstepK * nd, 256, smem>>>(n, d_S, a, eB, cB, d_K, variant == 2 ? d_d : nullptr, nd, variant, ts, tdis, d_pst,
                                                d_prior, (int)npst, d_lik, d_scr);
    XCK(cudaGetLastError());
    XCK(cudaMemcpy(lik_out, d_lik, (size_t)nstepK * nd * 8, cudaMemcpyDeviceToHost));
done:
    cudaFree(d_masks); cudaFree(d_pst); cudaFree(d_prior); cudaFree(d_S); cudaFree(d_K); cudaFree(d_d); cudaFree(d_lik); cudaFree(d_scr);
    return rc;
}

}  // extern "C"
