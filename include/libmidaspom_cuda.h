/* libmidaspom_cuda.h -- C ABI of the B200-native SPOM likelihood / MCMC engine.
 *
 * This is the drop-in boundary for ONE hot path of nalcala/MIDASPOM: the stochastic patch
 * occupancy model (SPOM) log-likelihood and its connectivity term.  The reference has no
 * plugin/FFI boundary (each program is one C translation unit); the entry points below sit on its
 * function-level seams and keep its conventions: plain C, caller-owned flat host arrays, years
 * major / patches minor (piobs[i][j], main_MIDASPOM.c:152-161), scalars by value.  Unlike the
 * reference they return an int status (0 = ok, <0 = error, text via mp_last_error) and never abort.
 *
 * Reference code each entry point replaces (paths under /root/reference/sources/):
 *   mp_set_landscape_linear   main_MIDASPOM.c:177-188          kernel matrix M[i][j]=exp(-a|i-j|d) (flag -d);
 *                             never materialised here: weights are evaluated on the fly
 *   mp_set_landscape_coords / _dense                            (extension: general d_ij, areas A_j^b)
 *   mp_set_observations       main_MIDASPOM.c:138-167           piobs table (-1/0/1), T x N
 *   mp_connectivity           main_MIDASPOM.c:350-358           S_k = sum_{l!=k} M[l][k] y_l  (scalar triple loop)
 *                             dieoff.c:73-79, loss.c:74-80,93-101, future.c:89-97
 *   mp_loglik                 main_MIDASPOM.c:18-50 (compPePc), dieoff.c:51-83 (pije/pijc),
 *                             loss.c:86-105 (pijcsource): the per-patch Bernoulli factors, summed as logs
 *   mp_flip_delta             (no counterpart) rank-1 incremental form of the same terms
 *   mp_sweep / mp_get_draws   replaces the (e,c) grid loops main_MIDASPOM.c:341-395 by data-augmented
 *                             Gibbs/Metropolis chains; MIDASPOM_MPI's static row split
 *                             (main_MIDASPOM_MPI.c:361-372,483-505) becomes chain sharding (chain_offset)
 *   mp_set_era / mp_params.K,Ksrc,dsrc   dieoff.c:56-57,78 ; loss.c:93-101,365 ; future.c:67,90-97
 *   mp_simulate               future.c:64-110 (simpij) + :359-386 (simulation loop)
 *
 * All device memory is owned by the engine handle; host buffers are owned by the caller.  One
 * handle per GPU / host thread (thread-compatible, not thread-safe).  There is no CPU fallback:
 * every entry point fails with MP_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef LIBMIDASPOM_CUDA_H
#define LIBMIDASPOM_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MP_ABI_VERSION 1

enum { MP_OK = 0, MP_ERR_ARG = -1, MP_ERR_CUDA = -2, MP_ERR_STATE = -3, MP_ERR_UNSUPPORTED = -4 };
enum { MP_FP32 = 0, MP_FP64 = 1 };                       /* arithmetic of the weight / log terms */
enum { MP_GEOM_LINEAR = 0, MP_GEOM_COORDS = 1, MP_GEOM_DENSE = 2 };

#define MP_NDRAW 11  /* per (sweep, chain): e, c, alpha, b, p, loglik, #y=1, #z=1, K, Ksrc, dsrc */
#define MP_NLSIG 8   /* log proposal scales: e, c, alpha, b, p, K, Ksrc, dsrc */
#define MP_NPART 4   /* loglik parts: extinction, colonisation, year-0 prior, detection */

typedef struct mp_engine mp_engine;

typedef struct {
    int32_t  n_patches;      /* N */
    int32_t  n_years;        /* T (rows of the observation table) */
    int32_t  n_chains;       /* chains resident on this GPU */
    int32_t  chain_offset;   /* global id of local chain 0 (RNG streams are keyed by global id) */
    int32_t  precision;      /* MP_FP32 | MP_FP64 */
    int32_t  device;         /* CUDA ordinal */
    int32_t  detect;         /* 0: perfect detection (reference) ; 1: obs=0 hides z=1 w.p. 1-p */
    int32_t  max_draws;      /* capacity (sweeps) of the device draw buffer */
    uint64_t seed;
    double   prior_occ;      /* flag -p: year-0 prior occupancy of latent cells (float in the reference) */
} mp_config;

/* e, c: flags/grid axes of the reference; alpha = 1/(flag -m); b: area exponent; p: detection;
 * K: pre-event scaling (E=min(1,e/K), C=min(1,c*(K*S+...))); Ksrc, dsrc: external source size and
 * distance unit (source term exp(-alpha*u_k*dsrc)*Ksrc, u_k = k+1 by default, loss.c:365). */
typedef struct { double e, c, alpha, b, p, K, Ksrc, dsrc; } mp_params;

typedef struct {
    double  e_min, e_max, c_min, c_max, alpha_min, alpha_max, b_min, b_max, p_min, p_max;
    double  K_min, K_max, Ksrc_min, Ksrc_max, dsrc_min, dsrc_max;   /* variant parameters (dieoff.c:283-286, loss.c:319-322 grids) */
    int32_t sample_e, sample_c, sample_alpha, sample_b, sample_p;
    int32_t n_e_steps, n_c_steps;   /* Metropolis sub-steps per sweep */
    int32_t n_adapt;                /* sweeps during which proposal scales adapt */
    int32_t update_z, update_y;     /* Gibbs updates of latent occupancy / intermediate state */
    int32_t sample_K, sample_Ksrc, sample_dsrc;   /* Metropolis on the pre-event scaling / external source (log-uniform K, Ksrc; uniform dsrc) */
    int32_t n_v_steps;              /* sub-steps per sweep for each sampled variant parameter */
} mp_sampler_config;

/* kernel categories for mp_get_timing */
enum { MP_K_CONN = 0, MP_K_COL = 1, MP_K_SWEEP_Y = 2, MP_K_SWEEP_Z = 3, MP_K_SMALL = 4, MP_K_SIM = 5, MP_K_NCAT = 6 };
/* device buffers for mp_device_ptr (zero-copy hand-off to NCCL / torch) */
enum { MP_BUF_DRAWS = 0, MP_BUF_Z = 1, MP_BUF_Y = 2, MP_BUF_S = 3, MP_BUF_PARAMS = 4, MP_BUF_S_PROP = 5 };

const char *mp_version(void);
int mp_device_count(void);

int mp_create(const mp_config *cfg, mp_engine **out);
int mp_destroy(mp_engine *h);
const char *mp_last_error(const mp_engine *h);   /* h may be NULL: error of the last failed mp_create */

/* ---- landscape (replaces the dense kernel matrix M) ---- */
int mp_set_landscape_linear(mp_engine *h, double spacing, const double *area /* N or NULL */);
int mp_set_landscape_coords(mp_engine *h, const double *x, const double *y, const double *area);
int mp_set_landscape_dense(mp_engine *h, const double *dist /* N*N, [source][target] */, const double *area);
int mp_set_source_units(mp_engine *h, const double *src_unit /* N or NULL => k+1 */);
/* Order in which the y scan visits the patches of a year (slot -> patch): the Morton order of planar
 * coordinates, index order for linear and dense landscapes (oracle: spom_scan_order). */
int mp_get_scan_order(mp_engine *h, int32_t *order /* N */);
/* Block grid of the y scan (planar coordinates only; no reference counterpart -- the reference has no sampler).
 * The bounding box of the patches is cut into nx x ny cells and the cells are coloured k x k periodically.  The scan
 * then visits the patches colour by colour, block by block, in Morton order inside a block (oracle: spom_scan_order
 * with blk_nx, blk_ny, blk_k -- any fixed order is a valid Gibbs scan).  The FP32 engine scans the blocks of one colour
 * CONCURRENTLY, one cluster per (chain, year, block), with the block's own patches and every patch within `halo` of
 * its cell as targets: same-colour cells must lie farther apart than 2 halo ((k - 1) x cell width > 2 halo, else
 * MP_ERR_ARG), so their target sets are disjoint.  Each sweep a (chain, year) is scanned by blocks only if the largest
 * dispersal weight at the halo distance is below 2^-36 of the year's smallest S -- the commit rule the whole-year culled
 * scan applies anyway -- and the year holds at least 64 occupied patches; otherwise that year falls back to the
 * whole-year scan in the same order.  nx * ny <= 1 switches the grid off.  The FP64 engine only takes the order. */
int mp_set_scan_blocks(mp_engine *h, int nx, int ny, int k, double halo);

/* ---- data ---- */
int mp_set_observations(mp_engine *h, const int8_t *obs /* T*N, -1/0/1 */);
int mp_set_era(mp_engine *h, const uint8_t *era /* T-1 flags or NULL */);

/* ---- chain state ---- */
int mp_set_params(mp_engine *h, const mp_params *par /* n_chains */);
int mp_get_params(mp_engine *h, mp_params *par);
int mp_set_state(mp_engine *h, const uint8_t *z /* C*T*N */, const uint8_t *y /* C*(T-1)*N */);
int mp_get_state(mp_engine *h, uint8_t *z, uint8_t *y);
int mp_set_scales(mp_engine *h, const double *lsig /* C*MP_NLSIG */);
int mp_get_scales(mp_engine *h, double *lsig);

/* ---- likelihood ---- */
/* recompute S from the current y and parameters; S_out (host, C*(T-1)*N, nullable).  FP64 engines: 1e-9 path.  FP32 engines: the
 * evaluation form of k_conn (or the tensor-core path, see mp_get_conn_path), inside the FP32 path's 1e-5 */
int mp_connectivity(mp_engine *h, double *S_out);
int mp_get_connectivity(mp_engine *h, double *S_out);   /* copy of the resident S (no recompute) */
/* complete-data log-likelihood per chain (recomputes S); parts: C*MP_NPART, nullable */
int mp_loglik(mp_engine *h, double *ll, double *parts);
/* one-call form: upload parameters and state, evaluate, download (host buffers in and out) */
int mp_loglik_host(mp_engine *h, const mp_params *par, const uint8_t *z, const uint8_t *y, double *ll, double *parts);
/* log-odds of flipping y[chain][t][k] by the rank-1 update of the resident S */
int mp_flip_delta(mp_engine *h, int chain, int t, int k, double *dll);

/* ---- sampler ---- */
int mp_init_chains(mp_engine *h, const mp_sampler_config *sc, int disperse);
int mp_set_sampler(mp_engine *h, const mp_sampler_config *sc);
int mp_sweep(mp_engine *h, int nsweeps);            /* asynchronous on the engine stream */
int mp_synchronize(mp_engine *h);
/* One chain sharded over several GPUs (large N): every rank holds a full replica and runs the four phases of
 * a sweep with a collective in between (midaspom_b200/distributed.py: ShardedChain).
 *   mp_set_shard: this engine evaluates the connectivity of the targets [conn_lo, conn_hi) only (conn_hi < 0: all;
 *     positions in the Morton order of the patches on landscapes with positions, patch numbers otherwise; multiples
 *     of 256 except for the last rank) and scans the (chain, year) tasks task_first, task_first + task_stride, ... only.
 *     The other targets' columns of S / S_prop are left zero so that a sum over ranks assembles them.
 *   mp_sweep_phase: 0 proposal + connectivity (sum S / S_prop over ranks afterwards; flags_out bit 0: resident S
 *     recomputed, bit 1: proposal computed), 1 Metropolis decisions + z update, 2 y scan of the owned tasks
 *     (exchange the owned rows of y and S afterwards), 3 e/p update + record.  mp_sweep == the four phases. */
int mp_set_shard(mp_engine *h, int conn_lo, int conn_hi, int task_first, int task_stride);
int mp_sweep_phase(mp_engine *h, int phase, int *flags_out);
/* ---- several GPUs behind the C ABI (NCCL; replaces MPI_Init / the row split / the MPI_Send-MPI_Recv gather of
 * main_MIDASPOM_MPI.c:344-372,483-505).  One process per GPU: rank 0 calls mp_comm_unique_id, ships the MP_COMM_ID_BYTES
 * to the other ranks by whatever launcher it runs under, and every rank calls mp_comm_init.  One process driving several
 * engines (one per device): mp_comm_init_all.  NCCL is loaded at the first of these calls (MP_ERR_UNSUPPORTED if absent).
 *   mp_gather_draws   every rank's draws [first, first + count): out[rank][sweep][chain][MP_NDRAW], same on every rank
 *                     (independent chains per rank, chain_offset = rank x n_chains)
 *   mp_sweep_sharded  nsweeps iterations of ONE set of chains replicated on every rank: connectivity split by target
 *                     patches (all-gather of the owned columns of S / S_prop), y scan split by (chain, year) tasks
 *                     (broadcast of the owned rows of y and S), decisions replicated; equals mp_sweep on one engine bit
 *                     for bit.  Everything is enqueued on the engine stream; no host synchronisation inside a sweep.
 *                     In a single process call it for every engine of the communicator between mp_comm_group_begin /
 *                     mp_comm_group_end is NOT needed: use mp_sweep_sharded_all instead. */
#define MP_COMM_ID_BYTES 128
int mp_comm_unique_id(void *id128);
int mp_comm_init(mp_engine *h, int nranks, int rank, const void *id128);
int mp_comm_init_all(mp_engine **engines, int n);
int mp_comm_destroy(mp_engine *h);
int mp_comm_rank(mp_engine *h);
int mp_comm_size(mp_engine *h);
const char *mp_comm_last_error(void);
int mp_gather_draws(mp_engine *h, int first, int count, double *out /* nranks*count*C*MP_NDRAW */);
int mp_gather_draws_all(mp_engine **engines, int n, int first, int count, double *out);   /* one process, all engines: one NCCL group */
int mp_sweep_sharded(mp_engine *h, int nsweeps);
int mp_sweep_sharded_all(mp_engine **engines, int n, int nsweeps);
int mp_num_draws(mp_engine *h);
int mp_get_draws(mp_engine *h, int first, int count, double *out /* count*C*MP_NDRAW */);
int mp_reset_draws(mp_engine *h);
int mp_sweep_index(mp_engine *h);

/* ---- forward simulator ---- */
/* nsims independent trajectories of nyears from z0 (N bytes); era_all != 0 applies K/Ksrc every
 * year (future.c).  z_out (nullable): nsims*(nyears+1)*N ; occupied_out (nullable): nsims*(nyears+1) counts */
int mp_simulate(mp_engine *h, const mp_params *par, const uint8_t *z0, int nyears, int nsims, uint64_t seed,
                int era_all, uint8_t *z_out, int32_t *occupied_out);
/* same, with one parameter set and one start state per trajectory (future.c:359-381 draws (e,c) from
 * the posterior table and a completion of the last survey for every simulation) */
int mp_simulate_ensemble(mp_engine *h, const mp_params *par_per_sim /* nsims */, const uint8_t *z0_per_sim /* nsims*N */,
                         int nyears, int nsims, uint64_t seed, int era_all, uint8_t *z_out, int32_t *occupied_out);

/* ---- exact small-n engine: the reference's own algorithm (state enumeration on an (e,c) grid) ----
 * Replaces the whole hot loop of MIDASPOM.out, main_MIDASPOM.c:198-290 (state tables) and :341-395
 * (compPePc, Pe.Pc, forward recursion): loglik_out[ie*nstep+ic] = log L(e_ie, c_ic) on the grid
 * ecmin + i*(ecmax-ecmin)/(nstep-1); ltot_out (nullable) = the "Total log-likelihood" of :414-425;
 * state_info (nullable, 4 ints) = variable patches, enumerated states, short-list states, max states/year.
 * Linear habitat only (flags -m -> a = 1/m, -d, -p, -s, -l, -u).  At most 24 variable patches, 12 missing cells in one
 * year and 4,096 distinct observation-compatible rows (P is formed in 32 x 32 tiles, the grid in batches of <= 2 GB). */
int mp_exact_posterior(int device, const int8_t *obs, int n_years, int n_patches, double a, double d, double prior_occ,
                       int nstep, double ecmin, double ecmax, double *loglik_out, double *ltot_out, int *state_info);
/* MIDASPOM_dieoff.out (variant 1, main_MIDASPOM_dieoff.c:307-351) / MIDASPOM_loss.out (variant 2,
 * main_MIDASPOM_loss.c:345-386): likelihood of the FIRST survey row after ts pre-event + tdis post-event
 * unobserved years, on the log-spaced K grid (x the d_L grid for loss).  lik_out: nstepK (dieoff) or
 * nstepK*nstepd (loss) raw likelihoods, in the layout the reference writes.  At most 16 patches (the reference's own
 * 2^n x 2^n matrices stop fitting memory at 15); up to 13 the state vectors stay in shared memory. */
int mp_exact_variant(int device, int variant, const int8_t *first_row, int n_patches, double a, double d, double prior_occ,
                     double eB, double cB, int ts, int tdis, int nstepK, double Kmin, double Kmax, int nstepd, double dmin,
                     double dmax, double *lik_out);
const char *mp_exact_last_error(void);

/* ---- plumbing ---- */
int mp_device_ptr(mp_engine *h, int which, void **ptr, size_t *bytes);
/* the cudaStream_t every kernel of this engine is launched on (for CUDA-event timing by the caller) */
int mp_get_stream(mp_engine *h, void **stream);
int mp_set_timing(mp_engine *h, int enabled);
/* accumulated since the last call with reset != 0: ms[MP_K_NCAT], launches[MP_K_NCAT] */
int mp_get_timing(mp_engine *h, double *ms, int64_t *launches, int reset);
/* Work actually executed by the hot kernels since the last reset (always on: warp-local counts, one atomic per warp and
 * launch).  Used by bench.py so that roofline fractions count executed work, not the dense algorithm's:
 *   MP_CNT_SCAN_TRIPS    warp-trips of the culled y scan (one trip = one exchange of SP candidate sums)
 *   MP_CNT_SCAN_EXEC     (candidate, group of 32 targets) evaluations executed, speculative ones included
 *   MP_CNT_SCAN_RETIRED  the part of MP_CNT_SCAN_EXEC that belongs to candidates decided in their trip
 *   MP_CNT_SCAN_COMMIT   (accepted flip, group of 32 targets) rank-1 updates of S
 *   MP_CNT_SCAN_DENSE    (candidate, target) pairs of the unculled scans (k_sweep_y_fast, k_sweep_y)
 *   MP_CNT_CONN_EXEC     (group of 32 targets, group of 32 sources) tiles of k_conn evaluated
 *   MP_CNT_CONN_TOTAL    the same, culled ones included
 *   MP_CNT_GEMM_TILES    128 x 128 x 64 tensor-core tiles issued by the fixed-(alpha, b) connectivity path
 *   MP_CNT_SCAN_BLOCKS   block tasks of the y scan executed (mp_set_scan_blocks; 0: every year was scanned as a whole) */
enum { MP_CNT_SCAN_TRIPS = 0, MP_CNT_SCAN_EXEC = 1, MP_CNT_SCAN_RETIRED = 2, MP_CNT_SCAN_COMMIT = 3, MP_CNT_SCAN_DENSE = 4,
       MP_CNT_CONN_EXEC = 5, MP_CNT_CONN_TOTAL = 6, MP_CNT_GEMM_TILES = 7, MP_CNT_SCAN_BLOCKS = 8, MP_CNT_N = 9 };
int mp_get_work_counters(mp_engine *h, uint64_t *out /* MP_CNT_N */, int reset);
/* which kernel evaluated the connectivity last: 0 = k_conn (per-chain parameters; the sampler's own S with the year contraction
 * in FP64, the evaluation calls mp_connectivity / mp_loglik / mp_loglik_host of an FP32 engine with FP32 partial sums per 32
 * sources joined in FP64, ~1e-7 of every S -- the sweep after such a call recomputes the resident S), 1 = the tensor-core
 * contraction k_conn_gemm (FP32 engines; taken by mp_connectivity / mp_loglik / mp_loglik_host when every chain holds the
 * same alpha and b as uploaded by mp_set_params, the matrix form c*M%*%pti of Rscript/simuls_traj.R:16,203,214) */
int mp_get_conn_path(mp_engine *h);
/* launch geometry of the last y scan: out4 = threads per (chain, year) task, CTAs per cluster, candidates evaluated per
 * trip (1 + speculative), 1 if the spatially culled kernel ran, 2 if it ran block by block (mp_set_scan_blocks; the first
 * three numbers then describe a block task) (0: k_sweep_y_fast; all 0: the generic FP64 k_sweep_y) */
int mp_get_scan_geometry(mp_engine *h, int *out4);
/* micro-benchmarks on the engine's device: out[0] MUFU.EX2 Gop/s, out[1] FP32 FMA GFMA/s,
 * out[2] FP64 add Gop/s, out[3] device copy GB/s (read+write) */
int mp_probe_peaks(mp_engine *h, double *out4);

#ifdef __cplusplus
}
#endif
#endif /* LIBMIDASPOM_CUDA_H */
