/* oracle/spom_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, FP64) of the SPOM hot path of nalcala/MIDASPOM, used ONLY as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * The product (midaspom_b200/, include/libmidaspom_cuda.h) never includes, links or calls it.
 *
 * What it restates, with the reference lines each function follows (paths under
 * /root/reference/sources/):
 *   spom_weight / spom_connectivity   main_MIDASPOM.c:180-188 (kernel matrix M), :350-358 (S, pC)
 *   spom_transition_prob              main_MIDASPOM.c:18-50 (compPePc), dieoff.c:51-83 (pije/pijc),
 *                                     loss.c:86-105 (pijcsource)
 *   spom_loglik                       the same factors, summed as logs per (year, patch) cell
 *   spom_marginal_loglik              main_MIDASPOM.c:363 (Pe.Pc = sum over the intermediate state),
 *                                     :368-392 (sum over completions of -1 cells, year-0 prior :248-250)
 *   spom_simulate                     main_MIDASPOM_future.c:64-110 (simpij)
 *   spom_sweep                        NO reference counterpart (the reference has no sampler):
 *                                     data-augmented Gibbs/Metropolis on the factorised likelihood
 *                                     above.  It is the CPU twin of the CUDA sampler (same Philox
 *                                     counters, same scan order) -- "port", parity unpinned except
 *                                     through the exact grid posterior of the reference.
 *
 * Pinned (tests/test_oracle_vs_reference.py) against oracle/_ref/libmidaspom_ref.so (the
 * reference's own functions) and oracle/_ref/MIDASPOM.out outputs / tests/golden/ fixtures.
 * Extensions the reference has no code for -- estimated alpha, areas with exponent b, planar or
 * dense distances, detection probability p<1 -- are PARITY UNPINNED except at the degenerate
 * point (alpha fixed, b=0 or A=1, p=1, linear geometry).
 */
#ifndef SPOM_ORACLE_H
#define SPOM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { SPOM_GEOM_LINEAR = 0, SPOM_GEOM_COORDS = 1, SPOM_GEOM_DENSE = 2 };

typedef struct {
    int32_t n;            /* patches */
    int32_t T;            /* years (rows of obs) */
    int32_t geom;         /* SPOM_GEOM_* */
    int32_t detect;       /* 0: perfect detection (reference); 1: obs=0 may hide z=1 w.p. 1-p */
    double  spacing;      /* linear geometry: distance between neighbouring patches (flag -d) */
    double  prior_occ;    /* year-0 prior occupancy of latent cells (flag -p; float in the reference) */
    const double *px, *py;    /* planar coordinates (geom=1) */
    const double *dist;       /* n*n row-major distances (geom=2) */
    const double *area;       /* patch areas, NULL => all 1 */
    const double *src_unit;   /* external-source distance multipliers u_k, NULL => k+1 (loss.c:365) */
    const int8_t *obs;        /* T*n, years major: -1 missing, 0, 1 (piobs, main_MIDASPOM.c:152-161) */
    const uint8_t *era;       /* T-1 flags: 1 => pre-event transition (K, Ksrc apply); NULL => all 0 */
    int32_t blk_nx, blk_ny, blk_k;   /* block grid of the y scan (spom_scan_order; 0 or 1 => one block: plain Morton order) */
} spom_model;

/* e, c: extinction / colonisation; alpha: 1/mean dispersal (flag -m); b: area exponent;
 * p: detection probability; K: pre-event scaling (dieoff.c:56-57,78); Ksrc, dsrc: external source
 * size and distance unit (loss.c:93-101,365; future.c:90-97,277). */
typedef struct { double e, c, alpha, b, p, K, Ksrc, dsrc; } spom_params;

typedef struct {
    double e_min, e_max, c_min, c_max, alpha_min, alpha_max, b_min, b_max, p_min, p_max;
    double K_min, K_max, Ksrc_min, Ksrc_max, dsrc_min, dsrc_max;
    int32_t sample_e, sample_c, sample_alpha, sample_b, sample_p;
    int32_t n_e_steps, n_c_steps;
    int32_t n_adapt;            /* sweeps during which proposal scales adapt */
    int32_t update_z, update_y;
    int32_t sample_K, sample_Ksrc, sample_dsrc, n_v_steps;
} spom_sampler_cfg;

#define SPOM_NDRAW 11  /* e, c, alpha, b, p, loglik, #y=1, #z=1, K, Ksrc, dsrc */
#define SPOM_NLSIG 8   /* log proposal scales: e, c, alpha, b, p, K, Ksrc, dsrc */

/* ---- Philox4x32-10 (Salmon et al. 2011), counter-based RNG shared with the CUDA engine ---- */
void   spom_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void   spom_rng(uint64_t seed, uint32_t chain, uint32_t sweep, uint32_t kind, uint32_t a, uint32_t b,
                uint32_t out[4]);
double spom_u01(uint32_t x);

/* ---- visiting order of the y scan: planar landscapes colour by colour, block by block of the blk_nx x blk_ny grid,
 * Morton order inside a block (one block: plain Morton order); index order otherwise ---- */
void   spom_scan_order(const spom_model *m, int32_t *order);

/* ---- likelihood pieces ---- */
double spom_weight(const spom_model *m, double alpha, double b, int target, int source);
void   spom_connectivity(const spom_model *m, double alpha, double b, const uint8_t *y_row, double *S_row);
void   spom_connectivity_targets(const spom_model *m, double alpha, double b, const uint8_t *y_row,
                                 const int32_t *targets, int ntargets, double *S_out /* ntargets */);
double spom_source_term(const spom_model *m, const spom_params *p, int k);
/* returns Pe*Pc; pe_out / pc_out (nullable) receive the two factors (Pee / Pcc entries of compPePc) */
double spom_transition_prob(const spom_model *m, const spom_params *p, int pre_event,
                            const uint8_t *z_old, const uint8_t *y_mid, const uint8_t *z_new,
                            double *pe_out, double *pc_out);
/* complete-data log-likelihood of (z, y); parts[4] = extinction, colonisation, prior, detection.
 * S_out (nullable) receives the (T-1)*n connectivity.  -INFINITY where the reference gives 0. */
double spom_loglik(const spom_model *m, const spom_params *p, const uint8_t *z, const uint8_t *y,
                   double *parts, double *S_out);
/* exact log-likelihood of the observations (sum over y and over completions of -1 cells);
 * Theta(2^#latent): small n only.  Perfect detection only. */
double spom_marginal_loglik(const spom_model *m, const spom_params *p);
/* Delta log-likelihood of flipping y[t][k], recomputed from scratch (2 full evaluations) */
double spom_flip_delta_bruteforce(const spom_model *m, const spom_params *p, const uint8_t *z,
                                  const uint8_t *y, int t, int k);

/* ---- sampler (CPU twin of the CUDA engine) ---- */
void spom_init_chain(const spom_model *m, const spom_sampler_cfg *cfg, uint64_t seed, uint32_t chain,
                     int disperse, spom_params *par, double *lsig, uint8_t *z, uint8_t *y, double *S);
void spom_refresh_S(const spom_model *m, const spom_params *par, const uint8_t *y, double *S);
/* one MCMC iteration; if y_flip_limit >= 0 only that many candidate cells per year are visited
 * (bounded sample for benchmarking); returns number of y candidates visited */
int64_t spom_sweep(const spom_model *m, const spom_sampler_cfg *cfg, uint64_t seed, uint32_t chain,
                   uint32_t sweep, spom_params *par, double *lsig, uint8_t *z, uint8_t *y, double *S,
                   double *draw, int64_t y_flip_limit, double *phase_s /* nullable: [other, y scan] seconds */);
/* rank-1 incremental log-odds of flipping y[t][k] given a consistent S (what the sweep uses) */
double spom_flip_delta(const spom_model *m, const spom_params *par, const uint8_t *z, const uint8_t *y,
                       const double *S, int t, int k);

/* ---- forward simulator (simpij) ---- */
void spom_simulate(const spom_model *m, const spom_params *p, uint64_t seed, uint32_t sim_id,
                   const uint8_t *z0, int nyears, uint8_t *z_out /* (nyears+1)*n */);

/* ---- multi-chain helpers for the CPU baseline (OpenMP over chains) ---- */
int64_t spom_sweep_chains(const spom_model *m, const spom_sampler_cfg *cfg, uint64_t seed, int nchains,
                          uint32_t chain0, uint32_t sweep, spom_params *par, double *lsig, uint8_t *z,
                          uint8_t *y, double *S, double *draws, int64_t y_flip_limit, int nthreads,
                          double *phase_s /* nullable: nchains x [other, y scan] seconds */);
int spom_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
