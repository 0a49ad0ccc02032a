/* oracle/spom_oracle.c -- TEST INFRASTRUCTURE ONLY.  See spom_oracle.h for scope and citations.
 * Plain C, FP64, single-threaded per chain (OpenMP only across chains in spom_sweep_chains).
 * Paths cited as file:line are under /root/reference/sources/.
 */
#include "spom_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

/* ------------------------------------------------------------------ Philox4x32-10 */
/* Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3" (SC'11).  Not in the
 * reference (it uses libc rand(), main_MIDASPOM_future.c:77,99); shared with the CUDA engine so a
 * chain's stream depends only on (seed, chain, sweep, kind, cell), never on launch geometry. */
static inline void philox_round(uint32_t c[4], const uint32_t k[2])
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
void spom_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = { ctr[0], ctr[1], ctr[2], ctr[3] };
    uint32_t k[2] = { key[0], key[1] };
    for (int r = 0; r < 10; r++) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof c);
}
enum { RK_INIT_PARAM = 1, RK_INIT_Z = 2, RK_SIM_EXT = 3, RK_SIM_COL = 4, RK_Z = 5, RK_Y = 6,
       RK_AB = 7, RK_C = 8, RK_E = 9, RK_P = 10, RK_K = 11, RK_KSRC = 12, RK_DSRC = 13 };
void spom_rng(uint64_t seed, uint32_t chain, uint32_t sweep, uint32_t kind, uint32_t a, uint32_t b,
              uint32_t out[4])
{
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) + chain };
    const uint32_t ctr[4] = { a, b, kind, sweep };
    spom_philox4x32(ctr, key, out);
}
double spom_u01(uint32_t x) { return ((double)x + 0.5) * (1.0 / 4294967296.0); }

static void box_muller(uint32_t x0, uint32_t x1, double *n1, double *n2)
{
    const double r = sqrt(-2.0 * log(spom_u01(x0)));
    const double th = 6.283185307179586476925286766559 * spom_u01(x1);
    *n1 = r * cos(th); *n2 = r * sin(th);
}

/* ------------------------------------------------------------------ scan order */
/* Order in which the systematic y scan visits the patches of a year.  Any fixed order is a valid
 * Gibbs scan; planar landscapes use the Z-order (Morton) curve of the coordinates, quantised to
 * 16 bits per axis over the bounding square, ties in index order, so that consecutive visits are
 * spatial neighbours (the CUDA engine exploits that; both sides must visit in the same order for
 * the draw-by-draw twin tests).  Linear and dense landscapes are visited in index order. */
static uint32_t morton_spread16(uint32_t v)
{
    v &= 0xffffu; v = (v | (v << 8)) & 0x00ff00ffu; v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u; v = (v | (v << 1)) & 0x55555555u;
    return v;
}
typedef struct { uint32_t group, code; int32_t idx; } scan_key;
static int scan_key_cmp(const void *a, const void *b)
{
    const scan_key *x = a, *y = b;
    if (x->group != y->group) return x->group < y->group ? -1 : 1;
    if (x->code != y->code) return x->code < y->code ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}
/* Block grid of the scan (blk_nx x blk_ny cells over the bounding box, blk_k x blk_k colours): cell and colour of a point.
 * The scan visits colour by colour, block by block, Morton order inside a block; with a 1 x 1 grid this is the plain
 * Morton order.  The engine (mp_set_scan_blocks) uses the identical arithmetic. */
static inline int scan_cell(double v, double lo, double hi, int ncell)
{
    const double w = (hi - lo) / ncell;
    if (!(w > 0.0)) return 0;
    int c = (int)((v - lo) / w);
    if (c < 0) c = 0;
    if (c > ncell - 1) c = ncell - 1;
    return c;
}
void spom_scan_order(const spom_model *m, int32_t *order)
{
    const int n = m->n;
    if (m->geom != SPOM_GEOM_COORDS || n == 0) { for (int k = 0; k < n; k++) order[k] = k; return; }
    double x0 = m->px[0], x1 = m->px[0], y0 = m->py[0], y1 = m->py[0];
    for (int k = 1; k < n; k++) {
        if (m->px[k] < x0) x0 = m->px[k];
        if (m->px[k] > x1) x1 = m->px[k];
        if (m->py[k] < y0) y0 = m->py[k];
        if (m->py[k] > y1) y1 = m->py[k];
    }
    double span = x1 - x0 > y1 - y0 ? x1 - x0 : y1 - y0;
    if (!(span > 1e-300)) span = 1e-300;
    const int nbx = m->blk_nx > 1 ? m->blk_nx : 1, nby = m->blk_ny > 1 ? m->blk_ny : 1, bk = m->blk_k > 1 ? m->blk_k : 1;
    scan_key *key = malloc((size_t)n * sizeof *key);
    for (int k = 0; k < n; k++) {
        double fx = (m->px[k] - x0) / span * 65535.0, fy = (m->py[k] - y0) / span * 65535.0;
        if (fx > 65535.0) fx = 65535.0;
        if (fy > 65535.0) fy = 65535.0;
        key[k].code = morton_spread16((uint32_t)fx) | (morton_spread16((uint32_t)fy) << 1);
        key[k].idx = k;
        const int bx = scan_cell(m->px[k], x0, x1, nbx), by = scan_cell(m->py[k], y0, y1, nby);
        key[k].group = (uint32_t)(((bx % bk) + bk * (by % bk)) * (nbx * nby) + by * nbx + bx);   /* (colour, block) */
    }
    qsort(key, (size_t)n, sizeof *key, scan_key_cmp);
    for (int s = 0; s < n; s++) order[s] = key[s].idx;
    free(key);
}

/* ------------------------------------------------------------------ model pieces */
static inline double pair_distance(const spom_model *m, int a, int b)
{
    switch (m->geom) {
    case SPOM_GEOM_LINEAR: return 0.0; /* handled in spom_weight to keep the reference's op order */
    case SPOM_GEOM_COORDS: { const double dx = m->px[a] - m->px[b], dy = m->py[a] - m->py[b];
                             return sqrt(dx * dx + dy * dy); }
    default:               return m->dist[(size_t)b * m->n + a]; /* [source][target], like M[l][k] */
    }
}
/* main_MIDASPOM.c:184  M[i][j]=exp(-a*(j-i)*d)  -- evaluated as ((-a)*(j-i))*d */
static inline double dist_factor(const spom_model *m, double alpha, int target, int source)
{
    if (m->geom == SPOM_GEOM_LINEAR) {
        const unsigned int gap = (unsigned int)abs(target - source);
        return exp(-alpha * gap * m->spacing);
    }
    return exp(-alpha * pair_distance(m, target, source));
}
static inline double area_factor(const spom_model *m, double b, int source)
{
    return (m->area && b != 0.0) ? pow(m->area[source], b) : 1.0;
}
double spom_weight(const spom_model *m, double alpha, double b, int target, int source)
{
    double w = dist_factor(m, alpha, target, source);
    if (m->area && b != 0.0) w *= area_factor(m, b, source);
    return w;
}
/* main_MIDASPOM.c:351-355  s1 += M[l][k]*piall[j][l], l ascending, l != k */
void spom_connectivity(const spom_model *m, double alpha, double b, const uint8_t *y, double *S)
{
    const int n = m->n;
    for (int k = 0; k < n; k++) {
        double s = 0.0;
        for (int l = 0; l < n; l++)
            if (l != k && y[l]) s += spom_weight(m, alpha, b, k, l);
        S[k] = s;
    }
}
/* the same sum for a list of target patches only (O(ntargets * n): checks of very large landscapes on a sample) */
void spom_connectivity_targets(const spom_model *m, double alpha, double b, const uint8_t *y, const int32_t *targets,
                               int ntargets, double *S)
{
    const int n = m->n;
    for (int i = 0; i < ntargets; i++) {
        const int k = targets[i];
        double s = 0.0;
        for (int l = 0; l < n; l++)
            if (l != k && y[l]) s += spom_weight(m, alpha, b, k, l);
        S[i] = s;
    }
}
/* loss.c:365  M[n][j]=exp(-a*(j+1)*d_L) ; future.c:277 likewise with the -s distance */
double spom_source_term(const spom_model *m, const spom_params *p, int k)
{
    const double u = m->src_unit ? m->src_unit[k] : (double)(k + 1);
    return exp(-p->alpha * u * p->dsrc);
}
static inline double ext_prob(const spom_params *p, int pre)
{   /* main_MIDASPOM.c:21-22 ; dieoff.c:56-57 */
    double E = pre ? p->e / p->K : p->e;
    if (E > 1.0) E = 1.0;
    return E;
}
static inline double col_prob(const spom_params *p, int pre, double S, double g)
{   /* main_MIDASPOM.c:356-357 ; dieoff.c:78-79 (c*S*K) ; loss.c:98-100 (S + M[n][k]*Ksrc) */
    double C = pre ? p->c * (p->K * S + p->Ksrc * g) : p->c * S;
    if (C > 1.0) C = 1.0;
    return C;
}
static inline int is_pre(const spom_model *m, int t) { return m->era ? m->era[t] != 0 : 0; }
static inline int needs_src(const spom_params *p, int pre) { return pre && p->Ksrc != 0.0; }

double spom_transition_prob(const spom_model *m, const spom_params *p, int pre, const uint8_t *zo,
                            const uint8_t *y, const uint8_t *zn, double *pe_out, double *pc_out)
{
    const int n = m->n;
    double *S = malloc((size_t)n * sizeof(double));
    spom_connectivity(m, p->alpha, p->b, y, S);
    const double E = ext_prob(p, pre);
    unsigned s1 = 0, s2 = 0;
    double pc = 1.0;
    int possible = 1;
    for (int k = 0; k < n; k++) {
        if (y[k] && (!zo[k] || !zn[k])) { possible = 0; break; }   /* compPePc:34 ; pijc dieoff.c:72 */
        s1 += (1 - y[k]) * zo[k];                                  /* compPePc:38 */
        s2 += y[k] * zo[k];                                        /* compPePc:39 */
        const double g = needs_src(p, pre) ? spom_source_term(m, p, k) : 0.0;
        const double C = col_prob(p, pre, S[k], g);
        pc *= y[k] + (1 - y[k]) * ((1 - zn[k]) * (1.0 - C) + zn[k] * C);   /* compPePc:40 */
    }
    free(S);
    double pe = possible ? pow(E, s1) * pow(1.0 - E, s2) : 0.0;    /* compPePc:43 */
    if (!possible) pc = 0.0;
    if (pe_out) *pe_out = pe;
    if (pc_out) *pc_out = pc;
    return pe * pc;
}

static inline double log_col(int zn, double C) { return zn ? log(C) : log1p(-C); }
static inline double prior_p0(const spom_model *m) { return (double)(float)m->prior_occ; } /* :66,218 float */
static inline int is_latent(const spom_model *m, int t, int k)
{
    const int o = m->obs[(size_t)t * m->n + k];
    return o == -1 || (m->detect && o == 0);
}

/* colonisation part for one year given S_t: sum over cells with y=0 */
static double ll_col_year(const spom_model *m, const spom_params *p, int t, const uint8_t *y_t,
                          const uint8_t *z_next, const double *S_t)
{
    const int n = m->n, pre = is_pre(m, t), src = needs_src(p, pre);
    double s = 0.0;
    for (int k = 0; k < n; k++) {
        if (y_t[k]) { if (!z_next[k]) s += -INFINITY; continue; }
        const double g = src ? spom_source_term(m, p, k) : 0.0;
        s += log_col(z_next[k], col_prob(p, pre, S_t[k], g));
    }
    return s;
}
static double ll_col_all(const spom_model *m, const spom_params *p, const uint8_t *z, const uint8_t *y,
                         const double *S)
{
    const int n = m->n;
    double s = 0.0;
    for (int t = 0; t + 1 < m->T; t++)
        s += ll_col_year(m, p, t, y + (size_t)t * n, z + (size_t)(t + 1) * n, S + (size_t)t * n);
    return s;
}
static void ext_counts(const spom_model *m, const uint8_t *z, const uint8_t *y, int64_t n10[2],
                       int64_t n11[2], int64_t *bad)
{
    const int n = m->n;
    n10[0] = n10[1] = n11[0] = n11[1] = 0; *bad = 0;
    for (int t = 0; t + 1 < m->T; t++) {
        const int pre = is_pre(m, t);
        for (int k = 0; k < n; k++) {
            const int zz = z[(size_t)t * n + k], yy = y[(size_t)t * n + k];
            if (zz) { if (yy) n11[pre]++; else n10[pre]++; }
            else if (yy) (*bad)++;
        }
    }
}
static inline double xlog(int64_t cnt, double v) { return cnt ? (double)cnt * log(v) : 0.0; }
static double ll_ext_counts(const spom_params *p, const int64_t n10[2], const int64_t n11[2], int64_t bad)
{
    if (bad) return -INFINITY;
    double s = 0.0;
    for (int pre = 0; pre < 2; pre++) {
        const double E = ext_prob(p, pre);
        s += xlog(n10[pre], E) + xlog(n11[pre], 1.0 - E);
    }
    return s;
}
static double ll_prior(const spom_model *m, const uint8_t *z)
{
    const double p0 = prior_p0(m);
    double s = 0.0;
    for (int k = 0; k < m->n; k++)
        if (is_latent(m, 0, k)) s += z[k] ? log(p0) : log(1.0 - p0);   /* main_MIDASPOM.c:249 */
    return s;
}
static void det_counts(const spom_model *m, const uint8_t *z, int64_t *nd, int64_t *nm, int64_t *bad)
{
    *nd = *nm = *bad = 0;
    const size_t tot = (size_t)m->T * m->n;
    for (size_t i = 0; i < tot; i++) {
        const int o = m->obs[i];
        if (o == 1) { if (z[i]) (*nd)++; else (*bad)++; }
        else if (o == 0) { if (z[i]) { if (m->detect) (*nm)++; else (*bad)++; } }
    }
}
static double ll_det_counts(const spom_model *m, const spom_params *p, int64_t nd, int64_t nm, int64_t bad)
{
    if (bad) return -INFINITY;
    if (!m->detect) return 0.0;
    return xlog(nd, p->p) + xlog(nm, 1.0 - p->p);
}

void spom_refresh_S(const spom_model *m, const spom_params *par, const uint8_t *y, double *S)
{
    /* S[t][k] = sum_{l != k, l ascending} w(k<-l) y[t][l]  (main_MIDASPOM.c:351-355 for every year);
     * the weight of a pair is evaluated once and reused by every year -- same per-cell addition
     * order as spom_connectivity, so the values are bit-identical to it. */
    const int n = m->n, nt = m->T - 1;
    uint8_t *any = calloc((size_t)n, 1);
    for (int t = 0; t < nt; t++) for (int l = 0; l < n; l++) any[l] |= y[(size_t)t * n + l];
    for (size_t i = 0; i < (size_t)nt * n; i++) S[i] = 0.0;
    const int scaled = m->area && par->b != 0.0;
    double *aw = malloc((size_t)n * sizeof(double));             /* A_l^b once per source: same value, same product */
    for (int l = 0; l < n; l++) aw[l] = area_factor(m, par->b, l);
    {
        /* y as 0.0/1.0 doubles, years contiguous per source, and a contiguous accumulator per
         * target: acc[t] += w * y01  adds exactly w or 0 (w*1 and w*0 are exact, also when fused),
         * i.e. the same additions in the same order (l ascending) as the scalar reference loop,
         * laid out so the compiler vectorises over t. */
        double *yb = malloc((size_t)n * nt * sizeof(double)), *acc = malloc((size_t)(nt ? nt : 1) * sizeof(double));
        for (int t = 0; t < nt; t++) for (int l = 0; l < n; l++) yb[(size_t)l * nt + t] = y[(size_t)t * n + l] ? 1.0 : 0.0;
        for (int k = 0; k < n; k++) {
            for (int t = 0; t < nt; t++) acc[t] = 0.0;
            for (int l = 0; l < n; l++) {
                if (l == k || !any[l]) continue;
                double w = dist_factor(m, par->alpha, k, l);
                if (scaled) w *= aw[l];
                const double *yl = yb + (size_t)l * nt;
                for (int t = 0; t < nt; t++) acc[t] += w * yl[t];
            }
            for (int t = 0; t < nt; t++) S[(size_t)t * n + k] = acc[t];
        }
        free(yb); free(acc);
    }
    free(any); free(aw);
}

double spom_loglik(const spom_model *m, const spom_params *p, const uint8_t *z, const uint8_t *y,
                   double *parts, double *S_out)
{
    const size_t cells = (size_t)(m->T - 1) * m->n;
    double *S = S_out ? S_out : malloc((cells ? cells : 1) * sizeof(double));
    spom_refresh_S(m, p, y, S);
    int64_t n10[2], n11[2], bad, nd, nm, bad2;
    ext_counts(m, z, y, n10, n11, &bad);
    det_counts(m, z, &nd, &nm, &bad2);
    const double le = ll_ext_counts(p, n10, n11, bad);
    const double lc = ll_col_all(m, p, z, y, S);
    const double lp = ll_prior(m, z);
    const double ld = ll_det_counts(m, p, nd, nm, bad2);
    if (parts) { parts[0] = le; parts[1] = lc; parts[2] = lp; parts[3] = ld; }
    if (!S_out) free(S);
    return le + lc + lp + ld;
}

/* exact data likelihood: forward recursion over years on the completions of the -1 cells
 * (main_MIDASPOM.c:368-392), each one-year transition marginalised over the intermediate state
 * y <= min(z_t, z_t+1) (main_MIDASPOM.c:363). */
static double trans_marginal(const spom_model *m, const spom_params *p, int t, const uint8_t *zo,
                             const uint8_t *zn)
{
    const int n = m->n;
    if (n <= 0) return 1.0;
    int freeidx[64], nf = 0;
    for (int k = 0; k < n; k++) if (zo[k] && zn[k]) { if (nf >= 30) return NAN; freeidx[nf++] = k; }
    uint8_t *y = calloc((size_t)n, 1);
    double tot = 0.0;
    for (uint32_t mask = 0; mask < (1u << nf); mask++) {
        for (int f = 0; f < nf; f++) y[freeidx[f]] = (mask >> f) & 1u;
        tot += spom_transition_prob(m, p, is_pre(m, t), zo, y, zn, NULL, NULL);
    }
    free(y);
    return tot;
}
double spom_marginal_loglik(const spom_model *m, const spom_params *p)
{
    const int n = m->n, T = m->T;
    if (m->detect) return NAN;
    const double p0 = prior_p0(m);
    int maxm = 0;
    for (int t = 0; t < T; t++) { int c = 0; for (int k = 0; k < n; k++) c += m->obs[(size_t)t * n + k] == -1; if (c > maxm) maxm = c; }
    if (maxm > 20) return NAN;
    const size_t cap = (size_t)1 << maxm;
    double *fa = malloc(cap * sizeof(double)), *fb = malloc(cap * sizeof(double));
    uint8_t *za = malloc((size_t)n), *zb = malloc((size_t)n);
    /* year 0 */
    int nm_prev = 0;
    for (int k = 0; k < n; k++) if (m->obs[k] == -1) nm_prev++;
    for (uint32_t s = 0; s < (1u << nm_prev); s++) {
        double pr = 1.0;
        for (int f = 0; f < nm_prev; f++) pr *= ((s >> f) & 1u) ? p0 : 1.0 - p0;
        fa[s] = pr;
    }
    for (int t = 1; t < T; t++) {
        int pidx[32], np = 0, cidx[32], nc = 0;
        for (int k = 0; k < n; k++) { if (m->obs[(size_t)(t - 1) * n + k] == -1) pidx[np++] = k;
                                      if (m->obs[(size_t)t * n + k] == -1) cidx[nc++] = k; }
        for (uint32_t s2 = 0; s2 < (1u << nc); s2++) {
            for (int k = 0; k < n; k++) zb[k] = m->obs[(size_t)t * n + k] == 1;
            for (int f = 0; f < nc; f++) zb[cidx[f]] = (s2 >> f) & 1u;
            double acc = 0.0;
            for (uint32_t s1 = 0; s1 < (1u << np); s1++) {
                for (int k = 0; k < n; k++) za[k] = m->obs[(size_t)(t - 1) * n + k] == 1;
                for (int f = 0; f < np; f++) za[pidx[f]] = (s1 >> f) & 1u;
                acc += fa[s1] * trans_marginal(m, p, t - 1, za, zb);
            }
            fb[s2] = acc;
        }
        double *tmp = fa; fa = fb; fb = tmp;
        nm_prev = nc;
    }
    double L = 0.0;
    for (uint32_t s = 0; s < (1u << nm_prev); s++) L += fa[s];
    free(fa); free(fb); free(za); free(zb);
    return log(L);
}

double spom_flip_delta_bruteforce(const spom_model *m, const spom_params *p, const uint8_t *z,
                                  const uint8_t *y, int t, int k)
{
    const size_t cells = (size_t)(m->T - 1) * m->n;
    uint8_t *y2 = malloc(cells);
    memcpy(y2, y, cells);
    y2[(size_t)t * m->n + k] ^= 1u;
    const double a = spom_loglik(m, p, z, y, NULL, NULL), b = spom_loglik(m, p, z, y2, NULL, NULL);
    free(y2);
    return b - a;
}

/* ------------------------------------------------------------------ sampler */
/* difference of two log terms with the conventions shared with the CUDA engine:
 * (-inf) - (-inf) := 0, finite - (-inf) := +inf */
static inline double ldiff(double alt, double cur)
{
    if (cur == -INFINITY) return alt == -INFINITY ? 0.0 : INFINITY;
    return alt - cur;
}

/* rank-1 evaluation shared by spom_flip_delta and the sweep.  L (nullable) caches the current
 * log_col per cell of the year (0 where y=1).  If commit, S/L/y are updated in place. */
static double flip_eval(const spom_model *m, const spom_params *par, int t, int k, const uint8_t *z_t,
                        const uint8_t *z_n, uint8_t *y_t, double *S_t, double *L_t, int *nocc,
                        double *S_alt_buf, double *L_alt_buf)
{
    const int n = m->n, pre = is_pre(m, t), src = needs_src(par, pre);
    const int cur = y_t[k];
    const int nocc_after = *nocc + (cur ? -1 : 1);
    double acc = 0.0;
    (void)z_t;
    const int scaled = m->area && par->b != 0.0;
    const double awk = area_factor(m, par->b, k);
    for (int q = 0; q < n; q++) {
        if (q == k) { S_alt_buf[q] = S_t[q]; L_alt_buf[q] = 0.0; continue; }
        double w = dist_factor(m, par->alpha, q, k);
        if (scaled) w *= awk;
        double sa = cur ? S_t[q] - w : S_t[q] + w;
        if (nocc_after == 0 || sa < 0.0) sa = 0.0;
        S_alt_buf[q] = sa;
        if (y_t[q]) { L_alt_buf[q] = 0.0; continue; }
        const double g = src ? spom_source_term(m, par, q) : 0.0;
        const double la = log_col(z_n[q], col_prob(par, pre, sa, g));
        const double lc = L_t ? L_t[q] : log_col(z_n[q], col_prob(par, pre, S_t[q], g));
        L_alt_buf[q] = la;
        acc += ldiff(la, lc);
    }
    /* own cell (candidate: z_t[k] = z_n[k] = 1): y=1 -> log(1-E); y=0 -> log E + log C_k */
    const double E = ext_prob(par, pre);
    const double gk = src ? spom_source_term(m, par, k) : 0.0;
    const double l1 = log(1.0 - E);
    const double l0 = log(E) + log(col_prob(par, pre, S_t[k], gk));
    acc += cur ? ldiff(l0, l1) : ldiff(l1, l0);
    if (isnan(acc)) acc = -INFINITY;
    return acc;
}

double spom_flip_delta(const spom_model *m, const spom_params *par, const uint8_t *z, const uint8_t *y,
                       const double *S, int t, int k)
{
    const int n = m->n;
    uint8_t *yt = malloc((size_t)n);
    double *St = malloc((size_t)n * sizeof(double)), *sa = malloc((size_t)n * sizeof(double)),
           *la = malloc((size_t)n * sizeof(double));
    memcpy(yt, y + (size_t)t * n, (size_t)n);
    memcpy(St, S + (size_t)t * n, (size_t)n * sizeof(double));
    int nocc = 0;
    for (int q = 0; q < n; q++) nocc += yt[q];
    const double d = flip_eval(m, par, t, k, z + (size_t)t * n, z + (size_t)(t + 1) * n, yt, St, NULL, &nocc, sa, la);
    free(yt); free(St); free(sa); free(la);
    return d;
}

void spom_init_chain(const spom_model *m, const spom_sampler_cfg *cfg, uint64_t seed, uint32_t chain,
                     int disperse, spom_params *par, double *lsig, uint8_t *z, uint8_t *y, double *S)
{
    const int n = m->n, T = m->T;
    uint32_t r[4];
    if (disperse) {
        spom_rng(seed, chain, 0, RK_INIT_PARAM, 0, 0, r);
        if (cfg->sample_e) par->e = cfg->e_min + spom_u01(r[0]) * (cfg->e_max - cfg->e_min);
        if (cfg->sample_c) par->c = cfg->c_min + spom_u01(r[1]) * (cfg->c_max - cfg->c_min);
        if (cfg->sample_alpha) par->alpha = cfg->alpha_min * pow(cfg->alpha_max / cfg->alpha_min, spom_u01(r[2]));
        if (cfg->sample_b) par->b = cfg->b_min + spom_u01(r[3]) * (cfg->b_max - cfg->b_min);
        spom_rng(seed, chain, 0, RK_INIT_PARAM, 1, 0, r);
        if (cfg->sample_p) par->p = cfg->p_min + spom_u01(r[0]) * (cfg->p_max - cfg->p_min);
        if (cfg->sample_K) par->K = cfg->K_min * pow(cfg->K_max / cfg->K_min, spom_u01(r[1]));
        if (cfg->sample_Ksrc) par->Ksrc = cfg->Ksrc_min * pow(cfg->Ksrc_max / cfg->Ksrc_min, spom_u01(r[2]));
        if (cfg->sample_dsrc) par->dsrc = cfg->dsrc_min + spom_u01(r[3]) * (cfg->dsrc_max - cfg->dsrc_min);
    }
    for (int t = 0; t < T; t++)
        for (int k = 0; k < n; k++) {
            const int o = m->obs[(size_t)t * n + k];
            uint8_t v;
            if (o == 1) v = 1;
            else if (o == 0) v = 0;
            else if (disperse) { spom_rng(seed, chain, 0, RK_INIT_Z, (uint32_t)k, (uint32_t)t, r); v = spom_u01(r[0]) < 0.5; }
            else v = 1;
            z[(size_t)t * n + k] = v;
        }
    for (int t = 0; t + 1 < T; t++)
        for (int k = 0; k < n; k++)
            y[(size_t)t * n + k] = z[(size_t)t * n + k] && z[(size_t)(t + 1) * n + k];
    spom_refresh_S(m, par, y, S);
    if (disperse) {
        /* a dispersed start must be a possible state: an empty cell next year (y=0, z'=0) needs C < 1,
         * i.e. c < 1 / max(K_t S + Ksrc_t g).  Pull c below that bound (no effect on feasible draws). */
        double smax = 0.0;
        for (int t = 0; t + 1 < T; t++) {
            const int pre = is_pre(m, t), src = needs_src(par, pre);
            for (int k = 0; k < n; k++) {
                const size_t i = (size_t)t * n + k;
                if (y[i] || z[i + n]) continue;
                const double g = src ? spom_source_term(m, par, k) : 0.0;
                const double v = pre ? par->K * S[i] + par->Ksrc * g : S[i];
                if (v > smax) smax = v;
            }
        }
        if (smax > 0.0 && par->c * smax >= 1.0) { par->c = 0.5 / smax; if (par->c < cfg->c_min) par->c = cfg->c_min; }
    }
    lsig[0] = log(0.05); lsig[1] = log(0.1 * par->c); lsig[2] = log(0.05); lsig[3] = log(0.05); lsig[4] = log(0.05);
    lsig[5] = log(0.1); lsig[6] = log(0.1); lsig[7] = log(0.1 * (cfg->dsrc_max > cfg->dsrc_min ? cfg->dsrc_max - cfg->dsrc_min : 1.0));
}

/* Gibbs draws compare logit(u) with the log-odds: u < 1/(1+exp(-d))  <=>  log(u/(1-u)) < d */
static inline double logit_u(uint32_t x) { const double u = spom_u01(x); return log(u) - log1p(-u); }
static inline double adapt_gain(uint32_t sweep) { return 1.0 / pow((double)sweep + 1.0, 0.6); }
static inline int mh_accept(double logu, double d) { if (isnan(d)) return 0; return logu < d; }

int64_t spom_sweep(const spom_model *m, const spom_sampler_cfg *cfg, uint64_t seed, uint32_t chain,
                   uint32_t sweep, spom_params *par, double *lsig, uint8_t *z, uint8_t *y, double *S,
                   double *draw, int64_t y_flip_limit, double *phase_s)
{
    const int n = m->n, T = m->T;
    const double t_begin = now_s();
    double t_yscan = 0.0;
    const size_t cells = (size_t)(T - 1) * n;
    const int adapting = sweep < (uint32_t)cfg->n_adapt;
    const double gain = adapt_gain(sweep);
    uint32_t r[4];
    int64_t visited = 0;

    /* A: refresh S; joint random-walk MH on (log alpha, b) */
    spom_refresh_S(m, par, y, S);
    double llc = ll_col_all(m, par, z, y, S);
    if (cfg->sample_alpha || cfg->sample_b) {
        double n1, n2;
        spom_rng(seed, chain, sweep, RK_AB, 0, 0, r);
        box_muller(r[0], r[1], &n1, &n2);
        const double logu = log(spom_u01(r[2]));
        spom_params prop = *par;
        if (cfg->sample_alpha) prop.alpha = par->alpha * exp(exp(lsig[2]) * n1);
        if (cfg->sample_b) prop.b = par->b + exp(lsig[3]) * n2;
        int acc = 0;
        if (prop.alpha >= cfg->alpha_min && prop.alpha <= cfg->alpha_max && prop.b >= cfg->b_min && prop.b <= cfg->b_max) {
            double *S2 = malloc((cells ? cells : 1) * sizeof(double));
            spom_refresh_S(m, &prop, y, S2);
            /* ridge move: (alpha, b) mostly rescale S, and c S is what the data pin down.  Propose
             * c' = c mean(S)/mean(S') together with (alpha', b'): a deterministic, reversible shift of log c
             * (mean S depends on (alpha, b, y) only), so the Hastings ratio only gains the Jacobian c'/c of the
             * uniform-in-c prior seen in log c. */
            double ljac = 0.0;
            if (cfg->sample_c && cells) {
                double m1 = 0.0, m2 = 0.0;
                for (size_t i = 0; i < cells; i++) { m1 += S[i]; m2 += S2[i]; }
                if (m1 > 0.0 && m2 > 0.0) { prop.c = par->c * (m1 / m2); ljac = log(prop.c / par->c); }
            }
            if (prop.c >= cfg->c_min && prop.c <= cfg->c_max) {
                const double llc2 = ll_col_all(m, &prop, z, y, S2);
                if (mh_accept(logu, llc2 - llc + ljac)) { acc = 1; *par = prop; memcpy(S, S2, cells * sizeof(double)); llc = llc2; }
            }
            free(S2);
        }
        if (adapting) {
            if (cfg->sample_alpha) lsig[2] += gain * (acc - 0.30);
            if (cfg->sample_b) lsig[3] += gain * (acc - 0.30);
        }
    }
    /* B: random-walk MH on c */
    if (cfg->sample_c)
        for (int s = 0; s < cfg->n_c_steps; s++) {
            double n1, n2;
            spom_rng(seed, chain, sweep, RK_C, (uint32_t)s, 0, r);
            box_muller(r[0], r[1], &n1, &n2);
            const double logu = log(spom_u01(r[2]));
            spom_params prop = *par;
            prop.c = par->c + exp(lsig[1]) * n1;
            int acc = 0;
            if (prop.c >= cfg->c_min && prop.c <= cfg->c_max) {
                const double llc2 = ll_col_all(m, &prop, z, y, S);
                if (mh_accept(logu, llc2 - llc)) { acc = 1; *par = prop; llc = llc2; }
            }
            if (adapting) lsig[1] += gain * (acc - 0.44);
        }
    /* V: variant parameters -- K (pre-event scaling, log-uniform), Ksrc (source size, log-uniform), dsrc (source
     * distance unit, uniform): random-walk MH, extinction part from the counts, colonisation part over the cells */
    if (cfg->sample_K || cfg->sample_Ksrc || cfg->sample_dsrc) {
        int64_t vn10[2], vn11[2], vbad;
        ext_counts(m, z, y, vn10, vn11, &vbad);
        for (int which = 0; which < 3; which++) {
            const int on = which == 0 ? cfg->sample_K : which == 1 ? cfg->sample_Ksrc : cfg->sample_dsrc;
            if (!on) continue;
            for (int s = 0; s < cfg->n_v_steps; s++) {
                double n1, n2;
                spom_rng(seed, chain, sweep, RK_K + (uint32_t)which, (uint32_t)s, 0, r);
                box_muller(r[0], r[1], &n1, &n2);
                const double logu = log(spom_u01(r[2]));
                spom_params prop = *par;
                int inb;
                if (which == 0) { prop.K = par->K * exp(exp(lsig[5]) * n1); inb = prop.K >= cfg->K_min && prop.K <= cfg->K_max; }
                else if (which == 1) { prop.Ksrc = par->Ksrc * exp(exp(lsig[6]) * n1); inb = prop.Ksrc >= cfg->Ksrc_min && prop.Ksrc <= cfg->Ksrc_max; }
                else { prop.dsrc = par->dsrc + exp(lsig[7]) * n1; inb = prop.dsrc >= cfg->dsrc_min && prop.dsrc <= cfg->dsrc_max; }
                int acc = 0;
                if (inb) {
                    const double llc2 = ll_col_all(m, &prop, z, y, S);
                    const double d = (llc2 - llc) + (ll_ext_counts(&prop, vn10, vn11, vbad) - ll_ext_counts(par, vn10, vn11, vbad));
                    if (mh_accept(logu, d)) { acc = 1; *par = prop; llc = llc2; }
                }
                if (adapting) lsig[5 + which] += gain * (acc - 0.44);
            }
        }
    }
    /* D: latent z cells -- conditionally independent given y (S depends on y only) */
    if (cfg->update_z) {
        const double p0 = prior_p0(m);
        for (int t = 0; t < T; t++)
            for (int k = 0; k < n; k++) {
                if (!is_latent(m, t, k)) continue;
                const size_t i = (size_t)t * n + k;
                const int forced = (t > 0 && y[i - n]) || (t + 1 < T && y[i]);
                if (forced) { z[i] = 1; continue; }
                double lo = 0.0;
                if (t > 0) {
                    const int pre = is_pre(m, t - 1);
                    const double g = needs_src(par, pre) ? spom_source_term(m, par, k) : 0.0;
                    const double C = col_prob(par, pre, S[i - n], g);
                    lo += log(C) - log1p(-C);
                }
                if (t + 1 < T) lo += log(ext_prob(par, is_pre(m, t)));
                if (t == 0) lo += log(p0) - log(1.0 - p0);
                if (m->detect && m->obs[i] == 0) lo += log(1.0 - par->p);
                spom_rng(seed, chain, sweep, RK_Z, (uint32_t)k, (uint32_t)t, r);
                z[i] = isnan(lo) ? 0 : (logit_u(r[0]) < lo);
            }
    }
    /* E: y_t | z -- systematic scan over candidate cells, rank-1 update of S_t */
    if (cfg->update_y) {
        const double t_y0 = now_s();
        double *L = malloc((size_t)n * sizeof(double)), *sa = malloc((size_t)n * sizeof(double)),
               *la = malloc((size_t)n * sizeof(double));
        int32_t *order = malloc((size_t)n * sizeof(int32_t));
        spom_scan_order(m, order);
        for (int t = 0; t + 1 < T; t++) {
            uint8_t *y_t = y + (size_t)t * n;
            const uint8_t *z_t = z + (size_t)t * n, *z_n = z + (size_t)(t + 1) * n;
            double *S_t = S + (size_t)t * n;
            const int pre = is_pre(m, t), src = needs_src(par, pre);
            int nocc = 0;
            for (int k = 0; k < n; k++) {
                nocc += y_t[k];
                const double g = src ? spom_source_term(m, par, k) : 0.0;
                L[k] = y_t[k] ? 0.0 : log_col(z_n[k], col_prob(par, pre, S_t[k], g));
            }
            int64_t vis_t = 0;
            for (int sl = 0; sl < n; sl++) {
                const int k = order[sl];
                if (!(z_t[k] && z_n[k])) continue;
                if (y_flip_limit >= 0 && vis_t >= y_flip_limit) break;
                vis_t++;
                const double d = flip_eval(m, par, t, k, z_t, z_n, y_t, S_t, L, &nocc, sa, la);
                spom_rng(seed, chain, sweep, RK_Y, (uint32_t)k, (uint32_t)t, r);
                if (logit_u(r[0]) < d) {   /* u < sigmoid(d) */
                    const int cur = y_t[k];
                    memcpy(S_t, sa, (size_t)n * sizeof(double));
                    for (int q = 0; q < n; q++) if (q != k && !y_t[q]) L[q] = la[q];
                    y_t[k] = (uint8_t)!cur;
                    nocc += cur ? -1 : 1;
                    if (cur) { const double g = src ? spom_source_term(m, par, k) : 0.0;
                               L[k] = log_col(1, col_prob(par, pre, S_t[k], g)); }
                    else L[k] = 0.0;
                }
            }
            visited += vis_t;
        }
        free(L); free(sa); free(la); free(order);
        t_yscan = now_s() - t_y0;
    }
    /* C: random-walk MH on e from the sufficient counts */
    int64_t n10[2], n11[2], bad;
    ext_counts(m, z, y, n10, n11, &bad);
    if (cfg->sample_e) {
        double lle = ll_ext_counts(par, n10, n11, bad);
        for (int s = 0; s < cfg->n_e_steps; s++) {
            double n1, n2;
            spom_rng(seed, chain, sweep, RK_E, (uint32_t)s, 0, r);
            box_muller(r[0], r[1], &n1, &n2);
            const double logu = log(spom_u01(r[2]));
            spom_params prop = *par;
            prop.e = par->e + exp(lsig[0]) * n1;
            int acc = 0;
            if (prop.e >= cfg->e_min && prop.e <= cfg->e_max) {
                const double l2 = ll_ext_counts(&prop, n10, n11, bad);
                if (mh_accept(logu, l2 - lle)) { acc = 1; *par = prop; lle = l2; }
            }
            if (adapting) lsig[0] += gain * (acc - 0.44);
        }
    }
    /* F: detection probability */
    int64_t nd, nm, bad2;
    det_counts(m, z, &nd, &nm, &bad2);
    if (cfg->sample_p && m->detect) {
        double lld = ll_det_counts(m, par, nd, nm, bad2);
        for (int s = 0; s < cfg->n_e_steps; s++) {
            double n1, n2;
            spom_rng(seed, chain, sweep, RK_P, (uint32_t)s, 0, r);
            box_muller(r[0], r[1], &n1, &n2);
            const double logu = log(spom_u01(r[2]));
            spom_params prop = *par;
            prop.p = par->p + exp(lsig[4]) * n1;
            int acc = 0;
            if (prop.p >= cfg->p_min && prop.p <= cfg->p_max) {
                const double l2 = ll_det_counts(m, &prop, nd, nm, bad2);
                if (mh_accept(logu, l2 - lld)) { acc = 1; *par = prop; lld = l2; }
            }
            if (adapting) lsig[4] += gain * (acc - 0.44);
        }
    }
    /* record */
    if (draw) {
        int64_t sy = 0, sz = 0;
        for (size_t i = 0; i < cells; i++) sy += y[i];
        for (size_t i = 0; i < (size_t)T * n; i++) sz += z[i];
        draw[0] = par->e; draw[1] = par->c; draw[2] = par->alpha; draw[3] = par->b; draw[4] = par->p;
        draw[5] = ll_ext_counts(par, n10, n11, bad) + ll_col_all(m, par, z, y, S) + ll_prior(m, z)
                + ll_det_counts(m, par, nd, nm, bad2);
        draw[6] = (double)sy; draw[7] = (double)sz; draw[8] = par->K; draw[9] = par->Ksrc; draw[10] = par->dsrc;
    }
    if (phase_s) { phase_s[0] = now_s() - t_begin - t_yscan; phase_s[1] = t_yscan; }
    return visited;
}

int64_t spom_sweep_chains(const spom_model *m, const spom_sampler_cfg *cfg, uint64_t seed, int nchains,
                          uint32_t chain0, uint32_t sweep, spom_params *par, double *lsig, uint8_t *z,
                          uint8_t *y, double *S, double *draws, int64_t y_flip_limit, int nthreads, double *phase_s)
{
    const size_t zc = (size_t)m->T * m->n, yc = (size_t)(m->T - 1) * m->n;
    int64_t total = 0;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads > 0 ? nthreads : omp_get_max_threads()) reduction(+ : total)
#endif
    for (int c = 0; c < nchains; c++)
        total += spom_sweep(m, cfg, seed, chain0 + (uint32_t)c, sweep, par + c, lsig + (size_t)c * SPOM_NLSIG,
                            z + c * zc, y + c * yc, S + c * yc, draws ? draws + (size_t)c * SPOM_NDRAW : NULL,
                            y_flip_limit, phase_s ? phase_s + 2 * (size_t)c : NULL);
    return total;
}
int spom_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------ forward simulator (simpij) */
/* future.c:64-110: survive iff u > E (:78), colonise iff u < C (:100); C uses K*S + source (:90-97).
 * era flags (m->era, length nyears) select which years apply K / Ksrc; NULL => none. */
void spom_simulate(const spom_model *m, const spom_params *p, uint64_t seed, uint32_t sim_id,
                   const uint8_t *z0, int nyears, uint8_t *z_out)
{
    const int n = m->n;
    uint8_t *y = malloc((size_t)n);
    double *S = malloc((size_t)n * sizeof(double));
    uint32_t r[4];
    memcpy(z_out, z0, (size_t)n);
    for (int t = 0; t < nyears; t++) {
        const uint8_t *zo = z_out + (size_t)t * n;
        uint8_t *zn = z_out + (size_t)(t + 1) * n;
        const int pre = is_pre(m, t), src = needs_src(p, pre);
        const double E = ext_prob(p, pre);
        for (int k = 0; k < n; k++) {
            spom_rng(seed, sim_id, (uint32_t)t, RK_SIM_EXT, (uint32_t)k, 0, r);
            y[k] = zo[k] && (spom_u01(r[0]) > E);
        }
        spom_connectivity(m, p->alpha, p->b, y, S);
        for (int k = 0; k < n; k++) {
            const double g = src ? spom_source_term(m, p, k) : 0.0;
            const double C = col_prob(p, pre, S[k], g);
            spom_rng(seed, sim_id, (uint32_t)t, RK_SIM_COL, (uint32_t)k, 0, r);
            zn[k] = y[k] ? 1 : (spom_u01(r[0]) < C);
        }
    }
    free(y); free(S);
}
