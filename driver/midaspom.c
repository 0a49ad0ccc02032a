/* driver/midaspom.c -- command-line driver with the flags and file formats of MIDASPOM.out
 * (reference: sources/main_MIDASPOM.c), calling the B200 engine through the C ABI
 * (include/libmidaspom_cuda.h) where the reference has its grid loops, compPePc and cblas_dgemm
 * (main_MIDASPOM.c:341-395).
 *
 *   midaspom -m 400 -d 100 -i occupancies.txt -o posterior.txt          (same flags: m p d i o s l u)
 *
 * Default mode = the reference's computation: exact likelihood on the s x s grid of (e, c)
 * (mp_exact_posterior), trapezoid normalisation, table written with "%.20lf\t".
 * Extra flags (letters the reference does not use) select the sampler instead of the grid:
 *   -n SWEEPS   run data-augmented MCMC (mp_sweep) and histogram the (e, c) draws onto the same
 *               s x s grid, so the reference's R scripts read the file unchanged
 *   -c CHAINS   chains (default 8)      -b BURNIN  sweeps discarded (default SWEEPS/5)
 *   -r SEED     Philox seed (default 1) -g DEVICE  CUDA ordinal (default 0)
 *   -t FILE     also write the raw draws (one row per sweep and chain)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "../include/libmidaspom_cuda.h"
#include "mp_io.h"

static int run_grid(const int8_t *obs, int tmax, int n, double a, double d, float prioroc, int nstep, double ecmin,
                    double ecmax, int device, const char *fout)
{
    double *lik = (double *)malloc((size_t)nstep * nstep * sizeof(double));
    double ltot = 0.0;
    int info[4];
    printf("Starting parallel likelihood computation\n");
    int rc = mp_exact_posterior(device, obs, tmax, n, a, d, prioroc, nstep, ecmin, ecmax, lik, &ltot, info);
    if (rc != MP_OK) { fprintf(stderr, "mp_exact_posterior failed (%d): %s\n", rc, mp_exact_last_error()); free(lik); return 2; }
    printf("Number of states to compute: %d\n", info[1]);
    for (int ie = 0; ie < nstep; ie++) printf("%.2f%% done\n", ((float)ie + 1) * 100.0 / nstep);
    printf("end likelihood computation\n");
    printf("Total log-likelihood=%.5lf\n", ltot);
    printf("Writing output in file %s... ", fout);
    for (long i = 0; i < (long)nstep * nstep; i++) lik[i] = exp(lik[i] - ltot);     /* main_MIDASPOM.c:432 */
    rc = mp_write_table(fout, lik, nstep, nstep);
    free(lik);
    if (rc) return 3;
    printf("done\n");
    return 0;
}

static int run_mcmc(const int8_t *obs, int tmax, int n, double a, double d, float prioroc, int nstep, double ecmin,
                    double ecmax, int device, const char *fout, int nsweeps, int nchains, int burn, unsigned long long seed,
                    const char *fdraws)
{
    mp_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.n_patches = n; cfg.n_years = tmax; cfg.n_chains = nchains; cfg.precision = MP_FP64; cfg.device = device;
    cfg.max_draws = nsweeps; cfg.seed = seed; cfg.prior_occ = prioroc;
    mp_engine *h = NULL;
    if (mp_create(&cfg, &h) != MP_OK) { fprintf(stderr, "mp_create: %s\n", mp_last_error(NULL)); return 2; }
    int rc = 2;
    mp_sampler_config sc;
    memset(&sc, 0, sizeof sc);
    sc.e_min = ecmin; sc.e_max = ecmax; sc.c_min = ecmin; sc.c_max = ecmax;     /* uniform prior on the grid's square (:312-319) */
    sc.alpha_min = a; sc.alpha_max = a; sc.b_min = 0; sc.b_max = 0; sc.p_min = 1; sc.p_max = 1;
    sc.sample_e = 1; sc.sample_c = 1; sc.n_e_steps = 4; sc.n_c_steps = 2; sc.n_adapt = burn; sc.update_z = 1; sc.update_y = 1;
    mp_params *par = (mp_params *)calloc((size_t)nchains, sizeof(mp_params));
    for (int c = 0; c < nchains; c++) { par[c].e = 0.5 * (ecmin + ecmax); par[c].c = 0.5 * (ecmin + ecmax); par[c].alpha = a; par[c].p = 1; par[c].K = 1; }
    double *draws = (double *)malloc((size_t)nsweeps * nchains * MP_NDRAW * sizeof(double));
    double *dens = (double *)calloc((size_t)nstep * nstep, sizeof(double));
    if (mp_set_landscape_linear(h, d, NULL) || mp_set_source_units(h, NULL) || mp_set_observations(h, obs) ||
        mp_set_params(h, par) || mp_init_chains(h, &sc, 1)) { fprintf(stderr, "setup: %s\n", mp_last_error(h)); goto out; }
    printf("Starting MCMC: %d chains x %d sweeps (%d burn-in)\n", nchains, nsweeps, burn);
    if (mp_sweep(h, nsweeps) || mp_synchronize(h) || mp_get_draws(h, 0, nsweeps, draws)) { fprintf(stderr, "sweep: %s\n", mp_last_error(h)); goto out; }
    {
        /* histogram on the reference's grid: node i collects draws within half a window, edge nodes half
         * as wide -- the cells the trapezoid weights of :418-421 stand for */
        const double win = (ecmax - ecmin) / (nstep - 1);
        double me = 0, mc = 0; long cnt = 0;
        for (int s = burn; s < nsweeps; s++)
            for (int c = 0; c < nchains; c++) {
                const double *dr = draws + ((size_t)s * nchains + c) * MP_NDRAW;
                int ie = (int)floor((dr[0] - ecmin) / win + 0.5), ic = (int)floor((dr[1] - ecmin) / win + 0.5);
                ie = ie < 0 ? 0 : (ie > nstep - 1 ? nstep - 1 : ie);
                ic = ic < 0 ? 0 : (ic > nstep - 1 ? nstep - 1 : ic);
                dens[(size_t)ie * nstep + ic] += 1.0; me += dr[0]; mc += dr[1]; cnt++;
            }
        for (int i = 0; i < nstep; i++)
            for (int j = 0; j < nstep; j++) {
                double w = win * win;
                if (i == 0 || i == nstep - 1) w *= 0.5;
                if (j == 0 || j == nstep - 1) w *= 0.5;
                dens[(size_t)i * nstep + j] /= (double)cnt * w;        /* density: trapezoid sum * win^2 == 1 */
            }
        printf("posterior means: e=%.5lf c=%.5lf (%ld draws)\n", me / cnt, mc / cnt, cnt);
    }
    printf("Writing output in file %s... ", fout);
    if (mp_write_table(fout, dens, nstep, nstep)) goto out;
    printf("done\n");
    if (fdraws) {
        FILE *f = fopen(fdraws, "wb");
        if (f) {
            fprintf(f, "sweep\tchain\te\tc\talpha\tb\tp\tloglik\tny1\tnz1\n");
            for (int s = 0; s < nsweeps; s++)
                for (int c = 0; c < nchains; c++) {
                    const double *dr = draws + ((size_t)s * nchains + c) * MP_NDRAW;
                    fprintf(f, "%d\t%d\t%.10g\t%.10g\t%.10g\t%.10g\t%.10g\t%.10g\t%.0f\t%.0f\n", s, c, dr[0], dr[1], dr[2], dr[3], dr[4], dr[5], dr[6], dr[7]);
                }
            fclose(f);
        }
    }
    rc = 0;
out:
    free(par); free(draws); free(dens);
    mp_destroy(h);
    return rc;
}

int main(int argc, char **argv)
{
    printf("------ MIDASPOM on B200 (midaspom_b200; model and formats of MIDASPOM beta, Alcala, Cole & Rosenberg) ------\n");
    float prioroc = 0.5f;                 /* defaults of main_MIDASPOM.c:66-73 */
    const char *fname = "input.txt", *fout = "posterior.txt", *fdraws = NULL;
    double d = 100, a = 1.0 / 400, ecmin = 0.0, ecmax = 1.0;
    int nstep = 101, nsweeps = 0, nchains = 8, burn = -1, device = 0, c;
    unsigned long long seed = 1;
    opterr = 0;
    while ((c = getopt(argc, argv, "m:p:d:i:o:s:l:u:n:c:b:r:g:t:")) != -1)
        switch (c) {
        case 'm': a = 1.0 / atof(optarg); break;
        case 'p': prioroc = (float)atof(optarg); break;
        case 'd': d = atof(optarg); break;
        case 'i': fname = optarg; break;
        case 'o': fout = optarg; break;
        case 's': nstep = atoi(optarg); break;
        case 'l': ecmin = atof(optarg); break;
        case 'u': ecmax = atof(optarg); break;
        case 'n': nsweeps = atoi(optarg); break;
        case 'c': nchains = atoi(optarg); break;
        case 'b': burn = atoi(optarg); break;
        case 'r': seed = strtoull(optarg, NULL, 10); break;
        case 'g': device = atoi(optarg); break;
        case 't': fdraws = optarg; break;
        default: fprintf(stderr, "Unknown option `-%c'.\n", optopt); return 1;
        }
    if (nstep < 2) { fprintf(stderr, "-s must be at least 2\n"); return 1; }
    const double win = (ecmax - ecmin) / (nstep - 1);
    printf("Parameters for numerical approximation of the posterior density:\n\tWindow size=%lf, number of steps=%d\n", win, nstep);
    time_t start = time(NULL);
    printf("Reading observations from file %s... ", fname);
    int8_t *obs = NULL; int n = 0, tmax = 0;
    if (mp_read_occupancy(fname, &obs, &n, &tmax)) return 1;
    printf("done\n");
    printf("Number of habitat patches: %d\nNumber of sampled years: %d\n", n, tmax);
    if (tmax < 2) { fprintf(stderr, "need at least two sampled years\n"); free(obs); return 1; }
    printf("Dispersal matrix:\n");                                                  /* main_MIDASPOM.c:180-195 */
    for (int i = 0; i < n && n <= 64; i++) {
        for (int j = 0; j < n; j++) printf("%.3f ", i == j ? 0.0 : exp(-a * abs(j - i) * d));
        printf("\n");
    }
    printf("Input occupancy data:\n");
    for (int i = 0; i < tmax; i++) {
        printf("Year %d: ", i);
        for (int j = 0; j < n && n <= 256; j++) printf("%d ", obs[(size_t)i * n + j]);
        printf("\n");
    }
    printf("Number of possible states per year:\n");
    for (int i = 0; i < tmax; i++) {
        int s1 = 0;
        for (int j = 0; j < n; j++) s1 += obs[(size_t)i * n + j] == -1;
        if (s1 < 31) printf("Year %d: %d\n", i, 1 << s1); else printf("Year %d: 2^%d\n", i, s1);
    }
    setbuf(stdout, NULL);
    int rc;
    if (nsweeps > 0) {
        if (burn < 0) burn = nsweeps / 5;
        rc = run_mcmc(obs, tmax, n, a, d, prioroc, nstep, ecmin, ecmax, device, fout, nsweeps, nchains, burn, seed, fdraws);
    } else rc = run_grid(obs, tmax, n, a, d, prioroc, nstep, ecmin, ecmax, device, fout);
    free(obs);
    printf(" Total running time: %.2lf min\n", difftime(time(NULL), start) / 60.0);
    return rc;
}
