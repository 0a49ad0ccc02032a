/* driver/midaspom_future.c -- flags and file formats of MIDASPOM_future.out
 * (reference: sources/main_MIDASPOM_future.c), forward simulation on the B200 engine.
 *
 *   midaspom_future -a 50 -m 400 -d 100 -i occupancies.txt -q posterior.txt -o pext_future.txt [-S 1 -s 500] [-D K]
 *
 * As the reference (future.c:359-386): for each of -n simulations draw a grid cell of the (e, c)
 * posterior table with its trapezoid weights, start from a random completion of the LAST survey row,
 * simulate -a years with simpij (E = min(1, e/K_D), C = min(1, c (K_D S + exp(-a (k+1) d_S) K_S)))
 * and count, per future year, the simulations in which every patch is empty.  Output: -a integers "%d\t".
 * Differences: the grid value is ie * win (win from the table size) instead of the hard-coded
 * ie * 0.01 (future.c:370-371; identical for the default 101 x 101 table), and the random numbers are
 * Philox streams (-r seed) instead of srand(time(NULL)).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "../include/libmidaspom_cuda.h"
#include "mp_io.h"

int main(int argc, char **argv)
{
    printf("------ MIDASPOM future on B200 (midaspom_b200) ------\n");
    int tfut = 50, nsimul = 10000, device = 0, c;                       /* defaults of future.c:123-133 */
    double KS = 0, dS = 200, KD = 1, a = 1.0 / 400, d = 200;
    float prioroc = 0.5f;
    const char *finame = "posterior.txt", *fname = "input.txt", *fout = "pext_future.txt";
    unsigned long long seed = 1;
    opterr = 0;
    while ((c = getopt(argc, argv, "n:a:m:p:q:d:i:o:S:s:D:r:g:")) != -1)
        switch (c) {
        case 'n': nsimul = atoi(optarg); break;
        case 'a': tfut = atoi(optarg); break;
        case 'm': a = 1.0 / atof(optarg); break;
        case 'p': prioroc = (float)atof(optarg); break;
        case 'q': finame = optarg; break;
        case 'd': d = atof(optarg); break;
        case 'i': fname = optarg; break;
        case 'o': fout = optarg; break;
        case 'S': KS = atof(optarg); break;
        case 's': dS = atof(optarg); break;
        case 'D': KD = atof(optarg); break;
        case 'r': seed = strtoull(optarg, NULL, 10); break;
        case 'g': device = atoi(optarg); break;
        default: fprintf(stderr, "Unknown option `-%c'.\n", optopt); return 1;
        }
    printf("%d years in the future\n", tfut);
    time_t start = time(NULL);
    printf("Reading observations from file %s... ", fname);
    int8_t *obs = NULL; int n = 0, tmax = 0;
    if (mp_read_occupancy(fname, &obs, &n, &tmax)) return 1;
    const int8_t *pend = obs + (size_t)(tmax - 1) * n;                  /* the last row read (future.c:207-215) */
    printf("Last occupancy survey:\n");
    for (int j = 0; j < n; j++) printf("%d ", pend[j]);
    printf("\n\n done\n");
    printf("Number of habitat patches: %d\nNumber of sampled years: %d\n", n, tmax);
    printf("Reading posterior distribution from file %s... ", finame);
    double *post = NULL; int nec = 0;
    if (mp_read_square_table(finame, &post, &nec)) { free(obs); return 1; }
    printf("%dX%d posterior distribution\n", nec, nec);
    int miss[32], s1 = 0;
    for (int j = 0; j < n; j++) if (pend[j] == -1) { if (s1 >= 30) { fprintf(stderr, "too many missing cells\n"); return 1; } miss[s1++] = j; }
    const int npstates = 1 << s1;
    printf("npstates = %d\n", npstates);
    (void)prioroc;                                                      /* the reference computes priorst and frees it unused (:326,337) */

    /* cumulative trapezoid weights of the table (future.c:362-369) */
    double *cum = (double *)malloc((size_t)nec * nec * sizeof(double)), tot = 0;
    for (int ie = 0; ie < nec; ie++)
        for (int ic = 0; ic < nec; ic++) {
            double w = 1.0;
            if (ie == 0 || ie == nec - 1) w *= 0.5;
            if (ic == 0 || ic == nec - 1) w *= 0.5;
            tot += w * post[(size_t)ie * nec + ic];
            cum[(size_t)ie * nec + ic] = tot;
        }
    const double win = 1.0 / (nec - 1);
    mp_params *par = (mp_params *)calloc((size_t)nsimul, sizeof(mp_params));
    uint8_t *z0 = (uint8_t *)calloc((size_t)nsimul * n, 1);
    uint64_t rs = seed * 0x9E3779B97F4A7C15ull + 12345;
    for (int i = 0; i < nsimul; i++) {
        const double pec = tot * mp_unif(&rs);
        long lo = 0, hi = (long)nec * nec - 1;                          /* first cell with pec < cumulative weight */
        while (lo < hi) { long mid = (lo + hi) / 2; if (pec < cum[mid]) hi = mid; else lo = mid + 1; }
        par[i].e = (double)(lo / nec) * win; par[i].c = (double)(lo % nec) * win;
        par[i].alpha = a; par[i].b = 0; par[i].p = 1; par[i].K = KD; par[i].Ksrc = KS; par[i].dsrc = dS;
        const int init = (int)(mp_splitmix(&rs) % (uint64_t)npstates);  /* future.c:378 */
        for (int j = 0; j < n; j++) z0[(size_t)i * n + j] = pend[j] == 1;
        for (int m = 0; m < s1; m++) z0[(size_t)i * n + miss[m]] = (uint8_t)((init / (npstates >> (m + 1))) % 2);   /* :311-312 */
    }
    int rc = 2;
    int32_t *occ = (int32_t *)malloc((size_t)nsimul * (tfut + 1) * sizeof(int32_t));
    mp_config cfg; memset(&cfg, 0, sizeof cfg);
    cfg.n_patches = n; cfg.n_years = 2; cfg.n_chains = 1; cfg.precision = MP_FP64; cfg.device = device; cfg.seed = seed; cfg.prior_occ = prioroc;
    mp_engine *h = NULL;
    if (mp_create(&cfg, &h) != MP_OK) { fprintf(stderr, "mp_create: %s\n", mp_last_error(NULL)); goto out; }
    printf("Starting likelihood computation\n");
    if (mp_set_landscape_linear(h, d, NULL) || mp_set_source_units(h, NULL) ||
        mp_simulate_ensemble(h, par, z0, tfut, nsimul, seed, 1, NULL, occ)) { fprintf(stderr, "simulate: %s\n", mp_last_error(h)); goto out; }
    printf("end likelihood computation\n");
    printf("Writing on file %s... ", fout);
    {
        FILE *f = fopen(fout, "wb");
        if (!f) { fprintf(stderr, "cannot write %s\n", fout); goto out; }
        for (int t = 0; t < tfut; t++) {
            int ext = 0;
            for (int i = 0; i < nsimul; i++) ext += occ[(size_t)i * (tfut + 1) + t + 1] == 0;     /* jtmp==0 (future.c:383) */
            fprintf(f, "%d\t", ext);
        }
        fclose(f);
    }
    printf("done\n");
    rc = 0;
out:
    if (h) mp_destroy(h);
    free(obs); free(post); free(cum); free(par); free(z0); free(occ);
    printf("Finished. It took  %.2lf min\n", difftime(time(NULL), start) / 60.0);
    return rc;
}
