/* driver/midaspom_variant.c -- flags and file formats of MIDASPOM_dieoff.out and MIDASPOM_loss.out
 * (reference: sources/main_MIDASPOM_dieoff.c, sources/main_MIDASPOM_loss.c) on the B200 engine.
 * Compiled twice: -DMP_VARIANT=1 -> midaspom_dieoff, -DMP_VARIANT=2 -> midaspom_loss.
 *
 *   midaspom_dieoff -a 10 -e 0.71 -c 0.52 -m 400 -d 100 -s 151 -i occupancies.txt -o lh_dieoff.txt
 *   midaspom_loss   -a 10 -e 0.71 -c 0.52 -m 400 -d 100 -s 151 -i occupancies.txt -o lh_loss.txt
 * Flags (dieoff.c:125, loss.c:145): b (years before the event, default 20) a (years after) e c m p d i o
 * s (K steps, 151) l u (K range, 0.1..100); loss adds v (d_L steps, 20) L U (d_L range, 200..4000).
 * Only the FIRST line of the input is used (dieoff.c:183-201).  Output: dieoff one line of s values,
 * loss s lines of v values, "%.20lf\t" (dieoff.c:371-377, loss.c:408-413).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "../include/libmidaspom_cuda.h"
#include "mp_io.h"

#ifndef MP_VARIANT
#define MP_VARIANT 1
#endif

int main(int argc, char **argv)
{
    printf(MP_VARIANT == 1 ? "------ MIDASPOM, in situ die-off hypothesis, on B200 (midaspom_b200) ------\n"
                           : "------ MIDASPOM, habitat loss hypothesis, on B200 (midaspom_b200) ------\n");
    int ts = 20, tdis = 10, nstep = 151, nstepd = 20, device = 0, c;    /* the reference leaves tdis, eB, cB uninitialised */
    double eB = 0.5, cB = 0.5, Kmin = 0.1, Kmax = 100.0, dmin = 200, dmax = 4000, a = 1.0 / 400.0, d = 200;
    float prioroc = 0.5f;
    const char *fname = "input.txt", *fout = MP_VARIANT == 1 ? "lh_dieoff.txt" : "lh_loss.txt";
    opterr = 0;
    while ((c = getopt(argc, argv, "b:a:e:c:m:p:d:i:o:s:l:u:v:L:U:g:")) != -1)
        switch (c) {
        case 'b': ts = atoi(optarg); break;
        case 'a': tdis = atoi(optarg); break;
        case 'e': eB = atof(optarg); break;
        case 'c': cB = atof(optarg); break;
        case 'm': a = 1.0 / atof(optarg); break;
        case 'p': prioroc = (float)atof(optarg); break;
        case 'd': d = atof(optarg); break;
        case 'i': fname = optarg; break;
        case 'o': fout = optarg; break;
        case 's': nstep = atoi(optarg); break;
        case 'l': Kmin = atof(optarg); break;
        case 'u': Kmax = atof(optarg); break;
        case 'v': nstepd = atoi(optarg); break;
        case 'L': dmin = atof(optarg); break;
        case 'U': dmax = atof(optarg); break;
        case 'g': device = atoi(optarg); break;
        default: fprintf(stderr, "Unknown option `-%c'.\n", optopt); return 1;
        }
    printf("%d years before the event, %d years after the event\n", ts, tdis);
    time_t start = time(NULL);
    printf("Reading observations from file %s... ", fname);
    int8_t *obs = NULL; int n = 0, tmax = 0;
    if (mp_read_occupancy(fname, &obs, &n, &tmax)) return 1;
    printf("done\n%d patches\n", n);
    const int nd = MP_VARIANT == 2 ? nstepd : 1;
    double *lik = (double *)malloc((size_t)nstep * nd * sizeof(double));
    printf("Starting likelihood computation\n");
    int rc = mp_exact_variant(device, MP_VARIANT, obs /* first row */, n, a, d, prioroc, eB, cB, ts, tdis, nstep, Kmin, Kmax, nstepd,
                              dmin, dmax, lik);
    if (rc != MP_OK) { fprintf(stderr, "mp_exact_variant failed (%d): %s\n", rc, mp_exact_last_error()); free(obs); free(lik); return 2; }
    printf("end likelihood computation\n");
    printf("Writing on file %s... ", fout);
    FILE *f = fopen(fout, "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", fout); return 3; }
    if (MP_VARIANT == 1) for (int i = 0; i < nstep; i++) fprintf(f, "%.20lf\t", lik[i]);
    else for (int i = 0; i < nstep; i++) { for (int j = 0; j < nd; j++) fprintf(f, "%.20lf\t", lik[(size_t)i * nd + j]); fprintf(f, "\n"); }
    fclose(f);
    printf("done\n");
    free(obs); free(lik);
    printf("Finished. It took  %.2lf min\n", difftime(time(NULL), start) / 60.0);
    return 0;
}
