/* driver/midaspom_hypothesis.c -- the numbers of Rscript/hypothesis_test.R:29-46 without the plot:
 * AIC of H0 (no event), H1 (in-situ die-off), H2 (habitat loss) and the three log10 Bayes factors,
 * from the files written by midaspom_dieoff / midaspom_loss.  Host-side only (no GPU work).
 *
 *   midaspom_hypothesis lh_dieoff.txt lh_loss.txt N_PATCHES KMIN KMAX
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

static double *read_all(const char *path, long *count, int *first_line_fields)
{
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); return NULL; }
    long cap = 1024, n = 0; double *v = malloc(cap * sizeof(double)), x;
    int c, fields = 0, in_first = 1, in_tok = 0;
    while ((c = fgetc(f)) != EOF && in_first) {
        if (c == '\n') in_first = 0;
        else if (c == ' ' || c == '\t') in_tok = 0;
        else if (!in_tok) { in_tok = 1; fields++; }
    }
    rewind(f);
    while (fscanf(f, "%lf", &x) == 1) { if (n == cap) v = realloc(v, (cap *= 2) * sizeof(double)); v[n++] = x; }
    fclose(f);
    *count = n; *first_line_fields = fields;
    return v;
}

int main(int argc, char **argv)
{
    if (argc < 6) { fprintf(stderr, "usage: %s lh_dieoff.txt lh_loss.txt n Kmin Kmax\n", argv[0]); return 1; }
    long nd1, nl; int f1, stepd;
    double *die = read_all(argv[1], &nd1, &f1), *loss = read_all(argv[2], &nl, &stepd);
    if (!die || !loss || nd1 < 2 || stepd < 1) return 1;
    const double n = atof(argv[3]), Kmin = atof(argv[4]), Kmax = atof(argv[5]);
    const long stepKd = nd1, stepKl = nl / stepd;
    /* likelihood of the null model K = 1: the grid node if present, else linear interpolation (:29-35) */
    double postnull = NAN;
    for (long i = 0; i < stepKd; i++) {
        const double K = pow(10.0, log10(Kmin) + (log10(Kmax) - log10(Kmin)) * (double)i / (double)(stepKd - 1));
        if (K == 1.0) { postnull = die[i]; break; }
        if (K > 1.0) {
            const double Kp = pow(10.0, log10(Kmin) + (log10(Kmax) - log10(Kmin)) * (double)(i - 1) / (double)(stepKd - 1));
            postnull = (die[i] - die[i - 1]) / (K - Kp) * (1.0 - Kp) + die[i - 1];
            break;
        }
    }
    double maxd = die[0], sumd = 0, maxl = loss[0], suml = 0;
    for (long i = 0; i < stepKd; i++) { if (die[i] > maxd) maxd = die[i]; sumd += die[i]; }
    for (long i = 0; i < nl; i++) { if (loss[i] > maxl) maxl = loss[i]; suml += loss[i]; }
    const double two_n = pow(2.0, n);
    printf("AIC H0 %.10g\nAIC H1 %.10g\nAIC H2 %.10g\n", 2 * 1 - 2 * log(postnull / two_n), 2 * 2 - 2 * log(maxd / two_n),
           2 * 3 - 2 * log(maxl / two_n));                                                     /* :38-41 */
    printf("log10BF H0vsH1 %.10g\nlog10BF H0vsH2 %.10g\nlog10BF H1vsH2 %.10g\n", log10(postnull / (sumd / stepKl)),
           log10(postnull / (suml / stepKl / stepd)), log10(sumd / (suml / stepd)));           /* :44-46 */
    free(die); free(loss);
    return 0;
}
