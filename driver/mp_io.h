/* driver/mp_io.h -- input/output formats of the reference programs, written from their behaviour.
 *
 * Input (main_MIDASPOM.c:138-167): a whitespace table of 0 / 1 / -1, rows = years, columns = patches.
 * The reference takes n = 1 + number of blanks/tabs on the FIRST line and tmax = number of '\n', then
 * reads n*tmax integers in STREAM order with fscanf -- line breaks are ignored, so ragged rows wrap
 * (the bundled examples/input/occupancies.txt has 8,8,9,9,9,9,9 fields per row and is read as 8 x 7
 * with the last 5 integers dropped).  Reproduced exactly.
 * Output (main_MIDASPOM.c:427-436): nstep rows of nstep values, each "%.20lf\t", '\n' after each row.
 */
#ifndef MP_IO_H
#define MP_IO_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define MP_UNUSED __attribute__((unused))
MP_UNUSED static int mp_read_occupancy(const char *path, int8_t **obs_out, int *n_out, int *tmax_out)
{
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); return -1; }
    int c, n = 1, tmax = 0;
    while ((c = fgetc(f)) != EOF) {
        if (c == '\n') tmax++;
        if (tmax == 0 && (c == ' ' || c == '\t')) n++;
    }
    rewind(f);
    if (tmax < 1) { fclose(f); fprintf(stderr, "%s: no complete line\n", path); return -1; }
    int8_t *obs = (int8_t *)calloc((size_t)n * tmax, 1);
    for (long i = 0; i < (long)n * tmax; i++) {
        int v = 0;
        if (fscanf(f, "%d", &v) != 1) v = 0;       /* the reference leaves the cell uninitialised; 0 here */
        if (v < -1 || v > 1) { fprintf(stderr, "%s: value %d is not -1, 0 or 1\n", path, v); free(obs); fclose(f); return -1; }
        obs[i] = (int8_t)v;
    }
    fclose(f);
    *obs_out = obs; *n_out = n; *tmax_out = tmax;
    return 0;
}

MP_UNUSED static int mp_write_table(const char *path, const double *tab, int rows, int cols)
{
    FILE *f = fopen(path, "wb");
    if (!f) { fprintf(stderr, "cannot write %s\n", path); return -1; }
    for (int i = 0; i < rows; i++) {
        for (int j = 0; j < cols; j++) fprintf(f, "%.20lf\t", tab[(size_t)i * cols + j]);
        fprintf(f, "\n");
    }
    fclose(f);
    return 0;
}

/* square table of doubles as written by mp_write_table / the reference (future.c:238-262): size = blanks on line 1 */
MP_UNUSED static int mp_read_square_table(const char *path, double **tab_out, int *n_out)
{
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); return -1; }
    int c, n = 0;
    while ((c = fgetc(f)) != EOF) { if (c == '\n') break; if (c == ' ' || c == '\t') n++; }
    rewind(f);
    if (n < 2) { fclose(f); fprintf(stderr, "%s: not a posterior table\n", path); return -1; }
    double *t = (double *)calloc((size_t)n * n, sizeof(double));
    for (long i = 0; i < (long)n * n; i++) if (fscanf(f, "%lf", &t[i]) != 1) t[i] = 0.0;
    fclose(f);
    *tab_out = t; *n_out = n;
    return 0;
}

/* splitmix64: host-side draws of the driver (which grid cell, which completion) */
MP_UNUSED static uint64_t mp_splitmix(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
MP_UNUSED static double mp_unif(uint64_t *s) { return (double)(mp_splitmix(s) >> 11) * (1.0 / 9007199254740992.0); }
#endif
