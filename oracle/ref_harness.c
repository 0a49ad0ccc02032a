/* oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Flat-array wrappers around the REFERENCE'S OWN functions so tests can call them on arbitrary
 * inputs.  The reference sources are compiled where they lie (/root/reference/sources/*.c, see
 * oracle/Makefile) with -D renames (main -> refmain_*, compPePc -> refbase_compPePc, ...); nothing
 * is copied into this repo.  This file only declares those renamed symbols and adapts the
 * reference's `T**` arguments to the flat row-major arrays that ctypes/numpy hands over.
 *
 * Reference functions exposed:
 *   compPePc    main_MIDASPOM.c:18-50      (product form)   -> ref_base_compPePc
 *   compPePc    main_MIDASPOM_MPI.c:20-53  (sum-of-logs)    -> ref_mpi_compPePc
 *   pije/pijc   main_MIDASPOM_dieoff.c:51-83                -> ref_dieoff_pije / ref_dieoff_pijc
 *   pije/pijc/pijcsource main_MIDASPOM_loss.c:52-105        -> ref_loss_*
 *   simpij      main_MIDASPOM_future.c:64-110               -> ref_future_simpij
 *   matpow      main_MIDASPOM_dieoff.c:17-49                -> ref_dieoff_matpow
 */
#include <stdlib.h>

/* renamed reference symbols (see Makefile HARNESS_DEFS_*) */
void refbase_compPePc(double *, double *, unsigned int **, unsigned int *, double, double *, double **,
                      unsigned int, unsigned int, unsigned int, unsigned int);
void refmpi_compPePc(double *, double *, unsigned int **, unsigned int *, double, double *, double **,
                     unsigned int, unsigned int, unsigned int, unsigned int);
double refdieoff_pije(int *, int *, double, double, int);
double refdieoff_pijc(int *, int *, double, double, double **, int);
void   refdieoff_matpow(double *, int, int, double *);
double refloss_pije(int *, int *, double, int);
double refloss_pijc(int *, int *, double, double **, int);
double refloss_pijcsource(int *, int *, double, double, double **, int);
int    reffuture_simpij(int *, int *, double, double, double, double, double **, int);

static double **rows_of(const double *flat, int nrow, int ncol)
{
    double **r = malloc((size_t)nrow * sizeof(double *));
    for (int i = 0; i < nrow; i++) r[i] = (double *)flat + (size_t)i * ncol;
    return r;
}

/* Both compPePc variants: piall is nstates x n row-major; M is n x n row-major; fills column j
 * of Pe (nextid x nstates) and row j of Pc (nstates x nextid), exactly as the reference does. */
static void call_compPePc(int which, double *Pe, double *Pc, const unsigned int *piall_flat,
                          unsigned int *all2short, double e, double *pC, const double *M_flat,
                          unsigned int n, unsigned int nextid, unsigned int nstates, unsigned int j)
{
    unsigned int **piall = malloc((size_t)nstates * sizeof(unsigned int *));
    for (unsigned int i = 0; i < nstates; i++) piall[i] = (unsigned int *)piall_flat + (size_t)i * n;
    double **M = rows_of(M_flat, (int)n, (int)n);
    if (which == 0) refbase_compPePc(Pe, Pc, piall, all2short, e, pC, M, n, nextid, nstates, j);
    else            refmpi_compPePc(Pe, Pc, piall, all2short, e, pC, M, n, nextid, nstates, j);
    free(M); free(piall);
}
void ref_base_compPePc(double *Pe, double *Pc, const unsigned int *piall, unsigned int *all2short,
                       double e, double *pC, const double *M, unsigned int n, unsigned int nextid,
                       unsigned int nstates, unsigned int j)
{ call_compPePc(0, Pe, Pc, piall, all2short, e, pC, M, n, nextid, nstates, j); }
void ref_mpi_compPePc(double *Pe, double *Pc, const unsigned int *piall, unsigned int *all2short,
                      double e, double *pC, const double *M, unsigned int n, unsigned int nextid,
                      unsigned int nstates, unsigned int j)
{ call_compPePc(1, Pe, Pc, piall, all2short, e, pC, M, n, nextid, nstates, j); }

double ref_dieoff_pije(int *piold, int *pitmp, double e, double K, int n)
{ return refdieoff_pije(piold, pitmp, e, K, n); }
double ref_dieoff_pijc(int *pitmp, int *pinew, double c, double K, const double *M_flat, int n)
{ double **M = rows_of(M_flat, n, n); double r = refdieoff_pijc(pitmp, pinew, c, K, M, n); free(M); return r; }
void ref_dieoff_matpow(double *x, int n, int k, double *z) { refdieoff_matpow(x, n, k, z); }

double ref_loss_pije(int *piold, int *pitmp, double e, int n) { return refloss_pije(piold, pitmp, e, n); }
double ref_loss_pijc(int *pitmp, int *pinew, double c, const double *M_flat, int n)
{ double **M = rows_of(M_flat, n, n); double r = refloss_pijc(pitmp, pinew, c, M, n); free(M); return r; }
/* M_flat is (n+1) x n: row n is the external-source row (main_MIDASPOM_loss.c:365) */
double ref_loss_pijcsource(int *pitmp, int *pinew, double c, double Ksource, const double *M_flat, int n)
{ double **M = rows_of(M_flat, n + 1, n); double r = refloss_pijcsource(pitmp, pinew, c, Ksource, M, n); free(M); return r; }

/* M_flat is (n+1) x n (source row n, main_MIDASPOM_future.c:266-277). Uses libc rand(): seed
 * with srand() from the caller (ctypes: libc.srand) for repeatable draws. */
int ref_future_simpij(int *piold, int *pinew, double e, double c, double K, double Ksource,
                      const double *M_flat, int n)
{ double **M = rows_of(M_flat, n + 1, n); int r = reffuture_simpij(piold, pinew, e, c, K, Ksource, M, n); free(M); return r; }
void ref_srand(unsigned int s) { srand(s); }
