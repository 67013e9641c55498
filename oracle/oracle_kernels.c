/* oracle_kernels.c - plain-C, single-thread, fp64 restatement of the arithmetic on the
 * GraphLearning Poisson/Laplace hot path.
 *
 * TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py through oracle/c_oracle.py.
 * Nothing under graphlearning_b200/ may link or call this file.
 *
 * What each function follows (reference = jwcalder/GraphLearning v1.7.5):
 *   orc_csr_matvecs     scipy.sparse._sparsetools csr_matvecs (third-party, scipy 1.18.1 in this
 *                       image; called by the reference at graphlearning/ssl.py:668, utils.py:515,523):
 *                       for every row i, for every stored entry jj in stored order,
 *                       Y[i,:] += Ax[jj] * X[Aj[jj],:]  (axpy over the n_vecs columns).
 *   orc_poisson_iterate graphlearning/ssl.py:667-668   u <- Db + P*u, T times (P*u is computed
 *                       into a zeroed buffer by csr_matvecs, then added to Db: that order of
 *                       additions is reproduced so the result is bit-identical to scipy's).
 *   orc_mixing_iterate  graphlearning/ssl.py:667,669   v <- RW*v and max|v - vinf| (stopping rule).
 *   orc_lp_iterate      c_code/lp_iterate.cpp:35-125   Jacobi p-Laplace barrier sweeps.
 *   orc_lip_iterate     c_code/lp_iterate.cpp:129-187  Gauss-Seidel game-theoretic p-Laplace sweeps.
 *
 * Build: gcc -O2 -fPIC -shared (no -ffast-math: the checks against scipy are bit-level).
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>

void orc_csr_matvecs(int n_row, int n_vecs, const int *Ap, const int *Aj, const double *Ax,
                     const double *X, double *Y)
{
    for (int i = 0; i < n_row; i++) {
        double *y = Y + (size_t)n_vecs * i;
        for (int jj = Ap[i]; jj < Ap[i + 1]; jj++) {
            const double a = Ax[jj];
            const double *x = X + (size_t)n_vecs * Aj[jj];
            for (int k = 0; k < n_vecs; k++)
                y[k] += a * x[k];
        }
    }
}

/* u (n x c, row-major) is overwritten with the T-th iterate; returns 0, or -1 on alloc failure. */
int orc_poisson_iterate(int n, int c, const int *Pp, const int *Pj, const double *Px,
                        const double *Db, double *u, int T)
{
    size_t len = (size_t)n * c;
    double *tmp = (double *)malloc(len * sizeof(double));
    if (!tmp) return -1;
    for (int t = 0; t < T; t++) {
        memset(tmp, 0, len * sizeof(double));
        orc_csr_matvecs(n, c, Pp, Pj, Px, u, tmp);
        for (size_t i = 0; i < len; i++)
            u[i] = Db[i] + tmp[i];
    }
    free(tmp);
    return 0;
}

/* v <- RW*v, T times; returns max|v - vinf| after the last step (ssl.py:667,669). */
double orc_mixing_iterate(int n, const int *Rp, const int *Rj, const double *Rx,
                          const double *vinf, double *v, int T)
{
    double *tmp = (double *)malloc((size_t)n * sizeof(double));
    double err = 0.0;
    for (int t = 0; t < T; t++) {
        for (int i = 0; i < n; i++) {
            double s = 0.0;
            for (int jj = Rp[i]; jj < Rp[i + 1]; jj++)
                s += Rx[jj] * v[Rj[jj]];
            tmp[i] = s;
        }
        memcpy(v, tmp, (size_t)n * sizeof(double));
    }
    for (int i = 0; i < n; i++) {
        double d = fabs(v[i] - vinf[i]);
        if (d > err) err = d;
    }
    free(tmp);
    return err;
}

#define ORC_MIN(a, b) (((a) < (b)) ? (a) : (b))
#define ORC_MAX(a, b) (((a) > (b)) ? (a) : (b))

static void row_scan(const int *J, int n, int M, int *start, int *num)
{
    int j = 0;                                   /* lp_iterate.cpp:47-57 (bound checked first here) */
    for (int i = 0; i < n; i++) {
        start[i] = j;
        num[i] = 0;
        while (j < M && J[j] == i) { num[i]++; j++; }
    }
}

/* Returns the number of sweeps executed.  uu/ul are the CALLER's buffers and end up holding
 * what the reference leaves in them (it swaps local pointers each sweep, lp_iterate.cpp:116-123). */
int orc_lp_iterate(double *uu_c, double *ul_c, const int *I, const int *J, const double *W,
                   const int *ind, const double *val, double p, int T, double tol,
                   int n, int M, int m)
{
    double alpha = 1 / p, delta = 1 - 2 / p, dt = 0.9 / (alpha + 2 * delta);
    int *num = (int *)malloc(sizeof(int) * n), *start = (int *)malloc(sizeof(int) * n);
    double *invdeg = (double *)malloc(sizeof(double) * n);
    double *vu = (double *)calloc(n, sizeof(double)), *vl = (double *)calloc(n, sizeof(double));
    double *uu = uu_c, *ul = ul_c, *tmp;
    row_scan(J, n, M, start, num);
    for (int i = 0; i < n; i++) {
        double s = 0;
        for (int j = start[i]; j < start[i] + num[i]; j++) s += W[j];
        invdeg[i] = alpha / s;
    }
    double maxw_all = 0;
    for (int i = 0; i < M; i++) maxw_all = ORC_MAX(maxw_all, W[i]);
    dt = dt / maxw_all;
    int it, sweeps = 0;
    for (it = 0; it < T; it++) {
        sweeps++;
        double err = 0;
        for (int i = 0; i < n; i++) {
            double minw = 0, maxw = 0, sumw = 0;
            for (int j = start[i]; j < start[i] + num[i]; j++) {
                double d = W[j] * (uu[I[j]] - uu[i]);
                minw = ORC_MIN(d, minw); maxw = ORC_MAX(d, maxw); sumw += d;
            }
            vu[i] = uu[i] + dt * (invdeg[i] * sumw + delta * (minw + maxw));
            minw = 0; maxw = 0; sumw = 0;
            for (int j = start[i]; j < start[i] + num[i]; j++) {
                double d = W[j] * (ul[I[j]] - ul[i]);
                minw = ORC_MIN(d, minw); maxw = ORC_MAX(d, maxw); sumw += d;
            }
            vl[i] = ul[i] + dt * (invdeg[i] * sumw + delta * (minw + maxw));
            err = ORC_MAX(uu[i] - ul[i], err);
        }
        for (int j = 0; j < m; j++) { vu[ind[j]] = val[j]; vl[ind[j]] = val[j]; }
        if (err < tol && it > 10) break;
        tmp = uu; uu = vu; vu = tmp;
        tmp = ul; ul = vl; vl = tmp;
    }
    /* free only the scratch that is not the caller's memory */
    free((uu == uu_c) ? vu : uu);
    free((ul == ul_c) ? vl : ul);
    free(num); free(start); free(invdeg);
    return sweeps;
}

int orc_lip_iterate(double *u, const int *I, const int *J, const double *W, const int *ind,
                    const double *val, int T, double tol, int n, int M, int m,
                    double alpha, double beta)
{
    int *num = (int *)malloc(sizeof(int) * n), *start = (int *)malloc(sizeof(int) * n);
    char *mask = (char *)malloc(n);
    memset(mask, 1, n);
    row_scan(J, n, M, start, num);
    for (int j = 0; j < m; j++) { u[ind[j]] = val[j]; mask[ind[j]] = 0; }
    int it, sweeps = 0;
    for (it = 0; it < T; it++) {
        sweeps++;
        double err = 0;
        for (int i = 0; i < n; i++) {
            if (!mask[i]) continue;
            double minu = u[I[start[i]]], maxu = minu, sumu = 0.0, deg = 0.0;
            for (int j = start[i]; j < start[i] + num[i]; j++) {
                sumu += W[j] * u[I[j]];
                deg += W[j];
                minu = ORC_MIN(u[I[j]], minu);
                maxu = ORC_MAX(u[I[j]], maxu);
            }
            double ne = alpha * sumu / deg + beta * (minu + maxu) / 2;
            double d = fabs(u[i] - ne);
            err = ORC_MAX(d, err);
            u[i] = ne;
        }
        if (err < tol && it > 20) break;
    }
    free(num); free(start); free(mask);
    return sweeps;
}

/* c_code/lp_iterate.cpp:190-259 (lip_iterate_weighted_main): Gauss-Seidel sweeps of the weighted infinity
 * Laplacian; every row solves min_j w(t-u_j) + max_j w(t-u_j) = 0 for t by 30 bisection steps on [min u_j, max u_j]. */
int orc_lip_iterate_weighted(double *u, const int *I, const int *J, const double *W, const int *ind,
                             const double *val, int T, double tol, int n, int M, int m)
{
    int *num = (int *)malloc(sizeof(int) * n), *start = (int *)malloc(sizeof(int) * n);
    char *mask = (char *)malloc(n);
    memset(mask, 1, n);
    row_scan(J, n, M, start, num);
    for (int j = 0; j < m; j++) { u[ind[j]] = val[j]; mask[ind[j]] = 0; }
    int it, sweeps = 0;
    for (it = 0; it < T; it++) {
        sweeps++;
        double err = 0;
        for (int i = 0; i < n; i++) {
            if (!mask[i]) continue;
            double minu = u[I[start[i]]], maxu = minu;
            for (int j = start[i]; j < start[i] + num[i]; j++) {
                minu = ORC_MIN(u[I[j]], minu);
                maxu = ORC_MAX(u[I[j]], maxu);
            }
            double a = minu, b = maxu;
            for (int k = 0; k < 30; k++) {
                double minw = 0, maxw = 0, t = (a + b) / 2.0;
                for (int j = start[i]; j < start[i] + num[i]; j++) {
                    double d = W[j] * (t - u[I[j]]);
                    minw = ORC_MIN(d, minw);
                    maxw = ORC_MAX(d, maxw);
                }
                if (minw + maxw > 0) b = t; else a = t;
            }
            double ne = (a + b) / 2.0;
            double d = fabs(u[i] - ne);
            err = ORC_MAX(d, err);
            u[i] = ne;
        }
        if (err < tol && it > 20) break;
    }
    free(num); free(start); free(mask);
    return sweeps;
}
