// ref_shim.cpp - extern "C" doorway onto the UNMODIFIED reference p-Laplace solvers.
//
// TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile together with the reference's own
// sources where they lie (/root/reference/c_code/lp_iterate.cpp, memory_allocation.cpp) into
// oracle/_ref/liblp_ref.so.  No reference source is copied: the three prototypes below are the
// interface declared at c_code/lp_iterate.h:33-35, which the CPython shim
// c_code/cextensions.cpp:19-107 forwards to.
extern void lp_iterate_main(double *uu, double *ul, int *II, int *J, double *W, int *ind, double *val,
                            double p, int T, double tol, bool prog, int n, int M, int m);
extern void lip_iterate_main(double *u, int *II, int *J, double *W, int *ind, double *val, int T,
                             double tol, bool prog, int n, int M, int m, double alpha, double beta);
extern void lip_iterate_weighted_main(double *u, int *II, int *J, double *W, int *ind, double *val,
                                      int T, double tol, bool prog, int n, int M, int m);

extern "C" {
void ref_lp_iterate(double *uu, double *ul, int *II, int *J, double *W, int *ind, double *val,
                    double p, int T, double tol, int n, int M, int m)
{
    lp_iterate_main(uu, ul, II, J, W, ind, val, p, T, tol, false, n, M, m);
}
void ref_lip_iterate(double *u, int *II, int *J, double *W, int *ind, double *val, int T, double tol,
                     int n, int M, int m, double alpha, double beta)
{
    lip_iterate_main(u, II, J, W, ind, val, T, tol, false, n, M, m, alpha, beta);
}
void ref_lip_iterate_weighted(double *u, int *II, int *J, double *W, int *ind, double *val, int T,
                              double tol, int n, int M, int m)
{
    lip_iterate_weighted_main(u, II, J, W, ind, val, T, tol, false, n, M, m);
}
}
