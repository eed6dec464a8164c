/* CPU unit-test build of epimethex_b200/csrc/epx_math.h (the header the CUDA kernels include).
 * Built by tests/conftest.py with gcc; used only by tests/test_math_cpu.py. */
#include "../../epimethex_b200/csrc/epx_math.h"
double mt_t_two_sided_p(double t, double df) { return epx_t_two_sided_p(t, df); }
double mt_pkstwo(double x) { return epx_pkstwo(x); }
double mt_ks_p_asym(double dnum, double la, double lb) { return epx_ks_p_asym(dnum, la, lb); }
double mt_ks_p_exact(double dnum, int la, int lb, double *u) { return epx_ks_p_exact(dnum, la, lb, u); }
double mt_log_fc(double a, double b) { return epx_log_fc(a, b); }
double mt_linear_fc(double a, double b) { return epx_linear_fc(a, b); }
double mt_pearson_p(double r, double n) { return epx_pearson_p(r, n); }
double mt_lbeta_half(double a) { return epx_lbeta_half(a); }
int mt_welch(double mx, double vx, double nx, double my, double vy, double ny, double *t, double *df)
{ return epx_welch(mx, vx, nx, my, vy, ny, t, df); }
double mt_pearson_h(double r, double df) { return epx_pearson_h(r, df); }
int mt_pearson_grid(double df) { return epx_pearson_grid(df); }
double mt_pearson_p_tab(double r, const double *h, int N, double df) { return epx_pearson_p_tab(r, h, N, df); }
void mt_welch_grid(int m, int *N, int *K, double *df0, double *ddf) { epx_welch_grid(m, N, K, df0, ddf); }
double mt_welch_p_tab(double t, double df, const double *tab, int N, int ld, int K, double df0, double inv_ddf)
{ return epx_welch_p_tab(t, df, tab, N, ld, K, df0, inv_ddf); }
void mt_welch_table(int m, double *tab, int ld)
{
    int N, K; double df0, ddf;
    epx_welch_grid(m, &N, &K, &df0, &ddf);
    for (int k = 0; k < K; k++)
        for (int i = 0; i <= N; i++) tab[(long long)k * ld + i] = epx_pearson_h((double)i / (double)N, df0 + k * ddf);
}
