// k9_fit.cu -- host side of stage 2b: the four-parameter tanh fit of GaussianBandwidth.py:47-57,
//     logSumWij(x) ~ a3 + a2 tanh(a0 x + a1),   x = log(eps),
// which the reference solves with scipy.optimize.curve_fit(fun, logEps, logSumWij, p0=a0), i.e. MINPACK's lmdif
// (Levenberg-Marquardt with a forward-difference Jacobian) behind scipy.optimize.leastsq with scipy's defaults:
// ftol = xtol = 1.49012e-8, gtol = 0, maxfev = 200 (n + 1), epsfcn = machine epsilon, factor = 100, diag = None (mode 1).
// MINPACK is a dependency of the reference that is not in /root/reference (scipy, unpinned by the reference; 1.18.1 in
// the authoring container); this file restates its published algorithm (More, Garbow, Hillstrom: User Guide for
// MINPACK-1, ANL-80-74; routines lmdif, lmpar, qrfac, qrsolv, fdjac2, enorm) so that one C call can run a whole PD
// without entering the Python interpreter.  Pinned on the CPU against scipy itself (tests/test_host.py): popt agrees to
// ~1e-12 relative on the golden sweeps, far inside the 1e-7 of SURVEY appendix A.  The covariance follows
// scipy.optimize._minpack_py.leastsq (inverse of the permuted R factor) and curve_fit's s_sq = cost / (M - N) scaling.
// Pure host code: no device, no ctx.
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace esper {
namespace lmfit {

constexpr double EPSMCH = 2.220446049250313e-16;      // dpmpar(1)
constexpr double DWARF = 2.2250738585072014e-308;     // dpmpar(2)

// Euclidean norm with MINPACK's three accumulators (small / intermediate / large components).
static double enorm(int n, const double* x) {
    const double rdwarf = 3.834e-20, rgiant = 1.304e19;
    double s1 = 0.0, s2 = 0.0, s3 = 0.0, x1max = 0.0, x3max = 0.0;
    const double agiant = rgiant / (double)n;
    for (int i = 0; i < n; ++i) {
        const double xabs = fabs(x[i]);
        if (xabs > rdwarf && xabs < agiant) { s2 += xabs * xabs; continue; }
        if (xabs <= rdwarf) {
            if (xabs > x3max) { const double r = x3max / xabs; s3 = 1.0 + s3 * (r * r); x3max = xabs; }
            else if (xabs != 0.0) { const double r = xabs / x3max; s3 += r * r; }
        } else {
            if (xabs > x1max) { const double r = x1max / xabs; s1 = 1.0 + s1 * (r * r); x1max = xabs; }
            else { const double r = xabs / x1max; s1 += r * r; }
        }
    }
    if (s1 != 0.0) return x1max * sqrt(s1 + (s2 / x1max) / x1max);
    if (s2 != 0.0) {
        if (s2 >= x3max) return sqrt(s2 * (1.0 + (x3max / s2) * (x3max * s3)));
        return sqrt(x3max * ((s2 / x3max) + (x3max * s3)));
    }
    return x3max * sqrt(s3);
}

// Householder QR with column pivoting of the m x n matrix a (column-major, leading dimension lda).
static void qrfac(int m, int n, double* a, int lda, int* ipvt, double* rdiag, double* acnorm, double* wa) {
    for (int j = 0; j < n; ++j) {
        acnorm[j] = enorm(m, a + (size_t)j * lda);
        rdiag[j] = acnorm[j];
        wa[j] = rdiag[j];
        ipvt[j] = j;
    }
    const int minmn = std::min(m, n);
    for (int j = 0; j < minmn; ++j) {
        int kmax = j;
        for (int k = j; k < n; ++k) if (rdiag[k] > rdiag[kmax]) kmax = k;
        if (kmax != j) {
            for (int i = 0; i < m; ++i) std::swap(a[i + (size_t)j * lda], a[i + (size_t)kmax * lda]);
            rdiag[kmax] = rdiag[j];
            wa[kmax] = wa[j];
            std::swap(ipvt[j], ipvt[kmax]);
        }
        double* aj = a + (size_t)j * lda;
        double ajnorm = enorm(m - j, aj + j);
        if (ajnorm != 0.0) {
            if (aj[j] < 0.0) ajnorm = -ajnorm;
            for (int i = j; i < m; ++i) aj[i] /= ajnorm;
            aj[j] += 1.0;
            for (int k = j + 1; k < n; ++k) {
                double* ak = a + (size_t)k * lda;
                double sum = 0.0;
                for (int i = j; i < m; ++i) sum += aj[i] * ak[i];
                const double temp = sum / aj[j];
                for (int i = j; i < m; ++i) ak[i] -= temp * aj[i];
                if (rdiag[k] != 0.0) {
                    double t = ak[j] / rdiag[k];
                    rdiag[k] *= sqrt(std::max(0.0, 1.0 - t * t));
                    t = rdiag[k] / wa[k];
                    if (0.05 * (t * t) <= EPSMCH) {
                        rdiag[k] = enorm(m - j - 1, ak + j + 1);
                        wa[k] = rdiag[k];
                    }
                }
            }
        }
        rdiag[j] = -ajnorm;
    }
}

// Solve R z = Q^T b together with D z = 0 in the least-squares sense (Givens eliminations of the rows of D).
static void qrsolv(int n, double* r, int ldr, const int* ipvt, const double* diag, const double* qtb, double* x,
                   double* sdiag, double* wa) {
    for (int j = 0; j < n; ++j) {
        for (int i = j; i < n; ++i) r[i + (size_t)j * ldr] = r[j + (size_t)i * ldr];
        x[j] = r[j + (size_t)j * ldr];
        wa[j] = qtb[j];
    }
    for (int j = 0; j < n; ++j) {
        const int l = ipvt[j];
        if (diag[l] != 0.0) {
            for (int k = j; k < n; ++k) sdiag[k] = 0.0;
            sdiag[j] = diag[l];
            double qtbpj = 0.0;
            for (int k = j; k < n; ++k) {
                if (sdiag[k] == 0.0) continue;
                double& rkk = r[k + (size_t)k * ldr];
                double cs, sn;
                if (fabs(rkk) < fabs(sdiag[k])) {
                    const double cotan = rkk / sdiag[k];
                    sn = 0.5 / sqrt(0.25 + 0.25 * (cotan * cotan));
                    cs = sn * cotan;
                } else {
                    const double tn = sdiag[k] / rkk;
                    cs = 0.5 / sqrt(0.25 + 0.25 * (tn * tn));
                    sn = cs * tn;
                }
                rkk = cs * rkk + sn * sdiag[k];
                const double temp = cs * wa[k] + sn * qtbpj;
                qtbpj = -sn * wa[k] + cs * qtbpj;
                wa[k] = temp;
                for (int i = k + 1; i < n; ++i) {
                    double& rik = r[i + (size_t)k * ldr];
                    const double t2 = cs * rik + sn * sdiag[i];
                    sdiag[i] = -sn * rik + cs * sdiag[i];
                    rik = t2;
                }
            }
        }
        sdiag[j] = r[j + (size_t)j * ldr];
        r[j + (size_t)j * ldr] = x[j];
    }
    int nsing = n;
    for (int j = 0; j < n; ++j) {
        if (sdiag[j] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa[j] = 0.0;
    }
    for (int k = 1; k <= nsing; ++k) {
        const int j = nsing - k;
        double sum = 0.0;
        for (int i = j + 1; i < nsing; ++i) sum += r[i + (size_t)j * ldr] * wa[i];
        wa[j] = (wa[j] - sum) / sdiag[j];
    }
    for (int j = 0; j < n; ++j) x[ipvt[j]] = wa[j];
}

// Levenberg-Marquardt parameter: par such that |D x| is within 10 % of delta (or par = 0 for the Gauss-Newton step).
static void lmpar(int n, double* r, int ldr, const int* ipvt, const double* diag, const double* qtb, double delta,
                  double* par, double* x, double* sdiag, double* wa1, double* wa2) {
    int nsing = n;
    for (int j = 0; j < n; ++j) {
        wa1[j] = qtb[j];
        if (r[j + (size_t)j * ldr] == 0.0 && nsing == n) nsing = j;
        if (nsing < n) wa1[j] = 0.0;
    }
    for (int k = 1; k <= nsing; ++k) {
        const int j = nsing - k;
        wa1[j] /= r[j + (size_t)j * ldr];
        const double temp = wa1[j];
        for (int i = 0; i < j; ++i) wa1[i] -= r[i + (size_t)j * ldr] * temp;
    }
    for (int j = 0; j < n; ++j) x[ipvt[j]] = wa1[j];

    int iter = 0;
    for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
    double dxnorm = enorm(n, wa2);
    double fp = dxnorm - delta;
    if (fp <= 0.1 * delta) { *par = 0.0; return; }

    double parl = 0.0;
    if (nsing >= n) {
        for (int j = 0; j < n; ++j) { const int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
        for (int j = 0; j < n; ++j) {
            double sum = 0.0;
            for (int i = 0; i < j; ++i) sum += r[i + (size_t)j * ldr] * wa1[i];
            wa1[j] = (wa1[j] - sum) / r[j + (size_t)j * ldr];
        }
        const double temp = enorm(n, wa1);
        parl = ((fp / delta) / temp) / temp;
    }
    for (int j = 0; j < n; ++j) {
        double sum = 0.0;
        for (int i = 0; i <= j; ++i) sum += r[i + (size_t)j * ldr] * qtb[i];
        wa1[j] = sum / diag[ipvt[j]];
    }
    const double gnorm = enorm(n, wa1);
    double paru = gnorm / delta;
    if (paru == 0.0) paru = DWARF / std::min(delta, 0.1);
    *par = std::max(*par, parl);
    *par = std::min(*par, paru);
    if (*par == 0.0) *par = gnorm / dxnorm;

    for (;;) {
        ++iter;
        if (*par == 0.0) *par = std::max(DWARF, 0.001 * paru);
        double temp = sqrt(*par);
        for (int j = 0; j < n; ++j) wa1[j] = temp * diag[j];
        qrsolv(n, r, ldr, ipvt, wa1, qtb, x, sdiag, wa2);
        for (int j = 0; j < n; ++j) wa2[j] = diag[j] * x[j];
        dxnorm = enorm(n, wa2);
        temp = fp;
        fp = dxnorm - delta;
        if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || iter == 10) break;
        for (int j = 0; j < n; ++j) { const int l = ipvt[j]; wa1[j] = diag[l] * (wa2[l] / dxnorm); }
        for (int j = 0; j < n; ++j) {
            wa1[j] /= sdiag[j];
            const double t2 = wa1[j];
            for (int i = j + 1; i < n; ++i) wa1[i] -= r[i + (size_t)j * ldr] * t2;
        }
        temp = enorm(n, wa1);
        const double parc = ((fp / delta) / temp) / temp;
        if (fp > 0.0) parl = std::max(parl, *par);
        if (fp < 0.0) paru = std::min(paru, *par);
        *par = std::max(parl, *par + parc);
    }
    if (iter == 0) *par = 0.0;
}

struct TanhResidual {          // curve_fit's residual: fun(xdata, *params) - ydata   (GaussianBandwidth.py:18-24)
    const double* x; const double* y; int m;
    void operator()(const double* a, double* f) const {
        for (int i = 0; i < m; ++i) f[i] = (a[3] + a[2] * tanh(a[0] * x[i] + a[1])) - y[i];
    }
};

// lmdif with scipy.optimize.leastsq's defaults.  fjac (m x n, column-major) returns the R factor in its upper triangle.
template <class F>
static int lmdif(const F& fcn, int m, int n, double* x, double* fvec, double ftol, double xtol, double gtol, int maxfev,
                 double epsfcn, double factor, int* nfev_out, double* fjac, int* ipvt) {
    std::vector<double> diag(n), qtf(n), wa1(n), wa2(n), wa3(n), wa4(m);
    const int ldfjac = m;
    int info = 0, nfev = 0;
    if (n <= 0 || m < n || ftol < 0.0 || xtol < 0.0 || gtol < 0.0 || maxfev <= 0 || factor <= 0.0) { *nfev_out = 0; return 0; }
    fcn(x, fvec);
    nfev = 1;
    double fnorm = enorm(m, fvec);
    double par = 0.0, delta = 0.0, xnorm = 0.0;
    int iter = 1;
    for (;;) {
        // forward-difference Jacobian (fdjac2)
        {
            const double eps = sqrt(std::max(epsfcn, EPSMCH));
            for (int j = 0; j < n; ++j) {
                const double temp = x[j];
                double h = eps * fabs(temp);
                if (h == 0.0) h = eps;
                x[j] = temp + h;
                fcn(x, wa4.data());
                x[j] = temp;
                double* col = fjac + (size_t)j * ldfjac;
                for (int i = 0; i < m; ++i) col[i] = (wa4[i] - fvec[i]) / h;
            }
            nfev += n;
        }
        qrfac(m, n, fjac, ldfjac, ipvt, wa1.data(), wa2.data(), wa3.data());
        if (iter == 1) {
            for (int j = 0; j < n; ++j) { diag[j] = wa2[j]; if (wa2[j] == 0.0) diag[j] = 1.0; }
            for (int j = 0; j < n; ++j) wa3[j] = diag[j] * x[j];
            xnorm = enorm(n, wa3.data());
            delta = factor * xnorm;
            if (delta == 0.0) delta = factor;
        }
        // first n components of Q^T fvec
        for (int i = 0; i < m; ++i) wa4[i] = fvec[i];
        for (int j = 0; j < n; ++j) {
            double* col = fjac + (size_t)j * ldfjac;
            if (col[j] != 0.0) {
                double sum = 0.0;
                for (int i = j; i < m; ++i) sum += col[i] * wa4[i];
                const double temp = -sum / col[j];
                for (int i = j; i < m; ++i) wa4[i] += col[i] * temp;
            }
            col[j] = wa1[j];
            qtf[j] = wa4[j];
        }
        // norm of the scaled gradient
        double gnorm = 0.0;
        if (fnorm != 0.0) {
            for (int j = 0; j < n; ++j) {
                const int l = ipvt[j];
                if (wa2[l] == 0.0) continue;
                double sum = 0.0;
                for (int i = 0; i <= j; ++i) sum += fjac[i + (size_t)j * ldfjac] * (qtf[i] / fnorm);
                gnorm = std::max(gnorm, fabs(sum / wa2[l]));
            }
        }
        if (gnorm <= gtol) { info = 4; break; }
        for (int j = 0; j < n; ++j) diag[j] = std::max(diag[j], wa2[j]);

        double ratio = 0.0;
        do {
            lmpar(n, fjac, ldfjac, ipvt, diag.data(), qtf.data(), delta, &par, wa1.data(), wa2.data(), wa3.data(), wa4.data());
            for (int j = 0; j < n; ++j) {
                wa1[j] = -wa1[j];
                wa2[j] = x[j] + wa1[j];
                wa3[j] = diag[j] * wa1[j];
            }
            const double pnorm = enorm(n, wa3.data());
            if (iter == 1) delta = std::min(delta, pnorm);
            fcn(wa2.data(), wa4.data());
            ++nfev;
            const double fnorm1 = enorm(m, wa4.data());
            double actred = -1.0;
            if (0.1 * fnorm1 < fnorm) { const double t = fnorm1 / fnorm; actred = 1.0 - t * t; }
            for (int j = 0; j < n; ++j) {
                wa3[j] = 0.0;
                const double temp = wa1[ipvt[j]];
                for (int i = 0; i <= j; ++i) wa3[i] += fjac[i + (size_t)j * ldfjac] * temp;
            }
            const double temp1 = enorm(n, wa3.data()) / fnorm;
            const double temp2 = (sqrt(par) * pnorm) / fnorm;
            const double prered = temp1 * temp1 + temp2 * temp2 / 0.5;
            const double dirder = -(temp1 * temp1 + temp2 * temp2);
            ratio = 0.0;
            if (prered != 0.0) ratio = actred / prered;
            if (ratio <= 0.25) {
                double temp = (actred >= 0.0) ? 0.5 : 0.5 * dirder / (dirder + 0.5 * actred);
                if (0.1 * fnorm1 >= fnorm || temp < 0.1) temp = 0.1;
                delta = temp * std::min(delta, pnorm / 0.1);
                par /= temp;
            } else if (par == 0.0 || ratio >= 0.75) {
                delta = pnorm / 0.5;
                par *= 0.5;
            }
            if (ratio >= 1e-4) {
                for (int j = 0; j < n; ++j) { x[j] = wa2[j]; wa2[j] = diag[j] * x[j]; }
                for (int i = 0; i < m; ++i) fvec[i] = wa4[i];
                xnorm = enorm(n, wa2.data());
                fnorm = fnorm1;
                ++iter;
            }
            if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0) info = 1;
            if (delta <= xtol * xnorm) info = 2;
            if (fabs(actred) <= ftol && prered <= ftol && 0.5 * ratio <= 1.0 && info == 2) info = 3;
            if (info != 0) break;
            if (nfev >= maxfev) info = 5;
            if (fabs(actred) <= EPSMCH && prered <= EPSMCH && 0.5 * ratio <= 1.0) info = 6;
            if (delta <= EPSMCH * xnorm) info = 7;
            if (gnorm <= EPSMCH) info = 8;
            if (info != 0) break;
        } while (ratio < 1e-4);
        if (info != 0) break;
    }
    *nfev_out = nfev;
    return info;
}

}  // namespace lmfit

// popt[4], resnorm = sum_i sqrt|pcov_ii| (the retry criterion of GaussianBandwidth.py:50-52), R^2 (:54-57).
// Returns ESPER_OK, or ESPER_E_ARG for bad / non-finite input; *info = MINPACK's termination code (1-4 success).
int tanh_fit_host(const double* x, const double* y, int m, const double* a0, double* popt, double* resnorm, double* r_squared,
                  int* info_out, int* nfev_out) {
    using namespace lmfit;
    const int n = 4;
    if (!x || !y || !a0 || !popt || m < n) return ESPER_E_ARG;
    for (int i = 0; i < m; ++i) if (!std::isfinite(x[i]) || !std::isfinite(y[i])) return ESPER_E_ARG;
    for (int j = 0; j < n; ++j) if (!std::isfinite(a0[j])) return ESPER_E_ARG;
    std::vector<double> fvec(m), fjac((size_t)m * n);
    int ipvt[4];
    double a[4] = {a0[0], a0[1], a0[2], a0[3]};
    TanhResidual fcn{x, y, m};
    int nfev = 0;
    const double tol = 1.49012e-8;
    const int info = lmdif(fcn, m, n, a, fvec.data(), tol, tol, 0.0, 200 * (n + 1), EPSMCH, 100.0, &nfev, fjac.data(), ipvt);
    for (int j = 0; j < n; ++j) popt[j] = a[j];
    if (info_out) *info_out = info;
    if (nfev_out) *nfev_out = nfev;

    // covariance as scipy's leastsq + curve_fit build it: cov = P R^-1 R^-T P^T scaled by cost / (m - n)
    double rn = INFINITY;
    if (info >= 1 && info <= 4) {
        double R[4][4] = {}, inv[4][4] = {};
        bool singular = false;
        for (int j = 0; j < n; ++j) {
            for (int i = 0; i <= j; ++i) R[i][j] = fjac[i + (size_t)j * m];
            if (R[j][j] == 0.0) singular = true;
        }
        if (!singular) {
            for (int c = 0; c < n; ++c) {                     // back substitution: R inv = I
                for (int i = n - 1; i >= 0; --i) {
                    double s = (i == c) ? 1.0 : 0.0;
                    for (int k = i + 1; k < n; ++k) s -= R[i][k] * inv[k][c];
                    inv[i][c] = s / R[i][i];
                }
            }
            double cost = 0.0;
            for (int i = 0; i < m; ++i) cost += fvec[i] * fvec[i];
            const double s_sq = (m > n) ? cost / (double)(m - n) : 1.0;
            rn = 0.0;
            for (int i = 0; i < n; ++i) {                     // diagonal entry ipvt[i] of the covariance = |row i of R^-1|^2
                double d = 0.0;
                for (int k = 0; k < n; ++k) d += inv[i][k] * inv[i][k];
                rn += sqrt(fabs(d * s_sq));
            }
            if (!std::isfinite(rn)) rn = INFINITY;
        }
    }
    if (resnorm) *resnorm = rn;
    if (r_squared) {
        double mean = 0.0;
        for (int i = 0; i < m; ++i) mean += y[i];
        mean /= (double)m;
        double ss_res = 0.0, ss_tot = 0.0;
        for (int i = 0; i < m; ++i) {
            const double r = y[i] - (a[3] + a[2] * tanh(a[0] * x[i] + a[1]));
            ss_res += r * r;
            ss_tot += (y[i] - mean) * (y[i] - mean);
        }
        *r_squared = 1.0 - ss_res / ss_tot;
    }
    return ESPER_OK;
}

}  // namespace esper
