/*
 * CPU ORACLE -- TEST / BASELINE INFRASTRUCTURE ONLY.  Plain-C restatement of cv::solvePoly
 * (OpenCV modules/core/src/mathfuncs.cpp, 4.x), the Durand-Kerner root finder that OpenCV's five-point solver
 * calls with its default of 300 iterations (five-point.cpp: `solvePoly(coeffs, roots)`): start values (1+i)^k,
 * Gauss-Seidel sweeps in root order, the update num/denom with OpenCV's complex division, a factor skipped when two
 * iterates coincide exactly, stop only when a whole sweep changes nothing.  (OpenCV's extra handling of EXACTLY
 * coinciding iterates -- taking roots of the update -- is not restated: it needs two roots equal to the last bit.)
 * The number of sweeps matters: close pairs of real roots converge linearly for dozens of sweeps, and whether such a
 * pair ends up real (|imag| <= 1e-10 in five-point.cpp) decides which models a sample contributes.
 *
 * coeffs[0..n0]: ascending real coefficients; roots: n0 (re, im) pairs out; returns the degree actually solved.
 */
#include <float.h>
#include <math.h>

int oracle_solve_poly(const double* coeffs, int n0, int max_iters, double* roots) {
    int n = n0;
    double re[64], im[64];
    if (n0 > 64) return -1;
    for (; n > 1; --n)
        if (fabs(coeffs[n]) > DBL_EPSILON) break;
    {
        double pr = 1.0, pi = 0.0;
        for (int i = 0; i < n; ++i) {
            re[i] = pr; im[i] = pi;
            const double t = pr * 1.0 - pi * 1.0;      /* p = p * (1 + 1i) */
            pi = pr * 1.0 + pi * 1.0;
            pr = t;
        }
    }
    if (max_iters <= 0) max_iters = 1000;
    for (int iter = 0; iter < max_iters; ++iter) {
        double max_diff = 0.0;
        for (int i = 0; i < n; ++i) {
            const double xr = re[i], xi = im[i];
            double nr = coeffs[n], ni = 0.0, dr = coeffs[n], di = 0.0;
            for (int j = 0; j < n; ++j) {
                const double t = nr * xr - ni * xi + coeffs[n - j - 1];
                ni = nr * xi + ni * xr;
                nr = t;
                if (j != i) {
                    const double er = xr - re[j], ei = xi - im[j];
                    if (er != 0.0 || ei != 0.0) {
                        const double u = dr * er - di * ei;
                        di = dr * ei + di * er;
                        dr = u;
                    }
                }
            }
            {
                const double t = 1.0 / (dr * dr + di * di);
                const double qr = (nr * dr + ni * di) * t, qi = (-nr * di + ni * dr) * t;
                const double a = sqrt(qr * qr + qi * qi);
                re[i] = xr - qr;
                im[i] = xi - qi;
                if (a > max_diff) max_diff = a;
            }
        }
        if (max_diff <= 0) break;
    }
    for (int i = 0; i < n; ++i) {
        roots[2 * i] = re[i];
        roots[2 * i + 1] = fabs(im[i]) < 1e-100 ? 0.0 : im[i];
    }
    for (int i = n; i < n0; ++i) { roots[2 * i] = roots[2 * (n - 1)]; roots[2 * i + 1] = roots[2 * (n - 1) + 1]; }
    return n;
}
