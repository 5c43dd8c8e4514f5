/*
 * CPU ORACLE -- TEST / BASELINE INFRASTRUCTURE ONLY.  Plain-C restatement of the reference's
 * Levenberg-Marquardt (Ronnypetson/epivo jac_Rt_gen_.cpp:23-478), used (a) to cross-check the
 * numpy restatement at sizes numpy is slow at and (b) as the LM leg of the CPU baseline that
 * bench.py times next to the GPU path (the Eigen/Sophus original cannot be built here).
 * It follows the reference's data flow: dense J (rep_N x 6 n_zeta), H = J'J, damping on the
 * diagonal, delta = -H^-1 b through an explicit LU inverse, SE3::exp right-multiplied.
 * PARITY UNPINNED against a running reference binary (none can be built); pinned against the
 * numpy restatement and finite differences in tests/test_oracle_lm.py.
 *
 *   gcc -O2 -ffp-contract=off -shared -fPIC oracle/lm_c.c -o oracle/_build/liboracle_lm.so -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double H_norm, r_norm, lambda; } lm_res_t;

static void mat4_mul(const double* a, const double* b, double* o) {
    double r[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
            r[i * 4 + j] = s;
        }
    memcpy(o, r, sizeof(r));
}

/* general inverse by Gauss-Jordan with partial pivoting (the reference calls MatrixXd::inverse()) */
static int mat_inv(const double* a, int n, double* o) {
    double* m = (double*)malloc(sizeof(double) * n * 2 * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            m[i * 2 * n + j] = a[i * n + j];
            m[i * 2 * n + n + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(m[k * 2 * n + k]);
        for (int r = k + 1; r < n; ++r)
            if (fabs(m[r * 2 * n + k]) > best) { best = fabs(m[r * 2 * n + k]); p = r; }
        if (p != k)
            for (int c = 0; c < 2 * n; ++c) { double t = m[k * 2 * n + c]; m[k * 2 * n + c] = m[p * 2 * n + c]; m[p * 2 * n + c] = t; }
        double inv = 1.0 / m[k * 2 * n + k];          /* singular -> inf/nan, as Eigen propagates */
        for (int c = 0; c < 2 * n; ++c) m[k * 2 * n + c] *= inv;
        for (int r = 0; r < n; ++r) {
            if (r == k) continue;
            double f = m[r * 2 * n + k];
            if (f == 0.0) continue;
            for (int c = 0; c < 2 * n; ++c) m[r * 2 * n + c] -= f * m[k * 2 * n + c];
        }
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) o[i * n + j] = m[i * 2 * n + n + j];
    free(m);
    return 0;
}

/* jac_Rt_gen_.cpp:212-259 */
static void res(const double* T, const double* p, const double* p_, int n, double hd, double* r) {
    for (int i = 0; i < n; ++i) {
        const double* a = p + 3 * i;
        const double* b = p_ + 3 * i;
        double A0 = T[3] - b[0] * T[11], A1 = T[7] - b[1] * T[11];
        double q0 = T[0] * a[0] + T[1] * a[1] + T[2] * a[2];
        double q1 = T[4] * a[0] + T[5] * a[1] + T[6] * a[2];
        double q2 = T[8] * a[0] + T[9] * a[1] + T[10] * a[2];
        double B0 = q0 - b[0] * q2, B1 = q1 - b[1] * q2;
        double nb = sqrt(B0 * B0 + B1 * B1), d = 0.0;
        if (nb > 0) d = sqrt(A0 * A0 + A1 * A1) / nb;
        double X0 = q0 * d + T[3], X1 = q1 * d + T[7], X2 = q2 * d + T[11];
        double e0 = b[0] - X0 / X2, e1 = b[1] - X1 / X2, e2 = b[2] - X2 / X2;
        double ri = (e0 * e0 + e1 * e1 + e2 * e2) / 2.0;
        if (ri > hd) ri = hd * (sqrt(ri) - hd / 2.0);
        r[i] = ri;
    }
}

static const double GEN[6][16] = {
    {0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, -1, 0, 0, 1, 0, 0, 0, 0, 0, 0},
    {0, 0, 1, 0, 0, 0, 0, 0, -1, 0, 0, 0, 0, 0, 0, 0}, {0, -1, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}};

/* jac_Rt_gen_.cpp:23-209; J is n x 6 row-major with row stride ldj */
static void dr_deps(const double* Tl, const double* Tr, const double* p, const double* p_, int n, int reverse,
                    double hd, double* J, int ldj) {
    double T0[16], M[6][16], tmp[16];
    mat4_mul(Tl, Tr, T0);
    double s = reverse ? -1.0 : 1.0;
    for (int j = 0; j < 6; ++j) {
        mat4_mul(Tl, GEN[j], tmp);
        mat4_mul(tmp, Tr, M[j]);
        for (int k = 0; k < 16; ++k) M[j][k] *= s;
    }
    for (int i = 0; i < n; ++i) {
        const double* a = p + 3 * i;
        const double* b = p_ + 3 * i;
        double* row = J + (size_t)i * ldj;
        for (int j = 0; j < 6; ++j) row[j] = 0.0;
        double A0 = T0[3] - b[0] * T0[11], A1 = T0[7] - b[1] * T0[11];
        double q0 = T0[0] * a[0] + T0[1] * a[1] + T0[2] * a[2];
        double q1 = T0[4] * a[0] + T0[5] * a[1] + T0[6] * a[2];
        double q2 = T0[8] * a[0] + T0[9] * a[1] + T0[10] * a[2];
        double B0 = q0 - b[0] * q2, B1 = q1 - b[1] * q2;
        double ATA = A0 * A0 + A1 * A1, BTB = B0 * B0 + B1 * B1;
        if (ATA == 0 || BTB == 0) continue;
        double sa = sqrt(ATA), sb = sqrt(BTB);
        double d0 = sa / sb;
        double X0 = q0 * d0 + T0[3], X1 = q1 * d0 + T0[7], X2 = q2 * d0 + T0[11];
        double j00 = 0, j02 = 0, j12 = 0;
        if (X2 != 0) { j00 = 1.0 / X2; j02 = -X0 / (X2 * X2); j12 = -X1 / (X2 * X2); }
        double e0 = X0 / X2 - b[0], e1 = X1 / X2 - b[1], e2 = 1.0 - b[2];
        double ee = e0 * e0 + e1 * e1 + e2 * e2;
        double g0 = e0, g1 = e1;
        if (!(ee <= hd)) { double k = hd / sqrt(ee); g0 = k * e0; g1 = k * e1; }
        for (int j = 0; j < 6; ++j) {
            const double* m = M[j];
            double mp0 = 0, mp1 = 0, mp2 = 0;
            if (j >= 3) {
                mp0 = m[0] * a[0] + m[1] * a[1] + m[2] * a[2];
                mp1 = m[4] * a[0] + m[5] * a[1] + m[6] * a[2];
                mp2 = m[8] * a[0] + m[9] * a[1] + m[10] * a[2];
            }
            double dA0 = m[3] - b[0] * m[11], dA1 = m[7] - b[1] * m[11];
            double dB0 = mp0 - b[0] * mp2, dB1 = mp1 - b[1] * mp2;
            double jd = ((1.0 / sa) * sb * (A0 * dA0 + A1 * dA1) - (1.0 / sb) * sa * (B0 * dB0 + B1 * dB1)) / BTB;
            /* M_j [p d0; 1] uses the full 3x3 block of M_j (zero for j < 3) */
            double f0 = (m[0] * a[0] + m[1] * a[1] + m[2] * a[2]) * d0 + m[3] + q0 * jd;
            double f1 = (m[4] * a[0] + m[5] * a[1] + m[6] * a[2]) * d0 + m[7] + q1 * jd;
            double f2 = (m[8] * a[0] + m[9] * a[1] + m[10] * a[2]) * d0 + m[11] + q2 * jd;
            row[j] = g0 * (j00 * f0 + j02 * f2) + g1 * (j00 * f1 + j12 * f2);
        }
    }
}

static void se3_exp(const double* d, double* T) {
    double wx = d[3], wy = d[4], wz = d[5];
    double th2 = wx * wx + wy * wy + wz * wz, th = sqrt(th2), a, b, c;
    if (th < 1e-10) { a = 1.0; b = 0.5; c = 1.0 / 6.0; }
    else { a = sin(th) / th; b = (1.0 - cos(th)) / th2; c = (th - sin(th)) / (th2 * th); }
    double Om[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0}, Om2[9], V[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Om2[i * 3 + j] = Om[i * 3] * Om[j] + Om[i * 3 + 1] * Om[3 + j] + Om[i * 3 + 2] * Om[6 + j];
    memset(T, 0, 16 * sizeof(double));
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double I = (i == j) ? 1.0 : 0.0;
            T[i * 4 + j] = I + a * Om[i * 3 + j] + b * Om2[i * 3 + j];
            /* Sophus: V = so3.matrix() below epsilon = 1e-10 */
            V[i * 3 + j] = (th < 1e-10) ? T[i * 4 + j] : I + b * Om[i * 3 + j] + c * Om2[i * 3 + j];
        }
    for (int i = 0; i < 3; ++i) T[i * 4 + 3] = V[i * 3] * d[0] + V[i * 3 + 1] * d[1] + V[i * 3 + 2] * d[2];
    T[15] = 1.0;
}

/* jac_Rt_gen_.cpp:287-478.  T0s (n_zeta x 16) is updated in place.  Returns iterations started.
 * trace (optional, 2 x max_iters doubles): per started iteration |delta| and the candidate |r0| (NaN when the
 * loop broke before computing it) -- the same two numbers oracle/_ref exposes of the reference's trajectory. */
int oracle_lm_trace(int n_zeta, double epsilon, const int* reps, const double* wreps, int n_rep, double lambda0,
                    int max_iters, double hd, double* T0s, const double* pr, const double* p_r, int N, lm_res_t* out,
                    double* trace) {
    const int D = 6 * n_zeta, RN = n_rep * N;
    double* mem = (double*)calloc((size_t)n_zeta * n_zeta * 16, sizeof(double));
    double* r0 = (double*)calloc(RN, sizeof(double));
    double* J = (double*)calloc((size_t)RN * D, sizeof(double));
    double* H = (double*)calloc((size_t)D * D, sizeof(double));
    double* Hi = (double*)calloc((size_t)D * D, sizeof(double));
    double* b = (double*)calloc(D, sizeof(double));
    double* delta = (double*)calloc(D, sizeof(double));
    double* Tn = (double*)calloc((size_t)n_zeta * 16, sizeof(double));
    double* rr = (double*)calloc(N, sizeof(double));
    double lambda = lambda0, prevE = 1e10;
    int it = 0;
    for (int iter = 0; iter < max_iters; ++iter) {
        it = iter + 1;
        memset(r0, 0, sizeof(double) * RN);
        memset(J, 0, sizeof(double) * (size_t)RN * D);
        for (int j = 0; j < n_zeta; ++j) {                                   /* :328-335 */
            double sT[16];
            memcpy(sT, T0s + 16 * j, sizeof(sT));
            memcpy(mem + ((size_t)j * n_zeta + j) * 16, sT, sizeof(sT));
            for (int k = j + 1; k < n_zeta; ++k) {
                mat4_mul(T0s + 16 * k, sT, sT);
                memcpy(mem + ((size_t)j * n_zeta + k) * 16, sT, sizeof(sT));
            }
        }
        for (int j = 0; j < n_rep; ++j) {                                    /* :338-360 */
            int z0 = reps[2 * j], z1 = reps[2 * j + 1];
            double T[16];
            if (z0 <= z1) memcpy(T, mem + ((size_t)z0 * n_zeta + z1) * 16, sizeof(T));
            else mat_inv(mem + ((size_t)z1 * n_zeta + z0) * 16, 4, T);
            res(T, pr + (size_t)j * N * 3, p_r + (size_t)j * N * 3, N, hd, rr);
            for (int i = 0; i < N; ++i) r0[j * N + i] = wreps[j] * rr[i];
        }
        for (int j = 0; j < n_rep; ++j) {                                    /* :363-399 */
            int z0 = reps[2 * j], z1 = reps[2 * j + 1];
            int lo = z0 < z1 ? z0 : z1, hi = z0 < z1 ? z1 : z0;
            for (int k = lo; k <= hi; ++k) {
                double Tl[16], Tr[16];
                memset(Tr, 0, sizeof(Tr));
                Tr[0] = Tr[5] = Tr[10] = Tr[15] = 1.0;
                if (z0 <= z1) {                                              /* :271-275 */
                    if (z0 < k) memcpy(Tr, mem + ((size_t)z0 * n_zeta + (k - 1)) * 16, sizeof(Tr));
                    memcpy(Tl, mem + ((size_t)k * n_zeta + z1) * 16, sizeof(Tl));
                } else {                                                     /* :276-281 */
                    if (z0 > k) mat_inv(mem + ((size_t)(k + 1) * n_zeta + z0) * 16, 4, Tr);
                    mat_inv(mem + ((size_t)z1 * n_zeta + k) * 16, 4, Tl);
                }
                double* Jb = J + (size_t)j * N * D + 6 * k;
                dr_deps(Tl, Tr, pr + (size_t)j * N * 3, p_r + (size_t)j * N * 3, N, z0 > z1, hd, Jb, D);
                for (int i = 0; i < N; ++i)
                    for (int c = 0; c < 6; ++c) Jb[(size_t)i * D + c] *= wreps[j];
            }
        }
        for (int a = 0; a < D; ++a) {                                        /* :401-403 */
            double s = 0;
            for (int i = 0; i < RN; ++i) s += J[(size_t)i * D + a] * r0[i];
            b[a] = s;
            for (int c = a; c < D; ++c) {
                double h = 0;
                for (int i = 0; i < RN; ++i) h += J[(size_t)i * D + a] * J[(size_t)i * D + c];
                H[(size_t)a * D + c] = h;
                H[(size_t)c * D + a] = h;
            }
        }
        for (int a = 0; a < D; ++a) H[(size_t)a * D + a] += lambda * H[(size_t)a * D + a];
        mat_inv(H, D, Hi);                                                   /* :405 */
        int has_nan = 0;
        double dn = 0;
        for (int a = 0; a < D; ++a) {
            double s = 0;
            for (int c = 0; c < D; ++c) s += Hi[(size_t)a * D + c] * b[c];
            delta[a] = -s;
            if (!(delta[a] == delta[a])) has_nan = 1;                        /* hasNaN(): NaN only, inf passes */
            dn += delta[a] * delta[a];
        }
        if (trace) { trace[2 * iter] = sqrt(dn); trace[2 * iter + 1] = NAN; }
        if (has_nan) break;                                                  /* :407-410 */
        if (sqrt(dn) < epsilon) break;                                       /* :412-414 */
        for (int j = 0; j < n_zeta; ++j) {                                   /* :416-422 */
            double ex[16];
            se3_exp(delta + 6 * j, ex);
            mat4_mul(T0s + 16 * j, ex, Tn + 16 * j);
        }
        for (int j = 0; j < n_rep; ++j) {                                    /* :425-454 */
            int z0 = reps[2 * j], z1 = reps[2 * j + 1];
            double T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}, inv[16];
            if (z0 <= z1) for (int k = z0; k <= z1; ++k) mat4_mul(Tn + 16 * k, T, T);
            else for (int k = z0; k >= z1; --k) { mat_inv(Tn + 16 * k, 4, inv); mat4_mul(inv, T, T); }
            res(T, pr + (size_t)j * N * 3, p_r + (size_t)j * N * 3, N, hd, r0 + j * N);
        }
        double cs = 0;
        for (int i = 0; i < RN; ++i) cs += r0[i] * r0[i];
        double currE = sqrt(cs);                                             /* :456-467 */
        if (trace) trace[2 * iter + 1] = currE;
        if (currE < prevE) {
            prevE = currE;
            memcpy(T0s, Tn, sizeof(double) * (size_t)n_zeta * 16);
            lambda /= 2.0;
        } else {
            lambda *= 5.0;
        }
    }
    double hs = 0, rs = 0;
    for (size_t i = 0; i < (size_t)D * D; ++i) hs += H[i] * H[i];
    for (int i = 0; i < RN; ++i) rs += r0[i] * r0[i];
    out->H_norm = sqrt(hs);                                                  /* :473-475 */
    out->r_norm = sqrt(rs);
    out->lambda = lambda;
    free(mem); free(r0); free(J); free(H); free(Hi); free(b); free(delta); free(Tn); free(rr);
    return it;
}

int oracle_lm(int n_zeta, double epsilon, const int* reps, const double* wreps, int n_rep, double lambda0,
              int max_iters, double hd, double* T0s, const double* pr, const double* p_r, int N, lm_res_t* out) {
    return oracle_lm_trace(n_zeta, epsilon, reps, wreps, n_rep, lambda0, max_iters, hd, T0s, pr, p_r, N, out, 0);
}
