// Eigenvalues of a small dense real nonsymmetric matrix, one matrix per thread: balancing, reduction to Hessenberg
// form by stabilised elimination, then the implicit double-shift (Francis) QR iteration -- the textbook EISPACK
// sequence balanc / elmhes / hqr, written for compile-time N so that everything stays in registers / local memory.
// Replaces the per-point np.linalg.eigvals call of compute_einstein_eigenvalues
// (/root/reference/src/irrotational_warp/einstein.py:293-298), which runs LAPACK dgeev on each 4x4 G^mu_nu.
// Plain C++ (also compiles for the host: tests drive it on the CPU through IWB_HD).
#pragma once
#include <math.h>

#ifndef IWB_HD
#ifdef __CUDACC__
#define IWB_HD __host__ __device__ __forceinline__
#else
#define IWB_HD inline
#endif
#endif

namespace iwb {

template <int N>
IWB_HD void eig_balance(double (&a)[N][N])
{
    const double RADIX = 2.0, SQRDX = 4.0;
    bool done = false;
    for (int sweep = 0; sweep < 64 && !done; ++sweep) {
        done = true;
        for (int i = 0; i < N; ++i) {
            double r = 0.0, c = 0.0;
            for (int j = 0; j < N; ++j)
                if (j != i) { c += fabs(a[j][i]); r += fabs(a[i][j]); }
            if (c != 0.0 && r != 0.0 && isfinite(c) && isfinite(r)) {
                double g = r / RADIX, f = 1.0;
                const double s = c + r;
                while (c < g) { f *= RADIX; c *= SQRDX; }
                g = r * RADIX;
                while (c > g) { f /= RADIX; c /= SQRDX; }
                if ((c + r) / f < 0.95 * s) {
                    done = false;
                    g = 1.0 / f;
                    for (int j = 0; j < N; ++j) a[i][j] *= g;
                    for (int j = 0; j < N; ++j) a[j][i] *= f;
                }
            }
        }
    }
}

template <int N>
IWB_HD void eig_hessenberg(double (&a)[N][N])
{
    for (int m = 1; m < N - 1; ++m) {
        double x = 0.0;
        int piv = m;
        for (int j = m; j < N; ++j)
            if (fabs(a[j][m - 1]) > fabs(x)) { x = a[j][m - 1]; piv = j; }
        if (piv != m) {
            for (int j = m - 1; j < N; ++j) { const double t = a[piv][j]; a[piv][j] = a[m][j]; a[m][j] = t; }
            for (int j = 0; j < N; ++j) { const double t = a[j][piv]; a[j][piv] = a[j][m]; a[j][m] = t; }
        }
        if (x != 0.0) {
            for (int i = m + 1; i < N; ++i) {
                double y = a[i][m - 1];
                if (y != 0.0) {
                    y /= x;
                    a[i][m - 1] = y;
                    for (int j = m; j < N; ++j) a[i][j] -= y * a[m][j];
                    for (int j = 0; j < N; ++j) a[j][m] += y * a[j][i];
                }
            }
        }
    }
    for (int i = 2; i < N; ++i)
        for (int j = 0; j < i - 1; ++j) a[i][j] = 0.0;
}

IWB_HD double eig_sign(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

// Eigenvalues of an upper Hessenberg matrix (destroyed).  Returns false if an eigenvalue did not converge in 60
// iterations (the remaining wr / wi are then NaN) or if the matrix holds a non-finite entry.
template <int N>
IWB_HD bool eig_hqr(double (&a)[N][N], double (&wr)[N], double (&wi)[N])
{
    const double qnan = NAN;
    double anorm = 0.0;
    for (int i = 0; i < N; ++i) {
        wr[i] = qnan; wi[i] = qnan;
        for (int j = (i > 0 ? i - 1 : 0); j < N; ++j) anorm += fabs(a[i][j]);
    }
    if (!isfinite(anorm)) return false;
    int nn = N - 1;
    double t = 0.0;
    double p = 0.0, q = 0.0, r = 0.0;
    while (nn >= 0) {
        int its = 0, l;
        do {
            for (l = nn; l >= 1; --l) {
                double s = fabs(a[l - 1][l - 1]) + fabs(a[l][l]);
                if (s == 0.0) s = anorm;
                if (fabs(a[l][l - 1]) + s == s) { a[l][l - 1] = 0.0; break; }
            }
            double x = a[nn][nn];
            if (l == nn) {                         // one real root
                wr[nn] = x + t; wi[nn] = 0.0;
                --nn;
            } else {
                double y = a[nn - 1][nn - 1];
                double w = a[nn][nn - 1] * a[nn - 1][nn];
                if (l == nn - 1) {                 // a 2x2 block: two real roots or a conjugate pair
                    p = 0.5 * (y - x);
                    q = p * p + w;
                    double z = sqrt(fabs(q));
                    x += t;
                    if (q >= 0.0) {
                        z = p + eig_sign(z, p);
                        wr[nn - 1] = wr[nn] = x + z;
                        if (z != 0.0) wr[nn] = x - w / z;
                        wi[nn - 1] = wi[nn] = 0.0;
                    } else {
                        wr[nn - 1] = wr[nn] = x + p;
                        wi[nn - 1] = z; wi[nn] = -z;
                    }
                    nn -= 2;
                } else {
                    if (its == 60) return false;
                    if (its == 10 || its == 20 || its == 40) {      // exceptional shift
                        t += x;
                        for (int i = 0; i <= nn; ++i) a[i][i] -= x;
                        const double s = fabs(a[nn][nn - 1]) + fabs(a[nn - 1][nn - 2]);
                        y = x = 0.75 * s;
                        w = -0.4375 * s * s;
                    }
                    ++its;
                    int m;
                    double z;
                    for (m = nn - 2; m >= l; --m) {
                        z = a[m][m];
                        r = x - z;
                        double s = y - z;
                        p = (r * s - w) / a[m + 1][m] + a[m][m + 1];
                        q = a[m + 1][m + 1] - z - r - s;
                        r = a[m + 2][m + 1];
                        s = fabs(p) + fabs(q) + fabs(r);
                        p /= s; q /= s; r /= s;
                        if (m == l) break;
                        const double u = fabs(a[m][m - 1]) * (fabs(q) + fabs(r));
                        const double v = fabs(p) * (fabs(a[m - 1][m - 1]) + fabs(z) + fabs(a[m + 1][m + 1]));
                        if (u + v == v) break;
                    }
                    for (int i = m + 2; i <= nn; ++i) {
                        a[i][i - 2] = 0.0;
                        if (i != m + 2) a[i][i - 3] = 0.0;
                    }
                    for (int k = m; k <= nn - 1; ++k) {
                        if (k != m) {
                            p = a[k][k - 1];
                            q = a[k + 1][k - 1];
                            r = 0.0;
                            if (k != nn - 1) r = a[k + 2][k - 1];
                            if ((x = fabs(p) + fabs(q) + fabs(r)) != 0.0) { p /= x; q /= x; r /= x; }
                        }
                        const double s = eig_sign(sqrt(p * p + q * q + r * r), p);
                        if (s != 0.0) {
                            if (k == m) {
                                if (l != m) a[k][k - 1] = -a[k][k - 1];
                            } else
                                a[k][k - 1] = -s * x;
                            p += s;
                            x = p / s; y = q / s; z = r / s;
                            q /= p; r /= p;
                            for (int j = k; j <= nn; ++j) {
                                p = a[k][j] + q * a[k + 1][j];
                                if (k != nn - 1) { p += r * a[k + 2][j]; a[k + 2][j] -= p * z; }
                                a[k + 1][j] -= p * y;
                                a[k][j] -= p * x;
                            }
                            const int mmin = nn < k + 3 ? nn : k + 3;
                            for (int i = l; i <= mmin; ++i) {
                                p = x * a[i][k] + y * a[i][k + 1];
                                if (k != nn - 1) { p += z * a[i][k + 2]; a[i][k + 2] -= p * r; }
                                a[i][k + 1] -= p * q;
                                a[i][k] -= p;
                            }
                        }
                    }
                }
            }
        } while (l < nn - 1);
    }
    return true;
}

// All eigenvalues of a general real N x N matrix (destroyed).
template <int N>
IWB_HD bool eig_general(double (&a)[N][N], double (&wr)[N], double (&wi)[N])
{
    eig_balance<N>(a);
    eig_hessenberg<N>(a);
    return eig_hqr<N>(a, wr, wi);
}

}  // namespace iwb
