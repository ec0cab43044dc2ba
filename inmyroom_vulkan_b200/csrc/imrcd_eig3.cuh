// imrcd_eig3.cuh -- the reference's OBB fit on the device, bit for bit:
//   eigen_decomposition (eig3/eig3.cpp:256-265; tred2 :21-134, tql2 :138-254: the public-domain JAMA routines) in FP64, and
//   OBB::CreateOBBfromPoints / CreateAABBfromPoints (IMR/src/Geometry/OBB.cpp:33-166) including the rows-of-V axes (:80-87).
// Compiled with --fmad=false: every double operation below is one IEEE-754 binary64 operation in the order written, which is
// what g++ -O2 -ffp-contract=off emits for the reference.  CUDA's double sqrt and division are correctly rounded.
#pragma once
#include "imrcd_math.cuh"
#include <cfloat>

IMR_D double e3_hypot2(double x, double y) { return sqrt(x * x + y * y); }        // eig3.cpp:15-17

// eig3.cpp:21-134, n = 3
__device__ inline void e3_tred2(double V[3][3], double d[3], double e[3]) {
    const int n = 3;
    for (int j = 0; j < n; j++) d[j] = V[n - 1][j];
    for (int i = n - 1; i > 0; i--) {
        double scale = 0.0, h = 0.0;
        for (int k = 0; k < i; k++) scale = scale + fabs(d[k]);
        if (scale == 0.0) {
            e[i] = d[i - 1];
            for (int j = 0; j < i; j++) { d[j] = V[i - 1][j]; V[i][j] = 0.0; V[j][i] = 0.0; }
        } else {
            for (int k = 0; k < i; k++) { d[k] /= scale; h += d[k] * d[k]; }
            double f = d[i - 1];
            double g = sqrt(h);
            if (f > 0) g = -g;
            e[i] = scale * g;
            h = h - f * g;
            d[i - 1] = f - g;
            for (int j = 0; j < i; j++) e[j] = 0.0;
            for (int j = 0; j < i; j++) {
                f = d[j];
                V[j][i] = f;
                g = e[j] + V[j][j] * f;
                for (int k = j + 1; k <= i - 1; k++) { g += V[k][j] * d[k]; e[k] += V[k][j] * f; }
                e[j] = g;
            }
            f = 0.0;
            for (int j = 0; j < i; j++) { e[j] /= h; f += e[j] * d[j]; }
            const double hh = f / (h + h);
            for (int j = 0; j < i; j++) e[j] -= hh * d[j];
            for (int j = 0; j < i; j++) {
                f = d[j]; g = e[j];
                for (int k = j; k <= i - 1; k++) V[k][j] -= (f * e[k] + g * d[k]);
                d[j] = V[i - 1][j];
                V[i][j] = 0.0;
            }
        }
        d[i] = h;
    }
    for (int i = 0; i < n - 1; i++) {
        V[n - 1][i] = V[i][i];
        V[i][i] = 1.0;
        const double h = d[i + 1];
        if (h != 0.0) {
            for (int k = 0; k <= i; k++) d[k] = V[k][i + 1] / h;
            for (int j = 0; j <= i; j++) {
                double g = 0.0;
                for (int k = 0; k <= i; k++) g += V[k][i + 1] * V[k][j];
                for (int k = 0; k <= i; k++) V[k][j] -= g * d[k];
            }
        }
        for (int k = 0; k <= i; k++) V[k][i + 1] = 0.0;
    }
    for (int j = 0; j < n; j++) { d[j] = V[n - 1][j]; V[n - 1][j] = 0.0; }
    V[n - 1][n - 1] = 1.0;
    e[0] = 0.0;
}

// eig3.cpp:138-254, n = 3
__device__ inline void e3_tql2(double V[3][3], double d[3], double e[3]) {
    const int n = 3;
    for (int i = 1; i < n; i++) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    double f = 0.0, tst1 = 0.0;
    const double eps = 2.220446049250313e-16;      // pow(2.0, -52.0)
    for (int l = 0; l < n; l++) {
        const double t = fabs(d[l]) + fabs(e[l]);
        tst1 = (tst1 > t) ? tst1 : t;
        int m = l;
        while (m < n) { if (fabs(e[m]) <= eps * tst1) break; m++; }
        if (m > l) {
            do {
                double g = d[l];
                double p = (d[l + 1] - g) / (2.0 * e[l]);
                double r = e3_hypot2(p, 1.0);
                if (p < 0) r = -r;
                d[l] = e[l] / (p + r);
                d[l + 1] = e[l] * (p + r);
                const double dl1 = d[l + 1];
                double h = g - d[l];
                for (int i = l + 2; i < n; i++) d[i] -= h;
                f = f + h;
                p = d[m];
                double c = 1.0, c2 = c, c3 = c;
                const double el1 = e[l + 1];
                double s = 0.0, s2 = 0.0;
                for (int i = m - 1; i >= l; i--) {
                    c3 = c2; c2 = c; s2 = s;
                    g = c * e[i];
                    h = c * p;
                    r = e3_hypot2(p, e[i]);
                    e[i + 1] = s * r;
                    s = e[i] / r;
                    c = p / r;
                    p = c * d[i] - s * g;
                    d[i + 1] = h + s * (c * g + s * d[i]);
                    for (int k = 0; k < n; k++) {
                        h = V[k][i + 1];
                        V[k][i + 1] = s * V[k][i] + c * h;
                        V[k][i] = c * V[k][i] - s * h;
                    }
                }
                p = -s * s2 * c3 * el1 * e[l] / dl1;
                e[l] = s * p;
                d[l] = c * p;
            } while (fabs(e[l]) > eps * tst1);
        }
        d[l] = d[l] + f;
        e[l] = 0.0;
    }
    for (int i = 0; i < n - 1; i++) {
        int k = i;
        double p = d[i];
        for (int j = i + 1; j < n; j++) if (d[j] < p) { k = j; p = d[j]; }
        if (k != i) {
            d[k] = d[i]; d[i] = p;
            for (int j = 0; j < n; j++) { p = V[j][i]; V[j][i] = V[j][k]; V[j][k] = p; }
        }
    }
}

// eig3.cpp:256-265.  A, V row-major 3x3.
__device__ inline void e3_eigen_decomposition(const double A[9], double Vout[9], double d[3]) {
    double V[3][3], e[3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[i][j] = A[3 * i + j];
    e3_tred2(V, d, e);
    e3_tql2(V, d, e);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Vout[3 * i + j] = V[i][j];
}
