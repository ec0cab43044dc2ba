/* oracle/imr_oracle.c -- TEST INFRASTRUCTURE ONLY (see imr_oracle.h).
 *
 * Plain-C restatement of the reference's CPU collision path.  Compile with
 * -O2 -ffp-contract=off and WITHOUT -ffast-math (oracle/Makefile), so that
 * every float/double operation below is one IEEE-754 operation in exactly the
 * order written -- that is the reference's evaluation order (SURVEY.md
 * finding 2).  "IMR/" = /root/reference/inMyRoom_vulkan/.
 */
#include "imr_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* ------------------------------------------------------------------ */
/* glm arithmetic, restated (glm @99e83f5, scalar code paths)          */
/* ------------------------------------------------------------------ */
typedef struct { float x, y, z; } v3;

/* glm/detail/func_geometric.inl:48-55 : tmp = a*b; tmp.x + tmp.y + tmp.z */
static float dot3(v3 a, v3 b) { float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z; return tx + ty + tz; }
/* func_geometric.inl:68-79 */
static v3 cross3(v3 x, v3 y) {
    v3 r;
    r.x = x.y * y.z - y.y * x.z;
    r.y = x.z * y.x - y.z * x.x;
    r.z = x.x * y.y - y.x * x.y;
    return r;
}
static v3 sub3(v3 a, v3 b) { v3 r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }
static v3 add3(v3 a, v3 b) { v3 r = { a.x + b.x, a.y + b.y, a.z + b.z }; return r; }
static v3 scale3(v3 a, float s) { v3 r = { a.x * s, a.y * s, a.z * s }; return r; }
/* func_geometric.inl:8-14 : sqrt(dot(v,v)) */
static float length3(v3 a) { return sqrtf(dot3(a, a)); }
/* func_geometric.inl:82-90 + func_exponential.inl:134-139 : v * (1/sqrt(dot(v,v))) */
static v3 normalize3(v3 a) { float inv = 1.0f / sqrtf(dot3(a, a)); return scale3(a, inv); }

/* mat4 is column-major: m[4*c + r].  type_mat4x4.inl:561-572:
 * (m[0]*v0 + m[1]*v1) + (m[2]*v2 + m[3]*v3), then glm::vec3(...) drops w. */
static v3 mat4_mul_point(const float* m, v3 p, float w) {
    v3 r;
    r.x = (m[0] * p.x + m[4] * p.y) + (m[8]  * p.z + m[12] * w);
    r.y = (m[1] * p.x + m[5] * p.y) + (m[9]  * p.z + m[13] * w);
    r.z = (m[2] * p.x + m[6] * p.y) + (m[10] * p.z + m[14] * w);
    return r;
}
/* type_mat4x4.inl:630-648 : Result[c] = A0*B[c][0] + A1*B[c][1] + A2*B[c][2] + A3*B[c][3] (left to right) */
static void mat4_mul(const float* a, const float* b, float* out) {
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r)
            out[4 * c + r] = ((a[0 + r] * b[4 * c + 0] + a[4 + r] * b[4 * c + 1]) + a[8 + r] * b[4 * c + 2]) + a[12 + r] * b[4 * c + 3];
}
/* func_matrix.inl:347-405 (compute_inverse<4,4>) */
#define M(c, r) m[4 * (c) + (r)]
static void mat4_inverse(const float* m, float* out) {
    float Coef00 = M(2,2) * M(3,3) - M(3,2) * M(2,3);
    float Coef02 = M(1,2) * M(3,3) - M(3,2) * M(1,3);
    float Coef03 = M(1,2) * M(2,3) - M(2,2) * M(1,3);
    float Coef04 = M(2,1) * M(3,3) - M(3,1) * M(2,3);
    float Coef06 = M(1,1) * M(3,3) - M(3,1) * M(1,3);
    float Coef07 = M(1,1) * M(2,3) - M(2,1) * M(1,3);
    float Coef08 = M(2,1) * M(3,2) - M(3,1) * M(2,2);
    float Coef10 = M(1,1) * M(3,2) - M(3,1) * M(1,2);
    float Coef11 = M(1,1) * M(2,2) - M(2,1) * M(1,2);
    float Coef12 = M(2,0) * M(3,3) - M(3,0) * M(2,3);
    float Coef14 = M(1,0) * M(3,3) - M(3,0) * M(1,3);
    float Coef15 = M(1,0) * M(2,3) - M(2,0) * M(1,3);
    float Coef16 = M(2,0) * M(3,2) - M(3,0) * M(2,2);
    float Coef18 = M(1,0) * M(3,2) - M(3,0) * M(1,2);
    float Coef19 = M(1,0) * M(2,2) - M(2,0) * M(1,2);
    float Coef20 = M(2,0) * M(3,1) - M(3,0) * M(2,1);
    float Coef22 = M(1,0) * M(3,1) - M(3,0) * M(1,1);
    float Coef23 = M(1,0) * M(2,1) - M(2,0) * M(1,1);
    float Fac0[4] = { Coef00, Coef00, Coef02, Coef03 };
    float Fac1[4] = { Coef04, Coef04, Coef06, Coef07 };
    float Fac2[4] = { Coef08, Coef08, Coef10, Coef11 };
    float Fac3[4] = { Coef12, Coef12, Coef14, Coef15 };
    float Fac4[4] = { Coef16, Coef16, Coef18, Coef19 };
    float Fac5[4] = { Coef20, Coef20, Coef22, Coef23 };
    float Vec0[4] = { M(1,0), M(0,0), M(0,0), M(0,0) };
    float Vec1[4] = { M(1,1), M(0,1), M(0,1), M(0,1) };
    float Vec2[4] = { M(1,2), M(0,2), M(0,2), M(0,2) };
    float Vec3[4] = { M(1,3), M(0,3), M(0,3), M(0,3) };
    static const float SignA[4] = { +1.f, -1.f, +1.f, -1.f };
    static const float SignB[4] = { -1.f, +1.f, -1.f, +1.f };
    float inv[16];
    for (int i = 0; i < 4; ++i) {
        float Inv0 = (Vec1[i] * Fac0[i] - Vec2[i] * Fac1[i]) + Vec3[i] * Fac2[i];
        float Inv1 = (Vec0[i] * Fac0[i] - Vec2[i] * Fac3[i]) + Vec3[i] * Fac4[i];
        float Inv2 = (Vec0[i] * Fac1[i] - Vec1[i] * Fac3[i]) + Vec3[i] * Fac5[i];
        float Inv3 = (Vec0[i] * Fac2[i] - Vec1[i] * Fac4[i]) + Vec2[i] * Fac5[i];
        inv[0 + i]  = Inv0 * SignA[i];
        inv[4 + i]  = Inv1 * SignB[i];
        inv[8 + i]  = Inv2 * SignA[i];
        inv[12 + i] = Inv3 * SignB[i];
    }
    float Dot0x = M(0,0) * inv[0], Dot0y = M(0,1) * inv[4], Dot0z = M(0,2) * inv[8], Dot0w = M(0,3) * inv[12];
    float Dot1 = (Dot0x + Dot0y) + (Dot0z + Dot0w);
    float OneOverDeterminant = 1.0f / Dot1;
    for (int i = 0; i < 16; ++i) out[i] = inv[i] * OneOverDeterminant;
}
#undef M

/* ------------------------------------------------------------------ */
/* Paralgram / OBB (IMR/include/Geometry/Paralgram.h:29-35)            */
/* ------------------------------------------------------------------ */
typedef struct { v3 c, u, v, w; } box_t;

static box_t box_load(const float* f) {
    box_t b = { { f[0], f[1], f[2] }, { f[3], f[4], f[5] }, { f[6], f[7], f[8] }, { f[9], f[10], f[11] } };
    return b;
}
static void box_store(box_t b, float* f) {
    f[0] = b.c.x; f[1] = b.c.y; f[2] = b.c.z; f[3] = b.u.x; f[4] = b.u.y; f[5] = b.u.z;
    f[6] = b.v.x; f[7] = b.v.y; f[8] = b.v.z; f[9] = b.w.x; f[10] = b.w.y; f[11] = b.w.z;
}
/* IMR/src/Geometry/Paralgram.cpp:4-15 */
static box_t box_transform(const float* m, box_t b) {
    box_t r;
    r.c = mat4_mul_point(m, b.c, 1.f);
    r.u = mat4_mul_point(m, b.u, 0.f);
    r.v = mat4_mul_point(m, b.v, 0.f);
    r.w = mat4_mul_point(m, b.w, 0.f);
    return r;
}
/* Paralgram.cpp:175-196 */
static void box_minmax(const box_t* b, v3 axis, float* mn, float* mx) {
    float cp = dot3(b->c, axis);
    float pu = fabsf(dot3(axis, b->u));
    float pv = fabsf(dot3(axis, b->v));
    float pw = fabsf(dot3(axis, b->w));
    float sum = pu + pv + pw;
    *mn = cp - sum;
    *mx = cp + sum;
}
/* Paralgram.cpp:198-201 */
static int axis_overlap(const box_t* l, const box_t* r, v3 axis) {
    float lmn, lmx, rmn, rmx;
    box_minmax(l, axis, &lmn, &lmx);
    box_minmax(r, axis, &rmn, &rmx);
    return (lmx >= rmn) && (rmx >= lmn);
}
/* Paralgram.cpp:17-173 : 15 axes in the reference's fixed order */
static int box_sat(const box_t* l, const box_t* r) {
    if (!axis_overlap(l, r, cross3(l->v, l->w))) return 0;   /* :21 */
    if (!axis_overlap(l, r, cross3(l->u, l->w))) return 0;   /* :31 */
    if (!axis_overlap(l, r, cross3(l->u, l->v))) return 0;   /* :41 */
    if (!axis_overlap(l, r, cross3(r->v, r->w))) return 0;   /* :52 */
    if (!axis_overlap(l, r, cross3(r->u, r->w))) return 0;   /* :62 */
    if (!axis_overlap(l, r, cross3(r->u, r->v))) return 0;   /* :72 */
    if (!axis_overlap(l, r, cross3(l->u, r->u))) return 0;   /* :83 */
    if (!axis_overlap(l, r, cross3(l->u, r->v))) return 0;
    if (!axis_overlap(l, r, cross3(l->u, r->w))) return 0;
    if (!axis_overlap(l, r, cross3(l->v, r->u))) return 0;
    if (!axis_overlap(l, r, cross3(l->v, r->v))) return 0;
    if (!axis_overlap(l, r, cross3(l->v, r->w))) return 0;
    if (!axis_overlap(l, r, cross3(l->w, r->u))) return 0;
    if (!axis_overlap(l, r, cross3(l->w, r->v))) return 0;
    if (!axis_overlap(l, r, cross3(l->w, r->w))) return 0;   /* :163 */
    return 1;
}
/* Paralgram.cpp:203-210 */
static float box_surface(const box_t* b) {
    float uv = length3(cross3(b->u, b->v));
    float uw = length3(cross3(b->u, b->w));
    float vw = length3(cross3(b->v, b->w));
    return 2.f * (uv + uw + vw);
}

/* ------------------------------------------------------------------ */
/* eig3 (public-domain JAMA port), restated: /root/reference/eig3/eig3.cpp */
/* ------------------------------------------------------------------ */
static double hypot2(double x, double y) { return sqrt(x * x + y * y); }   /* eig3.cpp:15-17 */

/* eig3.cpp:21-134 */
static void tred2(double V[3][3], double d[3], double e[3]) {
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
            double hh = f / (h + h);
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
        double h = d[i + 1];
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

/* eig3.cpp:138-254 */
static void tql2(double V[3][3], double d[3], double e[3]) {
    const int n = 3;
    for (int i = 1; i < n; i++) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    double f = 0.0, tst1 = 0.0;
    double eps = 2.220446049250313e-16;   /* pow(2.0,-52.0), exact */
    for (int l = 0; l < n; l++) {
        double t = fabs(d[l]) + fabs(e[l]);
        tst1 = (tst1 > t) ? tst1 : t;       /* MAX(a,b) ((a)>(b)?(a):(b)) */
        int m = l;
        while (m < n) { if (fabs(e[m]) <= eps * tst1) break; m++; }
        if (m > l) {
            do {
                double g = d[l];
                double p = (d[l + 1] - g) / (2.0 * e[l]);
                double r = hypot2(p, 1.0);
                if (p < 0) r = -r;
                d[l] = e[l] / (p + r);
                d[l + 1] = e[l] * (p + r);
                double dl1 = d[l + 1];
                double h = g - d[l];
                for (int i = l + 2; i < n; i++) d[i] -= h;
                f = f + h;
                p = d[m];
                double c = 1.0, c2 = c, c3 = c;
                double el1 = e[l + 1];
                double s = 0.0, s2 = 0.0;
                for (int i = m - 1; i >= l; i--) {
                    c3 = c2; c2 = c; s2 = s;
                    g = c * e[i];
                    h = c * p;
                    r = hypot2(p, e[i]);
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

/* eig3.cpp:256-265 */
void imro_eig3(const double A[9], double Vout[9], double d[3]) {
    double V[3][3], e[3];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[i][j] = A[3 * i + j];
    tred2(V, d, e);
    tql2(V, d, e);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Vout[3 * i + j] = V[i][j];
}

/* ------------------------------------------------------------------ */
/* OBB fit  (IMR/src/Geometry/OBB.cpp)                                 */
/* ------------------------------------------------------------------ */
/* OBB.cpp:168-178 */
static box_t empty_obb(void) {
    box_t b = { { 0.f, 0.f, 0.f }, { FLT_EPSILON, 0.f, 0.f }, { 0.f, FLT_EPSILON, 0.f }, { 0.f, 0.f, FLT_EPSILON } };
    return b;
}

/* OBB.cpp:105-166 ; the three axes are dvec3 */
static box_t aabb_along_axes(const v3* pts, uint64_t n, const double au[3], const double av[3], const double aw[3]) {
    box_t out;
    double center[3] = { 0., 0., 0. };
    const double* axes[3] = { au, av, aw };
    v3* sides[3] = { &out.u, &out.v, &out.w };
    for (int a = 0; a < 3; ++a) {
        const double* ax = axes[a];
        double mn = INFINITY, mx = -INFINITY;
        for (uint64_t i = 0; i < n; ++i) {
            /* glm::dot(dvec3,dvec3): tmp = a*b; tmp.x + tmp.y + tmp.z */
            double tx = ax[0] * (double)pts[i].x, ty = ax[1] * (double)pts[i].y, tz = ax[2] * (double)pts[i].z;
            double proj = tx + ty + tz;
            mn = (proj < mn) ? proj : mn;    /* std::min(this_projection, min) */
            mx = (mx < proj) ? proj : mx;    /* std::max(this_projection, max) */
        }
        double delta = (mx - mn) + 2. * (double)FLT_EPSILON;
        double mid = (mx + mn) / 2.;
        center[0] += mid * ax[0]; center[1] += mid * ax[1]; center[2] += mid * ax[2];
        double half = delta / 2.;
        sides[a]->x = (float)(half * ax[0]); sides[a]->y = (float)(half * ax[1]); sides[a]->z = (float)(half * ax[2]);
    }
    out.c.x = (float)center[0]; out.c.y = (float)center[1]; out.c.z = (float)center[2];
    return out;
}

/* OBB.cpp:33-89 */
static box_t obb_from_points(const v3* pts, uint64_t n) {
    /* unique-point probe (:35-41): stops once more than 3 distinct points were seen */
    v3 uniq[4]; int nu = 0;
    for (uint64_t i = 0; i < n && nu <= 3; ++i) {
        int seen = 0;
        for (int k = 0; k < nu; ++k)
            if (uniq[k].x == pts[i].x && uniq[k].y == pts[i].y && uniq[k].z == pts[i].z) { seen = 1; break; }
        if (!seen) uniq[nu++] = pts[i];
    }
    if (nu == 0) return empty_obb();
    if (nu == 1) { box_t b = empty_obb(); b.c = pts[0]; return b; }

    double sx = 0., sy = 0., sz = 0.;
    for (uint64_t i = 0; i < n; ++i) { sx = sx + (double)pts[i].x; sy = sy + (double)pts[i].y; sz = sz + (double)pts[i].z; }
    double dn = (double)n;
    double mx = sx / dn, my = sy / dn, mz = sz / dn;
    double cxx = 0., cyy = 0., czz = 0., cxy = 0., cxz = 0., cyz = 0.;
    for (uint64_t i = 0; i < n; ++i) cxx = cxx + ((double)pts[i].x - mx) * ((double)pts[i].x - mx);
    for (uint64_t i = 0; i < n; ++i) cyy = cyy + ((double)pts[i].y - my) * ((double)pts[i].y - my);
    for (uint64_t i = 0; i < n; ++i) czz = czz + ((double)pts[i].z - mz) * ((double)pts[i].z - mz);
    for (uint64_t i = 0; i < n; ++i) cxy = cxy + ((double)pts[i].x - mx) * ((double)pts[i].y - my);
    for (uint64_t i = 0; i < n; ++i) cxz = cxz + ((double)pts[i].x - mx) * ((double)pts[i].z - mz);
    for (uint64_t i = 0; i < n; ++i) cyz = cyz + ((double)pts[i].y - my) * ((double)pts[i].z - mz);
    double A[9] = { cxx, cxy, cxz, cxy, cyy, cyz, cxz, cyz, czz };
    for (int i = 0; i < 9; ++i) A[i] = A[i] / dn;     /* cov_mat /= double(points.size()) */
    double V[9], d[3];
    imro_eig3(A, V, d);
    /* OBB.cpp:80-87: eigenvectors[k] is glm column k of the dmat3 whose storage
     * eig3 filled as a C row-major V  ==> axis k = ROW k of V (SURVEY finding 3) */
    return aabb_along_axes(pts, n, &V[0], &V[3], &V[6]);
}

void imro_obb_from_points(const float* pts, uint64_t n, float* out12) {
    box_store(obb_from_points((const v3*)pts, n), out12);
}

/* ------------------------------------------------------------------ */
/* OBB tree  (IMR/src/Geometry/OBBtree.cpp)                            */
/* ------------------------------------------------------------------ */
struct imro_tree {
    uint64_t nv, n_tri;
    float* boxes; int32_t* left; int32_t* right; uint32_t* tri_off; uint32_t* tri_cnt;
    float* tri_pos; float* tri_nrm; uint32_t* tri_vid; uint32_t* tri_orig;
    uint32_t max_vid;
};

typedef struct bnode {
    box_t box; int leaf;
    struct bnode* l; struct bnode* r;
    uint32_t* tris; uint64_t n;     /* input-triangle indices, order preserved */
} bnode;

typedef struct { const float* pos; } build_ctx;

static v3 tri_p(const build_ctx* cx, uint32_t t, int k) { const float* p = cx->pos + 9 * (uint64_t)t + 3 * k; v3 r = { p[0], p[1], p[2] }; return r; }

/* Triangle.cpp:113-139 */
static void tri_minmax(const build_ctx* cx, uint32_t t, v3 axis, float* mn, float* mx) {
    float lo = INFINITY, hi = -INFINITY;
    for (int k = 0; k < 3; ++k) {
        float pr = dot3(axis, tri_p(cx, t, k));
        if (pr < lo) lo = pr;
        if (pr > hi) hi = pr;
    }
    *mn = lo; *mx = hi;
}

static bnode* build_node(const build_ctx* cx, uint32_t* tris, uint64_t n);

/* OBBtree.cpp:43-108 */
static void split_node(const build_ctx* cx, bnode* nd) {
    float len[3]; v3 ax[3];
    v3 side[3] = { nd->box.u, nd->box.v, nd->box.w };
    for (int k = 0; k < 3; ++k) { len[k] = length3(side[k]); ax[k] = normalize3(side[k]); }
    /* std::sort of 3 elements with a.first > b.first == insertion sort, ties keep order (:50-51) */
    int ord[3] = { 0, 1, 2 };
    for (int i = 1; i < 3; ++i) {
        int v = ord[i]; int j = i;
        while (j > 0 && len[v] > len[ord[j - 1]]) { ord[j] = ord[j - 1]; --j; }
        ord[j] = v;
    }
    uint64_t n = nd->n;
    uint32_t* L = (uint32_t*)malloc(sizeof(uint32_t) * n);
    uint32_t* R = (uint32_t*)malloc(sizeof(uint32_t) * n);
    uint64_t nl = 0, nr = 0;
    int chosen = 0;
    do {
        nl = nr = 0;
        v3 axis = ax[ord[chosen]];
        float cproj = dot3(nd->box.c, axis);                       /* GetCenterProjectionToAxis, Paralgram.cpp:192-196 */
        for (uint64_t i = 0; i < n; ++i) {
            float mn, mx;
            tri_minmax(cx, nd->tris[i], axis, &mn, &mx);
            float mean = (mn + mx) / 2.f;
            if (mean <= cproj) L[nl++] = nd->tris[i]; else R[nr++] = nd->tris[i];
        }
        chosen++;
    } while ((nl == 0 || nr == 0) && chosen < 3);
    if (nl == 0 || nr == 0) {                                       /* :83-95 */
        nl = nr = 0;
        for (uint64_t i = 0; i < n; ++i) { if (i < n / 2) L[nl++] = nd->tris[i]; else R[nr++] = nd->tris[i]; }
    }
    free(nd->tris); nd->tris = NULL; nd->n = 0;
    nd->leaf = 0;
    L = (uint32_t*)realloc(L, sizeof(uint32_t) * (nl ? nl : 1));
    R = (uint32_t*)realloc(R, sizeof(uint32_t) * (nr ? nr : 1));
    nd->l = build_node(cx, L, nl);
    nd->r = build_node(cx, R, nr);
    if (box_surface(&nd->r->box) > box_surface(&nd->l->box)) { bnode* t = nd->l; nd->l = nd->r; nd->r = t; }   /* :102-107 */
}

/* OBBtree.cpp:8-18 + OBB.cpp:91-103 */
static bnode* build_node(const build_ctx* cx, uint32_t* tris, uint64_t n) {
    bnode* nd = (bnode*)calloc(1, sizeof(bnode));
    nd->leaf = 1; nd->tris = tris; nd->n = n;
    v3* pts = (v3*)malloc(sizeof(v3) * (n ? 3 * n : 1));
    for (uint64_t i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) pts[3 * i + k] = tri_p(cx, tris[i], k);
    nd->box = obb_from_points(pts, 3 * n);
    free(pts);
    if (n > 4) split_node(cx, nd);       /* maxNumberOfTriangles = 4, OBBtree.h:49 */
    return nd;
}

static uint64_t count_nodes(const bnode* nd) { return nd->leaf ? 1 : 1 + count_nodes(nd->l) + count_nodes(nd->r); }

typedef struct { imro_tree* t; uint64_t nv; uint64_t nt; const float* pos; const float* nrm; const uint32_t* vid; } flat_ctx;

static v3 face_normal(const float* p) {    /* Triangle.cpp:141-147 */
    v3 p0 = { p[0], p[1], p[2] }, p1 = { p[3], p[4], p[5] }, p2 = { p[6], p[7], p[8] };
    return normalize3(cross3(sub3(p1, p0), sub3(p2, p0)));
}

/* pre-order flatten; leaf triangles appended in DFS order (OBBtree.cpp:179-233,385-394) */
static int32_t flatten_node(flat_ctx* f, bnode* nd) {
    imro_tree* t = f->t;
    uint64_t me = f->nv++;
    box_store(nd->box, t->boxes + 12 * me);
    if (nd->leaf) {
        t->left[me] = t->right[me] = -1;
        t->tri_off[me] = (uint32_t)f->nt; t->tri_cnt[me] = (uint32_t)nd->n;
        for (uint64_t i = 0; i < nd->n; ++i) {
            uint32_t src = nd->tris[i]; uint64_t dst = f->nt++;
            memcpy(t->tri_pos + 9 * dst, f->pos + 9 * (uint64_t)src, 36);
            if (f->nrm) memcpy(t->tri_nrm + 9 * dst, f->nrm + 9 * (uint64_t)src, 36);
            else { v3 fn = face_normal(f->pos + 9 * (uint64_t)src); for (int k = 0; k < 3; ++k) { t->tri_nrm[9 * dst + 3 * k] = fn.x; t->tri_nrm[9 * dst + 3 * k + 1] = fn.y; t->tri_nrm[9 * dst + 3 * k + 2] = fn.z; } }
            for (int k = 0; k < 3; ++k) t->tri_vid[3 * dst + k] = f->vid ? f->vid[3 * (uint64_t)src + k] : (uint32_t)(3 * src + k);
            t->tri_orig[dst] = src;
        }
    } else {
        t->tri_off[me] = 0; t->tri_cnt[me] = 0;
        int32_t l = flatten_node(f, nd->l);
        int32_t r = flatten_node(f, nd->r);
        t->left[me] = l; t->right[me] = r;
    }
    return (int32_t)me;
}

static void free_nodes(bnode* nd) { if (!nd) return; free_nodes(nd->l); free_nodes(nd->r); free(nd->tris); free(nd); }

static imro_tree* tree_alloc(uint64_t nv, uint64_t n_tri) {
    imro_tree* t = (imro_tree*)calloc(1, sizeof(imro_tree));
    t->nv = nv; t->n_tri = n_tri;
    uint64_t a = nv ? nv : 1, b = n_tri ? n_tri : 1;
    t->boxes = (float*)malloc(sizeof(float) * 12 * a);
    t->left = (int32_t*)malloc(sizeof(int32_t) * a); t->right = (int32_t*)malloc(sizeof(int32_t) * a);
    t->tri_off = (uint32_t*)malloc(sizeof(uint32_t) * a); t->tri_cnt = (uint32_t*)malloc(sizeof(uint32_t) * a);
    t->tri_pos = (float*)malloc(sizeof(float) * 9 * b); t->tri_nrm = (float*)malloc(sizeof(float) * 9 * b);
    t->tri_vid = (uint32_t*)malloc(sizeof(uint32_t) * 3 * b); t->tri_orig = (uint32_t*)malloc(sizeof(uint32_t) * b);
    return t;
}
static void tree_finish(imro_tree* t) {
    uint32_t mv = 0;
    for (uint64_t i = 0; i < 3 * t->n_tri; ++i) if (t->tri_vid[i] > mv) mv = t->tri_vid[i];
    t->max_vid = mv;
}

/* OBBtree.cpp:321-358 */
imro_tree* imro_tree_build(const float* positions, const float* normals, const uint32_t* vertex_ids, uint64_t n_tri) {
    build_ctx cx = { positions };
    uint32_t* all = (uint32_t*)malloc(sizeof(uint32_t) * (n_tri ? n_tri : 1));
    for (uint64_t i = 0; i < n_tri; ++i) all[i] = (uint32_t)i;
    bnode* root = build_node(&cx, all, n_tri);
    uint64_t nv = count_nodes(root);
    imro_tree* t = tree_alloc(nv, n_tri);
    flat_ctx f = { t, 0, 0, positions, normals, vertex_ids };
    flatten_node(&f, root);
    free_nodes(root);
    tree_finish(t);
    return t;
}

imro_tree* imro_tree_import(uint64_t nv, const float* boxes, const int32_t* left, const int32_t* right,
                            const uint32_t* tri_off, const uint32_t* tri_cnt, uint64_t n_tri,
                            const float* tri_pos, const float* tri_nrm, const uint32_t* tri_vid, const uint32_t* tri_orig) {
    imro_tree* t = tree_alloc(nv, n_tri);
    memcpy(t->boxes, boxes, sizeof(float) * 12 * nv);
    memcpy(t->left, left, sizeof(int32_t) * nv); memcpy(t->right, right, sizeof(int32_t) * nv);
    memcpy(t->tri_off, tri_off, sizeof(uint32_t) * nv); memcpy(t->tri_cnt, tri_cnt, sizeof(uint32_t) * nv);
    memcpy(t->tri_pos, tri_pos, sizeof(float) * 9 * n_tri);
    if (tri_nrm) memcpy(t->tri_nrm, tri_nrm, sizeof(float) * 9 * n_tri); else memset(t->tri_nrm, 0, sizeof(float) * 9 * n_tri);
    for (uint64_t i = 0; i < 3 * n_tri; ++i) t->tri_vid[i] = tri_vid ? tri_vid[i] : (uint32_t)i;
    for (uint64_t i = 0; i < n_tri; ++i) t->tri_orig[i] = tri_orig ? tri_orig[i] : (uint32_t)i;
    tree_finish(t);
    return t;
}

void imro_tree_free(imro_tree* t) {
    if (!t) return;
    free(t->boxes); free(t->left); free(t->right); free(t->tri_off); free(t->tri_cnt);
    free(t->tri_pos); free(t->tri_nrm); free(t->tri_vid); free(t->tri_orig); free(t);
}
uint64_t imro_tree_vertex_count(const imro_tree* t) { return t->nv; }
uint64_t imro_tree_tri_count(const imro_tree* t) { return t->n_tri; }
void imro_tree_flatten(const imro_tree* t, float* boxes, int32_t* left, int32_t* right, uint32_t* tri_off, uint32_t* tri_cnt,
                       float* tri_pos, float* tri_nrm, uint32_t* tri_vid, uint32_t* tri_orig) {
    if (boxes) memcpy(boxes, t->boxes, sizeof(float) * 12 * t->nv);
    if (left) memcpy(left, t->left, sizeof(int32_t) * t->nv);
    if (right) memcpy(right, t->right, sizeof(int32_t) * t->nv);
    if (tri_off) memcpy(tri_off, t->tri_off, sizeof(uint32_t) * t->nv);
    if (tri_cnt) memcpy(tri_cnt, t->tri_cnt, sizeof(uint32_t) * t->nv);
    if (tri_pos) memcpy(tri_pos, t->tri_pos, sizeof(float) * 9 * t->n_tri);
    if (tri_nrm) memcpy(tri_nrm, t->tri_nrm, sizeof(float) * 9 * t->n_tri);
    if (tri_vid) memcpy(tri_vid, t->tri_vid, sizeof(uint32_t) * 3 * t->n_tri);
    if (tri_orig) memcpy(tri_orig, t->tri_orig, sizeof(uint32_t) * t->n_tri);
}

/* ------------------------------------------------------------------ */
/* single predicates                                                    */
/* ------------------------------------------------------------------ */
int imro_sat(const float* lhs12, const float* rhs12, const float* m16) {
    box_t a = box_load(lhs12), b = box_load(rhs12);
    if (m16) b = box_transform(m16, b);
    return box_sat(&a, &b);
}
float imro_surface(const float* box12, const float* m16) {
    box_t a = box_load(box12);
    if (m16) a = box_transform(m16, a);
    return box_surface(&a);
}
void imro_box_transform(const float* box12, const float* m16, float* out12) { box_store(box_transform(m16, box_load(box12)), out12); }
void imro_pair_matrix(const float* a16, const float* b16, float* out16) {
    float inv[16];
    mat4_inverse(a16, inv);
    mat4_mul(inv, b16, out16);
}
/* IMR/src/CollisionDetection/CollisionDetection.cpp:9-13 */
static void sweep_axes(v3* U, v3* V, v3* W) {
    v3 a = { 0.8f, -0.2f, 0.f }, down = { 0.f, -1.f, 0.f };
    *U = normalize3(a);
    *W = normalize3(cross3(*U, down));
    *V = normalize3(cross3(*W, *U));
}
void imro_sweep_axes(float* out9) {
    v3 U, V, W; sweep_axes(&U, &V, &W);
    out9[0] = U.x; out9[1] = U.y; out9[2] = U.z; out9[3] = V.x; out9[4] = V.y; out9[5] = V.z; out9[6] = W.x; out9[7] = W.y; out9[8] = W.z;
}

/* ------------------------------------------------------------------ */
/* Moller tri-tri with intersection line (IMR/src/Geometry/Triangle.cpp) */
/* ------------------------------------------------------------------ */
#define TT_EPSILON 0.000001      /* a DOUBLE literal: Triangle.cpp:332 */

/* Triangle.cpp:402-419 EDGE_EDGE_TEST ; returns 1 on hit */
static int edge_edge(const float* V0, const float* U0, const float* U1, int i0, int i1, float Ax, float Ay) {
    float Bx = U0[i0] - U1[i0];
    float By = U0[i1] - U1[i1];
    float Cx = V0[i0] - U0[i0];
    float Cy = V0[i1] - U0[i1];
    float f = Ay * Bx - Ax * By;
    float d = By * Cx - Bx * Cy;
    if ((f > 0 && d >= 0 && d <= f) || (f < 0 && d <= 0 && d >= f)) {
        float e = Ax * Cy - Ay * Cx;
        if (f > 0) { if (e >= 0 && e <= f) return 1; }
        else { if (e <= 0 && e >= f) return 1; }
    }
    return 0;
}
/* Triangle.cpp:421-433 */
static int edge_against_tri(const float* V0, const float* V1, const float* U0, const float* U1, const float* U2, int i0, int i1) {
    float Ax = V1[i0] - V0[i0];
    float Ay = V1[i1] - V0[i1];
    if (edge_edge(V0, U0, U1, i0, i1, Ax, Ay)) return 1;
    if (edge_edge(V0, U1, U2, i0, i1, Ax, Ay)) return 1;
    if (edge_edge(V0, U2, U0, i0, i1, Ax, Ay)) return 1;
    return 0;
}
/* Triangle.cpp:435-458 */
static int point_in_tri(const float* V0, const float* U0, const float* U1, const float* U2, int i0, int i1) {
    float a, b, c, d0, d1, d2;
    a = U1[i1] - U0[i1]; b = -(U1[i0] - U0[i0]); c = -a * U0[i0] - b * U0[i1]; d0 = a * V0[i0] + b * V0[i1] + c;
    a = U2[i1] - U1[i1]; b = -(U2[i0] - U1[i0]); c = -a * U1[i0] - b * U1[i1]; d1 = a * V0[i0] + b * V0[i1] + c;
    a = U0[i1] - U2[i1]; b = -(U0[i0] - U2[i0]); c = -a * U2[i0] - b * U2[i1]; d2 = a * V0[i0] + b * V0[i1] + c;
    if (d0 * d1 > 0.0) { if (d0 * d2 > 0.0) return 1; }
    return 0;
}
/* Triangle.cpp:460-507 */
static int coplanar_tri_tri(const float* N, const float* V0, const float* V1, const float* V2,
                            const float* U0, const float* U1, const float* U2) {
    float A[3]; int i0, i1;
    A[0] = fabsf(N[0]); A[1] = fabsf(N[1]); A[2] = fabsf(N[2]);
    if (A[0] > A[1]) { if (A[0] > A[2]) { i0 = 1; i1 = 2; } else { i0 = 0; i1 = 1; } }
    else { if (A[2] > A[1]) { i0 = 0; i1 = 1; } else { i0 = 0; i1 = 2; } }
    if (edge_against_tri(V0, V1, U0, U1, U2, i0, i1)) return 1;
    if (edge_against_tri(V1, V2, U0, U1, U2, i0, i1)) return 1;
    if (edge_against_tri(V2, V0, U0, U1, U2, i0, i1)) return 1;
    if (point_in_tri(V0, U0, U1, U2, i0, i1)) return 1;
    if (point_in_tri(U0, V0, V1, V2, i0, i1)) return 1;
    return 0;
}
/* Triangle.cpp:764-778 */
static void isect2(const float* VTX0, const float* VTX1, const float* VTX2, float VV0, float VV1, float VV2,
                   float D0, float D1, float D2, float* isect0, float* isect1, float* ip0, float* ip1) {
    float tmp = D0 / (D0 - D1);
    float diff[3];
    *isect0 = VV0 + (VV1 - VV0) * tmp;
    diff[0] = VTX1[0] - VTX0[0]; diff[1] = VTX1[1] - VTX0[1]; diff[2] = VTX1[2] - VTX0[2];
    diff[0] = tmp * diff[0]; diff[1] = tmp * diff[1]; diff[2] = tmp * diff[2];
    ip0[0] = diff[0] + VTX0[0]; ip0[1] = diff[1] + VTX0[1]; ip0[2] = diff[2] + VTX0[2];
    tmp = D0 / (D0 - D2);
    *isect1 = VV0 + (VV2 - VV0) * tmp;
    diff[0] = VTX2[0] - VTX0[0]; diff[1] = VTX2[1] - VTX0[1]; diff[2] = VTX2[2] - VTX0[2];
    diff[0] = tmp * diff[0]; diff[1] = tmp * diff[1]; diff[2] = tmp * diff[2];
    ip1[0] = VTX0[0] + diff[0]; ip1[1] = VTX0[1] + diff[1]; ip1[2] = VTX0[2] + diff[2];
}
/* Triangle.cpp:795-830 */
static int compute_intervals_isectline(const float* VERT0, const float* VERT1, const float* VERT2,
                                       float VV0, float VV1, float VV2, float D0, float D1, float D2,
                                       float D0D1, float D0D2, float* isect0, float* isect1, float* ip0, float* ip1) {
    if (D0D1 > 0.0f)                     isect2(VERT2, VERT0, VERT1, VV2, VV0, VV1, D2, D0, D1, isect0, isect1, ip0, ip1);
    else if (D0D2 > 0.0f)                isect2(VERT1, VERT0, VERT2, VV1, VV0, VV2, D1, D0, D2, isect0, isect1, ip0, ip1);
    else if (D1 * D2 > 0.0f || D0 != 0.0f) isect2(VERT0, VERT1, VERT2, VV0, VV1, VV2, D0, D1, D2, isect0, isect1, ip0, ip1);
    else if (D1 != 0.0f)                 isect2(VERT1, VERT0, VERT2, VV1, VV0, VV2, D1, D0, D2, isect0, isect1, ip0, ip1);
    else if (D2 != 0.0f)                 isect2(VERT2, VERT0, VERT1, VV2, VV0, VV1, D2, D0, D1, isect0, isect1, ip0, ip1);
    else return 1;
    return 0;
}
#define TT_DOT(a, b) ((a)[0] * (b)[0] + (a)[1] * (b)[1] + (a)[2] * (b)[2])
#define TT_CROSS(d, a, b) do { (d)[0] = (a)[1] * (b)[2] - (a)[2] * (b)[1]; (d)[1] = (a)[2] * (b)[0] - (a)[0] * (b)[2]; (d)[2] = (a)[0] * (b)[1] - (a)[1] * (b)[0]; } while (0)
#define TT_SUB(d, a, b) do { (d)[0] = (a)[0] - (b)[0]; (d)[1] = (a)[1] - (b)[1]; (d)[2] = (a)[2] - (b)[2]; } while (0)
#define TT_SET(d, s) do { (d)[0] = (s)[0]; (d)[1] = (s)[1]; (d)[2] = (s)[2]; } while (0)

/* Triangle.cpp:866-1002 */
static int tri_tri_isectline(const float* V0, const float* V1, const float* V2,
                             const float* U0, const float* U1, const float* U2,
                             int* coplanar, float* isectpt1, float* isectpt2) {
    float E1[3], E2[3], N1[3], N2[3], d1, d2;
    float du0, du1, du2, dv0, dv1, dv2, D[3];
    /* isect2v / ipB*: the reference leaves these uninitialised when the second compute_intervals_isectline call
     * reports coplanar (Triangle.cpp:959-960, return value ignored) -- undefined behaviour there; zeros here and
     * in the CUDA path so that both are deterministic. */
    float isect1[2], isect2v[2] = { 0.f, 0.f };
    float ipA1[3], ipA2[3], ipB1[3] = { 0.f, 0.f, 0.f }, ipB2[3] = { 0.f, 0.f, 0.f };
    float du0du1, du0du2, dv0dv1, dv0dv2;
    int index; float vp0, vp1, vp2, up0, up1, up2, b, c, max;
    int smallest1, smallest2;

    TT_SUB(E1, V1, V0); TT_SUB(E2, V2, V0); TT_CROSS(N1, E1, E2);
    d1 = -TT_DOT(N1, V0);
    du0 = TT_DOT(N1, U0) + d1; du1 = TT_DOT(N1, U1) + d1; du2 = TT_DOT(N1, U2) + d1;
    if ((double)fabsf(du0) < TT_EPSILON) du0 = 0.0f;
    if ((double)fabsf(du1) < TT_EPSILON) du1 = 0.0f;
    if ((double)fabsf(du2) < TT_EPSILON) du2 = 0.0f;
    du0du1 = du0 * du1; du0du2 = du0 * du2;
    if (du0du1 > 0.0f && du0du2 > 0.0f) return 0;

    TT_SUB(E1, U1, U0); TT_SUB(E2, U2, U0); TT_CROSS(N2, E1, E2);
    d2 = -TT_DOT(N2, U0);
    dv0 = TT_DOT(N2, V0) + d2; dv1 = TT_DOT(N2, V1) + d2; dv2 = TT_DOT(N2, V2) + d2;
    if ((double)fabsf(dv0) < TT_EPSILON) dv0 = 0.0f;
    if ((double)fabsf(dv1) < TT_EPSILON) dv1 = 0.0f;
    if ((double)fabsf(dv2) < TT_EPSILON) dv2 = 0.0f;
    dv0dv1 = dv0 * dv1; dv0dv2 = dv0 * dv2;
    if (dv0dv1 > 0.0f && dv0dv2 > 0.0f) return 0;

    TT_CROSS(D, N1, N2);
    max = fabsf(D[0]); index = 0; b = fabsf(D[1]); c = fabsf(D[2]);
    if (b > max) { max = b; index = 1; }
    if (c > max) { max = c; index = 2; }
    vp0 = V0[index]; vp1 = V1[index]; vp2 = V2[index];
    up0 = U0[index]; up1 = U1[index]; up2 = U2[index];

    *coplanar = compute_intervals_isectline(V0, V1, V2, vp0, vp1, vp2, dv0, dv1, dv2, dv0dv1, dv0dv2, &isect1[0], &isect1[1], ipA1, ipA2);
    if (*coplanar) return coplanar_tri_tri(N1, V0, V1, V2, U0, U1, U2);
    compute_intervals_isectline(U0, U1, U2, up0, up1, up2, du0, du1, du2, du0du1, du0du2, &isect2v[0], &isect2v[1], ipB1, ipB2);

    /* SORT2, Triangle.cpp:751-761 */
    if (isect1[0] > isect1[1]) { float t = isect1[0]; isect1[0] = isect1[1]; isect1[1] = t; smallest1 = 1; } else smallest1 = 0;
    if (isect2v[0] > isect2v[1]) { float t = isect2v[0]; isect2v[0] = isect2v[1]; isect2v[1] = t; smallest2 = 1; } else smallest2 = 0;
    if (isect1[1] < isect2v[0] || isect2v[1] < isect1[0]) return 0;

    if (isect2v[0] < isect1[0]) {
        if (smallest1 == 0) TT_SET(isectpt1, ipA1); else TT_SET(isectpt1, ipA2);
        if (isect2v[1] < isect1[1]) { if (smallest2 == 0) TT_SET(isectpt2, ipB2); else TT_SET(isectpt2, ipB1); }
        else { if (smallest1 == 0) TT_SET(isectpt2, ipA2); else TT_SET(isectpt2, ipA1); }
    } else {
        if (smallest2 == 0) TT_SET(isectpt1, ipB1); else TT_SET(isectpt1, ipB2);
        if (isect2v[1] > isect1[1]) { if (smallest1 == 0) TT_SET(isectpt2, ipA2); else TT_SET(isectpt2, ipA1); }
        else { if (smallest2 == 0) TT_SET(isectpt2, ipB2); else TT_SET(isectpt2, ipB1); }
    }
    return 1;
}

/* Triangle.cpp:69-78 */
static void tri_transform(const float* m, const float* in9, float* out9) {
    for (int k = 0; k < 3; ++k) {
        v3 p = { in9[3 * k], in9[3 * k + 1], in9[3 * k + 2] };
        v3 q = mat4_mul_point(m, p, 1.f);
        out9[3 * k] = q.x; out9[3 * k + 1] = q.y; out9[3 * k + 2] = q.z;
    }
}

void imro_tri_tri(const float* a, const float* b, const float* m16, uint64_t n, uint8_t* flags, float* seg) {
    for (uint64_t i = 0; i < n; ++i) {
        float tb[9];
        if (m16) tri_transform(m16, b + 9 * i, tb); else memcpy(tb, b + 9 * i, 36);
        const float* ta = a + 9 * i;
        int cop = 0; float s[3], t[3];
        int hit = tri_tri_isectline(ta, ta + 3, ta + 6, tb, tb + 3, tb + 6, &cop, s, t);
        flags[i] = (uint8_t)((hit ? 1 : 0) | (cop ? 2 : 0));
        if (hit && !cop) { memcpy(seg + 6 * i, s, 12); memcpy(seg + 6 * i + 3, t, 12); }
    }
}

/* ------------------------------------------------------------------ */
/* broad phase (IMR/src/CollisionDetection/SweepAndPrune.cpp:15-88)     */
/* ------------------------------------------------------------------ */
void imro_extents(const float* mats, const float* root_boxes, uint64_t n, float* out) {
    v3 ax[3]; sweep_axes(&ax[0], &ax[1], &ax[2]);
    for (uint64_t i = 0; i < n; ++i) {
        box_t p = box_transform(mats + 16 * i, box_load(root_boxes + 12 * i));     /* :23 */
        for (int a = 0; a < 3; ++a) box_minmax(&p, ax[a], &out[6 * i + 2 * a], &out[6 * i + 2 * a + 1]);
    }
}

typedef struct { float mn, mx; uint32_t idx; } sap_entry;
static int sap_cmp(const void* a, const void* b) {
    const sap_entry* x = (const sap_entry*)a; const sap_entry* y = (const sap_entry*)b;
    if (x->mn < y->mn) return -1;
    if (x->mn > y->mn) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);     /* reference: std::sort, tie order unspecified; we fix it by index */
}

/* The reference keeps a pair iff it is found overlapping by the active-list
 * sweep on all three axes; the sweep on one axis reports (a,e), a earlier in
 * min-order, iff NOT(a.max < e.min) (:58).  Orientation = U-axis order (:63).
 * The pair store here is a sorted array instead of the unordered_map. */
uint64_t imro_broad(const float* mats, const float* root_boxes, const uint8_t* should_cb, uint64_t n,
                    uint32_t* pairs, uint64_t cap) {
    float* ext = (float*)malloc(sizeof(float) * 6 * (n ? n : 1));
    imro_extents(mats, root_boxes, n, ext);
    sap_entry* U = (sap_entry*)malloc(sizeof(sap_entry) * (n ? n : 1));
    for (uint64_t i = 0; i < n; ++i) { U[i].mn = ext[6 * i]; U[i].mx = ext[6 * i + 1]; U[i].idx = (uint32_t)i; }
    qsort(U, n, sizeof(sap_entry), sap_cmp);
    uint64_t count = 0;
    /* active list sweep on U, V/W overlap evaluated in closed form (equivalent, see header comment) */
    uint32_t* active = (uint32_t*)malloc(sizeof(uint32_t) * (n ? n : 1));
    uint64_t na = 0;
    for (uint64_t k = 0; k < n; ++k) {
        const sap_entry e = U[k];
        uint64_t keep = 0;
        for (uint64_t j = 0; j < na; ++j) {
            const sap_entry a = U[active[j]];
            if (a.mx < e.mn) continue;                         /* expired (:58) */
            active[keep++] = active[j];
            if (should_cb[a.idx] || should_cb[e.idx]) {        /* :60 */
                const float* A = ext + 6 * (uint64_t)a.idx; const float* E = ext + 6 * (uint64_t)e.idx;
                int v_ok = (A[2] <= E[2]) ? !(A[3] < E[2]) : !(E[3] < A[2]);
                int w_ok = (A[4] <= E[4]) ? !(A[5] < E[4]) : !(E[5] < A[4]);
                if (v_ok && w_ok) {
                    if (count < cap) { pairs[2 * count] = a.idx; pairs[2 * count + 1] = e.idx; }
                    ++count;
                }
            }
        }
        na = keep;
        active[na++] = (uint32_t)k;
    }
    free(active); free(U); free(ext);
    return count;
}

/* ------------------------------------------------------------------ */
/* mid phase (IMR/src/Geometry/OBBtree.cpp:396-477)                     */
/* ------------------------------------------------------------------ */
typedef struct {
    const imro_tree* a; const imro_tree* b; const float* rel;
    uint32_t* combos; uint64_t cap; uint64_t n;
    uint64_t visits, passes, tri_tests; uint32_t max_depth;
} mid_ctx;

static void mid_walk(mid_ctx* c, int32_t va, int32_t vb, uint32_t depth) {
    c->visits++;
    if (depth > c->max_depth) c->max_depth = depth;
    box_t first = box_load(c->a->boxes + 12 * (uint64_t)va);
    box_t second = box_transform(c->rel, box_load(c->b->boxes + 12 * (uint64_t)vb));      /* :420 */
    if (!box_sat(&first, &second)) return;                                                /* :422 */
    c->passes++;
    int la = c->a->left[va] < 0, lb = c->b->left[vb] < 0;
    if (!la && !lb) {
        if (box_surface(&first) >= box_surface(&second)) {                                /* :426 */
            mid_walk(c, c->a->left[va], vb, depth + 1); mid_walk(c, c->a->right[va], vb, depth + 1);
        } else {
            mid_walk(c, va, c->b->left[vb], depth + 1); mid_walk(c, va, c->b->right[vb], depth + 1);
        }
    } else if (la && !lb) {
        mid_walk(c, va, c->b->left[vb], depth + 1); mid_walk(c, va, c->b->right[vb], depth + 1);
    } else if (!la && lb) {
        mid_walk(c, c->a->left[va], vb, depth + 1); mid_walk(c, c->a->right[va], vb, depth + 1);
    } else {
        if (c->n < c->cap) {
            uint32_t* o = c->combos + 4 * c->n;
            o[0] = c->a->tri_off[va]; o[1] = c->a->tri_cnt[va]; o[2] = c->b->tri_off[vb]; o[3] = c->b->tri_cnt[vb];
        }
        c->n++;
        c->tri_tests += (uint64_t)c->a->tri_cnt[va] * (uint64_t)c->b->tri_cnt[vb];
    }
}

uint64_t imro_mid(const imro_tree* a, const float* mat_a, const imro_tree* b, const float* mat_b,
                  uint32_t* combos, uint64_t cap, uint64_t* stats5) {
    float rel[16];
    imro_pair_matrix(mat_a, mat_b, rel);                                                  /* OBBtreesCollision.cpp:15 */
    mid_ctx c = { a, b, rel, combos, cap, 0, 0, 0, 0, 0 };
    mid_walk(&c, 0, 0, 0);
    if (stats5) { stats5[0] = c.visits; stats5[1] = c.passes; stats5[2] = c.n; stats5[3] = c.tri_tests; stats5[4] = c.max_depth; }
    return c.n;
}

/* ------------------------------------------------------------------ */
/* narrow phase (IMR/src/CollisionDetection/CreateUncollideRays.cpp)    */
/* ------------------------------------------------------------------ */
typedef struct { uint32_t bits; v3 wavg; float weight; } cand_t;     /* TriangleCandidateRays :13-58 */

/* Plane.cpp:5-21 : Plane(normal, d) re-normalises */
typedef struct { v3 n; float d; } plane_t;
static plane_t plane_from_tri(const float* t9) {
    v3 p0 = { t9[0], t9[1], t9[2] }, p1 = { t9[3], t9[4], t9[5] }, p2 = { t9[6], t9[7], t9[8] };
    v3 nrm = normalize3(cross3(sub3(p1, p0), sub3(p2, p0)));          /* GetTriangleFaceNormal */
    float d = -dot3(p0, nrm);
    float len = length3(nrm);
    plane_t pl; pl.n.x = nrm.x / len; pl.n.y = nrm.y / len; pl.n.z = nrm.z / len; pl.d = d / len;
    return pl;
}
/* Plane.cpp:23-29 : OUTSIDE iff s > 0 */
static int plane_point_outside(const plane_t* pl, v3 p) { float s = dot3(p, pl->n) + pl->d; return s > 0; }

/* glm/gtc/matrix_inverse.inl:133-147 on mat3(m) ; result column-major 3x3 */
static void adjoint_transpose3(const float* m4, float* o) {
#define A(c, r) m4[4 * (c) + (r)]
    o[0] = +(A(1,1) * A(2,2) - A(2,1) * A(1,2));
    o[1] = -(A(1,0) * A(2,2) - A(2,0) * A(1,2));
    o[2] = +(A(1,0) * A(2,1) - A(2,0) * A(1,1));
    o[3] = -(A(0,1) * A(2,2) - A(2,1) * A(0,2));
    o[4] = +(A(0,0) * A(2,2) - A(2,0) * A(0,2));
    o[5] = -(A(0,0) * A(2,1) - A(2,0) * A(0,1));
    o[6] = +(A(0,1) * A(1,2) - A(1,1) * A(0,2));
    o[7] = -(A(0,0) * A(1,2) - A(1,0) * A(0,2));
    o[8] = +(A(0,0) * A(1,1) - A(1,0) * A(0,1));
#undef A
}
/* type_mat3x3.inl:468-474 */
static v3 mat3_mul(const float* m, v3 v) {
    v3 r;
    r.x = m[0] * v.x + m[3] * v.y + m[6] * v.z;
    r.y = m[1] * v.x + m[4] * v.y + m[7] * v.z;
    r.z = m[2] * v.x + m[5] * v.y + m[8] * v.z;
    return r;
}
/* Triangle.cpp:149-165 */
static void barycentric(const float* t9, v3 p, float* bx, float* by) {
    v3 p0 = { t9[0], t9[1], t9[2] }, p1 = { t9[3], t9[4], t9[5] }, p2 = { t9[6], t9[7], t9[8] };
    v3 v0 = sub3(p1, p0), v1 = sub3(p2, p0), v2 = sub3(p, p0);
    float d00 = dot3(v0, v0), d01 = dot3(v0, v1), d11 = dot3(v1, v1), d20 = dot3(v2, v0), d21 = dot3(v2, v1);
    float denom = d00 * d11 - d01 * d01;
    *bx = (d11 * d20 - d01 * d21) / denom;
    *by = (d00 * d21 - d01 * d20) / denom;
}

typedef struct { cand_t* map; uint8_t* present; uint32_t* order; uint64_t n_order; } cand_map;

static void cand_merge(cand_map* m, uint32_t idx, const cand_t* c) {      /* MergeWithMap :39-49 */
    if (!m->present[idx]) { m->present[idx] = 1; m->map[idx] = *c; m->order[m->n_order++] = idx; }
    else {
        m->map[idx].bits &= c->bits;
        m->map[idx].wavg = add3(m->map[idx].wavg, c->wavg);
        m->map[idx].weight += c->weight;
    }
}

/* find_rays_lambda, CreateUncollideRays.cpp:131-178 */
static uint64_t find_rays(const cand_map* cm, const imro_tree* t, int mul, const float* rel, const float* nmat,
                          float* rays, uint64_t ray_cap, v3* origin_sum) {
    uint64_t n_rays = 0;
    uint8_t* emplaced = (uint8_t*)calloc((uint64_t)t->max_vid + 1, 1);
    v3 sum = { 0.f, 0.f, 0.f };
    for (uint64_t k = 0; k < cm->n_order; ++k) {
        uint32_t ti = cm->order[k];
        const cand_t* c = &cm->map[ti];
        float tp[9];
        if (mul) tri_transform(rel, t->tri_pos + 9 * (uint64_t)ti, tp); else memcpy(tp, t->tri_pos + 9 * (uint64_t)ti, 36);
        const float* tn = t->tri_nrm + 9 * (uint64_t)ti;
        if (c->bits != 0) {
            for (int pi = 0; pi < 3; ++pi) {
                if (!((c->bits >> pi) & 1)) continue;
                uint32_t vid = t->tri_vid[3 * (uint64_t)ti + pi];
                if (emplaced[vid]) continue;
                emplaced[vid] = 1;
                v3 pos = { tp[3 * pi], tp[3 * pi + 1], tp[3 * pi + 2] };
                v3 nr = { tn[3 * pi], tn[3 * pi + 1], tn[3 * pi + 2] };
                v3 nn = mul ? normalize3(mat3_mul(nmat, nr)) : normalize3(nr);       /* Triangle.cpp:177-193 */
                if (n_rays < ray_cap) { float* o = rays + 6 * n_rays; o[0] = pos.x; o[1] = pos.y; o[2] = pos.z; o[3] = -nn.x; o[4] = -nn.y; o[5] = -nn.z; }
                n_rays++;
                sum = add3(sum, pos);
            }
        } else {
            v3 pos; pos.x = c->wavg.x / c->weight; pos.y = c->wavg.y / c->weight; pos.z = c->wavg.z / c->weight;   /* GetWeightedAvg :34-37 */
            float bx, by; barycentric(tp, pos, &bx, &by);
            v3 n0 = { tn[0], tn[1], tn[2] }, n1 = { tn[3], tn[4], tn[5] }, n2 = { tn[6], tn[7], tn[8] };
            float w0 = 1.f - bx - by;
            v3 in = add3(add3(scale3(n0, w0), scale3(n1, bx)), scale3(n2, by));      /* Triangle.cpp:195-212 */
            v3 nn = mul ? normalize3(mat3_mul(nmat, in)) : normalize3(in);
            if (n_rays < ray_cap) { float* o = rays + 6 * n_rays; o[0] = pos.x; o[1] = pos.y; o[2] = pos.z; o[3] = -nn.x; o[4] = -nn.y; o[5] = -nn.z; }
            n_rays++;
            sum = add3(sum, pos);
        }
    }
    free(emplaced);
    *origin_sum = sum;
    return n_rays;
}

void imro_pair(const imro_tree* a, const float* mat_a, const imro_tree* b, const float* mat_b,
               uint32_t* hit_ids, float* hit_seg, uint64_t cap, uint64_t* summary7, float* avg6,
               float* rays_first, float* rays_second, uint64_t ray_cap) {
    float rel[16], nmat[9];
    imro_pair_matrix(mat_a, mat_b, rel);                              /* :62 */
    adjoint_transpose3(rel, nmat);                                    /* :65 */
    /* mid phase */
    uint64_t stats[5];
    uint64_t ncomb = imro_mid(a, mat_a, b, mat_b, NULL, 0, stats);
    uint32_t* combos = (uint32_t*)malloc(sizeof(uint32_t) * 4 * (ncomb ? ncomb : 1));
    imro_mid(a, mat_a, b, mat_b, combos, ncomb, NULL);

    cand_map ma, mb;
    ma.map = (cand_t*)malloc(sizeof(cand_t) * (a->n_tri ? a->n_tri : 1)); ma.present = (uint8_t*)calloc(a->n_tri ? a->n_tri : 1, 1);
    ma.order = (uint32_t*)malloc(sizeof(uint32_t) * (a->n_tri ? a->n_tri : 1)); ma.n_order = 0;
    mb.map = (cand_t*)malloc(sizeof(cand_t) * (b->n_tri ? b->n_tri : 1)); mb.present = (uint8_t*)calloc(b->n_tri ? b->n_tri : 1, 1);
    mb.order = (uint32_t*)malloc(sizeof(uint32_t) * (b->n_tri ? b->n_tri : 1)); mb.n_order = 0;

    uint64_t n_tests = 0, n_hits = 0, n_copl = 0;
    for (uint64_t k = 0; k < ncomb; ++k) {
        uint32_t offA = combos[4 * k], cntA = combos[4 * k + 1], offB = combos[4 * k + 2], cntB = combos[4 * k + 3];
        cand_t ca[4], cb[4];
        for (int q = 0; q < 4; ++q) { ca[q].bits = 7; ca[q].wavg.x = ca[q].wavg.y = ca[q].wavg.z = 0.f; ca[q].weight = 0.f; cb[q] = ca[q]; }
        for (uint32_t i = 0; i != cntA; ++i)
            for (uint32_t j = 0; j != cntB; ++j) {
                const float* ta = a->tri_pos + 9 * (uint64_t)(offA + i);
                float tb[9];
                tri_transform(rel, b->tri_pos + 9 * (uint64_t)(offB + j), tb);             /* :84 */
                int cop = 0; float s[3], t[3];
                int hit = tri_tri_isectline(ta, ta + 3, ta + 6, tb, tb + 3, tb + 6, &cop, s, t);   /* :86 */
                ++n_tests;
                if (hit && cop) ++n_copl;
                if (hit && !cop) {                                                         /* :88 */
                    plane_t pa = plane_from_tri(ta), pb = plane_from_tri(tb);
                    v3 src = { s[0], s[1], s[2] }, tgt = { t[0], t[1], t[2] };
                    float weight = length3(sub3(src, tgt));                                /* :93 */
                    v3 sum = add3(src, tgt);
                    v3 avgp; avgp.x = (weight * sum.x) / 2.f; avgp.y = (weight * sum.y) / 2.f; avgp.z = (weight * sum.z) / 2.f;   /* :94 */
                    ca[i].wavg = add3(ca[i].wavg, avgp); ca[i].weight += weight;
                    cb[j].wavg = add3(cb[j].wavg, avgp); cb[j].weight += weight;
                    for (int pi = 0; pi < 3; ++pi) {
                        v3 pA = { ta[3 * pi], ta[3 * pi + 1], ta[3 * pi + 2] }, pB = { tb[3 * pi], tb[3 * pi + 1], tb[3 * pi + 2] };
                        if (plane_point_outside(&pb, pA)) ca[i].bits &= ~(1u << pi);
                        if (plane_point_outside(&pa, pB)) cb[j].bits &= ~(1u << pi);
                    }
                    if (n_hits < cap) {
                        hit_ids[2 * n_hits] = a->tri_orig[offA + i]; hit_ids[2 * n_hits + 1] = b->tri_orig[offB + j];
                        memcpy(hit_seg + 7 * n_hits, s, 12); memcpy(hit_seg + 7 * n_hits + 3, t, 12);
                        hit_seg[7 * n_hits + 6] = weight;
                    }
                    ++n_hits;
                }
            }
        for (uint32_t i = 0; i != cntA; ++i) if (!(ca[i].weight == 0.f)) cand_merge(&ma, offA + i, &ca[i]);   /* :117-121 */
        for (uint32_t j = 0; j != cntB; ++j) if (!(cb[j].weight == 0.f)) cand_merge(&mb, offB + j, &cb[j]);
    }

    v3 sum_a, sum_b;
    uint64_t ra = find_rays(&ma, a, 0, rel, nmat, rays_first, rays_first ? ray_cap : 0, &sum_a);
    uint64_t rb = find_rays(&mb, b, 1, rel, nmat, rays_second, rays_second ? ray_cap : 0, &sum_b);
    if (avg6) {
        /* :185-198 ; division by zero rays gives NaN exactly like the reference (it discards the pair anyway) */
        float fa = (float)ra, fb = (float)rb;
        avg6[0] = sum_a.x / fa; avg6[1] = sum_a.y / fa; avg6[2] = sum_a.z / fa;
        v3 sb = { sum_b.x / fb, sum_b.y / fb, sum_b.z / fb };
        float inv[16]; mat4_inverse(rel, inv);
        v3 back = mat4_mul_point(inv, sb, 1.f);
        avg6[3] = back.x; avg6[4] = back.y; avg6[5] = back.z;
    }
    summary7[0] = ncomb; summary7[1] = n_tests; summary7[2] = n_hits; summary7[3] = n_copl;
    summary7[4] = ra; summary7[5] = rb; summary7[6] = (ra || rb) ? 1 : 0;
    free(combos);
    free(ma.map); free(ma.present); free(ma.order); free(mb.map); free(mb.present); free(mb.order);
}

/* ------------------------------------------------------------------ */
/* response rays: Ray.cpp, ShootUncollideRays.cpp, CollisionDetection.cpp:80-103 */
/* ------------------------------------------------------------------ */
typedef struct { int hit, back; float dist; uint64_t tri; float bx, by; } rayhit_t;      /* RayOBBtreeIntersectInfo, Ray.h:17-24 */

/* glm fork, glm/gtx/intersect.inl:29-97 (intersectRayTriangle with itBackfaces) */
static int ray_triangle(v3 orig, v3 dir, v3 v0, v3 v1, v3 v2, float* bx, float* by, float* distance, int* back) {
    v3 edge1 = sub3(v1, v0), edge2 = sub3(v2, v0);
    v3 p = cross3(dir, edge2);
    float det = dot3(edge1, p);
    v3 perp = { 0.f, 0.f, 0.f };
    float x, y;
    if (det > FLT_EPSILON) {
        v3 dist = sub3(orig, v0);
        x = dot3(dist, p);
        if (x < 0.f || x > det) return 0;
        perp = cross3(dist, edge1);
        y = dot3(dir, perp);
        if (y < 0.f || (x + y) > det) return 0;
        *back = 0;
    } else if (det < -FLT_EPSILON) {
        v3 dist = sub3(orig, v0);
        x = dot3(dist, p);
        if (x > 0.f || x < det) return 0;
        perp = cross3(dist, edge1);
        y = dot3(dir, perp);
        if (y > 0.f || (x + y) < det) return 0;
        *back = 1;
    } else return 0;
    float inv_det = 1.f / det;
    *distance = dot3(edge2, perp) * inv_det;
    *bx = x * inv_det; *by = y * inv_det;
    return 1;
}

/* one slab of Ray::IntersectParalgram, Ray.cpp:49-75 (the U, V and W blocks are the same code) */
static int ray_slab(v3 a, v3 b, v3 side, v3 ray_origin, v3 dir, float* mn, float* mx) {
    v3 plane_dir = normalize3(cross3(a, b));
    float d = -fabsf(dot3(plane_dir, side));
    float c = dot3(plane_dir, ray_origin);
    float v_n1 = +c + d;
    float v_n2 = -c + d;
    float vd = dot3(plane_dir, dir);
    if (fabsf(vd) >= FLT_EPSILON) {
        float vd_inv = 1.f / vd;
        float t1 = -v_n1 * vd_inv;
        float t2 = +v_n2 * vd_inv;
        if (t1 > t2) { float t = t1; t1 = t2; t2 = t; }
        *mn = (t1 < *mn) ? *mn : t1;               /* std::max(t1, min_distance) */
        *mx = (*mx < t2) ? *mx : t2;               /* std::min(t2, max_distance) */
        if (*mn > *mx || *mx < 0.f) return 0;
    } else if (v_n1 > 0 || v_n2 > 0) return 0;
    return 1;
}
/* Ray::IntersectParalgram, Ray.cpp:38-134; the ray's origin is always (0,0,0) here (Ray.cpp:138,153-158) */
static int ray_box(v3 origin, v3 dir, const box_t* bx, float* tmin, float* tmax) {
    float mn = -INFINITY, mx = +INFINITY;
    v3 ro = sub3(origin, bx->c);
    if (!ray_slab(bx->v, bx->w, bx->u, ro, dir, &mn, &mx)) return 0;       /* U test :48-75 */
    if (!ray_slab(bx->w, bx->u, bx->v, ro, dir, &mn, &mx)) return 0;       /* V test :77-104 */
    if (!ray_slab(bx->u, bx->v, bx->w, ro, dir, &mn, &mx)) return 0;       /* W test :106-133 */
    *tmin = mn; *tmax = mx;
    return 1;
}

/* Ray::IntersectOBBtreeRecursive, Ray.cpp:163-236 */
static void ray_tree_rec(const imro_tree* t, int32_t v, const float* m, v3 dir, rayhit_t* best) {
    const v3 zero = { 0.f, 0.f, 0.f };
    if (t->left[v] >= 0) {
        int32_t l = t->left[v], r = t->right[v];
        box_t lb = box_transform(m, box_load(t->boxes + 12 * (uint64_t)l)), rb = box_transform(m, box_load(t->boxes + 12 * (uint64_t)r));
        float lmin = 0.f, lmax = 0.f, rmin = 0.f, rmax = 0.f;
        int lh = ray_box(zero, dir, &lb, &lmin, &lmax), rh = ray_box(zero, dir, &rb, &rmin, &rmax);
        if (lh && rh) {
            if (lmin < rmin) {
                if (lmin < best->dist && lmax >= 0.f) ray_tree_rec(t, l, m, dir, best);
                if (rmin < best->dist && rmax >= 0.f) ray_tree_rec(t, r, m, dir, best);
            } else {
                if (rmin < best->dist && rmax >= 0.f) ray_tree_rec(t, r, m, dir, best);
                if (lmin < best->dist && lmax >= 0.f) ray_tree_rec(t, l, m, dir, best);
            }
        } else if (lh) {
            if (lmin < best->dist && lmax >= 0.f) ray_tree_rec(t, l, m, dir, best);
        } else if (rh) {
            if (rmin < best->dist && rmax >= 0.f) ray_tree_rec(t, r, m, dir, best);
        }
    } else {
        for (uint64_t i = t->tri_off[v]; i != (uint64_t)t->tri_off[v] + t->tri_cnt[v]; ++i) {
            float tp[9];
            tri_transform(m, t->tri_pos + 9 * i, tp);
            v3 p0 = { tp[0], tp[1], tp[2] }, p1 = { tp[3], tp[4], tp[5] }, p2 = { tp[6], tp[7], tp[8] };
            float bx = 0.f, by = 0.f, dist = INFINITY; int back = 0;
            int hit = ray_triangle(zero, dir, p0, p1, p2, &bx, &by, &dist, &back);
            if (hit && dist > 0.f && dist < best->dist) { best->hit = 1; best->back = back; best->dist = dist; best->tri = i; best->bx = bx; best->by = by; }
        }
    }
}

/* Ray::IntersectOBBtree, Ray.cpp:136-161 */
static rayhit_t ray_tree(const imro_tree* t, const float* m16, v3 origin, v3 dir) {
    rayhit_t best = { 0, 0, INFINITY, (uint64_t)-1, 0.f, 0.f };
    float m[16]; memcpy(m, m16, 64);
    if (!(origin.x == 0.f && origin.y == 0.f && origin.z == 0.f)) {        /* centered_matrix[3] -= vec4(origin, 0) */
        m[12] = m[12] - origin.x; m[13] = m[13] - origin.y; m[14] = m[14] - origin.z; m[15] = m[15] - 0.f;
    }
    const v3 zero = { 0.f, 0.f, 0.f };
    box_t root = box_transform(m, box_load(t->boxes));
    float mn = 0.f, mx = 0.f;
    if (ray_box(zero, dir, &root, &mn, &mx) && mx >= 0.f) ray_tree_rec(t, 0, m, dir, &best);
    return best;
}

int imro_ray_tree(const imro_tree* t, const float* m16, const float* origin3, const float* dir3, float* out3, uint32_t* tri, int* back) {
    v3 o = { origin3[0], origin3[1], origin3[2] }, d = { dir3[0], dir3[1], dir3[2] };
    rayhit_t h = ray_tree(t, m16, o, d);
    out3[0] = h.dist; out3[1] = h.bx; out3[2] = h.by;
    *tri = (uint32_t)h.tri; *back = h.back;
    return h.hit;
}

typedef struct { int ok; v3 response, normal; } hermann_t;

/* ShootUncollideRays::HermannPass, ShootUncollideRays.cpp:116-148 (the ray is a copy: its origin is moved inside) */
static hermann_t hermann_pass(const imro_tree* A, const imro_tree* B, const float* matA, const float* matB, const float* nmatB, v3 origin, v3 dir) {
    hermann_t r; r.ok = 0; r.response.x = r.response.y = r.response.z = 0.f; r.normal = r.response;
    rayhit_t p2 = ray_tree(B, matB, origin, dir);
    if (p2.hit && p2.back) {
        /* Ray::MoveOriginEpsilonTowardsDirection(4.f), Ray.cpp:13-21 */
        float big = fabsf(origin.x); if (big < fabsf(origin.y)) big = fabsf(origin.y); if (big < fabsf(origin.z)) big = fabsf(origin.z);
        float scaled = big * FLT_EPSILON;
        v3 moved = add3(origin, scale3(dir, 4.f * scaled));
        float eps_dist = length3(sub3(moved, origin));
        rayhit_t p3 = ray_tree(A, matA, moved, dir);
        if (p2.dist <= p3.dist + eps_dist) {
            r.ok = 1;
            r.response = scale3(dir, p2.dist);                                 /* distanceFromOrigin * ray.GetDirection() */
            const float* tn = B->tri_nrm + 9 * p2.tri;
            v3 n0 = { tn[0], tn[1], tn[2] }, n1 = { tn[3], tn[4], tn[5] }, n2 = { tn[6], tn[7], tn[8] };
            float w0 = 1.f - p2.bx - p2.by;
            v3 in = add3(add3(scale3(n0, w0), scale3(n1, p2.bx)), scale3(n2, p2.by));   /* Triangle.cpp:197-204 */
            r.normal = normalize3(mat3_mul(nmatB, in));
        }
    }
    return r;
}

/* ShootUncollideRays::CalcForceResponse, ShootUncollideRays.cpp:104-114 */
static v3 force_response(const hermann_t* h) {
    v3 dirn = normalize3(h->response);
    float len = length3(h->response);
    float c = dot3(h->normal, dirn);
    return scale3(dirn, (c * c) * len);
}

typedef struct { v3 force; v3* resp; uint64_t n, cap; } shoot_acc;
static void shoot_push(shoot_acc* a, v3 v) {
    if (a->n == a->cap) { a->cap = a->cap ? 2 * a->cap : 64; a->resp = (v3*)realloc(a->resp, sizeof(v3) * a->cap); }
    a->resp[a->n++] = v;
}

/* ShootUncollideRays::ExecuteShootUncollideRays, ShootUncollideRays.cpp:14-93.  rays: 6 floats each (origin, direction), first's model
 * space.  Returns the world-space delta (first.current * vec4(localspace_response, 0)). */
void imro_shoot(const imro_tree* a, const float* mat_a, const imro_tree* b, const float* mat_b,
                const float* rays_first, uint64_t n_first, const float* rays_second, uint64_t n_second, float* delta3, uint64_t* n_responses) {
    static const float ident4[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
    static const float ident3[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    float rel[16], nmat[9];
    imro_pair_matrix(mat_a, mat_b, rel);
    adjoint_transpose3(rel, nmat);
    shoot_acc acc; memset(&acc, 0, sizeof(acc));
    for (int phase = 0; phase < 2; ++phase) {
        const float* rays = phase ? rays_second : rays_first;
        uint64_t n = phase ? n_second : n_first;
        for (uint64_t k = 0; k < n; ++k) {
            v3 o = { rays[6 * k], rays[6 * k + 1], rays[6 * k + 2] }, d = { rays[6 * k + 3], rays[6 * k + 4], rays[6 * k + 5] };
            for (int step = 0; step < 2; ++step) {
                /* first_to_second when (phase == 0) == (step == 0), else second_to_first (:29-71) */
                int f2s = (phase == 0) == (step == 0);
                hermann_t h = f2s ? hermann_pass(a, b, ident4, rel, nmat, o, d) : hermann_pass(b, a, rel, ident4, ident3, o, d);
                if (!h.ok) break;
                v3 fr = force_response(&h);
                if (f2s) { acc.force = sub3(acc.force, fr); v3 neg = { -h.response.x, -h.response.y, -h.response.z }; shoot_push(&acc, neg); }
                else { acc.force = add3(acc.force, fr); shoot_push(&acc, h.response); }
                /* ReflectHermannResult :95-101 */
                o = add3(o, h.response);
                d.x = -h.normal.x; d.y = -h.normal.y; d.z = -h.normal.z;
            }
        }
    }
    v3 local = { 0.f, 0.f, 0.f };
    if (!(acc.force.x == 0.f && acc.force.y == 0.f && acc.force.z == 0.f)) {
        v3 nf = normalize3(acc.force);
        /* FindResponse :150-172; edges: cos(65 deg) .. cos(40 deg) (CollisionDetection.cpp:22-24, ShootUncollideRays.cpp:8-9) */
        const float edge_b = cosf(0.01745329251994329576923690768489f * 40.f), edge_a = cosf(0.01745329251994329576923690768489f * 65.f);
        float max_len = 0.f;
        for (uint64_t k = 0; k < acc.n; ++k) {
            v3 nr = normalize3(acc.resp[k]);
            float c = dot3(nr, nf);
            float len = length3(acc.resp[k]);
            float need = len / c;
            float tq = (c - edge_a) / (edge_b - edge_a);
            float tmp = tq < 0.f ? 0.f : (1.f < tq ? 1.f : tq);                    /* std::clamp */
            float ss = tmp * tmp * tmp * (tmp * (tmp * 6 - 15) + 10);               /* SmootherStep :174-178 */
            float cand = ss * need;
            max_len = max_len < cand ? cand : max_len;                             /* std::max(max_length, cand) */
        }
        local = scale3(scale3(nf, max_len), 1.01f);                                /* ray_distance_bias_multiplier * (max_length * n) */
    }
    v3 w = mat4_mul_point(mat_a, local, 0.f);
    delta3[0] = w.x; delta3[1] = w.y; delta3[2] = w.z;
    if (n_responses) *n_responses = acc.n;
    free(acc.resp);
}

/* CollisionDetection::PointMovementBetweenFrames, CollisionDetection.cpp:143-150 */
static float point_movement(v3 p, const float* m_first, const float* m_second) {
    v3 a = mat4_mul_point(m_first, p, 1.f), b = mat4_mul_point(m_second, p, 1.f);
    return length3(sub3(a, b));
}

/* CollisionDetection.cpp:60-103 for one ordered pair: rays, then (when either entry moved) the response.  delta6 = deltaVector of first, of second.
 * Returns 1 when the pair is colliding (>= 1 ray). */
int imro_pair_delta(const imro_tree* a, const float* mat_a, const float* prev_a, const imro_tree* b, const float* mat_b, const float* prev_b,
                    float* delta6, uint64_t* n_responses) {
    uint64_t summ[7]; float avg[6];
    memset(delta6, 0, 24);
    if (n_responses) *n_responses = 0;
    imro_pair(a, mat_a, b, mat_b, NULL, NULL, 0, summ, avg, NULL, NULL, 0);
    if (!summ[6]) return 0;
    int moved = 0;                                     /* glm mat4 operator!= : any component differs (:80-81) */
    for (int i = 0; i < 16; ++i) if (mat_a[i] != prev_a[i] || mat_b[i] != prev_b[i]) moved = 1;
    if (!moved) return 1;
    uint64_t ra = summ[4], rb = summ[5];
    float* r1 = (float*)malloc(sizeof(float) * 6 * (ra ? ra : 1));
    float* r2 = (float*)malloc(sizeof(float) * 6 * (rb ? rb : 1));
    imro_pair(a, mat_a, b, mat_b, NULL, NULL, 0, summ, avg, r1, r2, ra > rb ? ra : rb);
    float d[3];
    imro_shoot(a, mat_a, b, mat_b, r1, ra, r2, rb, d, n_responses);
    v3 pa = { avg[0], avg[1], avg[2] }, pb = { avg[3], avg[4], avg[5] };
    float m1 = point_movement(pa, mat_a, prev_a), m2 = point_movement(pb, mat_b, prev_b);
    float total = m1 + m2;
    float f1 = m1 / total, f2 = m2 / total;
    delta6[0] = -d[0] * f1; delta6[1] = -d[1] * f1; delta6[2] = -d[2] * f1;       /* - delta * (first / total) :88 */
    delta6[3] = d[0] * f2; delta6[4] = d[1] * f2; delta6[5] = d[2] * f2;          /* + delta * (second / total) :89 */
    free(r1); free(r2);
    return 1;
}

/* ------------------------------------------------------------------ */
/* Triangle::CreateTriangleList (IMR/src/Geometry/Triangle.cpp:9-62,214-280) */
/* ------------------------------------------------------------------ */
static uint64_t triplet_count(uint32_t mode, uint64_t ni) {
    switch (mode) { case 0: return ni; case 1: return ni / 2; case 3: return ni ? ni - 1 : 0; case 4: return ni / 3; case 5: case 6: return ni >= 2 ? ni - 2 : 0; default: return 0; }
}
static void triplet(uint32_t mode, uint64_t i, uint64_t* a, uint64_t* b, uint64_t* c) {      /* CreateIndicesTriplets :14-59 */
    switch (mode) {
        case 0: *a = i; *b = i; *c = i; break;
        case 1: *a = 2 * i; *b = 2 * i; *c = 2 * i + 1; break;
        case 3: *a = i; *b = i; *c = i + 1; break;
        case 4: *a = 3 * i; *b = 3 * i + 1; *c = 3 * i + 2; break;
        case 5: *a = i; *b = i + (1 + i % 2); *c = i + (2 - i % 2); break;
        default: *a = i + 1; *b = i + 2; *c = 0; break;
    }
}
/* points / normals: n_points x 3 floats (normals may be NULL: face normals, :223-232); out arrays sized by the return value of a first call with NULLs */
uint64_t imro_triangle_list(const float* points, const float* normals, const uint32_t* indices, uint64_t n_indices, uint32_t mode,
                            float* pos9, float* nrm9, uint32_t* vid3) {
    uint64_t n = triplet_count(mode, n_indices);
    if (!pos9) return n;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t k[3]; triplet(mode, i, &k[0], &k[1], &k[2]);
        for (int q = 0; q < 3; ++q) {
            uint32_t v = indices[k[q]];
            memcpy(pos9 + 9 * i + 3 * q, points + 3 * (uint64_t)v, 12);
            if (normals) memcpy(nrm9 + 9 * i + 3 * q, normals + 3 * (uint64_t)v, 12);
            vid3[3 * i + q] = v;                                       /* CreateTriangleIndicesList :242-250: iota data, so the index itself */
        }
        if (!normals) { v3 fn = face_normal(pos9 + 9 * i); for (int q = 0; q < 3; ++q) { nrm9[9 * i + 3 * q] = fn.x; nrm9[9 * i + 3 * q + 1] = fn.y; nrm9[9 * i + 3 * q + 2] = fn.z; } }
    }
    return n;
}

/* Batch mid + narrow over a pair list: the loop of IMR/src/CollisionDetection/CollisionDetection.cpp:44-69 on
 * (first,second) entry-index pairs.  Re-entrant.  totals: [0] combos [1] tri-pair tests [2] colliding pairs
 * [3] pairs with >= 1 combo. */
void imro_frame_pairs(const float* mats, const imro_tree* const* trees, const uint32_t* pairs, uint64_t n_pairs, uint64_t* totals) {
    uint64_t combos = 0, tests = 0, colliding = 0, with_combos = 0;
    for (uint64_t k = 0; k < n_pairs; ++k) {
        const uint32_t ia = pairs[2 * k], ib = pairs[2 * k + 1];
        uint64_t summ[7]; float avg[6];
        imro_pair(trees[ia], mats + 16 * (uint64_t)ia, trees[ib], mats + 16 * (uint64_t)ib, NULL, NULL, 0, summ, avg, NULL, NULL, 0);
        combos += summ[0]; tests += summ[1]; colliding += summ[6]; with_combos += summ[0] ? 1 : 0;
    }
    totals[0] = combos; totals[1] = tests; totals[2] = colliding; totals[3] = with_combos;
}


/* ---- re-posing (BASELINE config 5; parity = this port, the reference runs it as GLSL) --------------------------------------
 * IMR/shaders/dynamicMeshShader_glsl.comp:99-145 for the position stream (VEC = vec4, no NORMALIZE / ZERO_W / USE_NORMAL_MATRIX):
 *   :105-111  morphed = V[x (T + 1)];  for i < T: morphed += morph_weights[i] * V[x (T + 1) + i + 1]
 *   :121-123  result = morphed when jointsGroupsCount == 0
 *   :124-132  result = sum over groups, components c: weights.c * CalucateSkinJoint(joints.c + matrixOffset + 1, joints.c + inverseMatricesOffset, morphed)
 *   :79-84    CalucateSkinJoint = modelMatrices[m].positionMatrix * inverseModelMatrices[i].positionMatrix * vertex
 * GLSL does not fix the order of the sums inside a matrix product; taken here as (M * InvBind) * v in glm's order
 * (type_mat4x4.inl:630-648 for mat4 * mat4, :561-572 for mat4 * vec4), no contraction -- the order the device uses.
 * vertices: n_vertices * (n_targets + 1) * 4 floats; joints / weights: n_vertices * n_groups * 4; matrices: n_joints * 16 each (already
 * offset: joint j is modelMatrices[matrixOffset + 1 + j] / inverseModelMatrices[inverseMatricesOffset + j]); out: n_vertices * 4. */
static void mat4_mul_vec4(const float* m, const float* v, float* r) {
    for (int k = 0; k < 4; ++k) r[k] = (m[k] * v[0] + m[4 + k] * v[1]) + (m[8 + k] * v[2] + m[12 + k] * v[3]);
}
void imro_repose(uint64_t n_vertices, uint32_t n_targets, const float* vertices, uint32_t n_groups, const uint16_t* joints, const float* weights,
                 const float* morph_weights, uint32_t n_joints, const float* joint_matrices, const float* inverse_bind, float* out) {
    float* prod = n_joints ? (float*)malloc(64ull * n_joints) : NULL;
    for (uint32_t j = 0; j < n_joints; ++j) mat4_mul(joint_matrices + 16ull * j, inverse_bind + 16ull * j, prod + 16ull * j);
    for (uint64_t x = 0; x < n_vertices; ++x) {
        const float* V = vertices + 4ull * (n_targets + 1ull) * x;
        float m[4] = { V[0], V[1], V[2], V[3] };
        for (uint32_t i = 0; i < n_targets; ++i) {
            const float* t = V + 4ull * (i + 1u); const float w = morph_weights[i];
            for (int k = 0; k < 4; ++k) m[k] += w * t[k];
        }
        float r[4] = { m[0], m[1], m[2], m[3] };
        if (n_groups) {
            r[0] = r[1] = r[2] = r[3] = 0.f;
            for (uint32_t g = 0; g < n_groups; ++g) {
                const float* w = weights + 4ull * (x * n_groups + g); const uint16_t* jn = joints + 4ull * (x * n_groups + g);
                for (int c = 0; c < 4; ++c) {
                    float v[4];
                    mat4_mul_vec4(prod + 16ull * jn[c], m, v);
                    for (int k = 0; k < 4; ++k) r[k] += w[c] * v[k];
                }
            }
        }
        for (int k = 0; k < 4; ++k) out[4 * x + k] = r[k];
    }
    free(prod);
}
