// imrcd_math.cuh -- device restatement of the reference's FP32 predicates, in the reference's
// evaluation order.  Compiled with --fmad=false (no contraction), IEEE div/sqrt (nvcc defaults
// -prec-div=true -prec-sqrt=true -ftz=false), so every operation below is one IEEE-754 binary32
// operation exactly as g++ -O2 -ffp-contract=off emits for the reference (SURVEY.md finding 2).
// "IMR/" = inMyRoom_vulkan/ in the reference checkout; glm = thesmallcreeper/glm @99e83f5.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define IMR_HD __host__ __device__ __forceinline__
#define IMR_D  __device__ __forceinline__

struct V3 { float x, y, z; };

IMR_HD V3 mk3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
// glm/detail/func_geometric.inl:48-55 : tmp = a*b ; tmp.x + tmp.y + tmp.z
IMR_HD float dot3(V3 a, V3 b) { float tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z; return (tx + ty) + tz; }
// func_geometric.inl:68-79
IMR_HD V3 cross3(V3 x, V3 y) { return mk3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
IMR_HD V3 sub3(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
IMR_HD V3 add3(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
IMR_HD V3 scale3(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
// func_geometric.inl:8-14
IMR_HD float length3(V3 a) { return sqrtf(dot3(a, a)); }
// func_geometric.inl:82-90 with inversesqrt = 1/sqrt (func_exponential.inl:134-139)
IMR_HD V3 normalize3(V3 a) { float inv = 1.0f / sqrtf(dot3(a, a)); return scale3(a, inv); }

// Affine part of a column-major glm::mat4 stored by ROWS: r[i] = (m[0][i], m[1][i], m[2][i], m[3][i]).
// Row 3 of the matrix never reaches a vec3 result (glm::vec3(M * vec4) drops w), so 12 floats suffice.
struct Rel { float4 r0, r1, r2; };

// type_mat4x4.inl:561-572 : (m[0]*v0 + m[1]*v1) + (m[2]*v2 + m[3]*v3)
IMR_HD V3 rel_mul(const Rel& m, V3 p, float w) {
    V3 o;
    o.x = (m.r0.x * p.x + m.r0.y * p.y) + (m.r0.z * p.z + m.r0.w * w);
    o.y = (m.r1.x * p.x + m.r1.y * p.y) + (m.r1.z * p.z + m.r1.w * w);
    o.z = (m.r2.x * p.x + m.r2.y * p.y) + (m.r2.z * p.z + m.r2.w * w);
    return o;
}

// ---- full 4x4 (column-major float[16], element (c,r) at 4c+r) -------------------------------
// type_mat4x4.inl:630-648
IMR_HD void mat4_mul(const float* a, const float* b, float* out) {
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 4; ++r)
            out[4 * c + r] = ((a[0 + r] * b[4 * c + 0] + a[4 + r] * b[4 * c + 1]) + a[8 + r] * b[4 * c + 2]) + a[12 + r] * b[4 * c + 3];
}
// func_matrix.inl:347-405 (compute_inverse<4,4>)
IMR_HD void mat4_inverse(const float* m, float* out) {
#define M(c, r) m[4 * (c) + (r)]
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
    const float Fac0[4] = { Coef00, Coef00, Coef02, Coef03 };
    const float Fac1[4] = { Coef04, Coef04, Coef06, Coef07 };
    const float Fac2[4] = { Coef08, Coef08, Coef10, Coef11 };
    const float Fac3[4] = { Coef12, Coef12, Coef14, Coef15 };
    const float Fac4[4] = { Coef16, Coef16, Coef18, Coef19 };
    const float Fac5[4] = { Coef20, Coef20, Coef22, Coef23 };
    const float Vec0[4] = { M(1,0), M(0,0), M(0,0), M(0,0) };
    const float Vec1[4] = { M(1,1), M(0,1), M(0,1), M(0,1) };
    const float Vec2[4] = { M(1,2), M(0,2), M(0,2), M(0,2) };
    const float Vec3[4] = { M(1,3), M(0,3), M(0,3), M(0,3) };
    float inv[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float sa = (i & 1) ? -1.f : 1.f, sb = (i & 1) ? 1.f : -1.f;
        float Inv0 = (Vec1[i] * Fac0[i] - Vec2[i] * Fac1[i]) + Vec3[i] * Fac2[i];
        float Inv1 = (Vec0[i] * Fac0[i] - Vec2[i] * Fac3[i]) + Vec3[i] * Fac4[i];
        float Inv2 = (Vec0[i] * Fac1[i] - Vec1[i] * Fac3[i]) + Vec3[i] * Fac5[i];
        float Inv3 = (Vec0[i] * Fac2[i] - Vec1[i] * Fac4[i]) + Vec2[i] * Fac5[i];
        inv[0 + i] = Inv0 * sa; inv[4 + i] = Inv1 * sb; inv[8 + i] = Inv2 * sa; inv[12 + i] = Inv3 * sb;
    }
    float Dot0x = M(0,0) * inv[0], Dot0y = M(0,1) * inv[4], Dot0z = M(0,2) * inv[8], Dot0w = M(0,3) * inv[12];
    float Dot1 = (Dot0x + Dot0y) + (Dot0z + Dot0w);
    float OneOverDeterminant = 1.0f / Dot1;
#pragma unroll
    for (int i = 0; i < 16; ++i) out[i] = inv[i] * OneOverDeterminant;
#undef M
}

// ---- Paralgram / OBB (IMR/include/Geometry/Paralgram.h:29-35) ---------------------------------
struct Box { V3 c, u, v, w; };

// IMR/src/Geometry/Paralgram.cpp:4-15
IMR_HD Box box_transform(const Rel& m, const Box& b) {
    Box r;
    r.c = rel_mul(m, b.c, 1.f);
    r.u = rel_mul(m, b.u, 0.f);
    r.v = rel_mul(m, b.v, 0.f);
    r.w = rel_mul(m, b.w, 0.f);
    return r;
}
// Paralgram.cpp:175-196
IMR_HD void box_minmax(const Box& b, V3 axis, float& mn, float& mx) {
    float cp = dot3(b.c, axis);
    float pu = fabsf(dot3(axis, b.u));
    float pv = fabsf(dot3(axis, b.v));
    float pw = fabsf(dot3(axis, b.w));
    float sum = (pu + pv) + pw;
    mn = cp - sum;
    mx = cp + sum;
}
// Paralgram.cpp:198-201 (closed intervals)
IMR_HD bool axis_overlap(const Box& l, const Box& r, V3 axis) {
    float lmn, lmx, rmn, rmx;
    box_minmax(l, axis, lmn, lmx);
    box_minmax(r, axis, rmn, rmx);
    return (lmx >= rmn) & (rmx >= lmn);
}
// The k-th separating axis in the reference's fixed order (Paralgram.cpp:21,31,41,52,62,72,83-163).
IMR_HD V3 sat_axis(const Box& l, const Box& r, int k) {
    switch (k) {
        case 0: return cross3(l.v, l.w);
        case 1: return cross3(l.u, l.w);
        case 2: return cross3(l.u, l.v);
        case 3: return cross3(r.v, r.w);
        case 4: return cross3(r.u, r.w);
        case 5: return cross3(r.u, r.v);
        case 6: return cross3(l.u, r.u);
        case 7: return cross3(l.u, r.v);
        case 8: return cross3(l.u, r.w);
        case 9: return cross3(l.v, r.u);
        case 10: return cross3(l.v, r.v);
        case 11: return cross3(l.v, r.w);
        case 12: return cross3(l.w, r.u);
        case 13: return cross3(l.w, r.v);
        default: return cross3(l.w, r.w);
    }
}
// Paralgram.cpp:17-173.  The verdict does not depend on evaluation order or early exit
// (each axis test is a pure function of the two boxes), so the device evaluates all 15 and ANDs.
// MODE 0: lane-level early exit after every axis (branches); 1: straight-line, all 15 axes; 2: the six face-normal axes straight-line,
// one exit, then the nine edge-edge axes straight-line (most separated pairs are caught by a face normal).
// the two halves of the verdict: the six face-normal axes (Paralgram.cpp:21-81) and the nine edge-edge axes (:83-163)
IMR_HD bool box_sat_faces(const Box& l, const Box& r) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 6; ++k) ok = ok & axis_overlap(l, r, sat_axis(l, r, k));
    return ok;
}
IMR_HD bool box_sat_edges(const Box& l, const Box& r) {
    bool ok = true;
#pragma unroll
    for (int k = 6; k < 15; ++k) ok = ok & axis_overlap(l, r, sat_axis(l, r, k));
    return ok;
}
template <int MODE>
IMR_HD bool box_sat_t(const Box& l, const Box& r) {
    bool ok = true;
    if (MODE == 2) {
        ok = box_sat_faces(l, r);
        if (ok) ok = box_sat_edges(l, r);
    } else {
#pragma unroll
        for (int k = 0; k < 15; ++k) {
            if (MODE == 1) ok = ok & axis_overlap(l, r, sat_axis(l, r, k));     // straight-line: no per-axis branches
            else ok = ok && axis_overlap(l, r, sat_axis(l, r, k));             // lane-level early exit (branches)
        }
    }
    return ok;
}
IMR_HD bool box_sat(const Box& l, const Box& r) { return box_sat_t<0>(l, r); }
// The same 15 axis tests dealt to `g` cooperating lanes (g a power of two): lane `sub` of the group evaluates axes sub, sub + g, ... and
// returns the AND of ITS verdicts; the caller ANDs over the group.  The axis is picked by index with selects (the code is the same in every
// lane, only the operands differ), and each axis is evaluated by exactly the operations of sat_axis / axis_overlap, so the combined verdict is
// box_sat's bit for bit.  Used when a warp has fewer node pairs than lanes: the latency of one visit drops by about the group size.
IMR_HD V3 sat_pick6(int i, const Box& l, const Box& r) {
    const V3 a = i == 0 ? l.u : (i == 1 ? l.v : l.w), b = i == 3 ? r.u : (i == 4 ? r.v : r.w);
    return i < 3 ? a : b;
}
IMR_HD bool box_sat_part(const Box& l, const Box& r, int sub, int g) {
    // operand indices of axis k into (l.u, l.v, l.w, r.u, r.v, r.w), 4 bits each, in the reference's axis order (Paralgram.cpp:21-163)
    const unsigned long long xs = 0x222111000334001ull, ys = 0x543543543455122ull;
    bool ok = true;
    for (int k = sub; k < 15; k += g) {
        const V3 x = sat_pick6((int)((xs >> (4 * k)) & 15ull), l, r), y = sat_pick6((int)((ys >> (4 * k)) & 15ull), l, r);
        ok = ok & axis_overlap(l, r, cross3(x, y));
    }
    return ok;
}
// Paralgram.cpp:203-210
IMR_HD float box_surface(const Box& b) {
    float uv = length3(cross3(b.u, b.v));
    float uw = length3(cross3(b.u, b.w));
    float vw = length3(cross3(b.v, b.w));
    return 2.f * ((uv + uw) + vw);
}

// ---- Moller tri-tri with intersection line (IMR/src/Geometry/Triangle.cpp) -------------------
// (double)fabsf(x) < 0.000001  <=>  fabsf(x) <= float(1e-6) = 0x358637BD, because the float
// nearest to 1e-6 lies below it (Triangle.cpp:332,899-901,922-924; SURVEY trap 8).
#define IMR_TT_EPS_F 9.99999997e-07f

struct Tri { float v[3][3]; };   // v[k] = vertex k

// Triangle.cpp:402-419
IMR_HD bool tt_edge_edge(const float* V0, const float* U0, const float* U1, int i0, int i1, float Ax, float Ay) {
    float Bx = U0[i0] - U1[i0];
    float By = U0[i1] - U1[i1];
    float Cx = V0[i0] - U0[i0];
    float Cy = V0[i1] - U0[i1];
    float f = Ay * Bx - Ax * By;
    float d = By * Cx - Bx * Cy;
    if ((f > 0 && d >= 0 && d <= f) || (f < 0 && d <= 0 && d >= f)) {
        float e = Ax * Cy - Ay * Cx;
        if (f > 0) { if (e >= 0 && e <= f) return true; }
        else { if (e <= 0 && e >= f) return true; }
    }
    return false;
}
// Triangle.cpp:421-433
IMR_HD bool tt_edge_tri(const float* V0, const float* V1, const float* U0, const float* U1, const float* U2, int i0, int i1) {
    float Ax = V1[i0] - V0[i0];
    float Ay = V1[i1] - V0[i1];
    if (tt_edge_edge(V0, U0, U1, i0, i1, Ax, Ay)) return true;
    if (tt_edge_edge(V0, U1, U2, i0, i1, Ax, Ay)) return true;
    if (tt_edge_edge(V0, U2, U0, i0, i1, Ax, Ay)) return true;
    return false;
}
// Triangle.cpp:435-458
IMR_HD bool tt_point_in_tri(const float* V0, const float* U0, const float* U1, const float* U2, int i0, int i1) {
    float a, b, c, d0, d1, d2;
    a = U1[i1] - U0[i1]; b = -(U1[i0] - U0[i0]); c = -a * U0[i0] - b * U0[i1]; d0 = (a * V0[i0] + b * V0[i1]) + c;
    a = U2[i1] - U1[i1]; b = -(U2[i0] - U1[i0]); c = -a * U1[i0] - b * U1[i1]; d1 = (a * V0[i0] + b * V0[i1]) + c;
    a = U0[i1] - U2[i1]; b = -(U0[i0] - U2[i0]); c = -a * U2[i0] - b * U2[i1]; d2 = (a * V0[i0] + b * V0[i1]) + c;
    if (d0 * d1 > 0.0f) { if (d0 * d2 > 0.0f) return true; }
    return false;
}
// Triangle.cpp:460-507.  i0/i1 are selected with register-friendly selects instead of indexing.
__host__ __device__ inline bool tt_coplanar(const float* N, const Tri& V, const Tri& U) {
    float A0 = fabsf(N[0]), A1 = fabsf(N[1]), A2 = fabsf(N[2]);
    int i0, i1;
    if (A0 > A1) { if (A0 > A2) { i0 = 1; i1 = 2; } else { i0 = 0; i1 = 1; } }
    else { if (A2 > A1) { i0 = 0; i1 = 1; } else { i0 = 0; i1 = 2; } }
    if (tt_edge_tri(V.v[0], V.v[1], U.v[0], U.v[1], U.v[2], i0, i1)) return true;
    if (tt_edge_tri(V.v[1], V.v[2], U.v[0], U.v[1], U.v[2], i0, i1)) return true;
    if (tt_edge_tri(V.v[2], V.v[0], U.v[0], U.v[1], U.v[2], i0, i1)) return true;
    if (tt_point_in_tri(V.v[0], U.v[0], U.v[1], U.v[2], i0, i1)) return true;
    if (tt_point_in_tri(U.v[0], V.v[0], V.v[1], V.v[2], i0, i1)) return true;
    return false;
}
// Triangle.cpp:764-778
IMR_HD void tt_isect2(V3 X0, V3 X1, V3 X2, float VV0, float VV1, float VV2, float D0, float D1, float D2,
                      float& isect0, float& isect1, V3& ip0, V3& ip1) {
    float tmp = D0 / (D0 - D1);
    isect0 = VV0 + (VV1 - VV0) * tmp;
    V3 diff = sub3(X1, X0);
    diff = mk3(tmp * diff.x, tmp * diff.y, tmp * diff.z);
    ip0 = add3(diff, X0);
    tmp = D0 / (D0 - D2);
    isect1 = VV0 + (VV2 - VV0) * tmp;
    diff = sub3(X2, X0);
    diff = mk3(tmp * diff.x, tmp * diff.y, tmp * diff.z);
    ip1 = add3(X0, diff);
}
// Triangle.cpp:795-830 ; returns true when coplanar.  The reference's five cases call isect2 with one of three orders of the vertices
// (the vertex alone on its side of the other triangle's plane first); here the order is SELECTED and isect2 runs once: the same
// operations on the same operands, without five copies of the code that the lanes of a warp would walk through one after the other.
IMR_HD bool tt_intervals(V3 X0, V3 X1, V3 X2, float VV0, float VV1, float VV2, float D0, float D1, float D2,
                         float D0D1, float D0D2, float& isect0, float& isect1, V3& ip0, V3& ip1) {
    int first;                                  // which vertex goes first: 0 -> (0,1,2), 1 -> (1,0,2), 2 -> (2,0,1)
    bool coplanar = false;
    if (D0D1 > 0.0f)                        first = 2;
    else if (D0D2 > 0.0f)                   first = 1;
    else if (D1 * D2 > 0.0f || D0 != 0.0f)  first = 0;
    else if (D1 != 0.0f)                    first = 1;
    else if (D2 != 0.0f)                    first = 2;
    else { first = 0; coplanar = true; }
    const V3 A = first == 0 ? X0 : (first == 1 ? X1 : X2), B = first == 0 ? X1 : X0, C = first == 2 ? X1 : X2;
    const float VA = first == 0 ? VV0 : (first == 1 ? VV1 : VV2), VB = first == 0 ? VV1 : VV0, VC = first == 2 ? VV1 : VV2;
    const float DA = first == 0 ? D0 : (first == 1 ? D1 : D2), DB = first == 0 ? D1 : D0, DC = first == 2 ? D1 : D2;
    float t0, t1; V3 p0, p1;
    tt_isect2(A, B, C, VA, VB, VC, DA, DB, DC, t0, t1, p0, p1);
    if (!coplanar) { isect0 = t0; isect1 = t1; ip0 = p0; ip1 = p1; }           // coplanar: the outputs stay as the caller left them (:830)
    return coplanar;
}
IMR_HD float tt_pick(V3 a, int index) { return index == 0 ? a.x : (index == 1 ? a.y : a.z); }

// ---- tri_tri_intersect_with_isectline (Triangle.cpp:866-1002) in three pure pieces, so that the narrow-phase kernel can
// hoist the plane of each triangle out of the pair loop and run the cheap rejection and the expensive segment
// computation as two dense passes.  Every piece is a pure function of its inputs, evaluated in the reference's order,
// so any composition gives the reference's bits.
#define TT_CROSS(a, b) mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x)   /* CROSS macro, Triangle.cpp:336-339 */
#define TT_DOT(a, b) ((a.x * b.x + a.y * b.y) + a.z * b.z)

// plane equation of a triangle: N = E1 x E2, d = -N.P0   (Triangle.cpp:877-882 and :905-910)
IMR_HD void tt_plane(V3 P0, V3 P1, V3 P2, V3& N, float& d) {
    V3 E1 = sub3(P1, P0), E2 = sub3(P2, P0);
    N = TT_CROSS(E1, E2);
    d = -TT_DOT(N, P0);
}
// signed distances of X0..X2 to the plane with the coplanarity clamp (:884-903 / :912-926); true = all on one side (reject)
IMR_HD bool tt_side(V3 N, float d, V3 X0, V3 X1, V3 X2, float& s0, float& s1, float& s2, float& s0s1, float& s0s2) {
    s0 = TT_DOT(N, X0) + d; s1 = TT_DOT(N, X1) + d; s2 = TT_DOT(N, X2) + d;
    if (fabsf(s0) <= IMR_TT_EPS_F) s0 = 0.0f;
    if (fabsf(s1) <= IMR_TT_EPS_F) s1 = 0.0f;
    if (fabsf(s2) <= IMR_TT_EPS_F) s2 = 0.0f;
    s0s1 = s0 * s1; s0s2 = s0 * s2;
    return s0s1 > 0.0f && s0s2 > 0.0f;
}
// everything after the two rejection tests (:928-1001).  Returns bit0 = intersect, bit1 = coplanar; src/tgt valid iff 1.
__host__ __device__ inline int tt_segment(V3 V0, V3 V1, V3 V2, V3 U0, V3 U1, V3 U2, V3 N1, V3 N2,
                                          float du0, float du1, float du2, float du0du1, float du0du2,
                                          float dv0, float dv1, float dv2, float dv0dv1, float dv0dv2, V3& src, V3& tgt) {
    V3 D = TT_CROSS(N1, N2);
    float mx = fabsf(D.x); int index = 0;
    float b = fabsf(D.y), c = fabsf(D.z);
    if (b > mx) { mx = b; index = 1; }
    if (c > mx) { mx = c; index = 2; }
    float vp0 = tt_pick(V0, index), vp1 = tt_pick(V1, index), vp2 = tt_pick(V2, index);
    float up0 = tt_pick(U0, index), up1 = tt_pick(U1, index), up2 = tt_pick(U2, index);

    // If the second interval computation finds U coplanar (all du == 0) while the first did not, the reference reads
    // isect2[] / isectpointB* uninitialised (Triangle.cpp:959-960): undefined behaviour.  Defined here as zeros.
    float i1a, i1b, i2a = 0.f, i2b = 0.f; V3 A1, A2, B1 = mk3(0.f, 0.f, 0.f), B2 = mk3(0.f, 0.f, 0.f);
    bool cop = tt_intervals(V0, V1, V2, vp0, vp1, vp2, dv0, dv1, dv2, dv0dv1, dv0dv2, i1a, i1b, A1, A2);
    if (cop) {
        // rare path: only here do the vertices go to indexable (local-memory) arrays
        float N[3] = { N1.x, N1.y, N1.z };
        Tri TV, TU;
        TV.v[0][0] = V0.x; TV.v[0][1] = V0.y; TV.v[0][2] = V0.z; TV.v[1][0] = V1.x; TV.v[1][1] = V1.y; TV.v[1][2] = V1.z;
        TV.v[2][0] = V2.x; TV.v[2][1] = V2.y; TV.v[2][2] = V2.z;
        TU.v[0][0] = U0.x; TU.v[0][1] = U0.y; TU.v[0][2] = U0.z; TU.v[1][0] = U1.x; TU.v[1][1] = U1.y; TU.v[1][2] = U1.z;
        TU.v[2][0] = U2.x; TU.v[2][1] = U2.y; TU.v[2][2] = U2.z;
        return tt_coplanar(N, TV, TU) ? 3 : 2;
    }
    tt_intervals(U0, U1, U2, up0, up1, up2, du0, du1, du2, du0du1, du0du2, i2a, i2b, B1, B2);
    int smallest1 = 0, smallest2 = 0;
    if (i1a > i1b) { float t = i1a; i1a = i1b; i1b = t; smallest1 = 1; }
    if (i2a > i2b) { float t = i2a; i2a = i2b; i2b = t; smallest2 = 1; }
    if (i1b < i2a || i2b < i1a) return 0;
    if (i2a < i1a) {
        src = (smallest1 == 0) ? A1 : A2;
        if (i2b < i1b) tgt = (smallest2 == 0) ? B2 : B1;
        else           tgt = (smallest1 == 0) ? A2 : A1;
    } else {
        src = (smallest2 == 0) ? B1 : B2;
        if (i2b > i1b) tgt = (smallest1 == 0) ? A2 : A1;
        else           tgt = (smallest2 == 0) ? B2 : B1;
    }
    return 1;
}
// Triangle.cpp:866-1002 as one call (unit-test hook).
__host__ __device__ inline int tri_tri_isectline(V3 V0, V3 V1, V3 V2, V3 U0, V3 U1, V3 U2, V3& src, V3& tgt) {
    V3 N1, N2; float d1, d2;
    float du0, du1, du2, du0du1, du0du2, dv0, dv1, dv2, dv0dv1, dv0dv2;
    tt_plane(V0, V1, V2, N1, d1);
    if (tt_side(N1, d1, U0, U1, U2, du0, du1, du2, du0du1, du0du2)) return 0;
    tt_plane(U0, U1, U2, N2, d2);
    if (tt_side(N2, d2, V0, V1, V2, dv0, dv1, dv2, dv0dv1, dv0dv2)) return 0;
    return tt_segment(V0, V1, V2, U0, U1, U2, N1, N2, du0, du1, du2, du0du1, du0du2, dv0, dv1, dv2, dv0dv1, dv0dv2, src, tgt);
}

// IMR/src/CollisionDetection/CollisionDetection.cpp:9-13
IMR_HD void sweep_axes(V3& U, V3& V, V3& W) {
    U = normalize3(mk3(0.8f, -0.2f, 0.f));
    W = normalize3(cross3(U, mk3(0.f, -1.f, 0.f)));
    V = normalize3(cross3(W, U));
}

// float -> uint32 with the same total order (for radix sort)
IMR_HD uint32_t float_orderable(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } cv; cv.f = f; uint32_t u = cv.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// ---- normals of the "uncollide" rays (CreateUncollideRays.cpp:150-163, IMR/src/Geometry/Triangle.cpp:148-212) ---------------
// glm::adjointTranspose(mat3(m)) (glm/gtc/matrix_inverse.inl:133-147; Triangle.cpp:175-178), column-major: a[3 * c + r]
struct M3 { float a[9]; };
IMR_HD M3 m3_identity() { M3 o; o.a[0] = 1.f; o.a[1] = 0.f; o.a[2] = 0.f; o.a[3] = 0.f; o.a[4] = 1.f; o.a[5] = 0.f; o.a[6] = 0.f; o.a[7] = 0.f; o.a[8] = 1.f; return o; }
IMR_HD M3 adjoint_transpose3(const Rel& m) {
    // A(c, r) = m[c][r]; rows of Rel: r0 = (m[0][0], m[1][0], m[2][0], .), r1 = (m[0][1], ...), r2 = (m[0][2], ...)
    const float a00 = m.r0.x, a01 = m.r1.x, a02 = m.r2.x, a10 = m.r0.y, a11 = m.r1.y, a12 = m.r2.y, a20 = m.r0.z, a21 = m.r1.z, a22 = m.r2.z;
    M3 o;
    o.a[0] = +(a11 * a22 - a21 * a12);
    o.a[1] = -(a10 * a22 - a20 * a12);
    o.a[2] = +(a10 * a21 - a20 * a11);
    o.a[3] = -(a01 * a22 - a21 * a02);
    o.a[4] = +(a00 * a22 - a20 * a02);
    o.a[5] = -(a00 * a21 - a20 * a01);
    o.a[6] = +(a01 * a12 - a11 * a02);
    o.a[7] = -(a00 * a12 - a10 * a02);
    o.a[8] = +(a00 * a11 - a10 * a01);
    return o;
}
// type_mat3x3.inl:468-474 : (m[0] * v.x + m[1] * v.y) + m[2] * v.z
IMR_HD V3 m3_mul(const M3& m, V3 v) {
    return mk3((m.a[0] * v.x + m.a[3] * v.y) + m.a[6] * v.z, (m.a[1] * v.x + m.a[4] * v.y) + m.a[7] * v.z, (m.a[2] * v.x + m.a[5] * v.y) + m.a[8] * v.z);
}
// TrianglePosition::GetBarycentricOfPoint, Triangle.cpp:148-163
IMR_HD void tri_barycentric(V3 p0, V3 p1, V3 p2, V3 point, float& bx, float& by) {
    const V3 v0 = sub3(p1, p0), v1 = sub3(p2, p0), v2 = sub3(point, p0);
    const float d00 = dot3(v0, v0), d01 = dot3(v0, v1), d11 = dot3(v1, v1), d20 = dot3(v2, v0), d21 = dot3(v2, v1);
    const float denom = d00 * d11 - d01 * d01;
    bx = (d11 * d20 - d01 * d21) / denom;
    by = (d00 * d21 - d01 * d20) / denom;
}
// TriangleNormal::GetNormal(baryCoords[, corrected_matrix]), Triangle.cpp:197-212 (before the normalisation)
IMR_HD V3 tri_interp_normal(V3 n0, V3 n1, V3 n2, float bx, float by) {
    const float w0 = (1.f - bx) - by;
    return add3(add3(scale3(n0, w0), scale3(n1, bx)), scale3(n2, by));
}
