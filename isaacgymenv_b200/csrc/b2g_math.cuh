// Small fixed-size linear algebra for the articulated-body kernels: 3-vectors, 3x3 matrices,
// spatial (6-D) motion/force vectors and symmetric 6x6 articulated inertias stored as 21 floats.
// Everything is scalar float code meant to live in registers (static indexing only).
//
// The same header compiles for the device (nvcc, sm_100a) and, with B2G_HOST_EMU defined, for the
// host lane emulator used by the CPU tests (tests/emu) -- the product library never builds that mode.
#pragma once

#include <math.h>

#if defined(B2G_HOST_EMU)
#define B2G_HD
#define B2G_INL inline
#else
#define B2G_HD __device__
#define B2G_INL __forceinline__
#endif

namespace b2g {

struct V3 {
    float x, y, z;
};
B2G_HD B2G_INL V3 v3(float x, float y, float z) { return V3{x, y, z}; }
B2G_HD B2G_INL V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
B2G_HD B2G_INL V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
B2G_HD B2G_INL V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
B2G_HD B2G_INL V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
B2G_HD B2G_INL V3 operator*(float s, V3 a) { return V3{a.x * s, a.y * s, a.z * s}; }
B2G_HD B2G_INL V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
B2G_HD B2G_INL V3& operator-=(V3& a, V3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
B2G_HD B2G_INL float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
B2G_HD B2G_INL V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// 3x3 row-major
struct M3 {
    float m[9];
};
B2G_HD B2G_INL V3 mul(const M3& a, V3 v) {
    return V3{a.m[0] * v.x + a.m[1] * v.y + a.m[2] * v.z, a.m[3] * v.x + a.m[4] * v.y + a.m[5] * v.z,
              a.m[6] * v.x + a.m[7] * v.y + a.m[8] * v.z};
}
B2G_HD B2G_INL V3 mulT(const M3& a, V3 v) {
    return V3{a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z,
              a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z};
}
B2G_HD B2G_INL M3 mul(const M3& a, const M3& b) {
    M3 o;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) o.m[i * 3 + j] = a.m[i * 3] * b.m[j] + a.m[i * 3 + 1] * b.m[3 + j] + a.m[i * 3 + 2] * b.m[6 + j];
    return o;
}
B2G_HD B2G_INL M3 quat_to_m3(float x, float y, float z, float w) {
    M3 o;
    o.m[0] = 1 - 2 * (y * y + z * z); o.m[1] = 2 * (x * y - z * w); o.m[2] = 2 * (x * z + y * w);
    o.m[3] = 2 * (x * y + z * w); o.m[4] = 1 - 2 * (x * x + z * z); o.m[5] = 2 * (y * z - x * w);
    o.m[6] = 2 * (x * z - y * w); o.m[7] = 2 * (y * z + x * w); o.m[8] = 1 - 2 * (x * x + y * y);
    return o;
}
B2G_HD B2G_INL M3 axis_angle_m3(V3 a, float ang) {
    float s, c;
#if defined(B2G_HOST_EMU)
    s = sinf(ang); c = cosf(ang);
#else
    sincosf(ang, &s, &c);
#endif
    float t = 1 - c;
    M3 o;
    o.m[0] = t * a.x * a.x + c; o.m[1] = t * a.x * a.y - s * a.z; o.m[2] = t * a.x * a.z + s * a.y;
    o.m[3] = t * a.x * a.y + s * a.z; o.m[4] = t * a.y * a.y + c; o.m[5] = t * a.y * a.z - s * a.x;
    o.m[6] = t * a.x * a.z - s * a.y; o.m[7] = t * a.y * a.z + s * a.x; o.m[8] = t * a.z * a.z + c;
    return o;
}

// symmetric 3x3: xx, xy, xz, yy, yz, zz
struct S3 {
    float xx, xy, xz, yy, yz, zz;
};
B2G_HD B2G_INL V3 mul(const S3& a, V3 v) {
    return V3{a.xx * v.x + a.xy * v.y + a.xz * v.z, a.xy * v.x + a.yy * v.y + a.yz * v.z, a.xz * v.x + a.yz * v.y + a.zz * v.z};
}
B2G_HD B2G_INL S3& operator+=(S3& a, const S3& b) {
    a.xx += b.xx; a.xy += b.xy; a.xz += b.xz; a.yy += b.yy; a.yz += b.yz; a.zz += b.zz;
    return a;
}
// R * diag/sym(I) * R^T for a symmetric I
B2G_HD B2G_INL S3 rotate_sym(const M3& r, const S3& i) {
    // t = R * I
    float t[9];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float r0 = r.m[a * 3], r1 = r.m[a * 3 + 1], r2 = r.m[a * 3 + 2];
        t[a * 3 + 0] = r0 * i.xx + r1 * i.xy + r2 * i.xz;
        t[a * 3 + 1] = r0 * i.xy + r1 * i.yy + r2 * i.yz;
        t[a * 3 + 2] = r0 * i.xz + r1 * i.yz + r2 * i.zz;
    }
    S3 o;
    o.xx = t[0] * r.m[0] + t[1] * r.m[1] + t[2] * r.m[2];
    o.xy = t[0] * r.m[3] + t[1] * r.m[4] + t[2] * r.m[5];
    o.xz = t[0] * r.m[6] + t[1] * r.m[7] + t[2] * r.m[8];
    o.yy = t[3] * r.m[3] + t[4] * r.m[4] + t[5] * r.m[5];
    o.yz = t[3] * r.m[6] + t[4] * r.m[7] + t[5] * r.m[8];
    o.zz = t[6] * r.m[6] + t[7] * r.m[7] + t[8] * r.m[8];
    return o;
}

// spatial vector: w = angular (motion) / moment (force), v = linear (motion, of the point at O) / force
struct SV {
    V3 w, v;
};
B2G_HD B2G_INL SV sv0() { return SV{V3{0, 0, 0}, V3{0, 0, 0}}; }
B2G_HD B2G_INL SV operator+(SV a, SV b) { return SV{a.w + b.w, a.v + b.v}; }
B2G_HD B2G_INL SV operator-(SV a, SV b) { return SV{a.w - b.w, a.v - b.v}; }
B2G_HD B2G_INL SV operator-(SV a) { return SV{-a.w, -a.v}; }
B2G_HD B2G_INL SV operator*(SV a, float s) { return SV{a.w * s, a.v * s}; }
B2G_HD B2G_INL SV& operator+=(SV& a, SV b) { a.w += b.w; a.v += b.v; return a; }
B2G_HD B2G_INL float dot(SV a, SV b) { return dot(a.w, b.w) + dot(a.v, b.v); }
// motion cross product a x b
B2G_HD B2G_INL SV crm(SV a, SV b) { return SV{cross(a.w, b.w), cross(a.w, b.v) + cross(a.v, b.w)}; }
// force cross product a x* f
B2G_HD B2G_INL SV crf(SV a, SV f) { return SV{cross(a.w, f.w) + cross(a.v, f.v), cross(a.w, f.v)}; }

// symmetric 6x6 [[A, B], [B^T, C]] with A, C symmetric, B general
struct SI {
    S3 A;
    M3 B;
    S3 C;
};
B2G_HD B2G_INL SV mul(const SI& i, SV s) { return SV{mul(i.A, s.w) + mul(i.B, s.v), mulT(i.B, s.w) + mul(i.C, s.v)}; }
B2G_HD B2G_INL SI& operator+=(SI& a, const SI& b) {
    a.A += b.A; a.C += b.C;
#pragma unroll
    for (int k = 0; k < 9; k++) a.B.m[k] += b.B.m[k];
    return a;
}
// i - u u^T * s
B2G_HD B2G_INL SI rank1_sub(const SI& i, SV u, float s) {
    SI o = i;
    V3 a = u.w * s, c = u.v * s;
    o.A.xx -= a.x * u.w.x; o.A.xy -= a.x * u.w.y; o.A.xz -= a.x * u.w.z; o.A.yy -= a.y * u.w.y; o.A.yz -= a.y * u.w.z; o.A.zz -= a.z * u.w.z;
    o.C.xx -= c.x * u.v.x; o.C.xy -= c.x * u.v.y; o.C.xz -= c.x * u.v.z; o.C.yy -= c.y * u.v.y; o.C.yz -= c.y * u.v.z; o.C.zz -= c.z * u.v.z;
    o.B.m[0] -= a.x * u.v.x; o.B.m[1] -= a.x * u.v.y; o.B.m[2] -= a.x * u.v.z;
    o.B.m[3] -= a.y * u.v.x; o.B.m[4] -= a.y * u.v.y; o.B.m[5] -= a.y * u.v.z;
    o.B.m[6] -= a.z * u.v.x; o.B.m[7] -= a.z * u.v.y; o.B.m[8] -= a.z * u.v.z;
    return o;
}
// rigid-body spatial inertia about O: mass m, com position c (relative to O), rotational inertia about the com iw (world axes)
B2G_HD B2G_INL SI rigid_inertia(float m, V3 c, const S3& iw) {
    SI o;
    float cc = dot(c, c);
    o.A.xx = iw.xx + m * (cc - c.x * c.x); o.A.xy = iw.xy - m * c.x * c.y; o.A.xz = iw.xz - m * c.x * c.z;
    o.A.yy = iw.yy + m * (cc - c.y * c.y); o.A.yz = iw.yz - m * c.y * c.z; o.A.zz = iw.zz + m * (cc - c.z * c.z);
    V3 h = c * m;
    o.B.m[0] = 0; o.B.m[1] = -h.z; o.B.m[2] = h.y;
    o.B.m[3] = h.z; o.B.m[4] = 0; o.B.m[5] = -h.x;
    o.B.m[6] = -h.y; o.B.m[7] = h.x; o.B.m[8] = 0;
    o.C.xx = m; o.C.xy = 0; o.C.xz = 0; o.C.yy = m; o.C.yz = 0; o.C.zz = m;
    return o;
}

// packed symmetric 6x6, lower triangle row-major: index(i,j) = i*(i+1)/2 + j  (j <= i)
struct P6 {
    float a[21];
};
B2G_HD B2G_INL constexpr int tri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
B2G_HD B2G_INL P6 pack6(const SI& s) {
    P6 p;
    p.a[tri(0, 0)] = s.A.xx; p.a[tri(1, 0)] = s.A.xy; p.a[tri(2, 0)] = s.A.xz; p.a[tri(1, 1)] = s.A.yy; p.a[tri(2, 1)] = s.A.yz; p.a[tri(2, 2)] = s.A.zz;
    p.a[tri(3, 3)] = s.C.xx; p.a[tri(4, 3)] = s.C.xy; p.a[tri(5, 3)] = s.C.xz; p.a[tri(4, 4)] = s.C.yy; p.a[tri(5, 4)] = s.C.yz; p.a[tri(5, 5)] = s.C.zz;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) p.a[tri(3 + j, i)] = s.B.m[i * 3 + j];   // row i (angular) x col 3+j (linear)
    return p;
}
B2G_HD B2G_INL SV mul(const P6& p, SV s) {
    float x[6] = {s.w.x, s.w.y, s.w.z, s.v.x, s.v.y, s.v.z}, o[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        float acc = 0;
#pragma unroll
        for (int j = 0; j < 6; j++) acc += p.a[tri(i, j)] * x[j];
        o[i] = acc;
    }
    return SV{V3{o[0], o[1], o[2]}, V3{o[3], o[4], o[5]}};
}
// inverse of a symmetric positive definite 6x6 (Cholesky L L^T, then L^-1, then L^-T L^-1). ok=false if not SPD.
B2G_HD B2G_INL P6 spd_inverse6(const P6& a, bool& ok) {
    float L[21];
    ok = true;
#pragma unroll
    for (int j = 0; j < 6; j++) {
        float d = a.a[tri(j, j)];
#pragma unroll
        for (int k = 0; k < j; k++) d -= L[tri(j, k)] * L[tri(j, k)];
        if (!(d > 0.0f)) { ok = false; d = 1.0f; }
        float s = sqrtf(d), is = 1.0f / s;
        L[tri(j, j)] = is;   // store the reciprocal of the diagonal
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            float t = a.a[tri(i, j)];
#pragma unroll
            for (int k = 0; k < j; k++) t -= L[tri(i, k)] * L[tri(j, k)];
            L[tri(i, j)] = t * is;
        }
    }
    // W = L^-1 (lower triangular), diagonal of L holds reciprocals already
    float W[21];
#pragma unroll
    for (int c = 0; c < 6; c++) {
        W[tri(c, c)] = L[tri(c, c)];
#pragma unroll
        for (int i = c + 1; i < 6; i++) {
            float t = 0;
#pragma unroll
            for (int k = c; k < i; k++) t -= L[tri(i, k)] * W[tri(k, c)];
            W[tri(i, c)] = t * L[tri(i, i)];
        }
    }
    P6 inv;
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            float t = 0;
#pragma unroll
            for (int k = i; k < 6; k++) t += W[tri(k, i)] * W[tri(k, j)];
            inv.a[tri(i, j)] = t;
        }
    return inv;
}

}  // namespace b2g
