// f64 vector helpers and the exact (reference-arithmetic) primitive tests.
//
// Everything in this header is compiled with -fmad=false: each + - * / sqrt is one IEEE-754
// binary64 operation in the reference's source order, so the ray parameter t of an accepted hit
// is bit-identical to the Rust code's (sphere.rs:30-69 + math.rs:7-30, cuboid.rs:55-95,
// triangle.rs:179-251).  cgmath conventions: dot = (x*x' + y*y') + z*z', normalize = v * (1/|v|).
#pragma once
#include "lgb_types.cuh"

namespace lgb {

struct D3 { double x, y, z; };
__device__ __forceinline__ D3 d3(double x, double y, double z) { D3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 operator-(D3 a) { return d3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return d3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ D3 operator*(double s, D3 a) { return d3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ D3 operator/(D3 a, double s) { return d3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ D3 mul_el(D3 a, D3 b) { return d3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 cross(D3 a, D3 b) {
    return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ D3 normalize(D3 a) { return a * (1.0 / sqrt(dot(a, a))); }
__device__ __forceinline__ double comp(D3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
__device__ __forceinline__ D3 face_forward(D3 n, D3 v) { return dot(n, v) < 0.0 ? -n : n; }   // normal.rs:37-40
__device__ __forceinline__ D3 axis_vec(int i) { return d3(i == 0 ? 1.0 : 0.0, i == 1 ? 1.0 : 0.0, i == 2 ? 1.0 : 0.0); }

struct Ray64 { D3 o, d; };

// ---- fast f64 forms for code that only colours (never for a decision the reference takes): explicit FMAs (this file is
// compiled with -fmad=false) and MUFU-seeded reciprocal / reciprocal square root refined by one third-order step each
// (seed 2^-20 relative -> ~2^-60; no IEEE rounding fix-up, no slow path: 4 / 6 instructions instead of ~25 for 1.0 / x and
// ~45 for 1.0 / sqrt(x)).  Arguments must be finite, normal and > 0 -- callers guard zeros.
__device__ __forceinline__ double fdot(D3 a, D3 b) { return fma(a.z, b.z, fma(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ D3 fmadd(D3 a, double s, D3 b) { return d3(fma(a.x, s, b.x), fma(a.y, s, b.y), fma(a.z, s, b.z)); }   // a s + b
__device__ __forceinline__ D3 fcross(D3 a, D3 b) {
    return d3(fma(a.y, b.z, -(a.z * b.y)), fma(a.z, b.x, -(a.x * b.z)), fma(a.x, b.y, -(a.y * b.x)));
}
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    return fma(r, fma(e, e, e), r);
}
__device__ __forceinline__ double fast_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = fma(-(x * y), y, 1.0);
    return fma(y * h, fma(0.375, h, 0.5), y);
}
__device__ __forceinline__ double fast_sqrt(double x) { return x > 0.0 ? x * fast_rsqrt(x) : 0.0; }

// ---- sphere.rs:30-69, 79-86 with math.rs:7-30 inlined -----------------------------------
// The `t >= isect.t` rejection (sphere.rs:86, cuboid.rs:95, triangle.rs:251) is applied by the caller,
// together with the reference-order tie rule.
__device__ __forceinline__ bool sphere_exact(D3 c, double rad, const Ray64& ray, double& t_out, bool& inside) {
    D3 d = ray.d;
    D3 l = ray.o - c;
    double a = dot(d, d);
    double b = 2.0 * dot(d, l);
    double cc = dot(l, l) - rad * rad;
    double t;
    inside = false;
    if (a == 0.0) {
        if (b == 0.0) return false;
        t = -cc / b;
    } else {
        // Origin outside the sphere and moving away from it (cc > 0, b > 0): whatever the discriminant, q = -(b + sqrt) / 2 < 0, so
        // r0 = q / a < 0 and r1 = cc / q < 0, t = max(r0, r1) < 0 and the reference returns None.  Decided on the reference's own b and
        // cc, bit for bit; the bounds keep every quotient clear of underflow to -0.0 (which would pass `t < 0.0`).  This is the test every
        // shadow ray leaving a sphere runs against its own sphere: it now ends before the square root and the two divisions.
        if (cc > 1e-200 && b > 1e-100 && b < 1e100 && a < 1e100) return false;
        double disc = b * b - 4.0 * a * cc;
        if (disc < 0.0) return false;
        double sg = copysign(1.0, b);
        double q = -(b + sg * sqrt(disc)) / 2.0;
        double r0 = q / a;
        double r1 = (q == 0.0) ? r0 : cc / q;
        double t0 = fmin(r0, r1), t1 = fmax(r0, r1);
        if (t0 < 0.0) { t = t1; inside = true; } else { t = t0; }
    }
    if (t < 0.0) return false;
    t_out = t;
    return true;
}

// ---- cuboid.rs:55-95.  (u_axis, v_axis): dpdu = e_u, dpdv = e_v of the chosen face -------
__device__ __forceinline__ bool cuboid_exact(const double* mn, const double* mx, const Ray64& ray,
                                             double& t_out, int& u_axis, int& v_axis) {
    double tnear = -CUDART_INF, tfar = CUDART_INF;
    int nu = 1, nv = 2, fu = 1, fv = 2;          // CUBE_DIFFERENTIALS[0] = (y, z)
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double o = comp(ray.o, i), dinv = 1.0 / comp(ray.d, i);   // Ray::new, ray.rs:28-33
        double t1 = (mn[i] - o) * dinv;
        double t2 = (mx[i] - o) * dinv;
        int a1 = (i + 1) % 3, a2 = (i + 2) % 3;  // CUBE_DIFFERENTIALS[i] = (e_a1, e_a2)
        double tmin, tmax; int dp0, dp1;
        if (t1 < t2) { tmin = t1; tmax = t2; dp0 = a2; dp1 = a1; }
        else { tmin = t2; tmax = t1; dp0 = a1; dp1 = a2; }
        if (tmin > tnear) { nu = dp0; nv = dp1; }
        if (tmax < tfar) { fu = dp1; fv = dp0; }
        tnear = fmax(tnear, tmin);
        tfar = fmin(tfar, tmax);
    }
    if (tnear > tfar || tfar <= 0.0) return false;
    double t;
    if (tnear <= 0.0) { t = tfar; u_axis = fu; v_axis = fv; } else { t = tnear; u_axis = nu; v_axis = nv; }
    t_out = t;
    return true;
}

// ---- triangle.rs:179-251 ---------------------------------------------------------------
__device__ __forceinline__ int max_dimension_abs(D3 d) {                       // space/mod.rs:23-36
    double x = fabs(d.x), y = fabs(d.y), z = fabs(d.z);
    if (x > y) return x > z ? 0 : 2;
    return y > z ? 1 : 2;
}
__device__ __forceinline__ bool triangle_exact(D3 p0, D3 p1, D3 p2, const Ray64& ray, double& t_out,
                                               double& b0, double& b1, double& b2) {
    D3 p0t = p0 - ray.o, p1t = p1 - ray.o, p2t = p2 - ray.o;
    int kz = max_dimension_abs(ray.d);
    int kx = (kz + 1) % 3;
    int ky = (kx + 1) % 3;
    double dx = comp(ray.d, kx), dy = comp(ray.d, ky), dz = comp(ray.d, kz);
    p0t = d3(comp(p0t, kx), comp(p0t, ky), comp(p0t, kz));
    p1t = d3(comp(p1t, kx), comp(p1t, ky), comp(p1t, kz));
    p2t = d3(comp(p2t, kx), comp(p2t, ky), comp(p2t, kz));
    double sx = -dx / dz, sy = -dy / dz, sz = 1.0 / dz;
    p0t.x += sx * p0t.z; p0t.y += sy * p0t.z;
    p1t.x += sx * p1t.z; p1t.y += sy * p1t.z;
    p2t.x += sx * p2t.z; p2t.y += sy * p2t.z;
    double e0 = p1t.x * p2t.y - p1t.y * p2t.x;
    double e1 = p2t.x * p0t.y - p2t.y * p0t.x;
    double e2 = p0t.x * p1t.y - p0t.y * p1t.x;
    if ((e0 < 0.0 || e1 < 0.0 || e2 < 0.0) && (e0 > 0.0 || e1 > 0.0 || e2 > 0.0)) return false;
    double det = e0 + e1 + e2;
    if (det == 0.0) return false;
    p0t.z *= sz; p1t.z *= sz; p2t.z *= sz;
    double tscaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
    if ((det < 0.0 && tscaled >= 0.0) || (det > 0.0 && tscaled <= 0.0)) return false;
    double invdet = 1.0 / det;
    b0 = e0 * invdet; b1 = e1 * invdet; b2 = e2 * invdet;
    double t = tscaled * invdet;
    t_out = t;
    return true;
}

__device__ __forceinline__ void coordinate_system(D3 v1, D3& v2, D3& v3) {     // space/mod.rs:39-47
    if (fabs(v1.x) > fabs(v1.y)) v2 = d3(-v1.z, 0.0, v1.x) / sqrt(v1.x * v1.x + v1.z * v1.z);
    else v2 = d3(0.0, v1.z, -v1.y) / sqrt(v1.y * v1.y + v1.z * v1.z);
    v3 = cross(v1, v2);
}

}  // namespace lgb
