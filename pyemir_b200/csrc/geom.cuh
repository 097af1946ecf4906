// Geometry shared by the plan-building kernels and by the host-side self check.
//
// Everything here is float64 and written with explicit, individually rounded
// multiplies and adds (no fused multiply-add) so that the 2-D polynomial map gives
// bit-identical results to numpy's evaluation order in the reference's
// numina.array.distortion.fmap (called from slitlet2d.py:421-423): that is what
// makes the nearest-neighbour index maps bit-exact.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define EMIRWV_HD __host__ __device__ __forceinline__
#else
#define EMIRWV_HD inline
#endif

namespace emirwv {

EMIRWV_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;   // host build also uses -ffp-contract=off
    return r;
#endif
}

EMIRWV_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}

EMIRWV_HD double sub_rn(double a, double b) { return add_rn(a, -b); }

EMIRWV_HD double div_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    volatile double r = a / b;
    return r;
#endif
}

// order from the number of coefficients, -1 when ncoef is not triangular
EMIRWV_HD int order_from_ncoef(int ncoef) {
    for (int order = 1; order <= 5; ++order)
        if ((order + 1) * (order + 2) / 2 == ncoef) return order;
    return -1;
}

// (u, v) = sum_k (a_k, b_k) x^(i-j) y^j, k running over i = 0..order, j = 0..i.
// Each term is (c * x^p) * y^q added to the accumulator, one rounding per operation,
// powers by repeated multiplication (exact for the integer / half-integer pixel
// coordinates of a 2048-wide frame up to the 4th power).
EMIRWV_HD void fmap(int order, const double *a, const double *b, double x, double y,
                    double &u, double &v) {
    double xp[6], yp[6];
    xp[0] = 1.0;
    yp[0] = 1.0;
    for (int p = 1; p <= order; ++p) {
        xp[p] = mul_rn(xp[p - 1], x);
        yp[p] = mul_rn(yp[p - 1], y);
    }
    u = 0.0;
    v = 0.0;
    int k = 0;
    for (int i = 0; i <= order; ++i) {
        for (int j = 0; j <= i; ++j) {
            u = add_rn(u, mul_rn(mul_rn(a[k], xp[i - j]), yp[j]));
            v = add_rn(v, mul_rn(mul_rn(b[k], xp[i - j]), yp[j]));
            ++k;
        }
    }
}

// ---------------------------------------------------------------------------------
// Area of (convex or mildly concave) polygon ∩ axis-aligned box.
// Successive clipping against the four half-planes of the box; vertices are kept in
// two fixed arrays (a quad clipped four times has at most 8 vertices).
// ---------------------------------------------------------------------------------
struct Poly {
    double x[10];
    double y[10];
    int n;
};

// keep the part of `in` where sgn * (coord - bound) >= 0; axis 0 = x, 1 = y.
// Individually rounded operations in the reference's order (absolute image
// coordinates, no fused multiply-add): the float64 areas then agree with the CPU
// restatement of numina's clipping to the last bit instead of to ~ulp(2048).
EMIRWV_HD void clip_edge(const Poly &in, int axis, double bound, double sgn, Poly &out) {
    out.n = 0;
    for (int i = 0; i < in.n; ++i) {
        const int j = (i + 1 == in.n) ? 0 : i + 1;
        const double ci = axis == 0 ? in.x[i] : in.y[i];
        const double cj = axis == 0 ? in.x[j] : in.y[j];
        const double di = mul_rn(sgn, sub_rn(ci, bound));
        const double dj = mul_rn(sgn, sub_rn(cj, bound));
        const bool keep_i = di >= 0.0;
        const bool keep_j = dj >= 0.0;
        if (keep_i) {
            out.x[out.n] = in.x[i];
            out.y[out.n] = in.y[i];
            ++out.n;
        }
        if (keep_i != keep_j) {
            const double f = div_rn(di, sub_rn(di, dj));
            double nx = add_rn(in.x[i], mul_rn(f, sub_rn(in.x[j], in.x[i])));
            double ny = add_rn(in.y[i], mul_rn(f, sub_rn(in.y[j], in.y[i])));
            if (axis == 0) nx = bound; else ny = bound;
            out.x[out.n] = nx;
            out.y[out.n] = ny;
            ++out.n;
        }
    }
}

EMIRWV_HD double quad_box_area(const double *qx, const double *qy, double x0, double x1,
                               double y0, double y1) {
    Poly p, q;
    p.n = 4;
    for (int i = 0; i < 4; ++i) {
        p.x[i] = qx[i];
        p.y[i] = qy[i];
    }
    clip_edge(p, 0, x0, 1.0, q);
    if (q.n < 3) return 0.0;
    clip_edge(q, 0, x1, -1.0, p);
    if (p.n < 3) return 0.0;
    clip_edge(p, 1, y0, 1.0, q);
    if (q.n < 3) return 0.0;
    clip_edge(q, 1, y1, -1.0, p);
    if (p.n < 3) return 0.0;
    double twice = 0.0;
    for (int i = 0; i < p.n; ++i) {
        const int j = (i + 1 == p.n) ? 0 : i + 1;
        twice = add_rn(twice, sub_rn(mul_rn(p.x[i], p.y[j]), mul_rn(p.x[j], p.y[i])));
    }
    return 0.5 * fabs(twice);
}

// Footprint of a mapped quad in source pixels: pixel c covers [c-.5, c+.5].
// lo = first pixel touched, hi = last pixel with a non-degenerate overlap.
EMIRWV_HD void quad_extent(const double *q, int &lo, int &hi) {
    double mn = q[0], mx = q[0];
    for (int i = 1; i < 4; ++i) {
        mn = fmin(mn, q[i]);
        mx = fmax(mx, q[i]);
    }
    lo = (int)floor(mn + 0.5);
    hi = (int)ceil(mx - 0.5);
    if (hi < lo) hi = lo;
}

// Steffen (1990) limited slope at an interior knot from the two adjacent secant
// slopes; r = h_right / (h_left + h_right) weights the left slope:
//   p = sl*r + sr*(1-r);  y' = (sign(sl)+sign(sr)) * min(|sl|, |sr|, |p|/2)
// Branch-free: when the signs differ the result is 0; when either slope is 0 the
// minimum is 0.  Positively homogeneous in (sl, sr).
template <typename T>
EMIRWV_HD T steffen_slope(T sl, T sr, T r) {
    const T p = sr + r * (sl - sr);
    const T m = fmin(fmin(fabs(sl), fabs(sr)), T(0.5) * fabs(p));
    const T twice = copysign(m + m, sr);
    const bool opposite = (sl < T(0)) != (sr < T(0));
    return opposite ? T(0) : twice;
}

// Cumulative flux inside one source interval, measured from its left knot:
// P(t) = S(x_i + t) - y_i for the cubic Hermite piece with end slopes yp0, yp1 and
// secant slope s; tau = t / h.  P(0) = 0, P(h) = s * h.
template <typename T>
EMIRWV_HD T steffen_partial(T s, T yp0, T yp1, T t, T tau) {
    const T A = yp0 + yp1 - T(2) * s;
    const T B = s - yp0 - A;
    return t * (yp0 + tau * (B + tau * A));
}

// Same piece in flux units, as the kernels use it: rect = s*h is the flux of the
// source pixel, ya = h*yp0 and yb = h*yp1 its end slopes scaled by its width; only
// the fractional position tau = t/h is needed.  P(0) = 0, P(1) = rect.
// (Steffen's limiter is positively homogeneous, so ya, yb come from steffen_slope
// applied to fluxes scaled by ratios of neighbouring pixel widths.)
template <typename T>
EMIRWV_HD T steffen_partial_flux(T rect, T ya, T yb, T tau) {
    const T A = ya + yb - T(2) * rect;
    const T B = rect - ya - A;
    return tau * (ya + tau * (B + tau * A));
}

}  // namespace emirwv
