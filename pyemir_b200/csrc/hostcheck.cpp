// Host build of geom.cuh: lets the CPU test-suite exercise the exact functions the
// plan-building and streaming kernels use (polynomial map, polygon clipping,
// Steffen slopes), against the oracle, without a GPU.  Test aid only — the
// product library (libemirwv.so) does not contain or call this file.
#include <stdint.h>

#include "geom.cuh"

extern "C" {

void hostcheck_fmap(int order, const double *a, const double *b, const double *x,
                    const double *y, int64_t n, double *u, double *v) {
    for (int64_t i = 0; i < n; ++i) emirwv::fmap(order, a, b, x[i], y[i], u[i], v[i]);
}

double hostcheck_quad_box_area(const double *qx, const double *qy, double x0, double x1,
                               double y0, double y1) {
    return emirwv::quad_box_area(qx, qy, x0, x1, y0, y1);
}

void hostcheck_quad_extent(const double *q, int *lo, int *hi) {
    emirwv::quad_extent(q, *lo, *hi);
}

double hostcheck_steffen_slope(double sl, double sr, double r) {
    return emirwv::steffen_slope<double>(sl, sr, r);
}

float hostcheck_steffen_slope_f(float sl, float sr, float r) {
    return emirwv::steffen_slope<float>(sl, sr, r);
}

double hostcheck_steffen_partial(double s, double yp0, double yp1, double t, double tau) {
    return emirwv::steffen_partial<double>(s, yp0, yp1, t, tau);
}

int hostcheck_order_from_ncoef(int ncoef) { return emirwv::order_from_ncoef(ncoef); }

}  // extern "C"

extern "C" double hostcheck_steffen_partial_flux(double rect, double ya, double yb, double tau) {
    return emirwv::steffen_partial_flux<double>(rect, ya, yb, tau);
}
