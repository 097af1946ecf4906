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

// ---- per-pixel arithmetic of the kernels either side of the hot path (pixel_ops.cuh)
#include "pixel_ops.cuh"

namespace {

// what k_stack_median does for one pixel: samples into N registers, padding, network
template <typename T, int N>
double median_bucket(const T *samples, int n, int64_t stride) {
    T v[N];
    for (int u = 0; u < N; ++u)
        v[u] = u < n ? samples[(int64_t)u * stride] : emirwv::stack_pad<T>(u, n, N);
    return emirwv::median_of_registers<T, N>(v, n);
}

template <typename T>
int stack_median_host(const T *frames, int n, int64_t npix, double *out) {
    if (n < 1 || n > 64) return -1;
    for (int64_t i = 0; i < npix; ++i) {
        const T *s = frames + i;
        if (n <= 4) out[i] = median_bucket<T, 4>(s, n, npix);
        else if (n <= 8) out[i] = median_bucket<T, 8>(s, n, npix);
        else if (n <= 16) out[i] = median_bucket<T, 16>(s, n, npix);
        else if (n <= 32) out[i] = median_bucket<T, 32>(s, n, npix);
        else out[i] = median_bucket<T, 64>(s, n, npix);
    }
    return 0;
}

// what k_prereduce does for every pixel of a stack
template <typename TIn, typename TCal, typename A>
void prereduce_host(const TIn *in, float *out, int nframes, int64_t npix, const TCal *bias,
                    const TCal *dark, const TCal *flat) {
    const bool hb = bias != nullptr, hd = dark != nullptr, hf = flat != nullptr;
    for (int64_t i = 0; i < npix; ++i) {
        const A b = hb ? (A)bias[i] : A(0), d = hd ? (A)dark[i] : A(0);
        A f = hf ? (A)flat[i] : A(1);
        f = f <= A(0) ? A(1) : f;
        for (int n = 0; n < nframes; ++n)
            out[(int64_t)n * npix + i] =
                emirwv::prereduce_px<TIn, A>(in[(int64_t)n * npix + i], hb, hd, hf, b, d, f);
    }
}

}  // namespace

extern "C" {

// frames [n][npix], out [npix]; dtype 0 = float32, 1 = float64
int hostcheck_stack_median(const void *frames, int dtype, int n, int64_t npix, double *out) {
    if (dtype == 0) return stack_median_host(static_cast<const float *>(frames), n, npix, out);
    return stack_median_host(static_cast<const double *>(frames), n, npix, out);
}

// in_dtype 0 float32, 1 float64, 2 uint16; cal_dtype 0 float32, 1 float64 (NULL skips a stage)
int hostcheck_prereduce(const void *in, int in_dtype, float *out, int nframes, int64_t npix,
                        const void *bias, const void *dark, const void *flat, int cal_dtype) {
#define PRE(TIN, TCAL, A)                                                                   \
    prereduce_host<TIN, TCAL, A>(static_cast<const TIN *>(in), out, nframes, npix,          \
                                 static_cast<const TCAL *>(bias),                           \
                                 static_cast<const TCAL *>(dark), static_cast<const TCAL *>(flat))
    if (in_dtype == 0 && cal_dtype == 0) PRE(float, float, float);
    else if (in_dtype == 0) PRE(float, double, double);
    else if (in_dtype == 1 && cal_dtype == 0) PRE(double, float, double);
    else if (in_dtype == 1) PRE(double, double, double);
    else if (in_dtype == 2 && cal_dtype == 0) PRE(uint16_t, float, float);
    else if (in_dtype == 2) PRE(uint16_t, double, double);
    else return -1;
#undef PRE
    return 0;
}

}  // extern "C"

// ---- sigma-clip combination, per-frame vertical shift, bad-pixel stage (pixel_ops.cuh)
namespace {

template <typename T, int N>
double sigmaclip_bucket(const T *samples, int n, int64_t stride, double low, double high) {
    T v[N];
    for (int u = 0; u < N; ++u) v[u] = u < n ? samples[(int64_t)u * stride] : T(0);
    return emirwv::sigmaclip_of_registers<T, N>(v, n, low, high);
}

template <typename T>
int sigmaclip_host(const T *frames, int n, int64_t npix, double low, double high, double *out) {
    if (n < 1 || n > 64) return -1;
    for (int64_t i = 0; i < npix; ++i) {
        const T *s = frames + i;
        if (n <= 4) out[i] = sigmaclip_bucket<T, 4>(s, n, npix, low, high);
        else if (n <= 8) out[i] = sigmaclip_bucket<T, 8>(s, n, npix, low, high);
        else if (n <= 16) out[i] = sigmaclip_bucket<T, 16>(s, n, npix, low, high);
        else if (n <= 32) out[i] = sigmaclip_bucket<T, 32>(s, n, npix, low, high);
        else out[i] = sigmaclip_bucket<T, 64>(s, n, npix, low, high);
    }
    return 0;
}

// what k_prereduce does for every pixel of a stack when a bad-pixel mask is given
template <typename TIn, typename TCal, typename A>
void prereduce_bpm_host(const TIn *in, float *out, int nframes, int naxis2, int naxis1,
                        const uint8_t *mask, const TCal *bias, const TCal *dark, const TCal *flat) {
    const bool hb = bias != nullptr, hd = dark != nullptr, hf = flat != nullptr;
    const int64_t npix = (int64_t)naxis1 * naxis2;
    for (int64_t i = 0; i < npix; ++i) {
        const A b = hb ? (A)bias[i] : A(0), d = hd ? (A)dark[i] : A(0);
        A f = hf ? (A)flat[i] : A(1);
        f = f <= A(0) ? A(1) : f;
        for (int n = 0; n < nframes; ++n) {
            const TIn *frame = in + (int64_t)n * npix;
            if (mask != nullptr && mask[i] != 0) {
                const double m = emirwv::bpm_window_median<TIn, 2>(
                    frame, mask, naxis1, naxis2, (int)(i % naxis1), (int)(i / naxis1), 0.0);
                out[(int64_t)n * npix + i] = emirwv::prereduce_from<TIn, A>(m, hb, hd, hf, b, d, f);
            } else {
                out[(int64_t)n * npix + i] = emirwv::prereduce_px<TIn, A>(frame[i], hb, hd, hf, b, d, f);
            }
        }
    }
}

}  // namespace

extern "C" {

int hostcheck_stack_sigmaclip(const void *frames, int dtype, int n, int64_t npix, double low,
                              double high, double *out) {
    if (dtype == 0) return sigmaclip_host(static_cast<const float *>(frames), n, npix, low, high, out);
    return sigmaclip_host(static_cast<const double *>(frames), n, npix, low, high, out);
}

// what k_shift_rows does for one frame (float64 in, float64 out)
void hostcheck_shift_rows(const double *img, int naxis2, int naxis1, double yoffset, double *out) {
    for (int y = 0; y < naxis2; ++y) {
        const emirwv::ShiftRow r = emirwv::shift_row(yoffset, y);
        for (int x = 0; x < naxis1; ++x) {
            const double xa = (double)x - 0.5, xb = (double)x + 0.5;
            double acc = 0.0;
            for (int ii = r.lo < 0 ? 0 : r.lo; ii <= (r.hi > naxis2 - 1 ? naxis2 - 1 : r.hi); ++ii) {
                const double area = emirwv::shift_area(xa, xb, r, ii);
                if (area != 0.0)
                    acc = emirwv::add_rn(acc, emirwv::mul_rn(area, img[(int64_t)ii * naxis1 + x]));
            }
            out[(int64_t)y * naxis1 + x] = acc;
        }
    }
}

int hostcheck_prereduce_bpm(const void *in, int in_dtype, float *out, int nframes, int naxis2,
                            int naxis1, const uint8_t *mask, const void *bias, const void *dark,
                            const void *flat, int cal_dtype) {
#define PRE(TIN, TCAL, A)                                                                       \
    prereduce_bpm_host<TIN, TCAL, A>(static_cast<const TIN *>(in), out, nframes, naxis2, naxis1, \
                                     mask, static_cast<const TCAL *>(bias),                     \
                                     static_cast<const TCAL *>(dark), static_cast<const TCAL *>(flat))
    if (in_dtype == 0 && cal_dtype == 0) PRE(float, float, float);
    else if (in_dtype == 0) PRE(float, double, double);
    else if (in_dtype == 1 && cal_dtype == 0) PRE(double, float, double);
    else if (in_dtype == 1) PRE(double, double, double);
    else if (in_dtype == 2 && cal_dtype == 0) PRE(uint16_t, float, float);
    else if (in_dtype == 2) PRE(uint16_t, double, double);
    else return -1;
#undef PRE
    return 0;
}

}  // extern "C"
