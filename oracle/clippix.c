/*
 * ORACLE — test infrastructure only (never linked or called by the product path).
 *
 * CPU restatement of the area-overlap resampler that the reference reaches through
 *   emirdrp/processing/wavecal/slitlet2d.py:421-423  rectify2d(..., resampling=2)
 * -> numina.array.distortion.rectify2d -> numina.array._clippix._resample (Cython).
 *
 * numina (pyproject.toml:32, "numina>=0.36.0", unpinned) is NOT present under
 * /root/reference nor installable here, so this follows the published algorithm
 * (SURVEY.md Appendix A.3): every output pixel is a quadrilateral in the source
 * image (its four corners mapped through the distortion polynomial); its value is
 *      sum over source pixels  area(quad ∩ pixel square) * image[pixel]
 * with no normalisation (flux preserving).  PARITY UNPINNED: the reference holds no
 * golden vector for this routine.
 *
 * Implementation: Sutherland–Hodgman clipping of the quad against the four
 * half-planes of each unit pixel square, shoelace area, float64 throughout.
 */
#include <math.h>
#include <stddef.h>

#define MAXV 16

/* clip polygon (n vertices) against the half-plane  s*(coord - bound) >= 0,
 * where coord is x (axis==0) or y (axis==1). Returns new vertex count. */
static int clip_halfplane(const double *px, const double *py, int n,
                          int axis, double bound, double s,
                          double *qx, double *qy)
{
    int m = 0;
    for (int i = 0; i < n; ++i) {
        int j = (i + 1 == n) ? 0 : i + 1;
        double ci = axis == 0 ? px[i] : py[i];
        double cj = axis == 0 ? px[j] : py[j];
        double di = s * (ci - bound);
        double dj = s * (cj - bound);
        int in_i = di >= 0.0;
        int in_j = dj >= 0.0;
        if (in_i) {
            qx[m] = px[i];
            qy[m] = py[i];
            ++m;
        }
        if (in_i != in_j) {
            double t = di / (di - dj);
            qx[m] = px[i] + t * (px[j] - px[i]);
            qy[m] = py[i] + t * (py[j] - py[i]);
            if (axis == 0) qx[m] = bound; else qy[m] = bound;
            ++m;
        }
    }
    return m;
}

/* area of the intersection of a polygon with the axis-aligned box */
double oracle_poly_box_area(const double *x, const double *y, int n,
                            double x0, double x1, double y0, double y1)
{
    double ax[MAXV], ay[MAXV], bx[MAXV], by[MAXV];
    int m = n;
    for (int i = 0; i < n; ++i) { ax[i] = x[i]; ay[i] = y[i]; }
    m = clip_halfplane(ax, ay, m, 0, x0, +1.0, bx, by); if (m < 3) return 0.0;
    m = clip_halfplane(bx, by, m, 0, x1, -1.0, ax, ay); if (m < 3) return 0.0;
    m = clip_halfplane(ax, ay, m, 1, y0, +1.0, bx, by); if (m < 3) return 0.0;
    m = clip_halfplane(bx, by, m, 1, y1, -1.0, ax, ay); if (m < 3) return 0.0;
    double twice = 0.0;
    for (int i = 0; i < m; ++i) {
        int j = (i + 1 == m) ? 0 : i + 1;
        twice += ax[i] * ay[j] - ax[j] * ay[i];
    }
    return 0.5 * fabs(twice);
}

/*
 * image2d      : (naxis2, naxis1) source image, C order
 * xxx, yyy     : 4*npix mapped corner coordinates (source array coordinates,
 *                pixel (i,j) covers [j-.5,j+.5] x [i-.5,i+.5])
 * ixx, iyy     : npix destination column/row of every quad
 * image2d_rect : (naxis2out, naxis1out) destination, accumulated into
 */
void oracle_resample(const double *image2d, long naxis2, long naxis1,
                     const double *xxx, const double *yyy, long npix,
                     const long *ixx, const long *iyy,
                     double *image2d_rect, long naxis2out, long naxis1out)
{
    for (long p = 0; p < npix; ++p) {
        const double *qx = xxx + 4 * p;
        const double *qy = yyy + 4 * p;
        long jo = ixx[p], io = iyy[p];
        if (jo < 0 || jo >= naxis1out || io < 0 || io >= naxis2out) continue;
        double xmin = qx[0], xmax = qx[0], ymin = qy[0], ymax = qy[0];
        for (int k = 1; k < 4; ++k) {
            if (qx[k] < xmin) xmin = qx[k];
            if (qx[k] > xmax) xmax = qx[k];
            if (qy[k] < ymin) ymin = qy[k];
            if (qy[k] > ymax) ymax = qy[k];
        }
        long jmin = (long)floor(xmin + 0.5), jmax = (long)floor(xmax + 0.5);
        long imin = (long)floor(ymin + 0.5), imax = (long)floor(ymax + 0.5);
        if (jmin < 0) jmin = 0;
        if (imin < 0) imin = 0;
        if (jmax > naxis1 - 1) jmax = naxis1 - 1;
        if (imax > naxis2 - 1) imax = naxis2 - 1;
        double acc = 0.0;
        for (long ii = imin; ii <= imax; ++ii) {
            for (long jj = jmin; jj <= jmax; ++jj) {
                double a = oracle_poly_box_area(qx, qy, 4,
                                                (double)jj - 0.5, (double)jj + 0.5,
                                                (double)ii - 0.5, (double)ii + 0.5);
                if (a != 0.0) acc += a * image2d[ii * naxis1 + jj];
            }
        }
        image2d_rect[io * naxis1out + jo] += acc;
    }
}

/* overlap areas of one quad with the (ny, nx) block of pixels whose first pixel is
 * (i0, j0); used by tests to check the product's weight tables one by one */
void oracle_quad_weights(const double *qx, const double *qy,
                         long i0, long j0, long ny, long nx, double *w)
{
    for (long a = 0; a < ny; ++a)
        for (long b = 0; b < nx; ++b)
            w[a * nx + b] = oracle_poly_box_area(
                qx, qy, 4,
                (double)(j0 + b) - 0.5, (double)(j0 + b) + 0.5,
                (double)(i0 + a) - 0.5, (double)(i0 + a) + 0.5);
}
