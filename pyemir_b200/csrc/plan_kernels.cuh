// Plan kernels (once per RectWaveCoeff): extent of the footprints, per-pixel table records,
// rectify2d on an explicit map, and the small helpers of the parity hooks.
// Included by emirwv.cu inside its anonymous namespace (one translation unit).

// ---------------------------------------------------------------- plan kernels
// Largest footprint (in source pixels) of any mapped output pixel, mode 2.
__global__ void k_extent(GeomParams p) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int s = blockIdx.z;
    const SlitDev &sd = p.slit[s];
    int need = 1;
    if (sd.valid && j < sd.bbw) {
        double qu[4], qv[4];
        map_corners(sd, j, sd.min_row + y, qu, qv);
        int lo, hi;
        quad_extent(qu, lo, hi);
        need = hi - lo + 1;
        quad_extent(qv, lo, hi);
        need = max(need, hi - lo + 1);
    }
    need = __reduce_max_sync(0xffffffffu, need);
    if ((threadIdx.x & 31) == 0) atomicMax(p.kmax, need);
}

// base[s][y][j] (frame coordinates of tap (0,0)), the full records w[s][y][j][RS], the 2 x 2
// block weights quad[s][y][j][4] and the packed meta[s][y][j].
template <typename T>
__global__ void k_build_geometry(GeomParams p, T *w, T *quad) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int s = blockIdx.z;
    const SlitDev &sd = p.slit[s];
    if (j >= W) return;
    const int K = p.K, KK = K * K;
    const size_t rowtab = (size_t)s * NROWS + y;
    const int RS = rec_size(K);
    T *wp = w + (rowtab * W + j) * RS;
    T *qp = quad + (rowtab * W + j) * 4;
    if (!sd.valid || j >= sd.bbw) {
        p.base[rowtab * W + j] = pack_base(0, 0);
        for (int tap = 0; tap < RS; ++tap) wp[tap] = T(0);
        *reinterpret_cast<int32_t *>(wp + KK) = pack_base(0, 0);
        for (int tap = 0; tap < 4; ++tap) qp[tap] = T(0);
        p.meta[rowtab * W + j] = META_COMPACT;
        p.blockrow[rowtab * W + j] = 16;
        return;
    }
    const int row = sd.min_row + y;
    double wt[MAXK * MAXK];
    int br, bc;
    if (p.mode == EMIRWV_AREA) {
        double qu[4], qv[4];
        map_corners(sd, j, row, qu, qv);
        int hr, hc;
        quad_extent(qu, bc, hc);
        quad_extent(qv, br, hr);
        // absolute bounding-box coordinates, like the reference's clipping
        for (int a = 0; a < K; ++a)
            for (int b = 0; b < K; ++b) {
                double area = 0.0;
                if (br + a <= hr && bc + b <= hc)
                    area = quad_box_area(qu, qv, (double)(bc + b) - 0.5, (double)(bc + b) + 0.5,
                                         (double)(br + a) - 0.5, (double)(br + a) + 0.5);
                wt[a * K + b] = area;
            }
    } else {
        double u, v;
        fmap(sd.order, sd.a, sd.b, (double)j, (double)row, u, v);
        u = clampd(u);
        v = clampd(v);
        if (p.mode == EMIRWV_NEAREST) {
            bc = (int)rint(u);     // round half to even, like numpy.rint
            br = (int)rint(v);
            wt[0] = 1.0;
        } else {                   // bilinear extension
            const double fu0 = floor(u), fv0 = floor(v);
            bc = (int)fu0;
            br = (int)fv0;
            const double fu = u - fu0, fv = v - fv0;
            wt[0] = (1.0 - fv) * (1.0 - fu);
            wt[1] = (1.0 - fv) * fu;
            wt[2] = fv * (1.0 - fu);
            wt[3] = fv * fu;
        }
    }
    // frame coordinates: bounding-box crop (slitlet2d.py:366-369) and the global
    // integer offsets (apply_rectwv_coeff.py:74-78) folded into the address
    const int fr0 = br + sd.ns1 - 1 - p.offy;
    const int fc0 = bc + sd.nc1 - 1 - p.offx;
    for (int a = 0; a < K; ++a)
        for (int b = 0; b < K; ++b) {
            const int r = br + a, c = bc + b;
            const int fr = fr0 + a, fc = fc0 + b;
            const bool inside = r >= 0 && r < sd.bbh && c >= 0 && c < sd.bbw && fr >= 0 &&
                                fr < p.naxis2 && fc >= 0 && fc < p.naxis1;
            wp[a * K + b] = inside ? (T)wt[a * K + b] : T(0);
        }
    for (int tap = KK; tap < RS; ++tap) wp[tap] = T(0);
    // the 2 x 2 block of K <= 2 is the footprint itself (K = 1: one weight)
    T q4[4] = {T(0), T(0), T(0), T(0)};
    bool compact = K <= 2;
    if (K == 1) q4[0] = wp[0];
    if (K == 2) q4[0] = wp[0], q4[1] = wp[1], q4[2] = wp[2], q4[3] = wp[3];
    // Compact form (K >= 3): almost every footprint fits a 2 x 2 block of source pixels.
    // Such a record holds the block's four weights in words 0..3 (zeros elsewhere) and the
    // coordinates of the block; the hot kernel then gathers 4 taps instead of K*K, with the
    // same result (the skipped taps have weight zero).
    int a0 = 0, b0 = 0;
    uint32_t flags = 0;
    if (K >= 3) {
        int alo = K, ahi = -1, blo = K, bhi = -1;
        for (int a = 0; a < K; ++a)
            for (int b = 0; b < K; ++b)
                if (wp[a * K + b] != T(0)) {
                    alo = min(alo, a), ahi = max(ahi, a);
                    blo = min(blo, b), bhi = max(bhi, b);
                }
        if (ahi < 0) alo = ahi = blo = bhi = 0;                // no contribution at all
        if (ahi - alo <= 1 && bhi - blo <= 1) {
            a0 = min(alo, K - 2);
            b0 = min(blo, K - 2);
            const T c00 = wp[a0 * K + b0], c01 = wp[a0 * K + b0 + 1];
            const T c10 = wp[(a0 + 1) * K + b0], c11 = wp[(a0 + 1) * K + b0 + 1];
            for (int tap = 0; tap < KK; ++tap) wp[tap] = T(0);
            wp[0] = c00, wp[1] = c01, wp[2] = c10, wp[3] = c11;
            flags = REC_COMPACT | ((uint32_t)a0 << 1) | ((uint32_t)b0 << 3);
            q4[0] = c00, q4[1] = c01, q4[2] = c10, q4[3] = c11;
            compact = true;
        }
    }
    for (int tap = 0; tap < 4; ++tap) qp[tap] = q4[tap];
    p.meta[rowtab * W + j] =
        ((uint32_t)pack_base(0, fc0 + b0) & 0xffffu) | (ring_slot(fr0 + a0) << 16) |
        (ring_slot(fr0 + a0 + 1) << 20) | (compact ? META_COMPACT : 0u);
    p.blockrow[rowtab * W + j] = (uint8_t)(a0 | (compact ? 16 : 0));
    *reinterpret_cast<int32_t *>(wp + KK) = pack_base(fr0 + a0, fc0 + b0);
    // ring slot of each of the K source rows in the hot kernel's rolling window
    uint32_t slots = flags << 24;
    for (int a = 0; a < K && a < 3; ++a)
        slots |= ring_slot(fr0 + a0 + a) << (8 * a);
    *reinterpret_cast<uint32_t *>(wp + KK + 1) = slots;
    if (K == 4)          // the fourth row of a 4 x 4 footprint has its own word
        *reinterpret_cast<uint32_t *>(wp + KK + 2) = ring_slot(fr0 + a0 + 3);
    p.base[rowtab * W + j] = pack_base(fr0, fc0);
}

// General rectify2d on an arbitrary image (numina.array.distortion.rectify2d as called
// by Slitlet2D.rectify, slitlet2d.py:421-423, direct or inverse map): one thread per
// output pixel, geometry evaluated on the fly, float64, accumulation in the reference's
// order.  Not the streaming path: it serves the per-slitlet API of the reference.
__global__ void k_rectify2d(const double *img, int naxis2, int naxis1, SlitDev sd, int mode,
                            double *out, int naxis2out, int naxis1out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= naxis1out || i >= naxis2out) return;
    double acc = 0.0;
    if (mode == EMIRWV_AREA) {
        double qu[4], qv[4];
        map_corners(sd, j, i, qu, qv);
        int c0, c1, r0, r1;
        quad_extent(qu, c0, c1);
        quad_extent(qv, r0, r1);
        for (int ii = max(r0, 0); ii <= min(r1, naxis2 - 1); ++ii)
            for (int jj = max(c0, 0); jj <= min(c1, naxis1 - 1); ++jj) {
                const double area = quad_box_area(qu, qv, (double)jj - 0.5, (double)jj + 0.5,
                                                  (double)ii - 0.5, (double)ii + 0.5);
                if (area != 0.0) acc = add_rn(acc, mul_rn(area, img[(size_t)ii * naxis1 + jj]));
            }
    } else {
        double u, v;
        fmap(sd.order, sd.a, sd.b, (double)j, (double)i, u, v);
        u = clampd(u);
        v = clampd(v);
        if (mode == EMIRWV_NEAREST) {
            const int ju = (int)rint(u), iv = (int)rint(v);
            if (ju >= 0 && ju < naxis1 && iv >= 0 && iv < naxis2) acc = img[(size_t)iv * naxis1 + ju];
        } else {
            const double u0 = floor(u), v0 = floor(v);
            const double fu = u - u0, fv = v - v0;
            const double wv[2] = {1.0 - fv, fv}, wu[2] = {1.0 - fu, fu};
            for (int dv = 0; dv < 2; ++dv)
                for (int du = 0; du < 2; ++du) {
                    const int cu = (int)u0 + du, cv = (int)v0 + dv;
                    const double val = (cu >= 0 && cu < naxis1 && cv >= 0 && cv < naxis2)
                                           ? img[(size_t)cv * naxis1 + cu] : 0.0;
                    acc = add_rn(acc, mul_rn(mul_rn(wv[dv], wu[du]), val));
                }
        }
    }
    out[(size_t)i * naxis1out + j] = acc;
}

// Tile-major copy of the staged tables: for every live tile and scanned row the block
// weights and the meta words of the tile's pixels (rounded out to multiples of 4 pixels,
// padded to PXCAP), laid out exactly as a stage of the hot kernel's shared memory, so that a
// row costs the copy warp ONE bulk copy.
template <typename T>
__global__ void k_pack_tiles(const Tile *tiles, const T *quad, const uint32_t *meta, char *pack) {
    const Tile &t = tiles[blockIdx.x];
    const int y = blockIdx.y;
    char *dst = pack + ((size_t)blockIdx.x * NROWS + y) * stage_bytes(sizeof(T));
    T *q = reinterpret_cast<T *>(dst);
    uint32_t *m = reinterpret_cast<uint32_t *>(dst + PXCAP * 4 * sizeof(T));
    const int jal = t.jlo & ~3;
    const size_t row = ((size_t)t.slit * NROWS + y) * W;
    for (int i = threadIdx.x; i < PXCAP; i += blockDim.x) {
        const int j = jal + i;
        const bool ok = j < W;
#pragma unroll
        for (int e = 0; e < 4; ++e) q[i * 4 + e] = ok ? quad[(row + j) * 4 + e] : T(0);
        m[i] = ok ? meta[row + j] : META_COMPACT;
    }
}

__global__ void k_init_jmm(int32_t *jmm, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) jmm[i] = (i & 1) ? -1 : INT_MAX;
}

// Kernel 1 alone, straight from global memory (parity hook / Slitlet2D.rectify).
template <typename T>
__global__ void k_rectify_only(const T *frame, T *rect, const int32_t *base, const T *w,
                               int s, int K, int bbw, int naxis1, int naxis2) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (j >= bbw) return;
    const size_t rowtab = (size_t)s * NROWS + y;
    const int32_t pk = base[rowtab * W + j];
    const int r0 = base_row(pk), c0 = base_col(pk);
    const T *wp = w + (rowtab * W + j) * rec_size(K);
    T acc = T(0);
    for (int a = 0; a < K; ++a)
        for (int b = 0; b < K; ++b) {
            const T wt = record_weight(wp, K, a, b);
            const int r = min(max(r0 + a, 0), naxis2 - 1);
            const int c = min(max(c0 + b, 0), naxis1 - 1);
            acc = fma(wt, frame[(size_t)r * naxis1 + c], acc);
        }
    rect[(size_t)y * bbw + j] = acc;
}
