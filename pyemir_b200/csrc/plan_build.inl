// Host side of a plan: launch configuration of the hot kernel, border tables in numpy's
// operation order, tiles, build_plan, workspace of the host-buffer entry.
// Included by emirwv.cu inside its anonymous namespace (one translation unit).

int env_int(const char *name, int fallback) {
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : fallback;
}

// launch configuration for a given number of frames per CTA
struct LaunchCfg {
    int G, threads;
    size_t smem;
    int *query_ctas_per_sm = nullptr;   // not null: report the occupancy, launch nothing
};

// shared memory of the hot kernel:
//   [3 tile descriptors][mbarriers, counters][G windows: WIN_ROWS x WIN_PITCH]
//   [per warp: 2 x (rect | ya) x 32 x G knots]
//   [nstage table rows: PXCAP x (4 block weights + 1 meta word)]
size_t smem_bytes(const emirwv_plan *pl, int G) {
    return 128 /* alignment slack */ + 3 * sizeof(Tile) + BAR_BYTES +
           (size_t)G * (WIN_ROWS * WIN_PITCH + NWARPS * 2 * 32) * pl->esize +
           (size_t)pl->nstage * stage_bytes((int)pl->esize) + ZFILL_BYTES;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (libcuda is not linked, so
// the library still loads on a box without a driver)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) !=
                cudaSuccess || qres != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(ptr);
    }();
    return fn;
}

// The frames of a launch as a 3-D tensor (column, row, frame); a box is WIN_PITCH columns of
// one row of G consecutive frames -- one ring row of the hot kernel's window.
int make_frames_tensor_map(CUtensorMap *tm, const void *d_in, int esize, int naxis1, int naxis2,
                           int nframes, int G) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (fn == nullptr) return fail(EMIRWV_E_CUDA, "cuTensorMapEncodeTiled is not available");
    const cuuint64_t dims[3] = {(cuuint64_t)naxis1, (cuuint64_t)naxis2, (cuuint64_t)nframes};
    const cuuint64_t strides[2] = {(cuuint64_t)naxis1 * esize, (cuuint64_t)naxis1 * naxis2 * esize};
    const cuuint32_t box[3] = {(cuuint32_t)WIN_PITCH, 1u, (cuuint32_t)G};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUresult rc = fn(tm, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64,
                           3, const_cast<void *>(d_in), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return fail(EMIRWV_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
    return EMIRWV_OK;
}

LaunchCfg choose_cfg(const emirwv_plan *pl, int nframes) {
    LaunchCfg cfg;
    cfg.threads = NTHREADS;
    int G = pl->tune_G > 0 ? pl->tune_G : env_int("EMIRWV_G", pl->esize == 4 ? 4 : 2);
    G = (G >= 4) ? 4 : (G >= 2 ? 2 : 1);
    while (G > 1 && (G > nframes || smem_bytes(pl, G) > (size_t)SMEM_LIMIT)) G >>= 1;
    cfg.G = G;
    cfg.smem = smem_bytes(pl, G);
    return cfg;
}

template <typename T, int K, int G, bool SPANLOOP, int NS>
int launch_final(const emirwv_plan *pl, RunParams<T> rp, const LaunchCfg &cfg, cudaStream_t st) {
    static_assert(WIN_PITCH <= 256 && (WIN_PITCH * sizeof(float)) % 16 == 0, "tensor-map box");
    auto kern = k_rectwv<T, K, G, SPANLOOP, NS>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)cfg.smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, cfg.threads, cfg.smem));
    if (per_sm < 1) return fail(EMIRWV_E_LIMIT, "kernel does not fit on an SM");
    if (cfg.query_ctas_per_sm != nullptr) {
        *cfg.query_ctas_per_sm = per_sm;
        return EMIRWV_OK;
    }
    rp.ngroups = (rp.nframes + G - 1) / G;
    const long long nwork = (long long)rp.nlive_tiles * rp.ngroups;
    if (nwork > 0) {
        if (nwork > INT_MAX) return fail(EMIRWV_E_LIMIT, "too many frames in one call");
        const int grid = (int)std::min<long long>(nwork, (long long)pl->num_sms * per_sm);
        alignas(64) CUtensorMap tmap;
        const int rc = make_frames_tensor_map(&tmap, rp.in, (int)sizeof(T), W, rp.naxis2,
                                              rp.nframes, G);
        if (rc) return rc;
        kern<<<grid, cfg.threads, cfg.smem, st>>>(rp, tmap);
        CUDA_TRY(cudaGetLastError());
    }
    return EMIRWV_OK;
}

// EMIRWV_SLIM (experiment builds, scripts/build_variants.py): only the instantiation the
// headline configuration uses (float32, K = 3, G = 4, no span loop) is compiled.
#ifdef EMIRWV_SLIM
template <typename T, int K, int G>
int launch_one(const emirwv_plan *pl, const RunParams<T> &rp, const LaunchCfg &cfg,
               cudaStream_t st) {
    if constexpr (sizeof(T) == 4 && K == 3 && G == 4) {
        if (pl->max_span <= 2) {
            if (pl->nstage == 2) return launch_final<T, K, G, false, 2>(pl, rp, cfg, st);
            return launch_final<T, K, G, false, NSTAGE_MAX>(pl, rp, cfg, st);
        }
    }
    return fail(EMIRWV_E_LIMIT, "slim experiment build: float32, K = 3, G = 4 only");
}
#else
template <typename T, int K, int G>
int launch_one(const emirwv_plan *pl, const RunParams<T> &rp, const LaunchCfg &cfg,
               cudaStream_t st) {
    if (pl->nstage == 2) {                   // the shallow pipeline (default for float64 plans)
        if (pl->max_span > 2) return launch_final<T, K, G, true, 2>(pl, rp, cfg, st);
        return launch_final<T, K, G, false, 2>(pl, rp, cfg, st);
    }
    if (pl->max_span > 2) return launch_final<T, K, G, true, NSTAGE_MAX>(pl, rp, cfg, st);
    return launch_final<T, K, G, false, NSTAGE_MAX>(pl, rp, cfg, st);
}
#endif

template <typename T, int K>
int launch_k(const emirwv_plan *pl, const RunParams<T> &rp, const LaunchCfg &cfg,
             cudaStream_t st) {
    switch (cfg.G) {
        case 4: return launch_one<T, K, 4>(pl, rp, cfg, st);
        case 2: return launch_one<T, K, 2>(pl, rp, cfg, st);
        default: return launch_one<T, K, 1>(pl, rp, cfg, st);
    }
}

template <typename T>
int launch_t(const emirwv_plan *pl, const void *d_in, void *d_out, int nframes,
             int32_t *d_jmm, cudaStream_t st, int *query_ctas_per_sm = nullptr,
             bool skip_fill = false) {
    LaunchCfg cfg = choose_cfg(pl, nframes);
    cfg.query_ctas_per_sm = query_ctas_per_sm;
    if (cfg.smem > (size_t)SMEM_LIMIT)
        return fail(EMIRWV_E_LIMIT, "kernel needs %zu bytes of shared memory", cfg.smem);
    if ((reinterpret_cast<uintptr_t>(d_in) & 15u) || (reinterpret_cast<uintptr_t>(d_out) & 15u))
        return fail(EMIRWV_E_VALUE, "device buffers must be 16-byte aligned");
    if (query_ctas_per_sm != nullptr) {      // occupancy of the kernel this launch would use
        RunParams<T> rq;
        memset(&rq, 0, sizeof(rq));
        switch (pl->K) {
            case 1: return launch_k<T, 1>(pl, rq, cfg, st);
            case 2: return launch_k<T, 2>(pl, rq, cfg, st);
            case 3: return launch_k<T, 3>(pl, rq, cfg, st);
            case 4: return launch_k<T, 4>(pl, rq, cfg, st);
            default: return fail(EMIRWV_E_LIMIT, "unsupported footprint %d", pl->K);
        }
    }
    RunParams<T> rp;
    rp.in = static_cast<const T *>(d_in);
    rp.out = static_cast<T *>(d_out);
    rp.jmm = d_jmm;
    rp.nframes = nframes;
    rp.tiles = pl->d_tiles;
    rp.base = pl->d_base;
    rp.w = static_cast<const T *>(pl->d_w);
    rp.tilepack = pl->d_tilepack;
    rp.border = static_cast<const Border<T> *>(pl->d_border);
    rp.pix = static_cast<const PixGeom<T> *>(pl->d_pix);
    rp.naxis2 = pl->desc.naxis2;
    rp.nout = pl->nout;
    rp.NB = pl->NB;
    rp.rps = pl->rps;
    rp.nrows_out = pl->nsl * pl->rps;
    rp.nsl = pl->nsl;
    rp.vec_ok = ((size_t)pl->nout * sizeof(T)) % 16 == 0;
    rp.nlive_tiles = pl->ntiles - pl->ntiles_empty;
    rp.ngroups = 0;
    rp.sched = nullptr;
    rp.early = pl->early_rows;
    if (pl->dynamic_sched && pl->d_sched != nullptr) {
        rp.sched = pl->d_sched + pl->sched_next.fetch_add(1) % SCHED_POOL;
        CUDA_TRY(cudaMemsetAsync(rp.sched, 0, sizeof(int), st));
    }
    if (d_jmm != nullptr) {
        const int n = nframes * pl->nsl * 2;
        k_init_jmm<<<(n + 255) / 256, 256, 0, st>>>(d_jmm, n);
        CUDA_TRY(cudaGetLastError());
    }
    // The zero-fill kernel is memory bound and the hot kernel issue bound: when a side
    // stream is available the hot kernel is enqueued first (it takes its shared memory
    // and registers on every SM) and the fill blocks run next to it in what is left.
    // (skip_fill: the caller's buffer already holds the zeros outside the coverage)
    // The zero runs are written by the copy warps of the hot kernel (TMA bulk stores) when
    // the output rows are 16-byte aligned and there is a hot kernel to ride on; otherwise by
    // k_zero_fill beside it (EMIRWV_FILL=kernel forces that).
    const char *fill_env = getenv("EMIRWV_FILL");
    const bool fill_in_hot = !skip_fill && pl->d_fillruns != nullptr && rp.vec_ok &&
                             rp.nlive_tiles > 0 && !(fill_env && strcmp(fill_env, "kernel") == 0);
    rp.fillruns = fill_in_hot ? pl->d_fillruns : nullptr;
    rp.nfillruns = pl->ntiles_empty * pl->rps;
    const int nfill = (skip_fill || fill_in_hot) ? 0 : pl->ntiles_empty;
    const bool forked = nfill > 0 && pl->aux != nullptr;
    rp.ndead_tiles = pl->ntiles_empty;
    const int zgrid = std::max(1, env_int("EMIRWV_FILL_BLOCKS", 4 * pl->num_sms));
    const int zthreads = std::max(32, std::min(128, env_int("EMIRWV_FILL_THREADS", 128) & ~31));
    if (nfill > 0 && !forked) {
        k_zero_fill<T><<<zgrid, zthreads, 0, st>>>(rp, rp.nlive_tiles);
        CUDA_TRY(cudaGetLastError());
    }
    std::unique_lock<std::mutex> lock(pl->aux_mu, std::defer_lock);
    if (forked) {
        lock.lock();
        CUDA_TRY(cudaEventRecord(pl->ev_fork, st));
    }
    int rc;
    switch (pl->K) {
        case 1: rc = launch_k<T, 1>(pl, rp, cfg, st); break;
        case 2: rc = launch_k<T, 2>(pl, rp, cfg, st); break;
        case 3: rc = launch_k<T, 3>(pl, rp, cfg, st); break;
        case 4: rc = launch_k<T, 4>(pl, rp, cfg, st); break;
        default: rc = fail(EMIRWV_E_LIMIT, "unsupported footprint %d", pl->K);
    }
    if (forked) {
        // the two kernels write disjoint parts of the output
        CUDA_TRY(cudaStreamWaitEvent(pl->aux, pl->ev_fork, 0));
        k_zero_fill<T><<<zgrid, zthreads, 0, pl->aux>>>(rp, rp.nlive_tiles);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(pl->ev_join, pl->aux));
        CUDA_TRY(cudaStreamWaitEvent(st, pl->ev_join, 0));
    }
    return rc;
}

// numpy.polynomial.polynomial.polyval: Horner, one rounding per operation
double polyval_np(double x, const double *c, int n) {
    double c0 = c[n - 1];
    for (int m = n - 2; m >= 0; --m) c0 = add_rn(c[m], mul_rn(c0, x));
    return c0;
}

template <typename T>
int upload_tables(emirwv_plan *pl, const std::vector<double> &tau, const std::vector<double> &h) {
    const int nsl = pl->nsl, NB = pl->NB;
    // both tables carry 64 extra entries: the kernel requests the row of its next
    // pass before it knows whether there is one
    std::vector<Border<T>> border((size_t)nsl * NB + 64);
    memset(border.data(), 0, border.size() * sizeof(Border<T>));
    for (size_t k = 0; k < (size_t)nsl * NB; ++k) {
        border[k].tau = (T)tau[k];
        border[k].idx = pl->h_idx[k];
    }
    std::vector<PixGeom<T>> pix((size_t)nsl * W + 64);
    memset(pix.data(), 0, pix.size() * sizeof(PixGeom<T>));
    for (int s = 0; s < nsl; ++s) {
        const emirwv_slitlet_desc &sl = pl->desc.slitlet[s];
        if (!sl.valid) continue;
        const int nchan = sl.bb_nc2_orig - sl.bb_nc1_orig + 1;
        const double *hs = h.data() + (size_t)s * W;
        for (int i = 0; i < nchan; ++i) {
            PixGeom<T> &pg = pix[(size_t)s * W + i];
            pg.qm = i > 0 ? (T)(hs[i] / hs[i - 1]) : T(0);
            pg.qp = i + 1 < nchan ? (T)(hs[i] / hs[i + 1]) : T(0);
            pg.r0 = i > 0 ? (T)(hs[i] / (hs[i - 1] + hs[i])) : T(0);
            pg.r1 = i + 1 < nchan ? (T)(hs[i + 1] / (hs[i] + hs[i + 1])) : T(0);
        }
    }
    CUDA_TRY(cudaMalloc(&pl->d_border, border.size() * sizeof(Border<T>)));
    CUDA_TRY(cudaMemcpy(pl->d_border, border.data(), border.size() * sizeof(Border<T>),
                        cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&pl->d_pix, pix.size() * sizeof(PixGeom<T>)));
    CUDA_TRY(cudaMemcpy(pl->d_pix, pix.data(), pix.size() * sizeof(PixGeom<T>),
                        cudaMemcpyHostToDevice));
    pl->table_bytes += (int64_t)(border.size() * sizeof(Border<T>) + pix.size() * sizeof(PixGeom<T>));
    return EMIRWV_OK;
}

// Border tables of every slitlet; float64 with numpy's operation order so that the
// interval of every border is the one numpy.searchsorted finds in the reference's
// resample_image2d_flux / SteffenInterpolator (apply_rectwv_coeff.py:152-159).
int build_border_tables(emirwv_plan *pl, std::vector<uint8_t> &covered) {
    const emirwv_plan_desc &d = pl->desc;
    const int nout = pl->nout, NB = pl->NB, nsl = pl->nsl;
    std::vector<double> nb(NB), wl(nout);
    for (int k = 0; k < nout; ++k) wl[k] = add_rn(d.crval1, mul_rn(d.cdelt1, (double)k));
    for (int k = 1; k < nout; ++k) nb[k] = mul_rn(0.5, add_rn(wl[k], wl[k - 1]));
    nb[0] = add_rn(mul_rn(2.0, wl[0]), -nb[1]);
    nb[nout] = add_rn(mul_rn(2.0, wl[nout - 1]), -nb[nout - 1]);

    pl->h_idx.assign((size_t)nsl * NB, 0);
    pl->h_t.assign((size_t)nsl * NB, 0.0);
    std::vector<double> h_tau((size_t)nsl * NB, 0.0);
    std::vector<double> h_h((size_t)nsl * W, 0.0);
    covered.assign((size_t)nsl * NB, 0);
    pl->max_span = 0;

    std::vector<double> xw(W + 1);
    for (int s = 0; s < nsl; ++s) {
        const emirwv_slitlet_desc &sl = d.slitlet[s];
        if (!sl.valid) continue;
        const int nchan = sl.bb_nc2_orig - sl.bb_nc1_orig + 1;
        double *h = h_h.data() + (size_t)s * W;
        for (int i = 0; i <= nchan; ++i) {
            const double xb = add_rn(add_rn(-0.5, (double)i), d.crpix1);
            xw[i] = polyval_np(xb, sl.wpoly, sl.nwpoly);
        }
        for (int i = 0; i < nchan; ++i) {
            h[i] = xw[i + 1] - xw[i];
            if (!(h[i] > 0.0))
                return fail(EMIRWV_E_VALUE,
                            "slitlet %d: wavelength polynomial is not increasing at pixel %d",
                            s + 1, i + 1);
        }
        int prev = 0;
        for (int k = 0; k < NB; ++k) {
            const double x = nb[k];
            int idx;
            double t, tau;
            uint8_t cov = 1;
            if (x < xw[0]) {                       // extrapolate='border': y_0
                idx = 0;
                t = 0.0;
                tau = 0.0;
                cov = 0;
            } else if (x > xw[nchan]) {            // extrapolate='border': y_N
                idx = nchan - 1;
                t = h[nchan - 1];
                tau = 1.0;
                cov = 0;
            } else {                               // searchsorted(side='left').clip(1, N) - 1
                const int pos = (int)(std::lower_bound(xw.begin(), xw.begin() + nchan + 1, x) -
                                      xw.begin());
                idx = std::min(std::max(pos, 1), nchan) - 1;
                t = x - xw[idx];
                tau = t / h[idx];
            }
            pl->h_idx[(size_t)s * NB + k] = idx;
            pl->h_t[(size_t)s * NB + k] = t;
            h_tau[(size_t)s * NB + k] = tau;
            covered[(size_t)s * NB + k] = cov;
            if (k > 0) pl->max_span = std::max(pl->max_span, idx - prev);
            prev = idx;
        }
    }
    if (pl->dtype == EMIRWV_F32) return upload_tables<float>(pl, h_tau, h_h);
    return upload_tables<double>(pl, h_tau, h_h);
}

// Work items: (slitlet, run of output columns), all 39 scanned rows.  A live tile is cut
// into at most NWARPS consecutive pieces, one per warp: at most BW output columns whose
// borders touch at most WARP_PX rectified pixels.  Columns without wavelength coverage
// (and missing slitlets) become wide runs of zeros for the fill kernel.
// Returns 1 when a tile does not fit the kernel's fixed shared-memory geometry with this
// BW (the caller retries with narrower pieces), 0 on success.
int try_build_tiles(emirwv_plan *pl, const std::vector<uint8_t> &covered, int BW) {
    const emirwv_plan_desc &d = pl->desc;
    const int nout = pl->nout, NB = pl->NB, K = pl->K;
    std::vector<Tile> live, dead;
    int wcols_max = 4, span_max = 1, npx_max = 1, nk_max = 0;
    for (int s = 0; s < pl->nsl; ++s) {
        const emirwv_slitlet_desc &sl = d.slitlet[s];
        const int nchan = sl.valid ? sl.bb_nc2_orig - sl.bb_nc1_orig + 1 : 0;
        const uint8_t *cov = covered.data() + (size_t)s * NB;
        const int32_t *idx = pl->h_idx.data() + (size_t)s * NB;
        // output column k carries flux when one of its borders lies inside the spectrum or
        // the spectrum lies between them
        auto is_live = [&](int k) {
            return sl.valid && (cov[k] != 0 || cov[k + 1] != 0 || idx[k] != idx[k + 1]);
        };
        // length of the run of dead columns starting at k
        auto dead_run = [&](int k) {
            int n = 0;
            while (k + n < nout && !is_live(k + n)) ++n;
            return n;
        };
        int k0 = 0;
        while (k0 < nout) {
            Tile t;
            memset(&t, 0, sizeof(t));
            t.slit = s;
            t.k0 = k0;
            t.nchan = nchan;
            int run = dead_run(k0);
            if (k0 + run < nout) run &= ~3;        // keep the next tile 16-byte aligned
            if (run >= 32 || k0 + run == nout) {
                t.empty = 1;
                t.nk = std::min(run, 1024);
                dead.push_back(t);
                k0 += t.nk;
                continue;
            }
            // live tile: one piece per warp
            int k = k0, nw = 0;
            t.kw[0] = 0;
            while (nw < NWARPS && k < nout) {
                if (nw > 0 && dead_run(k) >= 32) break;        // a fill run starts here
                int kend = std::min(k + BW, nout);
                while (kend > k + 1 && idx[kend] - idx[k] + 1 > WARP_PX) --kend;
                if (idx[kend] - idx[k] + 1 > WARP_PX) return 1;   // one column spans too much
                k = kend;
                t.kw[++nw] = (uint8_t)(k - k0);
            }
            if (k < nout && (k & 3)) {
                // end on a multiple of 4 so that a following fill run stays vectorised:
                // lengthen the last piece if it stays within its limits, else shorten it
                const int kprev = k0 + t.kw[nw - 1];
                const int up = std::min((k + 3) & ~3, nout), down = k & ~3;
                if (up - kprev <= WARP_BORDERS && idx[up] - idx[kprev] + 1 <= WARP_PX)
                    k = up;
                else if (down > kprev)
                    k = down;
                t.kw[nw] = (uint8_t)(k - k0);
            }
            for (int wq = nw + 1; wq <= NWARPS; ++wq) t.kw[wq] = t.kw[nw];
            t.nk = k - k0;
            const int ilo = idx[k0], ihi = idx[k0 + t.nk];
            t.jlo = std::max(0, ilo - 1);
            const int jhi = std::min(nchan - 1, ihi + 1);
            t.npx = jhi - t.jlo + 1;
            // (the staged table row is rounded out to multiples of 4 pixels)
            if (((t.jlo & 3) + t.npx + 3 & ~3) > PXCAP) return 1;
            int c0 = INT_MAX, c1 = INT_MIN;
            int rlo[NROWS], rhi[NROWS];
            for (int y = 0; y < NROWS; ++y) {
                const int32_t *row = pl->h_base.data() + ((size_t)s * NROWS + y) * W;
                const uint8_t *brow = pl->h_blockrow.data() + ((size_t)s * NROWS + y) * W;
                int r0 = INT_MAX, r1 = INT_MIN;
                for (int j = t.jlo; j <= jhi; ++j) {
                    // source rows with weight: the two rows of the 2 x 2 block, or K rows
                    const bool compact = (brow[j] & 16) != 0;
                    const int br = base_row(row[j]) + (brow[j] & 15), bc = base_col(row[j]);
                    r0 = std::min(r0, br);
                    r1 = std::max(r1, br + (compact ? std::min(K, 2) : K) - 1);
                    c0 = std::min(c0, bc);
                    c1 = std::max(c1, bc + K - 1);
                }
                rlo[y] = r0;
                rhi[y] = r1;
            }
            // rolling window: rows may only be added at the top and dropped at the bottom
            for (int y = 1; y < NROWS; ++y) rhi[y] = std::max(rhi[y], rhi[y - 1]);
            for (int y = NROWS - 2; y >= 0; --y) rlo[y] = std::min(rlo[y], rlo[y + 1]);
            for (int y = 0; y < NROWS; ++y) {
                const int next_hi = rhi[std::min(y + pl->nstage - 1, NROWS - 1)];
                span_max = std::max(span_max, next_hi - rlo[y] + 1);
                t.span[y] = (int32_t)(((uint32_t)(rhi[y] & 0xffff) << 16) | (uint32_t)(rlo[y] & 0xffff));
            }
            c0 = (int)(std::floor(c0 / 4.0) * 4.0);        // 16-byte aligned window rows
            t.c0 = c0;
            t.wcols = ((c1 - c0 + 1) + 3) & ~3;
            if (t.wcols > WIN_PITCH || span_max > WIN_ROWS) return 1;
            wcols_max = std::max(wcols_max, t.wcols);
            npx_max = std::max(npx_max, t.npx);
            nk_max = std::max(nk_max, t.nk);
            live.push_back(t);
            k0 += t.nk;
        }
    }
    pl->h_cover.assign((size_t)pl->nsl * 2, 0);
    for (const Tile &t : live) {
        int32_t *c = pl->h_cover.data() + 2 * t.slit;
        if (c[1] == 0) c[0] = t.k0;
        c[0] = std::min(c[0], t.k0);
        c[1] = std::max(c[1], t.k0 + t.nk);
    }
    // the widest (most expensive) tiles first, the narrow ones at the ends of the slitlets
    // last: the tail of the launch is made of short items
    if (env_int("EMIRWV_TILE_ORDER", 1))
        std::stable_sort(live.begin(), live.end(),
                         [](const Tile &a, const Tile &b) { return a.nk > b.nk; });
    pl->ntiles_empty = (int)dead.size();
    pl->h_tiles = live;
    pl->h_tiles.insert(pl->h_tiles.end(), dead.begin(), dead.end());
    pl->ntiles = (int)pl->h_tiles.size();
    pl->tile_k = nk_max;
    pl->wrows_max = span_max;
    pl->wcols_max = wcols_max;
    pl->npx_max = npx_max;
    return 0;
}

int build_tiles(emirwv_plan *pl, const std::vector<uint8_t> &covered) {
    // widest pieces whose source window fits the CTA
    int BW = std::max(4, std::min(pl->tile_k / NWARPS, WARP_BORDERS));
    for (; BW >= 4; BW -= (BW > 8 ? 4 : 1))
        if (try_build_tiles(pl, covered, BW) == 0) break;
    if (BW < 4)
        return fail(EMIRWV_E_LIMIT,
                    "distortion too strong: the source window of a 32-column block does not fit "
                    "(%d rows x %d columns available)", WIN_ROWS, WIN_PITCH);
    if (pl->ntiles_empty > 65535) return fail(EMIRWV_E_LIMIT, "too many tiles");
    pl->dynamic_sched = env_int("EMIRWV_DYNAMIC", 1);
    pl->early_rows = env_int("EMIRWV_EARLY", 1);
    CUDA_TRY(cudaMalloc(&pl->d_sched, SCHED_POOL * sizeof(int)));
    CUDA_TRY(cudaMemset(pl->d_sched, 0, SCHED_POOL * sizeof(int)));
    if (pl->ntiles_empty > 0) {
        // the zero runs of one frame, one per dead tile and output row: first element, bytes
        std::vector<int2> runs;
        runs.reserve((size_t)pl->ntiles_empty * pl->rps);
        for (int i = pl->ntiles - pl->ntiles_empty; i < pl->ntiles; ++i) {
            const Tile &t = pl->h_tiles[i];
            for (int y = 0; y < pl->rps; ++y)
                runs.push_back(make_int2((t.slit * pl->rps + y) * pl->nout + t.k0,
                                         (int)(t.nk * pl->esize)));
        }
        // bulk stores need 16-byte aligned runs; a plan with an odd one keeps k_zero_fill
        bool aligned = true;
        for (const int2 &r : runs)
            aligned = aligned && ((size_t)r.x * pl->esize) % 16 == 0 && r.y % 16 == 0;
        if (aligned) {
            CUDA_TRY(cudaMalloc(&pl->d_fillruns, runs.size() * sizeof(int2)));
            CUDA_TRY(cudaMemcpy(pl->d_fillruns, runs.data(), runs.size() * sizeof(int2),
                                cudaMemcpyHostToDevice));
            pl->table_bytes += (int64_t)(runs.size() * sizeof(int2));
        }
    }
    CUDA_TRY(cudaMalloc(&pl->d_tiles, pl->h_tiles.size() * sizeof(Tile)));
    CUDA_TRY(cudaMemcpy(pl->d_tiles, pl->h_tiles.data(), pl->h_tiles.size() * sizeof(Tile),
                        cudaMemcpyHostToDevice));
    pl->table_bytes += (int64_t)(pl->h_tiles.size() * sizeof(Tile));
    // the staged tables, tile-major; the pixel-major block weights and meta words are not
    // needed afterwards
    const int nlive = pl->ntiles - pl->ntiles_empty;
    const size_t pack_bytes = (size_t)nlive * NROWS * stage_bytes((int)pl->esize);
    CUDA_TRY(cudaMalloc(&pl->d_tilepack, pack_bytes + 64));
    if (nlive > 0) {
        const dim3 pgrid(nlive, NROWS);
        if (pl->esize == 4)
            k_pack_tiles<float><<<pgrid, 256>>>(pl->d_tiles, static_cast<const float *>(pl->d_quad),
                                                pl->d_meta, pl->d_tilepack);
        else
            k_pack_tiles<double><<<pgrid, 256>>>(pl->d_tiles, static_cast<const double *>(pl->d_quad),
                                                 pl->d_meta, pl->d_tilepack);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaDeviceSynchronize());
    }
    const size_t nrowtab = (size_t)pl->nsl * NROWS;
    pl->table_bytes += (int64_t)pack_bytes - (int64_t)(nrowtab * W * (4 * pl->esize + sizeof(uint32_t)));
    cudaFree(pl->d_quad);
    cudaFree(pl->d_meta);
    pl->d_quad = nullptr;
    pl->d_meta = nullptr;
    return EMIRWV_OK;
}

int validate(const emirwv_plan_desc *d) {
    // Slitlet2D.extract_slitlet2d accepts 2048 x 2048 frames only (slitlet2d.py:359-363)
    if (d->naxis1 != W) return fail(EMIRWV_E_VALUE, "Unexpected naxis1");
    if (d->naxis2 < 2 || d->naxis2 > 16384) return fail(EMIRWV_E_VALUE, "Unexpected naxis2");
    if (d->nslitlets < 1 || d->nslitlets > EMIRWV_MAX_SLITLETS)
        return fail(EMIRWV_E_VALUE, "nslitlets=%d out of range", d->nslitlets);
    if (d->rows_per_slitlet != NROWS - 1)
        return fail(EMIRWV_E_VALUE, "rows_per_slitlet must be %d", NROWS - 1);
    if (d->resampling < 1 || d->resampling > 3)
        return fail(EMIRWV_E_VALUE, "Unexpected resampling value=%d", d->resampling);
    if (d->dtype != EMIRWV_F32 && d->dtype != EMIRWV_F64)
        return fail(EMIRWV_E_VALUE, "unknown dtype %d", d->dtype);
    if (d->naxis1_out < 2 || d->naxis1_out > 65536)
        return fail(EMIRWV_E_VALUE, "naxis1_out=%d out of range", d->naxis1_out);
    if (!(d->cdelt1 > 0.0)) return fail(EMIRWV_E_VALUE, "cdelt1 must be positive");
    if (d->stages != 0 && d->stages != 2 && d->stages != NSTAGE_MAX)
        return fail(EMIRWV_E_VALUE, "stages must be 0 (default), 2 or %d", NSTAGE_MAX);
    const int cap = d->max_order > 0 ? d->max_order : 4;
    for (int s = 0; s < d->nslitlets; ++s) {
        const emirwv_slitlet_desc &sl = d->slitlet[s];
        if (!sl.valid) continue;
        const int order = order_from_ncoef(sl.ncoef);
        if (order < 0 || order > cap)
            return fail(EMIRWV_E_VALUE, "order > %d not implemented", cap);
        if (sl.nwpoly < 1 || sl.nwpoly > EMIRWV_MAX_WPOLY)
            return fail(EMIRWV_E_VALUE, "slitlet %d: %d wavelength coefficients", s + 1, sl.nwpoly);
        if (sl.bb_nc1_orig < 1 || sl.bb_nc2_orig > d->naxis1 || sl.bb_nc2_orig < sl.bb_nc1_orig + 1 ||
            sl.bb_ns1_orig < 1 || sl.bb_ns2_orig > d->naxis2 || sl.bb_ns2_orig < sl.bb_ns1_orig)
            return fail(EMIRWV_E_VALUE, "slitlet %d: bounding box outside the frame", s + 1);
        const int bbh = sl.bb_ns2_orig - sl.bb_ns1_orig + 1;
        if (sl.max_row_rectified - sl.min_row_rectified + 1 != d->rows_per_slitlet)
            return fail(EMIRWV_E_VALUE,
                        "could not broadcast input array from shape (%d,%d) into shape (%d,%d)",
                        sl.max_row_rectified - sl.min_row_rectified + 1, d->naxis1_out,
                        d->rows_per_slitlet, d->naxis1_out);
        if (sl.min_row_rectified < 0)
            return fail(EMIRWV_E_VALUE, "slitlet %d: negative min_row_rectified", s + 1);
        if (sl.max_row_rectified + 1 >= bbh)
            return fail(EMIRWV_E_INDEX, "index %d is out of bounds for axis 0 with size %d",
                        sl.max_row_rectified + 1, bbh);
    }
    return EMIRWV_OK;
}

int build_plan(emirwv_plan *pl) {
    const emirwv_plan_desc &d = pl->desc;
    CUDA_TRY(cudaGetDevice(&pl->device));
    CUDA_TRY(cudaDeviceGetAttribute(&pl->num_sms, cudaDevAttrMultiProcessorCount, pl->device));
    pl->dtype = d.dtype;
    pl->mode = d.resampling;
    pl->esize = d.dtype == EMIRWV_F32 ? 4 : 8;
    pl->nsl = d.nslitlets;
    pl->rps = d.rows_per_slitlet;
    pl->nout = d.naxis1_out;
    pl->NB = d.naxis1_out + 1;
    pl->tile_k = d.tile_k > 0 ? d.tile_k : env_int("EMIRWV_TILE_K", TILE_K_MAX);
    // float64: two staged table rows (36 bytes per pixel each) keep three CTAs per SM like
    // float32; measured faster than four stages with fewer CTAs, bit-identical output
    // (EMIRWV_NSTAGE_F64=4 brings the deep pipeline back)
    pl->nstage = NSTAGE_MAX;
    if (pl->esize == 8 && env_int("EMIRWV_NSTAGE_F64", 2) == 2) pl->nstage = 2;
    if (d.stages != 0) pl->nstage = d.stages;                  // validated: 2 or NSTAGE_MAX
    pl->tile_k = std::max(32, std::min(pl->tile_k, TILE_K_MAX));

    pl->h_slit.resize(pl->nsl);
    for (int s = 0; s < pl->nsl; ++s) {
        const emirwv_slitlet_desc &sl = d.slitlet[s];
        SlitDev &sd = pl->h_slit[s];
        memset(&sd, 0, sizeof(sd));
        sd.valid = sl.valid ? 1 : 0;
        if (!sl.valid) continue;
        sd.ns1 = sl.bb_ns1_orig;
        sd.nc1 = sl.bb_nc1_orig;
        sd.bbh = sl.bb_ns2_orig - sl.bb_ns1_orig + 1;
        sd.bbw = sl.bb_nc2_orig - sl.bb_nc1_orig + 1;
        sd.min_row = sl.min_row_rectified;
        sd.order = order_from_ncoef(sl.ncoef);
        memcpy(sd.a, sl.aij, sizeof(double) * sl.ncoef);
        memcpy(sd.b, sl.bij, sizeof(double) * sl.ncoef);
    }
    CUDA_TRY(cudaMalloc(&pl->d_slit, pl->nsl * sizeof(SlitDev)));
    CUDA_TRY(cudaMemcpy(pl->d_slit, pl->h_slit.data(), pl->nsl * sizeof(SlitDev),
                        cudaMemcpyHostToDevice));

    GeomParams gp;
    gp.slit = pl->d_slit;
    gp.nsl = pl->nsl;
    gp.naxis1 = d.naxis1;
    gp.naxis2 = d.naxis2;
    gp.offx = d.offx;
    gp.offy = d.offy;
    gp.mode = pl->mode;
    gp.kmax = nullptr;
    gp.meta = nullptr;
    gp.blockrow = nullptr;
    const dim3 block(128);
    const dim3 grid((W + 127) / 128, NROWS, pl->nsl);

    if (pl->mode == EMIRWV_AREA) {
        int *d_kmax = nullptr;
        int kmax = 1;
        CUDA_TRY(cudaMalloc(&d_kmax, sizeof(int)));
        CUDA_TRY(cudaMemcpy(d_kmax, &kmax, sizeof(int), cudaMemcpyHostToDevice));
        gp.kmax = d_kmax;
        gp.K = 0;
        gp.base = nullptr;
        k_extent<<<grid, block>>>(gp);
        cudaError_t err = cudaGetLastError();
        if (err == cudaSuccess)
            err = cudaMemcpy(&kmax, d_kmax, sizeof(int), cudaMemcpyDeviceToHost);
        cudaFree(d_kmax);
        if (err != cudaSuccess)
            return fail(EMIRWV_E_CUDA, "k_extent failed: %s", cudaGetErrorString(err));
        if (kmax > MAXK)
            return fail(EMIRWV_E_LIMIT,
                        "distortion too strong: an output pixel covers %d source pixels "
                        "along one axis (limit %d)", kmax, MAXK);
        pl->K = kmax;
    } else {
        pl->K = pl->mode == EMIRWV_NEAREST ? 1 : 2;
    }

    const size_t nrowtab = (size_t)pl->nsl * NROWS;
    // (+64 bytes: the kernel copies table rows in 16-byte chunks)
    CUDA_TRY(cudaMalloc(&pl->d_base, nrowtab * W * sizeof(int32_t) + 64));
    CUDA_TRY(cudaMalloc(&pl->d_w, nrowtab * W * rec_size(pl->K) * pl->esize + 64));
    CUDA_TRY(cudaMalloc(&pl->d_quad, nrowtab * W * 4 * pl->esize + 64));
    CUDA_TRY(cudaMalloc(&pl->d_meta, nrowtab * W * sizeof(uint32_t) + 64));
    uint8_t *d_blockrow = nullptr;
    CUDA_TRY(cudaMalloc(&d_blockrow, nrowtab * W));
    pl->table_bytes += (int64_t)(nrowtab * W * sizeof(int32_t));
    pl->table_bytes += (int64_t)(nrowtab * W * rec_size(pl->K) * pl->esize);
    pl->table_bytes += (int64_t)(nrowtab * W * (4 * pl->esize + sizeof(uint32_t)));
    gp.meta = pl->d_meta;
    gp.blockrow = d_blockrow;
    gp.K = pl->K;
    gp.base = pl->d_base;
    gp.kmax = nullptr;
    if (pl->dtype == EMIRWV_F32)
        k_build_geometry<float><<<grid, block>>>(gp, static_cast<float *>(pl->d_w),
                                                 static_cast<float *>(pl->d_quad));
    else
        k_build_geometry<double><<<grid, block>>>(gp, static_cast<double *>(pl->d_w),
                                                  static_cast<double *>(pl->d_quad));
    cudaError_t gerr = cudaGetLastError();
    pl->h_base.resize(nrowtab * W);
    pl->h_blockrow.resize(nrowtab * W);
    if (gerr == cudaSuccess)
        gerr = cudaMemcpy(pl->h_base.data(), pl->d_base, pl->h_base.size() * sizeof(int32_t),
                          cudaMemcpyDeviceToHost);
    if (gerr == cudaSuccess)
        gerr = cudaMemcpy(pl->h_blockrow.data(), d_blockrow, pl->h_blockrow.size(),
                          cudaMemcpyDeviceToHost);
    cudaFree(d_blockrow);
    if (gerr != cudaSuccess)
        return fail(EMIRWV_E_CUDA, "k_build_geometry failed: %s", cudaGetErrorString(gerr));

    std::vector<uint8_t> covered;
    int rc = build_border_tables(pl, covered);
    if (rc) return rc;
    rc = build_tiles(pl, covered);
    if (rc == EMIRWV_E_LIMIT && pl->nstage > 2 && d.stages == 0) {
        // fewer rows in flight need fewer source rows in the rolling window
        pl->nstage = 2;
        rc = build_tiles(pl, covered);
    }
    if (rc) return rc;
    if (env_int("EMIRWV_NO_AUX", 0) == 0) {
        CUDA_TRY(cudaStreamCreateWithFlags(&pl->aux, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&pl->ev_join, cudaEventDisableTiming));
    }
    return EMIRWV_OK;
}

void release_workspace(emirwv_plan *pl) {
    for (auto &sl : pl->slots) {
        if (sl.d_in) cudaFree(sl.d_in);
        if (sl.d_out) cudaFree(sl.d_out);
        if (sl.d_jmm) cudaFree(sl.d_jmm);
        if (sl.d_raw) cudaFree(sl.d_raw);
        if (sl.ev_ready) cudaEventDestroy(sl.ev_ready);
        if (sl.ev_free) cudaEventDestroy(sl.ev_free);
        if (sl.stream) cudaStreamDestroy(sl.stream);
        sl = emirwv_plan::Slot();
    }
    pl->ws_chunk = 0;
    pl->ws_raw_bytes = 0;
}

int ensure_workspace(emirwv_plan *pl, int chunk) {
    if (pl->ws_chunk >= chunk) return EMIRWV_OK;
    release_workspace(pl);
    const size_t in_b = (size_t)pl->desc.naxis1 * pl->desc.naxis2 * pl->esize;
    const size_t out_b = (size_t)pl->nsl * pl->rps * pl->nout * pl->esize;
    for (auto &sl : pl->slots) {
        CUDA_TRY(cudaMalloc(&sl.d_in, in_b * chunk));
        CUDA_TRY(cudaMalloc(&sl.d_out, out_b * chunk));
        CUDA_TRY(cudaMalloc(&sl.d_jmm, (size_t)chunk * pl->nsl * 2 * sizeof(int32_t)));
        CUDA_TRY(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&sl.ev_ready, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&sl.ev_free, cudaEventDisableTiming));
    }
    pl->ws_chunk = chunk;
    return EMIRWV_OK;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
