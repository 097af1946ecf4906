// The hot kernel k_rectwv (distortion map + rectify2d gather, Steffen-limited flux-conserving
// rebin, masking / assembly) and its tap gather.
// Included by emirwv.cu inside its anonymous namespace (one translation unit).

// Weighted gather of KT x KT taps for the G frames of an item: rowp[a] points at tap (a, 0)
// of frame 0 in the rolling window, frame g lies g * FRAME elements further.  float32 with
// an even G runs two frames per instruction.  Taps are accumulated row by row, left to right.
// (Lanes without a pixel of their own gather through a neighbour's record: a finite value
// that only ever reaches knots nobody reads.)
template <typename T, int KT, int G, int FRAME>
__device__ __forceinline__ void gather_taps(const T *const (&rowp)[KT], const T *wv, int wstride,
                                            T (&rect)[G]) {
    if constexpr (sizeof(T) == 4 && (G % 2) == 0) {
        float2 acc[G / 2];
#pragma unroll
        for (int h = 0; h < G / 2; ++h) acc[h] = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int a = 0; a < KT; ++a)
#pragma unroll
            for (int b = 0; b < KT; ++b) {
                const float2 wgt = dup2(wv[a * wstride + b]);
#pragma unroll
                for (int h = 0; h < G / 2; ++h)
                    acc[h] = fma2(wgt,
                                  make_float2(rowp[a][(2 * h) * FRAME + b],
                                              rowp[a][(2 * h + 1) * FRAME + b]),
                                  acc[h]);
            }
#pragma unroll
        for (int h = 0; h < G / 2; ++h) {
            rect[2 * h] = acc[h].x;
            rect[2 * h + 1] = acc[h].y;
        }
    } else {
#pragma unroll
        for (int g = 0; g < G; ++g) {
            T acc = T(0);
#pragma unroll
            for (int a = 0; a < KT; ++a)
#pragma unroll
                for (int b = 0; b < KT; ++b)
                    acc = fma(wv[a * wstride + b], rowp[a][g * FRAME + b], acc);
            rect[g] = acc;
        }
    }
}

// ---------------------------------------------------------------------------------
// The hot kernel.
//
// Persistent: CTA b processes work items b, b + gridDim.x, ...; an item is a live
// tile (slitlet x block of <= TILE_K output columns) for G consecutive frames, frame
// group fastest so that CTAs running at the same time share the tile's tables in L2.
//
// Inside an item the CTA marches down the 39 scanned rows of the slitlet.  Every
// compute thread keeps, for the whole item, ONE rectified pixel column (phases A and C)
// and ONE border of the output grid (phase D), so the per-column tables (pixel widths,
// border positions) sit in registers and only the per-row table records are loaded.
// A warp owns the borders of its piece of the tile and every pixel they touch, so the
// rows need no exchange between warps and no block barrier.  Per row and warp:
//   wait   full[y]: the source rows and the table row of row y have landed (TMA)
//   D      (row y-1) cumulative-flux difference at the thread's border and the next one
//          (shuffle), store, non-zero flag -- from the warp's own knot rows
//   A      rectified flux of G frames: K*K weighted gather from the rolling window
//   C      Steffen-limited end slopes from the neighbours' flux (warp shuffles: a warp
//          covers 32 pixels and produces 30), written to the warp's knot rows
//   arrive done[y]
// The copy warp waits for done[y] (all compute warps) and requests the table row (one bulk
// copy: the plan keeps the tables tile-major) and the new source rows of row y+NSTAGE (one
// TMA tensor copy per source row: a box of WIN_PITCH columns x G frames, zeros outside the
// detector) into the slots row y has just released.
// The source window is a ring of WIN_ROWS rows, each holding the G frames side by side.
//
// Shared memory (128-byte aligned): [window: WIN_ROWS ring rows x G frames x WIN_PITCH]
//                [per warp: (rect | ya) x 32 pixels x G]
//                [NSTAGE staged table rows: PXCAP x quad (4 weights), PXCAP meta words]
//                [zeros for the fill stores][tile descriptor ring: 3][mbarriers]
// ---------------------------------------------------------------------------------
template <typename T, int K, int G, bool SPANLOOP, int NSTAGE>
__global__ void __launch_bounds__(NTHREADS, sizeof(T) == 4 ? EMIRWV_MINBLOCKS : EMIRWV_MINBLOCKS_F64)
    k_rectwv(const RunParams<T> p, const __grid_constant__ CUtensorMap tmap) {
    static_assert(NSTAGE >= 2 && NSTAGE <= NSTAGE_MAX, "stage count");
    extern __shared__ unsigned char smem_dyn[];
    // (tensor copies want a 128-byte aligned destination)
    unsigned char *smem_raw = smem_dyn + ((128u - (smem_addr_of(smem_dyn) & 127u)) & 127u);
    constexpr int WROW = G * WIN_PITCH;                        // one ring row: G frames side by side
    constexpr int WSZ = WIN_ROWS * WROW;                       // the window
    constexpr int KNOT_WARP = 2 * 32 * G;                      // knot elements of one warp
    constexpr unsigned ROW_BYTES = WROW * sizeof(T);           // one tensor copy
    constexpr unsigned STAGE_BYTES = stage_bytes(sizeof(T));   // one staged table row

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const bool producer = warp == NWARPS;          // the copy warp computes nothing
    const size_t frame_out = (size_t)p.nrows_out * p.nout;
    const int nwork = p.nlive_tiles * p.ngroups;

    T *win = reinterpret_cast<T *>(smem_raw);                              // [WIN_ROWS][G][WIN_PITCH]
    T *knots = win + WSZ + min(warp, NWARPS - 1) * KNOT_WARP;   // this warp: [rect|ya][32][G]
    constexpr int RS = rec_size(K);
    unsigned char *stg = reinterpret_cast<unsigned char *>(win + WSZ + NWARPS * KNOT_WARP);
    // staged table row `st`: quad [PXCAP][4] T at stg + st * STAGE_BYTES, then meta [PXCAP]
    uint32_t *zbuf = reinterpret_cast<uint32_t *>(stg + NSTAGE * STAGE_BYTES);   // ZFILL_BYTES of zeros
    Tile *ring = reinterpret_cast<Tile *>(reinterpret_cast<unsigned char *>(zbuf) + ZFILL_BYTES);   // [3]
    // mbarriers: "full" and "done", one pair per rectified row in flight
    const unsigned bars = smem_addr_of(ring + 3);
    // float32 with an even number of frames: two frames per arithmetic instruction
    constexpr bool PACKED = sizeof(T) == 4 && (G % 2) == 0;
    using KV = KnotVec<T, G>;

    // Work distribution: the first two items of a CTA are b and b + gridDim.x; with p.sched the
    // following ones come from a counter shared by the CTAs (items differ in cost: tiles at the
    // ends of a slitlet are narrower), else the stride continues.  wq[] holds the item numbers
    // of iterations it .. it+2 (thread 0 draws the number of iteration it+3 before the barrier
    // that ends iteration it).
    __shared__ int wq[4];
    int w = blockIdx.x;
    if (w >= nwork) return;
    if (tid == 0) {
        wq[0] = w;
        wq[1] = w + (int)gridDim.x;
        wq[2] = p.sched != nullptr ? 2 * (int)gridDim.x + atomicAdd(p.sched, 1)
                                   : w + 2 * (int)gridDim.x;
    }

    // descriptors of the first two items
    constexpr int TCH = sizeof(Tile) / 16;
    if (tid < TCH)
        cp_async_16(reinterpret_cast<char *>(&ring[0]) + tid * 16,
                    reinterpret_cast<const char *>(p.tiles + w / p.ngroups) + tid * 16, true);
    if (w + (int)gridDim.x < nwork && tid >= 32 && tid < 32 + TCH)
        cp_async_16(reinterpret_cast<char *>(&ring[1]) + (tid - 32) * 16,
                    reinterpret_cast<const char *>(p.tiles + (w + gridDim.x) / p.ngroups) +
                        (tid - 32) * 16, true);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NSTAGE; ++i) {
            mbar_init(bars + 8 * i, 2);                 // "full": rows part + table part of a row
            mbar_init(bars + 8 * NSTAGE + 8 * i, NWARPS);       // "done": every compute warp is through
        }
        mbar_fence_init();
    }
    // the window starts out finite: taps with weight zero may read slots no copy has filled
    // (generic stores; the bulk copies that later land on the same addresses belong to the
    // async proxy, hence the proxy fence before the barrier that releases the copy warp)
    for (int e = tid; e < WSZ; e += NTHREADS) win[e] = T(0);
    if (tid == NTHREADS - 1) prefetch_tensormap(&tmap);
    for (int e = tid; e < ZFILL_BYTES / 4; e += NTHREADS) zbuf[e] = 0u;
    fence_proxy_async_smem();
    cp_async_wait_all();
    __syncthreads();

    unsigned fo = (unsigned)frame_out;                         // < 2^32 elements per frame
    pin(fo);
    unsigned rowseq = 0;            // rectified rows started so far (same in every thread)


    for (int it = 0;; ++it) {
        const Tile &tile = ring[it % 3];
        const int wnext = wq[(it + 1) & 3], wnn = wq[(it + 2) & 3];
        // descriptor of the item after the next one rides along with this item's copies
        if (wnn < nwork && tid < TCH)
            cp_async_16(reinterpret_cast<char *>(&ring[(it + 2) % 3]) + tid * 16,
                        reinterpret_cast<const char *>(p.tiles + wnn / p.ngroups) + tid * 16,
                        true);

        const int grp = w % p.ngroups;
        const int g0 = grp * G;
        const int nlive = min(G, p.nframes - g0);
        const int s = tile.slit;
        const int npx = tile.npx;
        const int k0 = tile.k0;

        // ---- per-item constants of this thread
        // The warp owns the borders kw[warp] .. kw[warp+1]-1 of the tile (at most 31) and
        // every rectified pixel they touch (at most 30, lanes 1..30; lanes 0 and 31 hold the
        // neighbours the slope limiter needs), so a row needs no exchange between warps.
        const int wc = min(warp, NWARPS - 1);
        const int kw0 = tile.kw[wc], nb = tile.kw[wc + 1] - kw0;
        const bool wwarp = nb > 0 && !producer;                // the warp has work at all
        const int q = kw0 + min(lane, nb);                     // border, tile relative
        const Border<T> bd = ldg_struct(p.border + (size_t)s * p.NB + k0 + q);
        const int ibl = bd.idx - tile.jlo;                     // source pixel, tile relative
        const int ibn = __shfl_down_sync(0xffffffffu, ibl, 1);
        const int plo = __shfl_sync(0xffffffffu, ibl, 0);      // first pixel the warp owns
        const int kl = ibl - plo + 1;                          // lane that holds pixel ibl
        const bool col_ok = lane < nb;
        const bool has1 = ibn > ibl, has2 = ibn > ibl + 1;
        const int kb = k0 + q;
        const int pp = plo - 1 + lane;                         // pixel column, tile relative
        const bool px_ok = pp >= 0 && pp < npx;
        PixGeom<T> pg;
        pg.qm = pg.qp = pg.r0 = pg.r1 = T(0);
        if (px_ok) pg = ldg_struct(p.pix + (size_t)s * W + tile.jlo + pp);
        const float qm2 = (float)(pg.qm + pg.qm), hr0 = (float)(T(0.5) * pg.r0);   // packed path
        // width ratio that turns the next pixel's left-knot slope into this border's pixel's
        // right-knot slope (the table holds 0 at the red end of the spectrum)
        const T qpb = __ldg(&p.pix[(size_t)s * W + bd.idx].qp);
        const unsigned stmask = col_ok ? (1u << nlive) - 1u : 0u;   // frames this lane stores
#if EMIRWV_DIAG & 2
        T dreg[G];
#pragma unroll
        for (int g = 0; g < G; ++g) dreg[g] = T(0);
#endif
        unsigned nzb[G];                                       // OR of the value bits seen
#pragma unroll
        for (int g = 0; g < G; ++g) nzb[g] = 0;

        // ---- the copy pipeline of the item (TMA copies completing on mbarriers): the new
        //   source rows that rectified row y needs, one tensor copy per row (all G frames)
        //   into the row's ring slot, and the staged table row of row y, one bulk copy into
        //   stage y % NSTAGE.
        constexpr int VN = Vec16<T>::N;
        const unsigned stg_a = smem_addr_of(stg), win_a = smem_addr_of(win);
        // what the copy warp needs to know about an item (this one, or the next one whose
        // first rows it requests as soon as this item's last row is through)
        struct CopyCtx {
            const char *trow;      // the tile's staged table rows (tile-major, STAGE_BYTES each)
            int c0, g0;            // first window column on the detector, first frame
        };
        auto make_ctx = [&](const Tile &t, int wi) {
            CopyCtx c;
            c.trow = p.tilepack + (size_t)(wi / p.ngroups) * NROWS * STAGE_BYTES;
            c.c0 = t.c0;
            c.g0 = (wi % p.ngroups) * G;
            return c;
        };

        // Row number n (counted across items) completes on mbarrier n % NSTAGE with two
        // arrivals: the source rows it adds and its table row.  Lane 0 posts the byte count,
        // lane l requests source row r_from + 1 + l: the box [c0, c0 + WIN_PITCH) x {r} x
        // [g0, g0 + G) of the (column, row, frame) tensor -- columns, rows and frames that do
        // not exist arrive as zeros, so detector edges and a short last group need no code.
        auto copy_window_rows = [&](const CopyCtx &c, unsigned seq, int r_from, int r_to) {
            const unsigned bar = bars + 8 * (seq % NSTAGE);
            const int n = r_to - r_from;
            if (lane == 0) mbar_expect_tx(bar, (unsigned)max(n, 0) * ROW_BYTES);
            __syncwarp();
            if (lane < n) {
                const int r = r_from + 1 + lane;
                tensor_g2s_3d(win_a + ring_slot(r) * ROW_BYTES, &tmap, c.c0, r, c.g0, bar);
            }
        };
        // table row y of the tile: block weights and meta words, one bulk copy
        auto copy_table_row = [&](const CopyCtx &c, unsigned seq, int y) {
            if (lane == 0) {
                const unsigned bar = bars + 8 * (seq % NSTAGE);
                mbar_expect_tx(bar, STAGE_BYTES);
                bulk_g2s(stg_a + (y % NSTAGE) * STAGE_BYTES, c.trow + (size_t)y * STAGE_BYTES,
                         STAGE_BYTES, bar);
            }
        };
        // the first NSTAGE rows of an item and their table rows
        auto copy_first_rows = [&](const CopyCtx &c, const Tile &t, unsigned seq0) {
            for (int y = 0; y < NSTAGE; ++y) {
                copy_window_rows(c, seq0 + y,
                                 y == 0 ? span_lo(t.span[0]) - 1 : span_hi(t.span[y - 1]),
                                 span_hi(t.span[y]));
                copy_table_row(c, seq0 + y, y);
            }
        };

        // The first NSTAGE rows and their table rows are requested up front -- for the first
        // item of the CTA here, for every later one by the copy warp at the end of the item
        // before it (below); once every compute warp is through row y the copy warp requests
        // the table row and the source rows of row y+NSTAGE (nobody reads the slots they
        // overwrite any more)
        if (producer && (it == 0 || !p.early)) copy_first_rows(make_ctx(tile, w), tile, rowseq);
        // this thread's slot in the staged table rows (lanes without a pixel read a
        // neighbour's entry; their result is discarded)
        const int ppc = min(max(pp, 0), npx - 1) + (tile.jlo & 3);      // column of the staged row
        const T *qsm = reinterpret_cast<const T *>(stg) + ppc * 4;      // this pixel, stage 0
        const uint32_t *msm = reinterpret_cast<const uint32_t *>(stg + PXCAP * 4 * sizeof(T)) + ppc;
        const T *grec = p.w + ((size_t)s * NROWS * W + tile.jlo + min(max(pp, 0), npx - 1)) * RS;
        pin(grec);
        T *orow = p.out + (size_t)g0 * frame_out + (size_t)s * p.rps * p.nout;
        int wb = -tile.c0;                                     // window addressed by frame coords
        pin(wb);

        // ---- D (row yr): cumulative flux at this thread's border and at the next one,
        //      difference = output pixel; store; remember non-zero values
        auto phase_d = [&](auto store_row) {                   // the 39th row is only scanned
            constexpr bool store = decltype(store_row)::value;
            const KV *kbase = reinterpret_cast<const KV *>(knots) + kl;
            const KV *kb_rect = kbase, *kb_ya = kbase + 32;
            // flux and left-knot slope of the border's pixel; its right-knot slope is the
            // next pixel's left-knot slope rescaled by the ratio of the pixel widths
#if EMIRWV_DIAG & 2
            KV kr, ka, kn, k1;                                 // DIAGNOSTIC: no knot loads
#pragma unroll
            for (int g = 0; g < G; ++g) kr.v[g] = dreg[g], ka.v[g] = dreg[g] * T(0.5), kn.v[g] = dreg[g] * T(0.25), k1.v[g] = T(0);
            (void)kb_rect; (void)kb_ya;
#else
            const KV kr = kb_rect[0], ka = kb_ya[0], kn = kb_ya[1];
            KV k1;
#pragma unroll
            for (int g = 0; g < G; ++g) k1.v[g] = T(0);
            if (has2) k1 = kb_rect[1];
#endif
            T o[G];
            if constexpr (PACKED) {
                const float2 tau2 = dup2(bd.tau);
#pragma unroll
                for (int h = 0; h < G / 2; ++h) {
                    const float2 r = make_float2(kr.v[2 * h], kr.v[2 * h + 1]);
                    const float2 a = make_float2(ka.v[2 * h], ka.v[2 * h + 1]);
                    const float2 c = mul2(make_float2(kn.v[2 * h], kn.v[2 * h + 1]), dup2(qpb));
                    // steffen_partial_flux, two frames at a time
                    const float2 A = fma2(r, dup2(-2.0f), add2(a, c));
                    const float2 B = sub2(sub2(r, a), A);
                    const float2 P = mul2(tau2, fma2(tau2, fma2(tau2, A, B), a));
                    const float2 Pn = make_float2(__shfl_down_sync(0xffffffffu, P.x, 1),
                                                  __shfl_down_sync(0xffffffffu, P.y, 1));
                    float2 whole = make_float2(has1 ? r.x : 0.0f, has1 ? r.y : 0.0f);
                    if (has2) whole = add2(whole, make_float2(k1.v[2 * h], k1.v[2 * h + 1]));
                    if (SPANLOOP)
                        for (int i = 2; i < ibn - ibl; ++i) {
                            const KV ki = kb_rect[i];
                            whole = add2(whole, make_float2(ki.v[2 * h], ki.v[2 * h + 1]));
                        }
                    const float2 o2 = add2(whole, sub2(Pn, P));
                    o[2 * h] = o2.x;
                    o[2 * h + 1] = o2.y;
                }
            } else {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const T P = steffen_partial_flux(kr.v[g], ka.v[g], kn.v[g] * qpb, bd.tau);
                    const T Pn = __shfl_down_sync(0xffffffffu, P, 1);
                    T whole = has1 ? kr.v[g] : T(0);
                    if (has2) whole += k1.v[g];
                    if (SPANLOOP)
                        for (int i = 2; i < ibn - ibl; ++i) whole += kb_rect[i].v[g];
                    o[g] = whole + (Pn - P);
                }
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
                T *dst = orow + ((unsigned)kb + (unsigned)g * fo);
                if (store && ((stmask >> g) & 1u)) __stcs(dst, o[g]);
                nzb[g] |= nz_bits(o[g]);       // lanes / frames that do not count are masked
            }                                  // once, after the last row
            orow += p.nout;
        };

        // the warp has read everything row y needs from the window and the table slot: the
        // copy warp may overwrite them
        auto row_done = [&]() {
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 8 * NSTAGE + 8 * (rowseq % NSTAGE));
            ++rowseq;
        };

        // ---- A, C (row y)
        auto phase_ac = [&](int y) {
            using V = typename Vec16<T>::type;
            constexpr int KK = K * K;
            // this pixel's meta word: source column of its 2 x 2 block, ring slots of the
            // block's two source rows, "the block is the whole footprint"
            const unsigned soff = (unsigned)(y % NSTAGE) * STAGE_BYTES;     // this row's stage
            const uint32_t meta =
                *reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(msm) + soff);
            const int fc = (int)(int16_t)(meta & 0xffffu) + wb;

            // ---- A (row y): rectified flux of this pixel column, G frames
            T rect[G];
            bool all_compact = true;
            if constexpr (K >= 3) all_compact = __all_sync(0xffffffffu, meta & META_COMPACT);
            if (all_compact) {
                // every footprint of the warp fits 2 x 2 source pixels: four taps (K = 1: one)
                T wt[4];
                const V *q = reinterpret_cast<const V *>(reinterpret_cast<const char *>(qsm) + soff);
#pragma unroll
                for (int i = 0; i < 4 / VN; ++i) *reinterpret_cast<V *>(wt + i * VN) = q[i];
#if EMIRWV_DIAG & 1
                constexpr int KB = 1;                          // DIAGNOSTIC: one tap instead of four
#else
                constexpr int KB = K == 1 ? 1 : 2;
#endif
                const T *rowp[KB];
                rowp[0] = win + ((meta >> 16) & 0xfu) * WROW + fc;
                if constexpr (KB == 2) rowp[1] = win + ((meta >> 20) & 0xfu) * WROW + fc;
                gather_taps<T, KB, G, WIN_PITCH>(rowp, wt, 2, rect);
            } else if constexpr (K >= 3) {
                // (rare: 12 % of the warp rows) the full K x K records, from global memory (L2)
                const V *rec = reinterpret_cast<const V *>(grec + (size_t)y * W * RS);
                T wt[RS];
#pragma unroll
                for (int i = 0; i < RS / VN; ++i) *reinterpret_cast<V *>(wt + i * VN) = __ldg(rec + i);
#if EMIRWV_RECPF
                // a warp that needs the full records in this row mostly needs them in the next
                // one too: ask for them now (L2 -> L1), so that the next row does not stand
                // still for an L2 round trip -- and with it the row's "done" barrier, the copy
                // warp and every other warp of the CTA
                if (y + 1 < NROWS) {
                    const char *nx = reinterpret_cast<const char *>(rec) + (size_t)W * RS * sizeof(T);
                    prefetch_l1(nx);
                    prefetch_l1(nx + RS * sizeof(T) - 4);
                }
#endif
                const int32_t pk = *reinterpret_cast<const int32_t *>(wt + KK);
                const uint32_t slots = *reinterpret_cast<const uint32_t *>(wt + KK + 1);
                const int fcf = base_col(pk) + wb;
                if ((slots >> 24) & REC_COMPACT) {             // spread the block over K x K
                    const T c00 = wt[0], c01 = wt[1], c10 = wt[2], c11 = wt[3];
#pragma unroll
                    for (int i = 0; i < KK; ++i) wt[i] = T(0);
                    wt[0] = c00, wt[1] = c01, wt[K] = c10, wt[K + 1] = c11;
                }
                const T *rowp[K];
#pragma unroll
                for (int a = 0; a < K; ++a) {
                    const uint32_t slot = a < 3 ? (slots >> (8 * a)) & 0xffu
                                                : *reinterpret_cast<const uint32_t *>(wt + KK + 2);
                    rowp[a] = win + slot * WROW + fcf;
                }
                gather_taps<T, K, G, WIN_PITCH>(rowp, wt, K, rect);
            }

            row_done();       // (phase C works on registers and the warp's own knot rows)

            // ---- C (row y): Steffen-limited slopes at both ends of the pixel, scaled
            //      by the pixel width (flux units).  qm = 0 / qp = 0 in the table give
            //      the zero slopes at the two ends of the spectrum.
            KV vr, va;
            if constexpr (PACKED) {
                // steffen_slope with everything doubled (exact): 2*min(|sl|,|sr|,|p|/2) =
                // min(|2 sl|, |2 sr|, |p|); the sign test folds into the minimum by
                // giving 2 sl the sign it has relative to sr.
#pragma unroll
                for (int h = 0; h < G / 2; ++h) {
                    const float2 r = make_float2(rect[2 * h], rect[2 * h + 1]);
                    const float2 l = make_float2(__shfl_up_sync(0xffffffffu, r.x, 1),
                                                 __shfl_up_sync(0xffffffffu, r.y, 1));
                    const float2 cl2 = mul2(l, dup2(qm2));
                    const float2 rr = add2(r, r);
                    const float2 pq = fma2(dup2(hr0), sub2(cl2, rr), r);
                    float ya[2];
                    const float clv[2] = {cl2.x, cl2.y}, rv[2] = {r.x, r.y}, rrv[2] = {rr.x, rr.y},
                                pv[2] = {pq.x, pq.y};
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const unsigned sgn = __float_as_uint(rv[e]) & 0x80000000u;
                        const float x = __uint_as_float(__float_as_uint(clv[e]) ^ sgn);
                        const float m = fmaxf(fminf(fminf(x, fabsf(rrv[e])), fabsf(pv[e])), 0.0f);
                        ya[e] = __uint_as_float(__float_as_uint(m) | sgn);
                    }
                    vr.v[2 * h] = r.x, vr.v[2 * h + 1] = r.y;
                    va.v[2 * h] = ya[0], va.v[2 * h + 1] = ya[1];
                }
            } else {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    // slope at the left knot from this pixel and its left neighbour (one
                    // limiter per pixel)
                    const T cl = __shfl_up_sync(0xffffffffu, rect[g], 1) * pg.qm;
                    vr.v[g] = rect[g];
                    va.v[g] = steffen_slope(cl, rect[g], pg.r0);
                }
            }
            // the warp's own knot rows (lane 0 holds an incomplete slope, never read)
            // (one buffer is enough: the warp's border phase of the previous row comes first
            // in program order)
#if EMIRWV_DIAG & 2
#pragma unroll
            for (int g = 0; g < G; ++g) dreg[g] = vr.v[g] + va.v[g];   // DIAGNOSTIC: no knot stores
#else
            KV *kdst = reinterpret_cast<KV *>(knots) + lane;
            kdst[0] = vr;
            kdst[32] = va;
#endif
        };

        // Row loop without block barriers.  A warp waits for the data of row y, evaluates
        // the borders of row y-1 from its own knots (phase D) and the flux and slopes of
        // row y (phases A, C: independent instruction streams in one basic block), then
        // reports row y done; warps may drift apart by up to NSTAGE-1 rows.
        if (producer) {
            const CopyCtx ctx = make_ctx(tile, w);
            // The runs without wavelength coverage (and the missing slitlets) are exact zeros:
            // the item's share of them -- runs (dead tile x output row) x frames -- is written
            // by TMA bulk stores from the zeros in shared memory, a few per rectified row so
            // that the writes trickle along with the item and the copy warp stays responsive.
            // The lanes fetch 32 runs at a time, a batch ahead of their use: the stores never
            // wait for a load.  (frem: runs left, fpos: lanes of the batch already used,
            // fper: runs per row)
            long long fu = 0;
            int frem = 0, fpos = 0, fper = 0;
            int2 frun = make_int2(0, 0);
            T *fdst = p.out;
            auto fill_fetch = [&]() {
                if (lane < frem) {
                    const long long u = fu + lane;
                    frun = __ldg(p.fillruns + (int)(u % p.nfillruns));
                    fdst = p.out + (size_t)(u / p.nfillruns) * frame_out + (unsigned)frun.x;
                }
            };
            auto fill_some = [&](int want) {
                const int take = min(want, min(32 - fpos, frem));
                if (lane >= fpos && lane < fpos + take) {
                    char *dst = reinterpret_cast<char *>(fdst);
                    const unsigned zsrc = smem_addr_of(zbuf);
                    for (unsigned off = 0; off < (unsigned)frun.y; off += ZFILL_BYTES)
                        bulk_s2g(dst + off, zsrc, min((unsigned)ZFILL_BYTES, (unsigned)frun.y - off));
                    bulk_commit();
                }
                fpos += take;
                frem -= take;
                if (fpos == 32 && frem > 0) {
                    fu += 32;
                    fpos = 0;
                    fill_fetch();
                }
            };
#if !EMIRWV_NOFILL
            if (p.fillruns != nullptr) {
                const long long total = (long long)p.nfillruns * p.nframes;
                fu = (long long)w * total / nwork;
                frem = (int)((long long)(w + 1) * total / nwork - fu);
                fper = (frem + NROWS - 3) / (NROWS - 2);
                fill_fetch();
            }
#endif
#pragma unroll 1
            for (int y = 0; y < NROWS; ++y) {
                mbar_wait(bars + 8 * NSTAGE + 8 * (rowseq % NSTAGE), (rowseq / NSTAGE) & 1);  // row y is done
                if (y + NSTAGE < NROWS) {
                    copy_table_row(ctx, rowseq + NSTAGE, y + NSTAGE);
                    copy_window_rows(ctx, rowseq + NSTAGE, span_hi(tile.span[y + NSTAGE - 1]),
                                     span_hi(tile.span[y + NSTAGE]));
                }
                ++rowseq;
#if !EMIRWV_NOFILL
                if (frem > 0) fill_some(fper);
#endif
            }
#if !EMIRWV_NOFILL
            while (frem > 0) fill_some(32);
#endif
            // Every compute warp has read what the last row needs: the window and the table
            // slots are free, so the first rows of the CTA's next item are requested now
            // -- while the compute warps still evaluate the borders of the last row -- and not
            // after the barrier that ends the item.  (Its descriptor landed one item ago; the
            // columns outside the detector that the next item zero-fills are not touched by
            // these copies.)
            if (p.early && wnext < nwork) {
                const Tile &tn = ring[(it + 1) % 3];
                copy_first_rows(make_ctx(tn, wnext), tn, rowseq);
            }
        } else {
            mbar_wait(bars + 8 * (rowseq % NSTAGE), (rowseq / NSTAGE) & 1);   // rows + table landed
            if (wwarp)
                phase_ac(0);
            else
                row_done();
#pragma unroll 1
            for (int y = 1; y < NROWS; ++y) {
                mbar_wait(bars + 8 * (rowseq % NSTAGE), (rowseq / NSTAGE) & 1);
                if (wwarp) {
                    phase_d(std::true_type{});
                    phase_ac(y);                               // reports the row done itself
                } else {
                    row_done();
                }
            }
        }
        if (wwarp) phase_d(std::false_type{});

        // ---- first / last non-zero channel of the slitlet (JMNSLTnn / JMXSLTnn)
        if (p.jmm != nullptr && wwarp) {
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const bool nz = col_ok && g < nlive && nzb[g] != 0;
                const int lo = __reduce_min_sync(0xffffffffu, nz ? kb : INT_MAX);
                const int hi = __reduce_max_sync(0xffffffffu, nz ? kb : -1);
                if (lane == 0 && hi >= 0) {
                    int32_t *dst = p.jmm + ((size_t)(g0 + g) * p.nsl + s) * 2;
                    atomicMin(dst, lo);
                    atomicMax(dst + 1, hi);
                }
            }
        }

        if (wnext >= nwork) break;
        w = wnext;
        if (tid == 0)
            wq[(it + 3) & 3] = p.sched != nullptr ? 2 * (int)gridDim.x + atomicAdd(p.sched, 1)
                                                  : wnn + (int)gridDim.x;
        cp_async_wait_all();
        __syncthreads();          // next descriptor landed; window and table slots are free
    }
    if (producer) bulk_wait_all();   // the zero runs are written before the CTA retires
}
