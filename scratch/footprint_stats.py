import sys, numpy as np
sys.path.insert(0, '.')
from pyemir_b200 import synthetic
from pyemir_b200.plan import get_plan
for cfg in (2, 3, 4):
    coeff = synthetic.make_config(cfg)
    plan = get_plan(coeff, 2, "float32")
    tot = fast = 0; wr_tot = wr_fast = 0
    for s in (1, 14, 28, 41, 55):
        row, col, ok, w = plan.debug_geometry(s, weights=True)
        nz = w != 0                     # [39][bbw][K][K]
        K = w.shape[-1]
        arows = nz.any(axis=3); bcols = nz.any(axis=2)
        def span(m):
            idx = np.arange(K)
            lo = np.where(m, idx, K).min(axis=-1); hi = np.where(m, idx, -1).max(axis=-1)
            return np.where(hi >= lo, hi - lo + 1, 0)
        f = (span(arows) <= 2) & (span(bcols) <= 2)
        tot += f.size; fast += f.sum()
        nb = f.shape[1] // 30
        blocks = f[:, :nb * 30].reshape(39, nb, 30).all(axis=2)
        wr_tot += blocks.size; wr_fast += blocks.sum()
    print("config", cfg, "K", K, "fast pixels %.4f" % (fast / tot), "fast 30-px warp rows %.4f" % (wr_fast / wr_tot))
