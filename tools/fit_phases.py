"""Where does an end-to-end fit(host CSR) spend its wall time?  (C2, 3 iterations)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from implicit_b200 import _lib, synthetic, AlternatingLeastSquares
import bench
ctx = _lib.Context(0)
Cui, X0, Y0, cfg = synthetic.config("C2")
Cpin = bench.pinned_csr(Cui)
X0p, Y0p = _lib.pinned_empty(X0.shape, np.float32), _lib.pinned_empty(Y0.shape, np.float32)
X0p[:], Y0p[:] = X0, Y0
def T(label, fn):
    ctx.sync(); t = time.perf_counter(); r = fn(); ctx.sync(); print(f"  {label:40s} {1e3*(time.perf_counter()-t):8.2f} ms", flush=True); return r
for rep in range(2):
    print("rep", rep)
    C = T("upload CSR (pinned) + schedule", lambda: _lib.DeviceCSR.upload(ctx, Cpin))
    C2 = T("upload CSR (pageable) + schedule", lambda: _lib.DeviceCSR.upload(ctx, Cui))
    Tt = T("device transpose + schedule", lambda: C.transpose())
    X = T("create+upload X (pageable)", lambda: _lib.DeviceFactors.from_host(ctx, X0))
    Y = T("create+upload Y", lambda: _lib.DeviceFactors.from_host(ctx, Y0))
    T("3 iterations", lambda: [(_lib.least_squares(ctx, C, X, Y, 0.01), _lib.least_squares(ctx, Tt, Y, X, 0.01)) for _ in range(3)])
    xs = T("download X", lambda: X.download())
    ys = T("download Y", lambda: Y.download())
    T("download X pinned", lambda: X.download(pinned=True))
    T("download Y pinned", lambda: Y.download(pinned=True))
    T("upload X pinned", lambda: _lib.DeviceFactors.from_host(ctx, X0p))
    T("has_nan X", lambda: X.has_nan())
    T("isnan check", lambda: (np.isnan(xs).any(), np.isnan(ys).any()))
    m = AlternatingLeastSquares(factors=64, use_cg=False, iterations=3)
    m._ctx = ctx
    m.user_factors, m.item_factors = X0p, Y0p
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    T("model.fit total", lambda: m.fit(Cpin, show_progress=False))
    pr.disable()
    if rep == 1:
        pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
    T("model factors to host", lambda: (m.user_factors, m.item_factors))
