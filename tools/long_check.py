"""Long-row path A/B on C2: the tcgen05 kernel (cholesky_tc.cu) against the mma.sync kernel (knob long_tc = 0 / 1) --
per-iteration Cholesky time, and the WARM user half of a second iteration against an fp64 solve on a row sample and,
row by row, against the other kernel.  SC_SCALE scales the configuration (default 1.0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("OPENBLAS_NUM_THREADS", "8")
import numpy as np
from helpers import row_err
from implicit_b200 import _lib, synthetic
ctx = _lib.Context(0)
scale = float(os.environ.get("SC_SCALE", "1.0"))
Cui, X0, Y0, cfg = synthetic.config("C2", scale=scale)
C = _lib.DeviceCSR.upload(ctx, Cui); T = C.transpose()
X, Y = _lib.DeviceFactors.from_host(ctx, X0), _lib.DeviceFactors.from_host(ctx, Y0)
deg = np.diff(Cui.indptr)
print(f"C2 x{scale}: user rows > 48 nnz: {(deg > 48).sum()} of {len(deg)} holding {deg[deg > 48].sum()} of {deg.sum()} nnz", flush=True)
sample = np.arange(0, cfg["users"], 197)
truth = None
base = None
for tc in (int(a) for a in (sys.argv[1:] or ["0", "1"])):
    ctx.set_knob("long_tc", tc)
    ctx.profile(True)
    for it in range(4):  # 3 timed cold-start iterations (same state every time, like bench.py's device arm)
        X.upload(X0); Y.upload(Y0)
        if it == 1: ctx.profile_read()
        _lib.least_squares(ctx, C, X, Y, 0.01); _lib.least_squares(ctx, T, Y, X, 0.01)
    p = ctx.profile_read()
    ctx.profile(False)
    ms = (p["cholesky"][0] + p["cholesky_finish"][0]) / 3
    Yin = Y.download()
    _lib.least_squares(ctx, C, X, Y, 0.01)
    got = X.download()
    if True:  # the truth of THIS setting's own inputs (they differ between settings by the rounding of two iterations)
        Y64 = Yin.astype(np.float64); G64 = Y64.T @ Y64
        truth = np.zeros((len(sample), 64))
        for n, u in enumerate(sample):
            s, t = Cui.indptr[u], Cui.indptr[u + 1]
            if s == t: continue
            Yu, c = Y64[Cui.indices[s:t]], Cui.data[s:t].astype(np.float64)
            truth[n] = np.linalg.solve(G64 + 0.01 * np.eye(64) + (Yu.T * (np.abs(c) - 1)) @ Yu, Yu.T @ np.where(c > 0, c, 0))
    e = row_err(got[sample], truth)
    lng = deg[sample] > 48
    msg = (f"long_tc={tc} cholesky {ms:.3f} ms/iter | warm user half vs fp64: max {e.max():.2e} median {np.median(e):.2e} "
           f"(rows > 48: max {e[lng].max():.2e} median {np.median(e[lng]):.2e})")
    d_s = deg[sample]
    msg += " | median by nnz: " + ", ".join(
        f"{lo}-{hi}: {np.median(e[(d_s > lo) & (d_s <= hi)]):.1e}" for lo, hi in ((48, 96), (96, 192), (192, 512), (512, 4096)) if ((d_s > lo) & (d_s <= hi)).any())
    if base is None:
        base = got
    else:
        d = row_err(got, base)
        msg += f" | vs first setting, all rows: max {d.max():.2e} median {np.median(d):.2e}; rows > 48: max {d[deg > 48].max():.2e}"
    print(msg, flush=True)
