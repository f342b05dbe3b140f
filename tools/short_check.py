"""Short-row path (cholesky_short.cu) A/B on C2: per-iteration Cholesky time and parity for
ALS_B200_SHORT_MAX in {0 (off), 16, 32, 48}.  The WARM user half of a second iteration is compared with an fp64
solve on a row sample and, row by row, with the limit-0 result."""
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
print(f"C2 x{scale}: user rows <=16: {(deg<=16).sum()}, <=32: {(deg<=32).sum()}, <=48: {(deg<=48).sum()} of {len(deg)}", flush=True)
sample = np.arange(0, cfg["users"], 197)
truth = None
base = None
for lim in (sys.argv[1:] or ["0", "32", "48s", "48"]):
    ctx.set_knob("short_max", int(lim.rstrip("s")))
    ctx.set_knob("short_serial", 1 if lim.endswith("s") else 0)  # "48s": short-row kernels on the compute stream
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
    if truth is None:  # inputs of this half differ between limits only by rounding; one truth serves all
        Y64 = Yin.astype(np.float64); G64 = Y64.T @ Y64
        truth = np.zeros((len(sample), 64))
        for n, u in enumerate(sample):
            s, t = Cui.indptr[u], Cui.indptr[u + 1]
            if s == t: continue
            Yu, c = Y64[Cui.indices[s:t]], Cui.data[s:t].astype(np.float64)
            truth[n] = np.linalg.solve(G64 + 0.01 * np.eye(64) + (Yu.T * (np.abs(c) - 1)) @ Yu, Yu.T @ np.where(c > 0, c, 0))
    e = row_err(got[sample], truth)
    short = deg[sample] <= 48
    msg = f"SHORT_MAX={lim:3s} cholesky {ms:.3f} ms/iter | warm user half vs fp64: max {e.max():.2e} median {np.median(e):.2e} (rows<=48: max {e[short].max():.2e}; longer: max {e[~short].max():.2e})"
    if base is None:
        base = got
    else:
        d = row_err(got, base)
        msg += f" | vs first setting, all rows: max {d.max():.2e} median {np.median(d):.2e}"
    print(msg, flush=True)
