"""Short-row path (cholesky_short.cu) A/B on C2: per-iteration Cholesky time and parity, for
ALS_B200_SHORT_MAX in {0 (off), 16, 32, 48}.  Every setting runs in its own process (the limit is read once);
a second, WARM iteration is compared row by row with the limit-0 result and with an fp64 solve of a row sample."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
from helpers import row_err
from implicit_b200 import _lib, synthetic
ctx = _lib.Context(0)
scale = float(os.environ.get("SC_SCALE", "1.0"))
Cui, X0, Y0, cfg = synthetic.config("C2", scale=scale)
C = _lib.DeviceCSR.upload(ctx, Cui); T = C.transpose()
X, Y = _lib.DeviceFactors.from_host(ctx, X0), _lib.DeviceFactors.from_host(ctx, Y0)
lim = os.environ.get("ALS_B200_SHORT_MAX", "default")
# timing: 3 cold-start iterations (same state every time, like bench.py's device arm)
ctx.profile(True)
for it in range(4):
    X.upload(X0); Y.upload(Y0)
    if it == 1: ctx.profile_read()
    _lib.least_squares(ctx, C, X, Y, 0.01); _lib.least_squares(ctx, T, Y, X, 0.01)
p = ctx.profile_read()
ms = (p["cholesky"][0] + p["cholesky_finish"][0]) / 3
# parity: the user half of the SECOND iteration (inputs = outputs of the first), vs fp64 on a sample
Xin, Yin = X.download(), Y.download()
_lib.least_squares(ctx, C, X, Y, 0.01)
got = X.download()
deg = np.diff(Cui.indptr)
sample = np.arange(0, cfg["users"], 197)
Y64 = Yin.astype(np.float64); G64 = Y64.T @ Y64
truth = np.zeros((len(sample), 64))
for n, u in enumerate(sample):
    s, t = Cui.indptr[u], Cui.indptr[u + 1]
    if s == t: continue
    Yu, c = Y64[Cui.indices[s:t]], Cui.data[s:t].astype(np.float64)
    truth[n] = np.linalg.solve(G64 + 0.01 * np.eye(64) + (Yu.T * (np.abs(c) - 1)) @ Yu, Yu.T @ np.where(c > 0, c, 0))
e = row_err(got[sample], truth)
short = deg[sample] <= 48
np.save(os.path.join(%r, "gpurun_out", "short_check_%%s.npy" %% lim), got[::7])
print("SHORT_MAX=%%-7s cholesky %%.3f ms/iter | warm user half vs fp64: max %%.2e median %%.2e (rows<=48 nnz: max %%.2e; longer: max %%.2e) | rows<=16: %%d, <=32: %%d, <=48: %%d of %%d"
      %% (lim, ms, e.max(), np.median(e), e[short].max(), e[~short].max(), (deg<=16).sum(), (deg<=32).sum(), (deg<=48).sum(), len(deg)))
''' % (ROOT, ROOT, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for lim in (sys.argv[1:] or ["0", "16", "32", "48"]):
    env = dict(os.environ, ALS_B200_SHORT_MAX=lim, OPENBLAS_NUM_THREADS="8")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    print((r.stdout.strip().splitlines() or ["?"])[-1], r.stderr.strip()[-600:] if r.returncode else "", flush=True)
import numpy as np
base = os.path.join(ROOT, "gpurun_out", "short_check_0.npy")
if os.path.exists(base):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import row_err
    ref = np.load(base)
    for lim in ("16", "32", "48"):
        f = os.path.join(ROOT, "gpurun_out", f"short_check_{lim}.npy")
        if os.path.exists(f):
            e = row_err(np.load(f), ref)
            print(f"SHORT_MAX={lim} vs 0 (every 7th row, warm half): max {e.max():.2e} median {np.median(e):.2e}")
