"""Parity of a library variant (ALS_B200_LIB) on the warm f=64 case + cold-start C2 sample vs fp64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
import numpy as np
import oracle
from helpers import row_err
from implicit_b200 import _lib, synthetic
ctx = _lib.Context(0)
orc = oracle.get("auto")
def half(Cui, X, Y):
    C = _lib.DeviceCSR.upload(ctx, Cui); dX, dY = _lib.DeviceFactors.from_host(ctx, X), _lib.DeviceFactors.from_host(ctx, Y)
    _lib.least_squares(ctx, C, dX, dY, 0.01); return dX.download()
Cui = synthetic.power_law_csr(700, 450, 9000, 164, 0.1); X, Y = synthetic.initial_factors(700, 450, 64)
oracle.fit(Cui, X, Y, iterations=2, use_cg=False)
exp = X.copy(); orc.least_squares(Cui, exp, Y, 0.01)
e = row_err(half(Cui, X, Y), exp)
print(os.environ.get("ALS_B200_LIB", "default"), f"warm f=64: max {e.max():.2e} median {np.median(e):.2e}", end="; ")
Cui, X0, Y0, cfg = synthetic.config("C2", scale=0.1)
got = half(Cui, X0, Y0)
sample = np.arange(0, cfg["users"], 97)
Y64 = Y0.astype(np.float64); G64 = Y64.T @ Y64
truth = np.zeros((len(sample), 64))
for n, u in enumerate(sample):
    s, t = Cui.indptr[u], Cui.indptr[u + 1]
    if s == t: continue
    Yu, c = Y64[Cui.indices[s:t]], Cui.data[s:t].astype(np.float64)
    truth[n] = np.linalg.solve(G64 + 0.01 * np.eye(64) + (Yu.T * (np.abs(c) - 1)) @ Yu, Yu.T @ np.where(c > 0, c, 0))
e = row_err(got[sample], truth)
print(f"cold C2/10 vs fp64: max {e.max():.2e} median {np.median(e):.2e}")
