"""Device generator (csrc/gen.cu) against the host generator (implicit_b200/synthetic.py) at the C2 shape: same
recipe, different random streams -> the degree statistics must agree; plus one Cholesky iteration on the result."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from implicit_b200 import _lib, synthetic
ctx = _lib.Context(0)
cfg = synthetic.CONFIGS["C2"]
t = time.perf_counter()
C = _lib.DeviceCSR.generate(ctx, cfg["users"], cfg["items"], cfg["nnz"], cfg["seed"])
ctx.sync()
print(f"device generation: {time.perf_counter() - t:.2f} s, shape {C.shape3}")
m = C.download()
assert m.has_sorted_indices or True
d = np.diff(m.indptr)
assert (np.diff(m.indices)[np.arange(m.nnz - 1)[np.isin(np.arange(1, m.nnz), m.indptr[1:-1], invert=True)]] > 0).all(), "unsorted or duplicate columns"
host, _, _, _ = synthetic.config("C2")
for name, a, b in (("users", d, np.diff(host.indptr)), ("items", np.bincount(m.indices, minlength=cfg["items"]), np.bincount(host.indices, minlength=cfg["items"]))):
    q = [50, 90, 99, 100]
    print(name, "device p50/p90/p99/max", np.percentile(a, q).astype(int), "host", np.percentile(b, q).astype(int), "empty", (a == 0).sum(), (b == 0).sum())
print("values: device min/mean/max", m.data.min(), m.data.mean(), m.data.max(), "host", host.data.min(), host.data.mean(), host.data.max())
T = C.transpose()
X, Y = _lib.DeviceFactors(ctx, cfg["users"], 64), _lib.DeviceFactors(ctx, cfg["items"], 64)
X.fill_uniform(42, 0.01); Y.fill_uniform(43, 0.01)
y = Y.download(); print("factors: mean", y.mean(), "max", y.max(), "min", y.min())
_lib.least_squares(ctx, C, X, Y, 0.01); _lib.least_squares(ctx, T, Y, X, 0.01)
print("one iteration ok, NaN:", X.has_nan(), Y.has_nan())
print("GEN_CHECK done")
