"""recommend() / fused top-k timing at a reduced C5 shape (users x items, f=64, k=10, liked filter)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from implicit_b200 import _lib, synthetic
ctx = _lib.Context(0)
Q, I, f, k = int(sys.argv[1]) if len(sys.argv) > 1 else 20000, int(sys.argv[2]) if len(sys.argv) > 2 else 1000000, 64, 10
rng = np.random.default_rng(5)
users = (rng.standard_normal((Q, f), dtype=np.float32) * 0.1)
items = (rng.standard_normal((I, f), dtype=np.float32) * 0.1)
liked = synthetic.power_law_csr(Q, I, 20 * Q, 5)
di, dq = _lib.DeviceFactors.from_host(ctx, items), _lib.DeviceFactors.from_host(ctx, users)
dl = _lib.DeviceCSR.upload(ctx, liked)
ctx.profile(True)
for rep in range(2):
    t = time.perf_counter()
    ids, sc = _lib.topk(ctx, di, dq, k, liked=dl)
    dt = time.perf_counter() - t
    p = ctx.profile_read()
    print(f"topk Q={Q} I={I} f={f} k={k}: wall {dt*1e3:.1f} ms kernel {p['topk'][0]:.1f} ms -> {2*Q*I*f/p['topk'][0]/1e9:.1f} TFLOP/s, {Q*I/p['topk'][0]/1e6:.1f} G candidates/s", flush=True)
# spot check against numpy on 64 rows
ref = users[:64] @ items.T
for r in range(64):
    ref[r, liked.indices[liked.indptr[r]:liked.indptr[r+1]]] = -np.finfo(np.float32).max
exp = np.argsort(-ref, axis=1, kind="stable")[:, :k]
print("ids match on 64-row sample:", (exp == ids[:64]).mean())
