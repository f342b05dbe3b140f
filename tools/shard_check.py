"""Does the short-row path still pay on an 8-way row shard?  One GPU, C2: the user half over rows [0, U/8) and the
item half over rows [0, I/8), with ALS_B200_SHORT_MAX = 0 and 48 (the whitening of Y is not sharded)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from implicit_b200 import _lib, synthetic
parts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = _lib.Context(0)
Cui, X0, Y0, cfg = synthetic.config("C2")
C = _lib.DeviceCSR.upload(ctx, Cui); T = C.transpose()
X, Y = _lib.DeviceFactors.from_host(ctx, X0), _lib.DeviceFactors.from_host(ctx, Y0)
Cs = C.slice_rows(0, cfg["users"] // parts); Ts = T.slice_rows(0, cfg["items"] // parts)
for lim in ("0", "48"):
    ctx.set_knob("short_max", int(lim))
    ctx.profile(True)
    for it in range(4):
        if it == 1: ctx.profile_read()
        _lib.gramian_shard(ctx, Y, 0, cfg["items"]); _lib.half_pregram(ctx, Cs, X, Y, 0.01, False, 3)
        _lib.gramian_shard(ctx, X, 0, cfg["users"]); _lib.half_pregram(ctx, Ts, Y, X, 0.01, False, 3)
    p = ctx.profile_read(); ctx.profile(False)
    print(f"1/{parts} shard, SHORT_MAX={lim}: cholesky {(p['cholesky'][0] + p['cholesky_finish'][0]) / 3:.3f} ms per iteration "
          f"(user rows {Cs.shape3[0]}, item rows {Ts.shape3[0]})", flush=True)
