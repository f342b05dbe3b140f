"""CG half-iteration timing for the lanes-per-row variants (ALS_B200_CG_NV) on C2 (f=64) and C3 (f=128)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys
sys.path.insert(0, %r)
from implicit_b200 import _lib, synthetic
ctx = _lib.Context(0)
name = sys.argv[1]
Cui, X0, Y0, cfg = synthetic.config(name)
C = _lib.DeviceCSR.upload(ctx, Cui); T = C.transpose()
X, Y = _lib.DeviceFactors.from_host(ctx, X0), _lib.DeviceFactors.from_host(ctx, Y0)
ctx.profile(True)
for it in range(3):
    ctx.timer_start()
    _lib.least_squares_cg(ctx, C, X, Y, 0.01, 3)
    _lib.least_squares_cg(ctx, T, Y, X, 0.01, 3)
    ms = ctx.timer_stop()
    p = ctx.profile_read()
print("%%s NV=%%s: %%.2f ms/iter  cg %%.2f  giant %%.2f  gramian %%.2f  -> %%.1f M rows/s" %% (name, os.environ.get("ALS_B200_CG_NV","default"), ms, p["cg"][0], p["cg_giant"][0], p["gramian"][0], (cfg["users"]+cfg["items"])/ms/1e3))
''' % ROOT
for name in ("C2", "C3"):
    for nv in ("1", "2", "4"):
        env = dict(os.environ, ALS_B200_CG_NV=nv)
        r = subprocess.run([sys.executable, "-c", code, name], env=env, capture_output=True, text=True, timeout=400)
        print((r.stdout.strip().splitlines() or ["?"])[-1], r.stderr.strip()[-300:] if r.returncode else "", flush=True)
