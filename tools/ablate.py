"""Timing ablations of the Cholesky half kernel on C2 (ALS_B200_DEBUG, results are wrong when set)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys
sys.path.insert(0, %r)
from implicit_b200 import _lib, synthetic
ctx = _lib.Context(0)
Cui, X0, Y0, cfg = synthetic.config("C2")
C = _lib.DeviceCSR.upload(ctx, Cui); T = C.transpose()
X, Y = _lib.DeviceFactors.from_host(ctx, X0), _lib.DeviceFactors.from_host(ctx, Y0)
ctx.profile(True)
for it in range(3):
    try:
        _lib.least_squares(ctx, C, X, Y, 0.01)
    except Exception: pass
    try:
        _lib.least_squares(ctx, T, Y, X, 0.01)
    except Exception: pass
    X.upload(X0); Y.upload(Y0)
    p = ctx.profile_read()
print("DEBUG=%%s cholesky ms/2 launches: %%.3f" %% (os.environ.get("ALS_B200_DEBUG","0"), p["cholesky"][0]))
''' % ROOT
variants = [v for v in sys.argv[1:]] or [""]
for lib, flags in [(v, f) for v in variants for f in ((0, 1, 2, 4) if len(variants) <= 2 else (0,))]:
    env = dict(os.environ, ALS_B200_DEBUG=str(flags))
    if lib:
        env["ALS_B200_LIB"] = os.path.join(ROOT, lib)
        print(lib, end=" ")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    print((r.stdout.strip().splitlines() or ["?"])[-1], r.stderr.strip()[-200:] if r.returncode else "", flush=True)
