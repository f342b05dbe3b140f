"""First-contact diagnostics on the GPU box: prints parity numbers for every kernel family without
stopping at the first failure (pytest -m gpu is the gate; this is the microscope)."""
import os
import sys
import time
import traceback

os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import oracle  # noqa: E402
from helpers import row_err  # noqa: E402
from implicit_b200 import _lib, synthetic  # noqa: E402

ctx = _lib.Context(0)
print(ctx.info())
orc = oracle.get("auto")
print("oracle:", orc.name)


def half(Cui, X, Y, reg, use_cg, steps=3):
    C = _lib.DeviceCSR.upload(ctx, Cui)
    dX, dY = _lib.DeviceFactors.from_host(ctx, X), _lib.DeviceFactors.from_host(ctx, Y)
    (_lib.least_squares_cg(ctx, C, dX, dY, reg, steps) if use_cg else _lib.least_squares(ctx, C, dX, dY, reg))
    return dX.download()


def attempt(name, fn):
    try:
        t = time.time()
        fn()
        print(f"[ok  ] {name} ({time.time() - t:.2f}s)")
    except Exception:
        print(f"[FAIL] {name}")
        traceback.print_exc()


def gram():
    for f in (16, 40, 64, 128):
        Y = (np.random.default_rng(f).standard_normal((5000, f)) * 0.3).astype(np.float32)
        G = _lib.gramian(ctx, _lib.DeviceFactors.from_host(ctx, Y))
        G64 = Y.astype(np.float64).T @ Y.astype(np.float64)
        print(f"  gramian f={f}: rel err {np.abs(G - G64).max() / np.abs(G64).max():.2e}")


def chol():
    for f in (16, 32, 40, 64):
        Cui = synthetic.power_law_csr(700, 450, 9000, 100 + f, 0.1)
        X, Y = synthetic.initial_factors(700, 450, f)
        oracle.fit(Cui, X, Y, iterations=2, use_cg=False)
        exp = X.copy()
        orc.least_squares(Cui, exp, Y, 0.01)
        got = half(Cui, X, Y, 0.01, False)
        e = row_err(got, exp)
        print(f"  cholesky f={f}: max {e.max():.2e} median {np.median(e):.2e} nan={np.isnan(got).sum()}")
        if e.max() > 1e-3:
            w = int(np.argmax(e))
            print("   worst row", w, "nnz", Cui.indptr[w + 1] - Cui.indptr[w], "\n   got", got[w][:8], "\n   exp", exp[w][:8])


def cg():
    for f in (16, 32, 40, 64, 100, 128):
        Cui = synthetic.power_law_csr(700, 450, 12000, 200 + f, 0.05)
        X, Y = synthetic.initial_factors(700, 450, f)
        oracle.fit(Cui, X, Y, iterations=2, use_cg=False)
        exp = X.copy()
        orc.least_squares_cg(Cui, exp, Y, 0.01, cg_steps=3)
        got = half(Cui, X, Y, 0.01, True)
        e = row_err(got, exp)
        print(f"  cg f={f}: max {e.max():.2e} median {np.median(e):.2e} nan={np.isnan(got).sum()}")


def perf():
    """Quick timing of the C2 halves at full size."""
    Cui, X0, Y0, cfg = synthetic.config("C2")
    C = _lib.DeviceCSR.upload(ctx, Cui)
    t = time.time()
    T = C.transpose()
    ctx.sync()
    print(f"  device transpose of {Cui.nnz} nnz: {time.time() - t:.3f}s")
    X, Y = _lib.DeviceFactors.from_host(ctx, X0), _lib.DeviceFactors.from_host(ctx, Y0)
    ctx.profile(True)
    for name, fn in (("cholesky", lambda A, B, M: _lib.least_squares(ctx, M, A, B, 0.01)),
                     ("cg", lambda A, B, M: _lib.least_squares_cg(ctx, M, A, B, 0.01, 3))):
        for it in range(3):
            ctx.timer_start()
            fn(X, Y, C)
            fn(Y, X, T)
            ms = ctx.timer_stop()
            print(f"  {name} C2 iteration {it}: {ms:.2f} ms -> {(cfg['users'] + cfg['items']) / ms / 1e3:.1f} M rows/s", ctx.profile_read())


attempt("gramian", gram)
attempt("cholesky", chol)
attempt("cg", cg)
if "--perf" in sys.argv:
    attempt("perf", perf)
