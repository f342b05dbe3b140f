"""Where the warps of the tcgen05 long-row kernel wait.  Needs the -DALS_TC_STATS variant:
    VARIANT_SRC=cholesky_tc tools/build_variant.sh tcstats -DALS_TC_STATS
    ALS_B200_LIB=variants/tcstats.so python tools/long_stats.py
Prints, per role, the mean share of the kernel's cycles spent in each wait (user half, then item half of C2)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from implicit_b200 import _lib, synthetic
ctx = _lib.Context(0)
ctx.set_knob("long_tc", 1)
Cui, X0, Y0, cfg = synthetic.config("C2", scale=float(os.environ.get("SC_SCALE", "1.0")))
C = _lib.DeviceCSR.upload(ctx, Cui); T = C.transpose()
X, Y = _lib.DeviceFactors.from_host(ctx, X0), _lib.DeviceFactors.from_host(ctx, Y0)
buf = (ctypes.c_ulonglong * (160 * 16 * 4))()
for name, (A, U, V) in (("user half", (C, X, Y)), ("item half", (T, Y, X))):
    for rep in range(2):
        _lib.least_squares(ctx, A, U, V, 0.01)
    ctx.sync()
    assert ctx.lib.als_debug_tc_stats(buf) == 0
    s = np.frombuffer(buf, dtype=np.uint64).reshape(160, 16, 4)[:148].astype(np.float64)
    tot = s[:, :, 3]
    t8 = tot[:, 8]
    q = np.percentile(t8, [0, 10, 50, 90, 100])
    print(f"{name}: kernel cycles per CTA: mean {t8.mean():.0f} max {t8.max():.0f}; percentiles 0/10/50/90/100: " + " ".join(f"{v:.0f}" for v in q))
    order = np.argsort(t8)
    print(f"  slowest CTAs (blockIdx): {order[-8:][::-1].tolist()}  fastest: {order[:8].tolist()}  corr(blockIdx, cycles) {np.corrcoef(np.arange(148), t8)[0, 1]:+.2f}")
    slow = order[-15:]
    print(f"  the 15 slowest CTAs: solvers in factor_solve {100 * (s[slow][:, :8, 2] / tot[slow][:, :8]).mean():.1f} %, waiting for rows {100 * (s[slow][:, :8, 0] / tot[slow][:, :8]).mean():.1f} %; "
          f"producers waiting for accumulators {100 * (s[slow][:, 9:, 1] / tot[slow][:, 9:]).mean():.1f} %, for stages {100 * (s[slow][:, 9:, 0] / tot[slow][:, 9:]).mean():.1f} %")
    print(f"  solvers   : wait row_done {100 * (s[:, :8, 0] / tot[:, :8]).mean():.1f} %, group barriers {100 * (s[:, :8, 1] / tot[:, :8]).mean():.1f} %, "
          f"factor_solve {100 * (s[:, :8, 2] / tot[:, :8]).mean():.1f} %")
    print(f"  MMA warp  : wait stage full {100 * (s[:, 8, 0] / tot[:, 8]).mean():.1f} %, wait accumulator free {100 * (s[:, 8, 1] / tot[:, 8]).mean():.1f} %")
    print(f"  producers : wait stage empty {100 * (s[:, 9:, 0] / tot[:, 9:]).mean():.1f} %, wait accumulator free {100 * (s[:, 9:, 1] / tot[:, 9:]).mean():.1f} %", flush=True)
