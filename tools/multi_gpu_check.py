"""N-GPU parity: the row-sharded fit (one process per GPU, solved rows mirrored into the peers' replicas by the solve
kernels, Gramian all-reduce between halves) against the single-GPU fit, on every rank.  Launch under torchrun / with
RANK, WORLD_SIZE, MASTER_* set:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py

Checks, in order:
  1. ONE half-iteration from a warm state, Cholesky and CG: only the summation order of the all-reduced Gramian differs
     from the single-GPU half, so rows must agree to ~1e-6 (an ordering / race bug in the peer stores would show here,
     not be mistaken for "CG amplifies rounding").  Also with ALS_B200_NO_P2P-style plain all-gather.
  2. 3-iteration fits from the cold start, gated against the single-GPU fit's OWN sensitivity: the same fit from initial
     factors perturbed in the last bit; the sharded fit must sit within 3x of that (and the training losses must agree).
  3. A NaN confidence in a row that only the last rank owns: every rank must raise in the same iteration (no rank may
     be left inside a collective), and a following fit must work.
"""
import os
import sys

os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from helpers import CHOL_MAX, row_err  # noqa: E402
from implicit_b200 import AlternatingLeastSquares, _lib, synthetic  # noqa: E402
from implicit_b200.distributed import init_process_group  # noqa: E402
from implicit_b200.utils import nnz_balanced_splits  # noqa: E402

pg = init_process_group()
rank, world, ctx = pg.rank, pg.world, pg.ctx
U, I, f = 30000, 20000, 64
Cui = synthetic.power_law_csr(U, I, 900000, 5)
X0, Y0 = synthetic.initial_factors(U, I, f)
ok = True


def say(msg):
    print(f"rank {rank}/{world} {msg}", flush=True)


def fit(use_cg, Xi, Yi, group, iters=3):
    m = AlternatingLeastSquares(factors=f, use_cg=use_cg, iterations=iters, calculate_training_loss=True,
                                process_group=group, device=ctx.device)
    if group is None:
        m._ctx = ctx
    m.user_factors, m.item_factors = Xi.copy(), Yi.copy()
    losses = []
    m.fit(Cui, show_progress=False, callback=lambda i, t, l: losses.append(l))
    return np.array(m.user_factors), np.array(m.item_factors), losses[-1]


# ---- 1. one half from a warm state -------------------------------------------------------------------------
Xw, Yw, _ = fit(False, X0, Y0, None, iters=1)
C = _lib.DeviceCSR.upload(ctx, Cui)
usplit = nnz_balanced_splits(Cui.indptr, world, 60)
Cs = C.slice_rows(usplit[rank], usplit[rank + 1])
for use_cg in (False, True):
    Xs, Ys = _lib.DeviceFactors.from_host(ctx, Xw), _lib.DeviceFactors.from_host(ctx, Yw)
    if use_cg:
        _lib.least_squares_cg(ctx, C, Xs, Ys, 0.01, 3)
    else:
        _lib.least_squares(ctx, C, Xs, Ys, 0.01)
    single = Xs.download()
    # the sharded half differs from this one only by the summation order of the Gramian: calibrate the gate with the
    # same half from item factors perturbed in the last bit (3 unconverged CG steps amplify that on ill-conditioned rows)
    Yq = _lib.DeviceFactors.from_host(ctx, (Yw * (1 + 1e-7 * np.random.default_rng(7).standard_normal(Yw.shape))).astype(np.float32))
    Xq = _lib.DeviceFactors.from_host(ctx, Xw)
    if use_cg:
        _lib.least_squares_cg(ctx, C, Xq, Yq, 0.01, 3)
    else:
        _lib.least_squares(ctx, C, Xq, Yq, 0.01)
    sens = row_err(Xq.download(), single)
    Xq.close()
    Yq.close()
    for p2p in (True, False):
        Xd, Yd = _lib.DeviceFactors.from_host(ctx, Xw), _lib.DeviceFactors.from_host(ctx, Yw)
        isplit = nnz_balanced_splits(Cui.T.tocsr().indptr, world, 60)
        if p2p:
            ctx.attach_peers(Xd)
        _lib.gramian_shard(ctx, Yd, isplit[rank], isplit[rank + 1] - isplit[rank])
        _lib.half_pregram_async(ctx, Cs, Xd, Yd, 0.01, use_cg, 3)
        if p2p:
            _lib.gramian_shard(ctx, Xd, usplit[rank], usplit[rank + 1] - usplit[rank])  # orders the peers' stores
            _lib.solver_status(ctx)
            ctx.barrier()
            ctx.detach_peers(Xd)
        else:
            _lib.solver_status(ctx)
            ctx.allgather_rows(Xd, usplit)
        e = row_err(Xd.download(), single)
        say(f"one {'cg' if use_cg else 'cholesky'} half, {'peer stores' if p2p else 'all-gather'}: row err max {e.max():.2e} "
            f"median {np.median(e):.2e} (single GPU, item factors perturbed by 1e-7: max {sens.max():.2e} median {np.median(sens):.2e})")
        ok &= bool(e.max() < max(2e-5, 3 * sens.max()) and np.median(e) < max(2e-6, 3 * np.median(sens)))
        Xd.close()
        Yd.close()
    Xs.close()
    Ys.close()
Cs.close()
C.close()

# ---- 2. three-iteration fits against the single-GPU fit and its own sensitivity ----------------------------------
rng = np.random.default_rng(97)
Xp = (X0 * (1 + 1e-7 * rng.standard_normal(X0.shape))).astype(np.float32)
Yp = (Y0 * (1 + 1e-7 * rng.standard_normal(Y0.shape))).astype(np.float32)
for use_cg in (False, True):
    xs, ys, ls = fit(use_cg, X0, Y0, pg)
    x1, y1, l1 = fit(use_cg, X0, Y0, None)
    xp, yp, _ = fit(use_cg, Xp, Yp, None)
    e = np.concatenate([row_err(xs, x1), row_err(ys, y1)])
    es = np.concatenate([row_err(xp, x1), row_err(yp, y1)])
    dl = abs(ls - l1) / abs(l1)
    say(f"{'cg' if use_cg else 'cholesky'} fit: sharded vs single row err max {e.max():.2e} p99 {np.quantile(e, 0.99):.2e} median "
        f"{np.median(e):.2e}; single vs itself from 1e-7-perturbed factors max {es.max():.2e} p99 {np.quantile(es, 0.99):.2e} "
        f"median {np.median(es):.2e}; loss {ls:.8f} vs {l1:.8f} (rel {dl:.1e})")
    ok &= bool(np.median(e) < max(1e-5, 3 * np.median(es)) and np.quantile(e, 0.99) < max(CHOL_MAX, 3 * np.quantile(es, 0.99))
               and e.max() < max(CHOL_MAX, 3 * es.max()) and dl < 1e-4)

# ---- 3. a failure on one rank raises everywhere ------------------------------------------------------------------------
bad = Cui.copy()
row = U - 1  # owned by the last rank only
while bad.indptr[row + 1] == bad.indptr[row]:
    row -= 1
bad.data[bad.indptr[row]] = np.nan
raised = None
try:
    m = AlternatingLeastSquares(factors=f, use_cg=False, iterations=2, process_group=pg)
    m.user_factors, m.item_factors = X0.copy(), Y0.copy()
    m.fit(bad, show_progress=False)
except (ValueError, _lib.AlsError, Exception) as exc:  # noqa: B014  (ModelFitError is also fine: NaN reached the factors)
    raised = type(exc).__name__ + ": " + str(exc)[:80]
say(f"NaN confidence in row {row}: {'raised ' + raised if raised else 'NO EXCEPTION'}")
ok &= raised is not None
pg.barrier()
xs, ys, ls = fit(False, X0, Y0, pg, iters=1)  # the group is still usable
ok &= bool(np.isfinite(xs).all() and np.isfinite(ys).all())
pg.barrier()
print(f"rank {rank}/{world} MULTI_GPU_CHECK", "OK" if ok else "FAILED", flush=True)
sys.exit(0 if ok else 1)
