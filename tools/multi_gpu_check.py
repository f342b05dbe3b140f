"""N-GPU parity: a row-sharded fit (one process per GPU, NCCL all-gather after each half) must equal the
single-GPU fit.  Launch under torchrun / with RANK, WORLD_SIZE, MASTER_* set:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
"""
import os
import sys

os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from helpers import row_err  # noqa: E402
from implicit_b200 import AlternatingLeastSquares, synthetic  # noqa: E402
from implicit_b200.distributed import init_process_group  # noqa: E402

pg = init_process_group()
Cui = synthetic.power_law_csr(30000, 20000, 900000, 5)
X0, Y0 = synthetic.initial_factors(30000, 20000, 64)
ok = True
for use_cg in (False, True):
    m = AlternatingLeastSquares(factors=64, use_cg=use_cg, iterations=3, calculate_training_loss=True, process_group=pg)
    m.user_factors, m.item_factors = X0.copy(), Y0.copy()
    losses = []
    m.fit(Cui, show_progress=False, callback=lambda i, t, l: losses.append(l))
    Xs, Ys = m.user_factors, m.item_factors
    # single-GPU result on this rank's device (no process group)
    s = AlternatingLeastSquares(factors=64, use_cg=use_cg, iterations=3, calculate_training_loss=True, device=pg.ctx.device)
    s._ctx = pg.ctx
    s.user_factors, s.item_factors = X0.copy(), Y0.copy()
    l1 = []
    s.fit(Cui, show_progress=False, callback=lambda i, t, l: l1.append(l))
    e = np.concatenate([row_err(Xs, s.user_factors), row_err(Ys, s.item_factors)])
    dl = abs(losses[-1] - l1[-1]) / abs(l1[-1])
    print(f"rank {pg.rank}/{pg.world} {'cg' if use_cg else 'cholesky'}: sharded vs single row err max {e.max():.2e} "
          f"median {np.median(e):.2e}; loss {losses[-1]:.8f} vs {l1[-1]:.8f} (rel {dl:.1e})", flush=True)
    # the all-reduced shard Gramians sum in a different order than the single-GPU Gramian (1e-7 relative), which
    # the Cholesky fit carries through at the 1e-6 level and truncated CG amplifies like any other rounding change
    ok &= (np.median(e) < 2e-3 and dl < 1e-4) if use_cg else (e.max() < 1e-4 and dl < 1e-5)
pg.barrier()
print("MULTI_GPU_CHECK", "OK" if ok else "FAILED", flush=True)
sys.exit(0 if ok else 1)
