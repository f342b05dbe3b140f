"""CPU-only host logic: input handling, shard planning, the rendezvous, and the N>1 path of the fit
loop exercised with world_size-2 gloo processes (the per-shard solves are done by the oracle there:
what is under test is the sharding / exchange logic, not the kernels)."""
import multiprocessing as mp
import os
import warnings

import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from helpers import row_err
from implicit_b200 import synthetic
from implicit_b200.utils import ParameterWarning, check_csr, check_random_state, nnz_balanced_splits


def test_check_csr_warns_and_converts():
    m = sp.random(10, 8, density=0.3, format="coo", dtype=np.float32, random_state=1)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        out = check_csr(m)
    assert isinstance(out, sp.csr_matrix) and any(issubclass(x.category, ParameterWarning) for x in w)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert check_csr(out) is out
    assert not w


def test_constructor_argument_checks():
    """What the device library cannot do is refused at construction, like the reference refuses unknown devices
    (implicit/als.py:60-85): no CPU path, float32 only, the Cholesky solver up to 128 factors, CG up to 1024."""
    from implicit_b200 import AlternatingLeastSquares

    for kw in (dict(use_gpu=False), dict(dtype=np.float64), dict(factors=129, use_cg=False), dict(factors=1025)):
        with pytest.raises(ValueError):
            AlternatingLeastSquares(**kw)
    for kw in (dict(factors=128, use_cg=False), dict(factors=1024, use_cg=True), dict(factors=200)):
        m = AlternatingLeastSquares(**kw)  # no device is touched before fit()
        assert m.factors == kw["factors"]


def test_check_random_state():
    a = check_random_state(42).random(3)
    b = check_random_state(42).random(3)
    np.testing.assert_array_equal(a, b)
    assert isinstance(check_random_state(np.random.RandomState(1)), np.random.Generator)
    g = np.random.default_rng(3)
    assert check_random_state(g) is g


def test_synthetic_generator_is_deterministic_and_well_formed():
    a = synthetic.power_law_csr(2000, 1500, 30000, 9)
    b = synthetic.power_law_csr(2000, 1500, 30000, 9)
    assert (a != b).nnz == 0 and a.nnz == 30000
    assert a.has_sorted_indices and a.data.min() >= 1.0 and a.data.max() < 5.0
    assert a.dtype == np.float32 and a.indices.dtype == np.int32
    n = synthetic.power_law_csr(2000, 1500, 30000, 9, negative_fraction=0.05)
    assert 0.02 < (n.data < 0).mean() < 0.08
    X, Y = synthetic.initial_factors(10, 7, 16)
    assert X.dtype == np.float32 and X.max() < 0.01 and Y.shape == (7, 16)


@pytest.mark.parametrize("parts", [1, 2, 3, 8])
def test_nnz_balanced_splits(parts):
    m = synthetic.power_law_csr(5000, 3000, 80000, 4)
    s = nnz_balanced_splits(m.indptr, parts)
    assert s[0] == 0 and s[-1] == 5000 and len(s) == parts + 1 and np.all(np.diff(s) >= 0)
    per = np.diff(m.indptr[s])
    assert per.sum() == m.nnz
    assert per.max() <= m.nnz / parts + np.diff(m.indptr).max()
    # degenerate: more parts than rows, empty matrix
    e = nnz_balanced_splits(np.zeros(4, dtype=np.int32), 8)
    assert e[0] == 0 and e[-1] == 3 and np.all(np.diff(e) >= 0)


def _rdzv_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from implicit_b200.distributed import exchange_bytes

    payload = bytes(range(128)) if rank == 0 else b""
    q.put((rank, exchange_bytes(rank, world, payload, timeout=60)))


def test_rendezvous_exchanges_the_unique_id():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_rdzv_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert all(v == bytes(range(128)) for v in got.values()) and set(got) == {0, 1, 2}


def _sharded_fit_worker(rank, world, port, out_dir):
    """The fit loop of implicit_b200.als with the device calls replaced by the oracle on this rank's shard
    and the NCCL all-gather replaced by gloo: validates shard planning + exchange order on CPU."""
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    port_impl = oracle.get("port")
    Cui = synthetic.power_law_csr(900, 600, 15000, 21)
    Ciu = Cui.T.tocsr()
    X, Y = synthetic.initial_factors(900, 600, 32)
    usplit = nnz_balanced_splits(Cui.indptr, world)
    isplit = nnz_balanced_splits(Ciu.indptr, world)

    def half(C, split, Xs, Ys):
        r0, r1 = split[rank], split[rank + 1]
        mine = np.zeros((r1 - r0, 32), dtype=np.float32)
        port_impl.least_squares(C[r0:r1], mine, Ys, 0.01)
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        for r in range(world):
            Xs[split[r]:split[r + 1]] = parts[r]

    for _ in range(2):
        half(Cui, usplit, X, Y)
        half(Ciu, isplit, Y, X)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), X=X, Y=Y)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_sharded_fit_equals_single_process(tmp_path):
    pytest.importorskip("torch")
    ctx = mp.get_context("spawn")
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_sharded_fit_worker, args=(r, 2, port, str(tmp_path))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    Cui = synthetic.power_law_csr(900, 600, 15000, 21)
    X, Y = synthetic.initial_factors(900, 600, 32)
    oracle.fit(Cui, X, Y, iterations=2, use_cg=False, kind="port")
    for r in range(2):
        z = np.load(tmp_path / f"rank{r}.npz")
        # identical per-row arithmetic -> identical results, whatever the sharding
        assert row_err(z["X"], X).max() < 1e-6 and row_err(z["Y"], Y).max() < 1e-6
