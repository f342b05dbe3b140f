"""One process per GPU: rendezvous + NCCL communicator for the row-sharded ALS fit.

The reference is single-process / single-GPU (``// TODO: multi-gpu support``, implicit/gpu/als.cu:169);
this module is new.  Launch with the usual environment (``torchrun``-compatible):

    RANK, LOCAL_RANK, WORLD_SIZE, MASTER_ADDR, MASTER_PORT

Only the 128-byte NCCL unique id crosses between the processes on the host side: rank 0 creates it
(als_comm_unique_id) and serves it over a loopback/TCP socket next to MASTER_PORT; every rank then
calls als_comm_init.  No PyTorch is involved; all collectives are NCCL calls inside libals_b200.so.
"""
import os
import socket
import struct
import time

from . import _lib

_MAGIC = b"ALSB200\x01"
_PORT_OFFSETS = (1, 2, 3, 5, 8, 13, 21, 34)


def _candidate_ports():
    base = int(os.environ.get("ALS_B200_RDZV_PORT", 0))
    if base:
        return [base]
    master = int(os.environ.get("MASTER_PORT", "29500"))
    return [master + o for o in _PORT_OFFSETS]


def _run_token():
    # distinguishes concurrent jobs that share a host: same for every rank of one launch
    return (os.environ.get("TORCHELASTIC_RUN_ID", "") + ":" + os.environ.get("MASTER_PORT", "")).encode()[:64]


def exchange_bytes(rank, world, payload, timeout=300.0):
    """Rank 0 hands `payload` (bytes) to every other rank; returns it on all ranks."""
    if world == 1:
        return payload
    addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
    token = _run_token()
    deadline = time.time() + timeout
    if rank == 0:
        srv = None
        for port in _candidate_ports():
            try:
                srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
                srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
                srv.bind(("", port))
                break
            except OSError:
                srv.close()
                srv = None
        if srv is None:
            raise RuntimeError("implicit_b200.distributed: no free rendezvous port near MASTER_PORT")
        srv.listen(world)
        srv.settimeout(1.0)
        served = 0
        while served < world - 1:
            if time.time() > deadline:
                raise TimeoutError("implicit_b200.distributed: rendezvous timed out on rank 0")
            try:
                conn, _ = srv.accept()
            except socket.timeout:
                continue
            with conn:
                conn.settimeout(10.0)
                try:
                    hello = _recv_exact(conn, len(_MAGIC) + 64)
                except (OSError, ConnectionError):
                    continue
                if hello[: len(_MAGIC)] != _MAGIC or hello[len(_MAGIC):].rstrip(b"\0") != token:
                    continue  # not one of ours
                conn.sendall(struct.pack("<I", len(payload)) + payload)
                served += 1
        srv.close()
        return payload
    ports = _candidate_ports()
    while True:
        for port in ports:
            try:
                with socket.create_connection((addr, port), timeout=2.0) as c:
                    c.settimeout(10.0)
                    c.sendall(_MAGIC + token.ljust(64, b"\0"))
                    (n,) = struct.unpack("<I", _recv_exact(c, 4))
                    return _recv_exact(c, n)
            except (OSError, ConnectionError, struct.error):
                continue
        if time.time() > deadline:
            raise TimeoutError(f"implicit_b200.distributed: rank {rank} could not reach rank 0")
        time.sleep(0.2)


def _recv_exact(conn, n):
    buf = b""
    while len(buf) < n:
        chunk = conn.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("peer closed")
        buf += chunk
    return buf


class ProcessGroup:
    """rank / world + the device context whose communicator spans all ranks."""

    def __init__(self, rank, world, ctx):
        self.rank, self.world, self.ctx = rank, world, ctx

    def barrier(self):
        self.ctx.barrier()

    def allreduce_max(self, value):
        return float(self.ctx.allreduce([value], "max")[0])

    def allreduce_sum(self, value):
        return float(self.ctx.allreduce([value], "sum")[0])


def init_process_group(device=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE, creates this rank's device context and joins the NCCL clique."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    ctx = _lib.Context(local if device is None else device)
    if world > 1:
        uid = _lib.comm_unique_id() if rank == 0 else b""
        uid = exchange_bytes(rank, world, uid)
        ctx.comm_init(rank, world, uid)
    return ProcessGroup(rank, world, ctx)
