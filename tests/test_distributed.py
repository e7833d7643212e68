"""The N > 1 path on CPU: world_size-2 gloo.  Environments are independent, so the multi-GPU driver only
(i) shards the ragged batch contiguously, balanced by neighbour count, and (ii) all-reduces the total energy.
The shards' energies/forces are produced here by the oracle (this is a test; the product path is the CUDA
library), and must reassemble to the single-process result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ace_jl_b200.descriptor import basis_descriptor
from ace_jl_b200.sharding import shard_bounds
from ace_jl_b200.utils import philox, rand_envs
from conftest import make_basis, rn_of


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, R, off, c, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import Oracle
    basis = make_basis("inv_simple_3_6")
    o = Oracle(basis_descriptor(basis, c.reshape(-1, 1)), threads=1)
    e0, e1 = shard_bounds(off, world)[rank]
    Rs, offs = R[off[e0]:off[e1]], off[e0:e1 + 1] - off[e0]
    E, G = o.energy_forces(Rs, offs)
    etot = torch.tensor([E.sum()], dtype=torch.float64)
    dist.all_reduce(etot)                      # the path's one collective
    out[rank] = (e0, e1, E, G, float(etot.item()))
    dist.destroy_process_group()


def test_shard_bounds_cover_and_balance():
    rng = philox(3)
    counts = rng.integers(1, 60, size=1000)
    off = np.concatenate(([0], np.cumsum(counts)))
    for w in (1, 2, 3, 8):
        b = shard_bounds(off, w)
        assert b[0][0] == 0 and b[-1][1] == 1000 and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        loads = [off[e1] - off[e0] for e0, e1 in b]
        assert max(loads) - min(loads) <= 2 * 60
    assert shard_bounds(np.array([0, 5]), 4) == [(0, 1), (1, 1), (1, 1), (1, 1)]


def test_two_rank_sharded_evaluation_matches_single_process():
    basis = make_basis("inv_simple_3_6")
    rng = philox(4)
    c = rng.random(len(basis)) - 0.5
    R, off, _ = rand_envs(rng, rn_of(basis), 24, rng.integers(1, 30, size=24))
    from oracle import Oracle
    Eref, Gref = Oracle(basis_descriptor(basis, c.reshape(-1, 1))).energy_forces(R, off)
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, R, off, c, out), nprocs=2, join=True)
    E = np.concatenate([out[r][2] for r in range(2)])
    G = np.concatenate([out[r][3] for r in range(2)])
    assert np.array_equal(E, Eref) and np.array_equal(G, Gref)
    assert abs(out[0][4] - Eref.sum()) < 1e-12 * abs(Eref.sum()) and out[0][4] == out[1][4]
