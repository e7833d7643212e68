"""aceb200_set_devices on real hardware: one handle, several GPUs, HOST batches sharded by neighbour count.
Skipped on a single-GPU box (the sharding logic itself is covered on the CPU by tests/test_emu_parity.py)."""
import numpy as np
import pytest

import ace_jl_b200 as ace
from ace_jl_b200 import _lib
from ace_jl_b200.utils import philox, rand_envs
from conftest import make_basis, rn_of

pytestmark = pytest.mark.gpu


def test_sharded_host_batch_equals_single_device():
    ndev = _lib.load().aceb200_device_count()
    if ndev < 2:
        pytest.skip("needs two GPUs")
    basis = make_basis("inv_sparse_3_12")
    rng = philox(51)
    c = rng.random(len(basis)) - 0.5
    h = ace.LinearACEModel(basis, c).evaluator.handle
    R, off, _ = rand_envs(rng, rn_of(basis), 20_000, rng.integers(1, 60, size=20_000))
    b = ace.B200Batch(R, off)
    E1, G1 = h.energy_forces(b)
    B1 = h.eval_B(b)
    h.set_devices(list(range(ndev)))
    E2, G2 = h.energy_forces(b)
    assert np.array_equal(E1, E2) and np.array_equal(G1, G2) and np.array_equal(B1, h.eval_B(b))
    h.set_devices([0])
    assert np.array_equal(h.energy_forces(b)[0], E1)
