"""CPU checks of the kernel core: tests/hostsim builds the __host__ __device__
game logic of skillshot_learning_b200/csrc/ss_env_core.cuh for the host (the
same source the sm_100a kernels inline) and runs the parity suite of
tests/parity.py against the golden vectors and the oracle.  The GPU run of the
same suite is tests/test_gpu_env_parity.py."""
import numpy as np
import pytest

from tests import parity
from tests.helpers import GOLDEN_FILES
from tests.hostsim.sim import HostSimEnvs


def make(n, **kw):
    return HostSimEnvs(n, **kw)


@pytest.mark.parametrize("name", GOLDEN_FILES)
def test_golden_lockstep(name):
    parity.check_golden_lockstep(make, name)


@pytest.mark.parametrize("name", ["lockstep_random", "close_hits"])
def test_golden_fused_ticks(name):
    parity.check_golden_fused(make, name, K=8)


def test_oracle_lockstep_random():
    parity.check_oracle_lockstep(make, n=512, T=128, seed=1)


def test_oracle_lockstep_hits_terminal_reward():
    hits = parity.check_oracle_lockstep(make, n=512, T=48, seed=2, close=True, reward_mode="terminal")
    assert hits > 50


def test_oracle_lockstep_fused_chunks():
    parity.check_oracle_lockstep(make, n=256, T=64, seed=3, close=True, reward_mode="simple", chunk=16)


def test_auto_reset():
    parity.check_auto_reset(make)


def test_random_reset():
    parity.check_random_reset_properties(make)


def test_speeds():
    parity.check_speeds(make)


def test_nan_action_sets_status():
    e = make(4)
    a = np.zeros((4, 2, 2), np.float32)
    a[2, 1, 0] = np.nan
    e.step(a)
    assert e.status_bits() & 1


def test_bench_shape_small():
    """bench.py's timed shape (fused ticks, Philox auto-reset, terminal reward) at a size the host simulation finishes in
    seconds; the GPU run of the same check is at the full 65,536 envs x 2,048 ticks x K = 128."""
    episodes, hits = parity.check_bench_shape(make, n=512, T=192, K=32, tick_limit=50)
    assert episodes > 512 * 2 and hits > 10
