"""GPU parity of full-grid trajectories (odl_trajectory) with scipy odeint at tight tolerance."""
import numpy as np
import pytest

from oracle import odelib_oracle as orc
from tests.helpers import device_model, golden, oracle_rhs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["zero_i", "one_i", "two_i"])
def test_trajectory_matches_odeint_on_the_output_grid(name):
    g = golden(name)
    dm, tab = device_model(name)
    theta = g["theta"][:4]
    traj, status, nsteps = dm.trajectory(theta, rtol=1e-12, atol=1e-12)
    assert traj.shape == (4, len(tab.times), dm.n_state) and np.all(status == 0)
    for k in range(4):
        ref = orc.integrate_grid(oracle_rhs(name), tab.y0, tab.times, theta[k], 1e-13, 1e-13, mxstep=200000)
        scale = np.abs(ref).max(axis=0)
        assert np.all(np.abs(traj[k] - ref) <= 5e-9 * scale + 1e-3)
    assert np.array_equal(traj[0][0], tab.y0)
