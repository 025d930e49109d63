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


@pytest.mark.parametrize("N", [7, 10])
def test_trajectories_of_systems_beyond_eight_states(N):
    """integrate() for n > 8 (the rolled-loop build of the trajectory kernel): every trajectory came back NaN in round 1
    (the parameter-load loop was miscompiled, see odl_traj_kernel); against the oracle's odeint at tight tolerance."""
    from odelib_b200 import workloads
    from odelib_b200.engine import DeviceModel
    from scipy.integrate import odeint
    rhs, n, P, groups = workloads.nclass(N, spec_only=True)
    dm = DeviceModel(rhs, n, P, groups)
    y0 = np.zeros(n); y0[0] = 5236900.0; y0[-1] = 10981000.0
    times = np.linspace(0, 3, 200)
    dm.set_grid(times, y0)
    center = np.array([0.3, 1.0e-7, 20.0, 2.0, 2.8 * N / 2])
    traj, status, nsteps = dm.trajectory(center[None], rtol=1e-10, atol=1e-10)
    assert status[0] == 0 and nsteps[0] > 100
    ref = odeint(rhs, y0, times, args=(list(center),), rtol=1e-12, atol=1e-12, mxstep=100000)
    np.testing.assert_allclose(traj[0], ref, rtol=2e-7, atol=1e-3)
