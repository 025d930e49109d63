"""CPU checks of the stepper code in odl_kernels.cuh (DOPRI5, ROS23, Radau5), compiled for the host by
tests/host_harness (g++; kernels and warp code compiled out) and compared with scipy's odeint.
These pin the integrator logic without a GPU; the GPU tests pin the kernels themselves."""
import numpy as np
import pytest
from scipy.integrate import odeint

from odelib_b200 import demo_models
from oracle import odelib_oracle as orc
from tests import host_harness as hh
from tests.helpers import golden, oracle_tables

TOL = 1.49012e-8


@pytest.fixture(scope="module")
def two_i():
    f, n, P, groups = demo_models.MODELS["two_i"]
    tab = oracle_tables("two_i")
    slots = tab.times[np.unique(np.concatenate([tab.tindex[s] for s in tab.obs_order]))]
    return hh.build(f, n, P, groups), tab, slots


def _ref(theta, tab, slots):
    return odeint(orc.two_i, list(tab.y0), slots, args=(list(theta),), rtol=1e-12, atol=1e-12, mxstep=500000)


def _relerr(out, ref):
    return np.nanmax(np.abs(out - ref) / (np.abs(ref) + 1.0))


@pytest.mark.parametrize("solver,bound,max_mean_steps", [("dopri5", 5e-7, 400), ("radau5", 1e-6, 500), ("bdf", 3e-6, 600),
                                                         ("ros23", 1e-4, 8000)])
def test_steppers_at_default_tolerance(two_i, solver, bound, max_mean_steps):
    lib, tab, slots = two_i
    g = golden("two_i")
    steps = []
    for th in g["theta"][:6]:
        out, st, ns = hh.solve(lib, solver, th, slots, tab.y0, TOL, TOL)
        assert st == 0
        assert _relerr(out, _ref(th, tab, slots)) < bound
        steps.append(ns)
    assert np.mean(steps) < max_mean_steps


def test_radau5_converges_with_tolerance(two_i):
    lib, tab, slots = two_i
    th = golden("two_i")["theta"][0]
    ref = _ref(th, tab, slots)
    errs = []
    for tol in (1e-5, 1e-8, 1e-11):
        out, st, _ = hh.solve(lib, "radau5", th, slots, tab.y0, tol, tol)
        assert st == 0
        errs.append(_relerr(out, ref))
    assert errs[0] > errs[1] > errs[2] and errs[2] < 1e-8


def test_stiff_variant_radau5_is_cheap_and_dopri5_is_not(two_i):
    """BASELINE config 4 point (tau=1e4, lam=1e-2): LSODA runs BDF; Radau5 needs a few hundred steps."""
    lib, tab, slots = two_i
    th = np.array([0.5, 1e-7, 50.0, 1e-2, 1e4])
    ref = _ref(th, tab, slots)
    out, st, ns = hh.solve(lib, "radau5", th, slots, tab.y0, TOL, TOL)
    assert st == 0 and ns < 1500 and _relerr(out, ref) < 1e-6
    out2, st2, ns2 = hh.solve(lib, "ros23", th, slots, tab.y0, TOL, TOL)
    assert st2 == 0 and _relerr(out2, ref) < 1e-4
    out3, st3, ns3 = hh.solve(lib, "dopri5", th, slots, tab.y0, TOL, TOL)
    assert ns3 > 5 * ns                      # stability-limited explicit steps
    assert st3 == 0 and _relerr(out3, ref) < 1e-5


def test_bdf_converges_with_tolerance_and_is_cheap_on_the_stiff_variant(two_i):
    """Variable-order BDF (LSODA's stiff family): error falls with the tolerance; on the config-4 point it needs a
    few hundred steps where the explicit method is stability-limited."""
    lib, tab, slots = two_i
    th = golden("two_i")["theta"][0]
    ref = _ref(th, tab, slots)
    errs = []
    for tol in (1e-5, 1e-8, 1e-11):
        out, st, _ = hh.solve(lib, "bdf", th, slots, tab.y0, tol, tol)
        assert st == 0
        errs.append(_relerr(out, ref))
    assert errs[0] > errs[1] > errs[2] and errs[2] < 1e-8
    stiff = np.array([0.5, 1e-7, 50.0, 1e-2, 1e4])
    out, st, ns = hh.solve(lib, "bdf", stiff, slots, tab.y0, TOL, TOL)
    assert st == 0 and ns < 600 and _relerr(out, _ref(stiff, tab, slots)) < 1e-6


def test_bdf_step_counts_on_the_stiff_tail_match_lsoda(two_i):
    """The stiff pass of the sweep is bound by the latency of its longest systems, so their STEP COUNT is what its time
    rests on: on the prior draws DOPRI5 does not finish in 512 attempts, the BDF stepper needs about as many steps as
    LSODA itself reports (nst) -- 883 vs 870 on the worst of 60,000 draws when this was written."""
    from tests.helpers import prior_draws
    lib, tab, slots = two_i
    theta = prior_draws("two_i", 12000, seed=0)
    tail = [i for i in range(len(theta)) if hh.solve(lib, "dopri5", theta[i], slots, tab.y0, TOL, TOL, max_steps=512)[1] != 0]
    assert 60 < len(tail) < 400                                  # ~1.2 % of the draws
    steps = np.array([hh.solve(lib, "bdf", theta[i], slots, tab.y0, TOL, TOL)[2] for i in tail])
    assert steps.max() < 1000 and np.median(steps) < 600
    worst = [tail[k] for k in np.argsort(steps)[-6:]]
    for i in worst:
        _, info = odeint(orc.two_i, list(tab.y0), slots, args=(list(theta[i]),), full_output=True, mxstep=500000)
        mine = hh.solve(lib, "bdf", theta[i], slots, tab.y0, TOL, TOL)[2]
        assert mine < 1.25 * info["nst"][-1] + 50, (i, mine, info["nst"][-1])


def test_step_budget_and_nonfinite_inputs_end_in_status_words(two_i):
    lib, tab, slots = two_i
    th = golden("two_i")["theta"][0]
    for solver in ("dopri5", "ros23", "radau5", "bdf"):
        _, st, ns = hh.solve(lib, solver, th, slots, tab.y0, TOL, TOL, max_steps=7)
        assert st == 1 and ns == 7
        bad = th.copy(); bad[1] = np.nan
        _, st, _ = hh.solve(lib, solver, bad, slots, tab.y0, TOL, TOL, max_steps=100000)
        assert st != 0


@pytest.mark.parametrize("model", ["one_i", "two_i", "n_class_10"])
def test_in_register_lu_with_pivoting(model):
    """The LU the ROS23 / Radau5 steppers use (unrolled in registers for n <= 8, rolled for larger n), real and
    complex, on matrices that force row exchanges -- against numpy.linalg.solve."""
    import ctypes as C
    if model == "n_class_10":
        f, n, P, groups = demo_models.n_class(10), 12, 5, [tuple(range(11)), (11,)]
    else:
        f, n, P, groups = demo_models.MODELS[model]
    lib = hh.build(f, n, P, groups)
    rng = np.random.default_rng(0)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    for trial in range(400):
        A = rng.standard_normal((n, n))
        mode = trial % 4
        if mode >= 1:
            A[0, 0] = 0.0
        if mode == 2:
            A[1, 1] = 0.0
        if mode == 3:
            A = A[rng.permutation(n)] * (10.0 ** rng.integers(-6, 4, n))[:, None]
        b = rng.standard_normal(n)
        x = np.empty(n)
        lib.harness_lu(ptr(A), ptr(b), ptr(x))
        assert np.abs(A @ x - b).max() <= 1e-11 * (np.abs(A).max() * np.abs(x).max() + 1)
        Ai, bi = rng.standard_normal((n, n)), rng.standard_normal(n)
        xr, xi = np.empty(n), np.empty(n)
        lib.harness_clu(ptr(A), ptr(Ai), ptr(b), ptr(bi), ptr(xr), ptr(xi))
        M, z = A + 1j * Ai, xr + 1j * xi
        assert np.abs(M @ z - (b + 1j * bi)).max() <= 1e-11 * (np.abs(M).max() * np.abs(z).max() + 1)


def test_dense_output_in_observed_space_equals_the_observed_dense_states(two_i):
    """The sweep / chain kernels interpolate the OBSERVED columns (sums of state groups) instead of the states -- the
    continuous extension is linear, so both give the same numbers up to the order of the additions; the steps taken are
    the same steps."""
    lib, tab, slots = two_i
    g = golden("two_i")
    for th in g["theta"][:8]:
        full, st, ns = hh.solve(lib, "dopri5", th, slots, tab.y0, TOL, TOL)
        obs, st2, ns2 = hh.solve_observed(lib, th, slots, tab.y0, TOL, TOL)
        assert st == 0 and st2 == 0 and ns == ns2
        want = np.column_stack([full[:, 0] + full[:, 1] + full[:, 2], full[:, 3]])      # H = S + I1 + I2, V
        np.testing.assert_allclose(obs, want, rtol=1e-13)
        assert not np.array_equal(obs, full[:, :2])                                      # (it really is the observed path)


def test_hand_over_from_dopri5_to_bdf(two_i):
    """The AUTO sweep's hand-over, on the host: on prior draws the capped DOPRI5 pass gives up on, the BDF stepper
    continues from where it stopped.  (1) a stopped solve stands still while its lane waits (0 or 7 further attempts: the
    same state bit for bit, hence the same result); (2) the result is the solution (odeint at 1e-12); (3) far fewer BDF
    steps than the BDF solve from t0."""
    from tests.helpers import prior_draws
    lib, tab, slots = two_i
    theta = prior_draws("two_i", 4000, seed=0)
    handed, steps_cont, steps_t0 = 0, 0, 0
    for th in theta:
        a, st, nd, nb, slot, t_hand = hh.solve_handover(lib, th, slots, tab.y0, TOL, TOL, idle=0)
        if nb == 0:
            continue                                             # DOPRI5 finished this row itself
        assert st == 0 and 0 < t_hand < slots[-1]
        b, st2, nd2, nb2, slot2, _ = hh.solve_handover(lib, th, slots, tab.y0, TOL, TOL, idle=7)
        assert st2 == 0 and (nd2, nb2, slot2) == (nd, nb, slot) and np.array_equal(a, b)
        ref = _ref(th, tab, slots)
        ok = ref > 1.0                                           # (below one cell/ml a state is integrator noise)
        if ok.any():
            assert np.max(np.abs(a[ok] - ref[ok]) / np.abs(ref[ok])) < 5e-6
        _, st3, n0 = hh.solve(lib, "bdf", th, slots, tab.y0, TOL, TOL)
        assert st3 == 0
        handed += 1; steps_cont += nb; steps_t0 += n0
        if handed == 25:
            break
    assert handed >= 20 and steps_cont < 0.5 * steps_t0, (handed, steps_cont, steps_t0)
