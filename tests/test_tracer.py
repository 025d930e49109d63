"""RHS tracer: the emitted SSA evaluates exactly like the Python model, Jacobian is right."""
import numpy as np
import pytest

from odelib_b200.tracer import TraceError, trace
from oracle import odelib_oracle as orc


@pytest.mark.parametrize("f,n,P", [(orc.zero_i, 2, 3), (orc.one_i, 3, 4), (orc.two_i, 4, 5)])
def test_demo_models_trace_bit_exact(f, n, P):
    m = trace(f, n, P)
    rng = np.random.default_rng(1)
    for _ in range(20):
        y = rng.lognormal(10, 3, n)
        p = rng.lognormal(-5, 4, P)
        np.testing.assert_array_equal(m.evaluate(m.outputs, y, 0.3, p), f(y, 0.3, p))
    assert m.autonomous


def _fd_jac(f, y, t, p, eps=1e-6):
    n = len(y)
    J = np.zeros((n, n))
    for j in range(n):
        d = np.zeros(n); d[j] = eps * max(1.0, abs(y[j]))
        J[:, j] = (np.asarray(f(y + d, t, p), float) - np.asarray(f(y - d, t, p), float)) / (2 * d[j])
    return J


def test_jacobian_against_finite_differences():
    def f(y, t, ps):
        return np.array([np.exp(-ps[0] * t) * y[0] ** 2 - np.sqrt(y[1]) / (1 + y[0]),
                         np.maximum(y[0], y[1]) * ps[1] + 3.0 + y[0] ** 1.5 * np.log(y[1]),
                         y[2] / (y[0] + ps[2]) - np.tanh(y[2]) + 2 ** y[1]])
    m = trace(f, 3, 3)
    y = np.array([1.3, 0.7, 2.1]); p = np.array([0.5, 1.7, 0.9]); t = 0.4
    J = m.jacobian()
    flat = [J[i][j] for i in range(3) for j in range(3)]
    Jv = np.array(m.evaluate(flat, y, t, p)).reshape(3, 3)
    np.testing.assert_allclose(Jv, _fd_jac(f, y, t, p), rtol=1e-6, atol=1e-8)
    assert not m.autonomous
    ft = np.array(m.evaluate(m.dfdt(), y, t, p))
    fd = (np.asarray(f(y, t + 1e-6, p), float) - np.asarray(f(y, t - 1e-6, p), float)) / 2e-6
    np.testing.assert_allclose(ft, fd, rtol=1e-6, atol=1e-9)


def test_two_i_jacobian_sparsity_and_source():
    m = trace(orc.two_i, 4, 5)
    sp = m.jacobian_sparsity()
    assert sum(map(sum, sp)) == 10                      # 10/16 non-zeros
    src = m.cuda_source(fmad=False, observe_groups=[(0, 1, 2), (3,)])
    assert "#define ODL_N 4" in src and "#define ODL_NOUT 2" in src
    assert "__dmul_rn" in src and "odl_jac" in src and "odl_observe" in src


def test_control_flow_is_rejected_loudly():
    def bad(y, t, ps):
        return np.array([ps[0] * y[0] if y[0] > 0 else 0.0])
    with pytest.raises(TraceError):
        trace(bad, 1, 1)


def test_wrong_arity_is_rejected():
    with pytest.raises(TraceError):
        trace(orc.zero_i, 3, 3)


def test_equality_on_traced_values_is_refused():
    """`==` / `!=` on a proxy must raise like the ordering comparisons do (object identity would silently pick a branch)."""
    from odelib_b200.tracer import TraceError, trace

    def rhs_eq(y, t, ps):
        if y[0] == 0:
            return [ps[0] * y[0]]
        return [-ps[0] * y[0]]

    def rhs_ne(y, t, ps):
        return [ps[0] * y[0] if ps[0] != 0 else y[0]]

    for f in (rhs_eq, rhs_ne):
        with pytest.raises(TraceError):
            trace(f, 1, 1)


@pytest.mark.parametrize("which,lanes", [("network", 8), ("network", 16), ("n_class_10", 4), ("two_i", 4)])
def test_slice_plan_reproduces_every_output(which, lanes):
    """tracer.slice_plan: outputs grouped into classes of identical shape, laid out class by class, `lanes` per round;
    evaluating every output through its class code + index-table row gives exactly the traced value, every component
    appears once, and the emitted odl_rhs_slice covers every round."""
    from odelib_b200 import demo_models
    from odelib_b200.tracer import trace
    f, n, P = {"network": demo_models.network(5, 5)[:3], "n_class_10": (demo_models.n_class(10), 12, 5),
               "two_i": demo_models.MODELS["two_i"][:3]}[which]
    tm = trace(f, n, P)
    plan = tm.slice_plan(lanes)
    assert sorted(k for k in plan["perm"] if k >= 0) == list(range(n)) and len(plan["perm"]) == plan["rounds"] * lanes
    rng = np.random.default_rng(0)
    y, p = rng.uniform(0.5, 2.0, n), rng.uniform(0.1, 1.0, P)
    assert tm.evaluate_sliced(plan, y, 0.3, p) == tm.evaluate(tm.outputs, y, 0.3, p)          # bit for bit
    if which == "network":
        assert len(plan["classes"]) == 3                              # dS_i, dI_ij, dV_j
    src = tm.cuda_source(coop_lanes=lanes)
    assert f"#define ODL_CS {plan['rounds']}" in src and src.count("f[") >= plan["rounds"]
    assert "ODL_COOP_SLICED" not in tm.cuda_source()
