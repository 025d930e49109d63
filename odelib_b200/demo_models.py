"""The demo's host-virus infection-state models (workload definitions for tests, smoke and bench).

Written from the model equations of the reference demo (Demo_InfectionStates.ipynb:60-128): susceptible
hosts S grow at rate mu and are infected at rate phi*S*V; infected classes I1..In progress at rate tau and
lyse at rate lam releasing beta virions.  ``n_class`` generalises one_i/two_i to N latent infected classes
(BASELINE.json config 3).
"""
import numpy as np


def zero_i(y, t, ps):
    mu, phi, beta = ps[0], ps[1], ps[2]
    S, V = y[0], y[1]
    dSdt = mu * S - phi * S * V
    dVdt = beta * phi * S * V - phi * S * V
    return np.array([dSdt, dVdt])


def one_i(y, t, ps):
    mu, phi, beta, lam = ps[0], ps[1], ps[2], ps[3]
    S, I1, V = y[0], y[1], y[2]
    dSdt = mu * S - phi * S * V
    dI1dt = phi * S * V - lam * I1
    dVdt = beta * lam * I1 - phi * S * V
    return np.array([dSdt, dI1dt, dVdt])


def two_i(y, t, ps):
    mu, phi, beta, lam, tau = ps[0], ps[1], ps[2], ps[3], ps[4]
    S, I1, I2, V = y[0], y[1], y[2], y[3]
    dSdt = mu * S - phi * S * V
    dI1dt = phi * S * V - tau * I1
    dI2dt = tau * I1 - lam * I2
    dVdt = beta * lam * I2 - phi * S * V
    return np.array([dSdt, dI1dt, dI2dt, dVdt])


def n_class(N):
    """S, I1..IN, V with N >= 2 infected classes; parameters mu, phi, beta, lam, tau."""
    if N < 2:
        raise ValueError("use zero_i / one_i for N < 2")

    def model(y, t, ps):
        mu, phi, beta, lam, tau = ps[0], ps[1], ps[2], ps[3], ps[4]
        S, V = y[0], y[N + 1]
        inf = phi * S * V
        d = [mu * S - inf, inf - tau * y[1]]
        for k in range(2, N):
            d.append(tau * y[k - 1] - tau * y[k])
        d.append(tau * y[N - 1] - lam * y[N])
        d.append(beta * lam * y[N] - inf)
        return np.array(d)

    model.__name__ = f"n_class_{N}"
    return model


# name -> (rhs, n_state, n_param, observe groups (state indices summed into each output column))
MODELS = {
    "zero_i": (zero_i, 2, 3, [(0,), (1,)]),
    "one_i": (one_i, 3, 4, [(0, 1), (2,)]),
    "two_i": (two_i, 4, 5, [(0, 1, 2), (3,)]),
}
PARAMETER_NAMES = {"zero_i": ["mu", "phi", "beta"], "one_i": ["mu", "phi", "beta", "lam"],
                   "two_i": ["mu", "phi", "beta", "lam", "tau"]}
STATE_NAMES = {"zero_i": ["S", "V"], "one_i": ["S", "I1", "V"], "two_i": ["S", "I1", "I2", "V"]}
# lognorm (s, scale) priors of the demo notebook (:885-891, :8575-8578, :17472-17476)
PRIORS = {
    "zero_i": {"mu": (3, 1e-8), "phi": (3, 1e-8), "beta": (1, 25)},
    "one_i": {"mu": (3, 1e-8), "phi": (3, 1e-8), "beta": (1, 20), "lam": (2, 0.1)},
    "two_i": {"mu": (3, 1e-8), "phi": (3, 1e-8), "beta": (1, 20), "lam": (2, 0.1), "tau": (2, 1)},
}


def network(n_host=5, n_virus=5):
    """Multi-strain host-virus network (BASELINE.json config 5).

    States: S_i (n_host), I_ij (n_host*n_virus, row-major), V_j (n_virus).
    Parameters: mu_i (n_host), phi_ij (n_host*n_virus), beta_j (n_virus), lam_j (n_virus).
    dS_i = mu_i S_i - sum_j phi_ij S_i V_j;  dI_ij = phi_ij S_i V_j - lam_j I_ij;
    dV_j = sum_i beta_j lam_j I_ij - sum_i phi_ij S_i V_j.
    Observables: H_i = S_i + sum_j I_ij and V_j."""
    H, V = n_host, n_virus

    def model(y, t, ps):
        S = [y[i] for i in range(H)]
        I = [[y[H + i * V + j] for j in range(V)] for i in range(H)]
        Vv = [y[H + H * V + j] for j in range(V)]
        mu = [ps[i] for i in range(H)]
        phi = [[ps[H + i * V + j] for j in range(V)] for i in range(H)]
        beta = [ps[H + H * V + j] for j in range(V)]
        lam = [ps[H + H * V + V + j] for j in range(V)]
        inf = [[phi[i][j] * S[i] * Vv[j] for j in range(V)] for i in range(H)]
        dS = []
        for i in range(H):
            acc = mu[i] * S[i]
            for j in range(V):
                acc = acc - inf[i][j]
            dS.append(acc)
        dI = [inf[i][j] - lam[j] * I[i][j] for i in range(H) for j in range(V)]
        dV = []
        for j in range(V):
            acc = beta[j] * lam[j] * I[0][j]
            for i in range(1, H):
                acc = acc + beta[j] * lam[j] * I[i][j]
            for i in range(H):
                acc = acc - inf[i][j]
            dV.append(acc)
        return np.array(dS + dI + dV)

    model.__name__ = f"network_{H}x{V}"
    n_state = H + H * V + V
    n_param = H + H * V + V + V
    groups = [tuple([i] + [H + i * V + j for j in range(V)]) for i in range(H)] + [(H + H * V + j,) for j in range(V)]
    return model, n_state, n_param, groups
