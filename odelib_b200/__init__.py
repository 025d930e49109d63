"""odelib_b200 -- B200-native (sm_100a) hot path of ODElib behind ODElib's own Python surface.

    import odelib_b200 as ODElib
    m = ODElib.ModelFramework(ODE=f, parameter_names=[...], state_names=[...], dataframe=df, ...)
    posterior = m.MCMC(chain_inits=32, iterations_per_chain=1000, fitsurvey_samples=10000, sd_fitdistance=6.0)
"""
__version__ = "0.1.0"

from .Framework import ModelFramework, parameter  # noqa: E402,F401
from .Statistics import Samplers, stats  # noqa: E402,F401
