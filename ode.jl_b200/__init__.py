"""ode.jl_b200 -- host mirror of ODE.jl's solver surface over libode_b200.so.

Only what the RK / ode23s hot path needs: the odeXX front-ends, the
common-interface `solve`, the registered right-hand sides, and drivers for the
ensemble and large-system regimes.  All arithmetic happens in the CUDA library
(csrc/); importing this package without the built library raises.
"""
from ._lib import lib, LIB_PATH, OdeB200Error, STATUS  # noqa: F401
from .rhs import DeviceRHS, register_rhs, register_stencil, register_field  # noqa: F401
from .api import (ode1, ode2_midpoint, ode2_heun, ode4, ode21, ode23, ode45_fe, ode45_dp, ode45, ode45_ck, ode78, ode23s,  # noqa: F401
                  ode4s, ode4s_s, ode4s_kr, ode4ms, ode5ms, ode_ms, ode_ms1, ode_ms2, ode_ms3, solve_ensemble, make_opts, Solver, pinned_empty)
from .common import ODEProblem, ODESolution, solve  # noqa: F401
from .large import (solve_large, solve_large_distributed, shard_bounds, integrate_slab,  # noqa: F401
                    CudaSlabEngine, connect_peers)
from .sharding import solve_ensemble_distributed, SharedEnsembleResult  # noqa: F401

lib()  # fail loudly at import if the CUDA library is missing
