"""trixi.jl_b200 -- B200-native drop-in for Trixi.jl's DGSEM hot path.

Host-side mirror of the reference API (names as exported by ``src/Trixi.jl:204-396``) above the
C ABI of ``libtrixi_b200.so`` (``include/trixi_b200.h``).  Import it as ``trixi_b200`` (the directory
name contains a dot; ``trixi_b200.py`` at the repository root is the loader).
"""
from .basis import LobattoLegendreBasis, SolutionAnalyzer, gauss_lobatto_nodes_weights  # noqa: F401
from .callbacks import (AliveCallback, AnalysisCallback, GlmSpeedCallback, StepsizeCallback,  # noqa: F401
                        SummaryCallback, calc_error_norms, calc_error_norms_device, integrate_device)
from .equations import *  # noqa: F401,F403
from .mesh import CartesianBoxMesh, TreeMesh  # noqa: F401
from .p4est import P4estMesh  # noqa: F401
from .structured import StructuredMesh  # noqa: F401
from .semidiscretization import (ODEProblem, SemidiscretizationHyperbolic, compute_coefficients,  # noqa: F401
                                 rhs_hyperbolic, semidiscretize)
from .solver import (DGSEM, IndicatorHennemannGassner, SurfaceIntegralWeakForm,  # noqa: F401
                     VolumeIntegralFluxDifferencing, VolumeIntegralPureLGLFiniteVolume, VolumeIntegralShockCapturingHG, VolumeIntegralWeakForm)
from .time_integration import (CallbackSet, CarpenterKennedy2N43, CarpenterKennedy2N54,  # noqa: F401
                               ParsaniKetchesonDeconinck3Sstar32, ParsaniKetchesonDeconinck3Sstar94,
                               SimpleSSPRK33, init, solve, step, step_2n_host)
