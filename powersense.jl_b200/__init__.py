"""powersense_b200: B200-native evaluator for the Powersense.jl AC-OPF hot path.

Host-side mirror of the reference's public API for this path (src/Powersense.jl:12 exports):
create_PowersenseData, create_opf_model!, run_opf! -- plus the MOI evaluator object the
solvers consume.  All numeric work is done by the CUDA library behind include/powersense_b200.h;
there is no CPU fallback (importing `.capi` raises if the library is missing)."""
from .types import *  # noqa: F401,F403
from .data import Network, PowersenseData, create_PowersenseData  # noqa: F401
from .evaluator import GpuNLPEvaluator  # noqa: F401,E402
from .opf import OPFModel, create_opf_model, run_opf  # noqa: F401,E402
