"""GpuNLPEvaluator: the MOI.AbstractNLPEvaluator method set (enumerated by the reference's EmptyNLPEvaluator,
src/opt/MOI_wrapper.jl:56-77) implemented on the CUDA library.  Same method names, argument order and
in-place output semantics as the Julia interface, so the tests read like calls through MathOptInterface.
Index tuples are 1-based like MOI's `Vector{Tuple{Int64,Int64}}`."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

from . import capi
from .types import ACOPF_fromulation


class GpuNLPEvaluator:
    def __init__(self, data, formulation: ACOPF_fromulation, FACTS_bShunt: bool = True, obj_linear: bool = True,
                 device: int = 0):
        self.handle = capi.Handle(data, formulation.code, FACTS_bShunt, obj_linear, device=device)
        self._initialized: List[str] = []

    # -- MOI.features_available(d) (MOI_wrapper.jl:1111)
    def features_available(self) -> List[str]:
        return ["Grad", "Jac", "Hess"]

    # -- MOI.initialize(d, requested_features) (MOI_wrapper.jl:1117)
    def initialize(self, requested_features: List[str]) -> None:
        for f in requested_features:
            if f not in self.features_available():
                raise ValueError(f"Unsupported feature {f}")
        self._initialized = list(requested_features)

    # -- objective: Powersense OPF models have no NL objective (has_objective = false, MOI_wrapper.jl:897,938)
    def eval_objective(self, x) -> float:
        raise RuntimeError("the NLP block has no objective (has_objective = false)")

    def eval_objective_gradient(self, g, x) -> None:
        raise RuntimeError("the NLP block has no objective (has_objective = false)")

    # -- MOI.NLPBlockData.constraint_bounds (MOI_wrapper.jl:1091-1094)
    def constraint_bounds(self) -> Tuple[np.ndarray, np.ndarray]:
        return self.handle.constraint_bounds()

    # -- MOI.eval_constraint(d, g, x) (MOI_wrapper.jl:969): fills g in place (g may be a view of a larger array)
    def eval_constraint(self, g: np.ndarray, x) -> None:
        self.handle.eval_constraint(x, g)

    # -- MOI.jacobian_structure(d) (MOI_wrapper.jl:813)
    def jacobian_structure(self) -> List[Tuple[int, int]]:
        r, c = self.handle.jacobian_structure()
        return list(zip(r.tolist(), c.tolist()))

    # -- MOI.eval_constraint_jacobian(d, J, x) (MOI_wrapper.jl:1026)
    def eval_constraint_jacobian(self, J: np.ndarray, x) -> None:
        self.handle.eval_jacobian(x, J)

    # -- MOI.hessian_lagrangian_structure(d) (MOI_wrapper.jl:855)
    def hessian_lagrangian_structure(self) -> List[Tuple[int, int]]:
        if "Hess" not in self._initialized:
            raise RuntimeError("Hessian requested but not initialized")
        r, c = self.handle.hessian_structure()
        return list(zip(r.tolist(), c.tolist()))

    # -- MOI.eval_hessian_lagrangian(d, H, x, sigma, mu) (MOI_wrapper.jl:1061)
    def eval_hessian_lagrangian(self, H: np.ndarray, x, sigma: float, mu) -> None:
        if "Hess" not in self._initialized:
            raise RuntimeError("Hessian requested but not initialized")
        self.handle.eval_hessian(x, sigma, mu, H)

    @property
    def num_variables(self) -> int:
        return self.handle.n_var

    @property
    def num_constraints(self) -> int:
        return self.handle.m_nl

    def close(self):
        self.handle.close()
