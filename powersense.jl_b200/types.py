"""Formulation tags (mirror of reference src/opf/type.jl:1-19).

The reference uses nine identity-compared singletons of an empty struct; the C ABI
(include/powersense_b200.h) uses the integer codes below in the same order as type.jl:7-15.
"""
from __future__ import annotations


class ACOPF_fromulation:  # sic: the reference spells it this way (type.jl:1)
    __slots__ = ("name", "code")

    def __init__(self, name: str, code: int):
        self.name = name
        self.code = code

    def __repr__(self) -> str:  # pragma: no cover - cosmetic
        return self.name


PBRAPVmodel = ACOPF_fromulation("PBRAPVmodel", 0)  # Power Branch Flow Rectangular Admittance Polar Voltage
PBRARVmodel = ACOPF_fromulation("PBRARVmodel", 1)  # ... Rectangular Voltage
PBRAWVmodel = ACOPF_fromulation("PBRAWVmodel", 2)  # ... W-matrix Voltage
CBRARVmodel = ACOPF_fromulation("CBRARVmodel", 3)  # Current Branch Flow ... Rectangular Voltage
CBRAWVmodel = ACOPF_fromulation("CBRAWVmodel", 4)  # Current Branch Flow ... W-matrix Voltage
PNPAPVmodel = ACOPF_fromulation("PNPAPVmodel", 5)  # Power Nodal Flow Polar Admittance Polar Voltage
PNRAPVmodel = ACOPF_fromulation("PNRAPVmodel", 6)  # Power Nodal Flow Rectangular Admittance Polar Voltage
PNRARVmodel = ACOPF_fromulation("PNRARVmodel", 7)  # Power Nodal Flow Rectangular Admittance Rectangular Voltage
PNRAWVmodel = ACOPF_fromulation("PNRAWVmodel", 8)  # Power Nodal Flow Rectangular Admittance W-matrix Voltage

ALL_FORMULATIONS = [PBRAPVmodel, PBRARVmodel, PBRAWVmodel, CBRARVmodel, CBRAWVmodel,
                    PNPAPVmodel, PNRAPVmodel, PNRARVmodel, PNRAWVmodel]
# the seven with non-linear nodal balances (README.md:13 "7 different AC-OPF formulations")
MAIN_FORMULATIONS = [PBRAPVmodel, PBRARVmodel, CBRARVmodel, CBRAWVmodel,
                     PNPAPVmodel, PNRAPVmodel, PNRARVmodel]
BY_NAME = {f.name: f for f in ALL_FORMULATIONS}
BY_CODE = {f.code: f for f in ALL_FORMULATIONS}

SRC_MATPOWER = 0
SRC_ARPAE = 1

FLAG_FACTS_BSHUNT = 1   # run_opf!(FACTS_bShunt=true) adds bss variables (formulations.jl:65-67)
FLAG_OBJ_LINEAR = 2     # obj_type=:linear adds cg (formulations.jl:60-62) and tgh (:121) variables

EVAL_G = 1
EVAL_J = 2
EVAL_H = 4


def source_code(source_type: str) -> int:
    """formulations.jl:299 selects the MATPOWER limit family iff source_type == "matpower"."""
    return SRC_MATPOWER if source_type == "matpower" else SRC_ARPAE
