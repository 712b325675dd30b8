"""porthamiltonian_fem_b200 - B200-native hot path of FinbarArgus/portHamiltonian_FEM.

Public surface (mirrors the reference's ``from wave_2D import *``):
    wave_2D_solve            drop-in for fenics_projects/wave_eqn/wave_2D.py::wave_2D_solve
    Passive_control_input    parameter holder mirroring control_input.py (the controller itself runs on the device)
Lower level: ``Mesh`` / ``Solver`` (handles of the C ABI in include/phwave.h), ``mesh`` generators.
"""
from .wave_2D import wave_2D_solve            # noqa: F401
from .control_input import Passive_control_input   # noqa: F401
from .solver import Mesh, Solver              # noqa: F401
