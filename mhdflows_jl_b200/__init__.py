"""mhdflows_jl_b200 -- B200-native (sm_100a) implementation of MHDFlows.jl's 3D periodic pseudospectral
right-hand side and RK4/LSRK54 time step, behind the reference's problem API.

The compute lives in libmhdflows_b200.so (hand-written CUDA, C ABI in include/mhdflows_b200.h); this package
is the thin host mirror of the Julia API.  No CPU fallback exists.
"""
from ._lib import EMHD, F32, F64, FRESH, HD, LSRK54, MHD, RK4, STAGE, STALE, MHDFlowsError  # noqa: F401
from .problem import (Cylindrical_Mask_Function, A99GPU, A99_vars, A99ForceDriving, DivBCorrection, DivVCorrection, GetA99vars_And_function, SetUpFk,  # noqa: F401
                      CPU, GPU, Diagnostic, DivFreeSpectraMap, GetN97vars_And_function, N97ForceDriving,  # noqa: F401
                      ProbDiagnostic, Problem, SetUpN97, SetUpProblemIC, TimeIntegrator, getCFL, increment,
                      nothingfunction, spectralline, stepforward, SetUpRandomPhaseIC, DFSM_CALL, ScaleDecomposition, VectorPotential, CF, SFC, SF2_1D,
                      NDForceDriving, GetNDvars_And_function, SetUpND, ND_vars)

from .io import Restart, readMHDFlows, savefile  # noqa: F401,E402

__all__ = ["savefile", "Restart", "readMHDFlows", "Problem", "SetUpProblemIC", "stepforward", "TimeIntegrator", "getCFL", "ProbDiagnostic", "Diagnostic",
           "increment", "DivFreeSpectraMap", "SetUpRandomPhaseIC", "ScaleDecomposition", "VectorPotential", "CF", "SFC", "SF2_1D", "NDForceDriving", "GetNDvars_And_function", "SetUpND", "ND_vars", "spectralline", "N97ForceDriving", "GetN97vars_And_function", "SetUpN97", "A99ForceDriving", "GetA99vars_And_function", "SetUpFk", "A99GPU", "A99_vars",
           "DivVCorrection", "DivBCorrection", "Cylindrical_Mask_Function", "CPU", "GPU", "nothingfunction", "MHDFlowsError",
           "FRESH", "STALE", "STAGE"]
