"""B200-native multigrid solve phase behind the API of JuliaInv/Multigrid.jl.

Host side (this package): the reference's setup restated with scipy.sparse
(geometric and SA-AMG), producing hierarchies in the reference's storage
convention.  Device side (csrc/ -> libmgb200.so, C ABI in include/mgb200.h):
every operation of the V/F/W/K cycle and the Krylov drivers as hand-written
CUDA for sm_100a.  There is no CPU fallback for the solve phase: the solve
functions raise if the CUDA library is missing.
"""
from .mesh import (RegularMesh, getRegularMesh, getNodalGradientMatrix, getNodalLaplacianMatrix,
                   getNodalDivSigGradMatrix, nodal_stencil_matrix, poisson_shifted,
                   helmholtz_shifted, edge_weights_from_cells)
from .transfer import getFWInterp, get1DFWInterp
from .mgdef import (MGparam, getMGparam, hierarchyExists, clear, copySolver, destroyCoarsestLU,
                    multilevelOperatorConstructor, getMultilevelOperatorConstructor,
                    Float64, ComplexF64, Float32, ComplexF32, Int64)
from .mgsetup import (MGsetup, getRelaxPrec, getSPAIprec, adjustMemoryForNumRHS,
                      replaceMatrixInHierarchy, transposeHierarchy, defineCoarsestAinv)
from .sa_amg import (SA_AMGsetup, getAggregation, getStrengthMatrix, neighborhoodAggregationNew,
                     aggrArray2P)
from .classical_amg import (ClassicalAMGsetup, getStrengthMatrixClassical, getColoringFirst, getColoringSecond,
                            getInterpolation)
from .systems import (getLinearOperatorsSystemsFaces, getInjectionOperatorsSystemsFaces, getLinearInterpolationFacesUj,
                      getRestrictionFacesFullWeightUj, getRestrictionFacesInjectionUj, getRestrictionCellCentered,
                      getLinearInterpolationCellCentered, faces_size)
from .device import DeviceHierarchy, MultiDeviceHierarchy, uploadHierarchy, MGB200Error, LIB_PATH
from .solve import (solveMG, solveCG_MG, solveGMRES_MG, solveBiCGSTAB_MG, getMultigridPreconditioner, recursiveCycle,
                    SpMatMul)
from .dist_setup import (slab_planes, setup_slab_hierarchy, poisson_window_operator, DistHierarchy, DistLevel)
from .wrappers import (MGsolver, SA_AMGsolver, getMGsolver, getSA_AMGsolver, solveLinearSystem, setupSolver)
from . import wrappers
