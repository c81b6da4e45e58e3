# ccall shim that re-points the solve phase of Multigrid.jl at libmgb200.so.
#
# It cannot be executed in this repository's build image (no Julia); it is the binding a
# maintainer adds next to src/Multigrid/SolveFuncs.jl.  The host-side setup (MGsetup,
# SA_AMGsetup, getMGparam, ...) stays untouched; only the functions below change.
#
# The library is located like the reference's own native libs (src/Multigrid/parRelax.jl:3,
# Vanka.jl:7): deps/builds/<name>.
module MultigridB200

using SparseArrays, LinearAlgebra
using Multigrid   # the reference package: MGparam, MGsetup, SA_AMGsetup, hierarchyExists, ...

const libmgb200 = joinpath(dirname(pathof(Multigrid)), "..", "deps", "builds", "libmgb200")

const MGB200_FP64  = Cint(0)
const MGB200_CFP64 = Cint(1)
const MGB200_FP32  = Cint(2)   # getMGparam(Float32, ...) / getMGparam(ComplexF32, ...): `singlePrecision`
const MGB200_CFP32 = Cint(3)   # hierarchies (MGdef.jl:119,151; MGsetup.jl:31-33,79-82,108-110)
valtype_code(::Type{Float64})    = MGB200_FP64
valtype_code(::Type{ComplexF64}) = MGB200_CFP64
valtype_code(::Type{Float32})    = MGB200_FP32
valtype_code(::Type{ComplexF32}) = MGB200_CFP32

check(status::Cint) = status == 0 ? nothing :
    error("mgb200 status $status: ", unsafe_string(ccall((:mgb200_last_error, libmgb200), Cstring, ())))

mutable struct DeviceHierarchy
    handle::Ptr{Cvoid}
end

"""
    uploadHierarchy(param; device=0) -> DeviceHierarchy

Hands the raw CSC arrays of param.As / Ps / Rs (adjoint storage, 1-based Int64: MGdef.jl:75-77)
and param.relaxPrecs to the device.  Julia owns the arrays; the library copies during the call.
"""
function uploadHierarchy(param::MGparam{VAL,Int64}; device::Integer=0) where {VAL}
    hierarchyExists(param) || error("You have to do a setup first.")
    levels = length(param.As)
    nrhs = length(param.memCycle) > 0 ? size(param.memCycle[1].x, 2) : 1
    pre  = Int64[param.relaxPre(l)  for l = 1:levels]
    post = Int64[param.relaxPost(l) for l = 1:levels]
    relaxKind = param.relaxType == "Jac-GMRES" ? Cint(1) : Cint(0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:mgb200_create, libmgb200), Cint,
                (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cchar, Cint, Ptr{Int64}, Ptr{Int64}, Cint),
                h, valtype_code(VAL), levels, nrhs, Cchar(param.cycleType), relaxKind, pre, post, device))
    for l = 1:levels-1
        AT, PT, RT = param.As[l], param.Ps[l], param.Rs[l]
        if length(param.Meshes) > l      # geometric hierarchy: the meshes are a hint for the transfer kernels (optional)
            nf = Int64.(param.Meshes[l].n .+ 1); nc = Int64.(param.Meshes[l+1].n .+ 1)
            check(ccall((:mgb200_set_level_grid, libmgb200), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Ptr{Int64}),
                        h[], l, length(nf), nf, nc))
        end
        d = convert(Vector{VAL}, param.relaxPrecs[l])
        check(ccall((:mgb200_upload_level, libmgb200), Cint,
                    (Ptr{Cvoid}, Cint, Int64, Int64,
                     Ptr{Int64}, Ptr{Int64}, Ptr{VAL},
                     Ptr{Int64}, Ptr{Int64}, Ptr{real(VAL)},      # Ps, Rs hold real(VAL) values
                     Ptr{Int64}, Ptr{Int64}, Ptr{real(VAL)},
                     Ptr{VAL}, Cint),
                    h[], l, size(AT, 2), size(param.As[l+1], 2),
                    AT.colptr, AT.rowval, AT.nzval,
                    PT.colptr, PT.rowval, PT.nzval,
                    RT.colptr, RT.rowval, RT.nzval,
                    d, 1))
    end
    AL = param.As[end]
    if param.coarseSolveType == "GMRES"      # MGsetup.jl:333-334: param.LU holds the Jacobi weights
        check(ccall((:mgb200_upload_coarsest_gmres, libmgb200), Cint,
                    (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{VAL}, Ptr{VAL}, Cint),
                    h[], size(AL, 2), AL.colptr, AL.rowval, AL.nzval, param.LU, 1))
    else
        check(ccall((:mgb200_upload_coarsest, libmgb200), Cint,
                    (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{VAL}, Cint),
                    h[], size(AL, 2), AL.colptr, AL.rowval, AL.nzval, 1))
    end
    dev = DeviceHierarchy(h[])
    finalizer(d -> ccall((:mgb200_destroy, libmgb200), Cint, (Ptr{Cvoid},), d.handle), dev)
    return dev
end

# ---- src/Multigrid/SolveFuncs.jl:3-39 -----------------------------------------------------------
function solveMG(param::MGparam{VAL,Int64}, dev::DeviceHierarchy, b::Array{VAL}, x::Array{VAL},
                 verbose::Bool) where {VAL}
    check(ccall((:mgb200_adjust_nrhs, libmgb200), Cint, (Ptr{Cvoid}, Cint), dev.handle, size(b, 2)))
    iter = Ref{Cint}(0)
    resvec = zeros(param.maxOuterIter + 1)
    check(ccall((:mgb200_solveMG, libmgb200), Cint,
                (Ptr{Cvoid}, Ptr{VAL}, Ptr{VAL}, Cdouble, Cint, Ref{Cint}, Ptr{Cdouble}),
                dev.handle, b, x, param.relativeTol, param.maxOuterIter, iter, resvec))
    if verbose
        for k = 1:iter[]
            println("Cycle ", k, " done with relres: ", resvec[k+1] / resvec[1],
                    ". Convergence factor: ", resvec[k+1] / resvec[k])
        end
    end
    return x, param, Int(iter[])
end

# ---- src/Multigrid/SolveFuncs.jl:103-116 ----------------------------------------------------------
function solveCG_MG(AT::SparseMatrixCSC{VAL,Int64}, param::MGparam{VAL,Int64}, dev::DeviceHierarchy,
                    b::Array{VAL}, x0::Array{VAL}, verbose::Bool=false) where {VAL}
    check(ccall((:mgb200_adjust_nrhs, libmgb200), Cint, (Ptr{Cvoid}, Cint), dev.handle, size(b, 2)))
    if AT !== param.As[1]
        check(ccall((:mgb200_set_krylov_matrix, libmgb200), Cint,
                    (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{VAL}, Cint),
                    dev.handle, size(AT, 2), AT.colptr, AT.rowval, AT.nzval, 1))
    end
    iter = Ref{Cint}(0); flag = Ref{Cint}(0)
    resvec = zeros(param.maxOuterIter * size(b, 2))
    check(ccall((:mgb200_solveCG, libmgb200), Cint,
                (Ptr{Cvoid}, Ptr{VAL}, Ptr{VAL}, Cdouble, Cint, Ref{Cint}, Ref{Cint}, Ptr{Cdouble}),
                dev.handle, b, x0, param.relativeTol, param.maxOuterIter, iter, flag, resvec))
    return x0, param, Int(iter[])
end

# ---- mixed precision (SolveFuncs.jl:52-60: VAL != eltype(B)) --------------------------------------
# A Float32 / ComplexF32 hierarchy under Float64 / ComplexF64 Krylov vectors: the outer handle holds only the Krylov
# matrix AT (double precision) and the Krylov vectors, its preconditioner is one cycle of `dev` on a rounded copy of
# the residual.  The Krylov front ends below are then called with the outer handle and double-precision b, x0:
#     dev32 = uploadHierarchy(param32);  outer = mixedPrecisionHandle(dev32, AT64)
#     solveCG_MG(AT64, param32, outer, b64, x64)        # dispatches on eltype(b), like the reference
function mixedPrecisionHandle(dev::DeviceHierarchy, AT::SparseMatrixCSC{VALD,Int64}) where {VALD<:Union{Float64,ComplexF64}}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:mgb200_create_mixed, libmgb200), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cvoid}), h, dev.handle))
    check(ccall((:mgb200_set_krylov_matrix, libmgb200), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{VALD}, Cint),
                h[], size(AT, 2), AT.colptr, AT.rowval, AT.nzval, 1))
    outer = DeviceHierarchy(h[])
    finalizer(d -> ccall((:mgb200_destroy, libmgb200), Cint, (Ptr{Cvoid},), d.handle), outer)
    return outer     # `dev` must outlive it
end
function solveCG_MG(AT::SparseMatrixCSC{VALD,Int64}, param::MGparam{VAL,Int64}, outer::DeviceHierarchy,
                    inner::DeviceHierarchy, b::Array{VALD}, x0::Array{VALD}) where {VAL<:Union{Float32,ComplexF32},VALD<:Union{Float64,ComplexF64}}
    for h in (inner.handle, outer.handle)
        check(ccall((:mgb200_adjust_nrhs, libmgb200), Cint, (Ptr{Cvoid}, Cint), h, size(b, 2)))
    end
    iter = Ref{Cint}(0); flag = Ref{Cint}(0)
    resvec = zeros(param.maxOuterIter * size(b, 2))
    check(ccall((:mgb200_solveCG, libmgb200), Cint,
                (Ptr{Cvoid}, Ptr{VALD}, Ptr{VALD}, Cdouble, Cint, Ref{Cint}, Ref{Cint}, Ptr{Cdouble}),
                outer.handle, b, x0, param.relativeTol, param.maxOuterIter, iter, flag, resvec))
    return x0, param, Int(iter[])
end

# ---- src/Multigrid/SolveFuncs.jl:85-99 ------------------------------------------------------------
function solveBiCGSTAB_MG(AT::SparseMatrixCSC{VAL,Int64}, param::MGparam{VAL,Int64}, dev::DeviceHierarchy,
                          b::Array{VAL}, x0::Array{VAL}, verbose::Bool=false) where {VAL}
    check(ccall((:mgb200_adjust_nrhs, libmgb200), Cint, (Ptr{Cvoid}, Cint), dev.handle, size(b, 2)))
    iter = Ref{Cint}(0); flag = Ref{Cint}(0); nprec = Ref{Cint}(0)
    resvec = zeros(param.maxOuterIter + 1)
    check(ccall((:mgb200_solveBiCGSTAB, libmgb200), Cint,
                (Ptr{Cvoid}, Ptr{VAL}, Ptr{VAL}, Cdouble, Cint, Ref{Cint}, Ref{Cint}, Ptr{Cdouble}, Ref{Cint}),
                dev.handle, b, x0, param.relativeTol, param.maxOuterIter, iter, flag, resvec, nprec))
    return x0, param, Int(iter[]), Int(nprec[])
end

# ---- src/Multigrid/SolveFuncs.jl:120-132 ----------------------------------------------------------
function solveGMRES_MG(AT::SparseMatrixCSC{VAL,Int64}, param::MGparam{VAL,Int64}, dev::DeviceHierarchy,
                       b::Array{VAL}, x0::Array{VAL}, flexible::Bool, inner::Int64,
                       verbose::Bool=false) where {VAL}
    check(ccall((:mgb200_adjust_nrhs, libmgb200), Cint, (Ptr{Cvoid}, Cint), dev.handle, size(b, 2)))
    iter = Ref{Cint}(0); flag = Ref{Cint}(0); nres = Ref{Cint}(0)
    resvec = zeros(inner * param.maxOuterIter)
    check(ccall((:mgb200_solveFGMRES, libmgb200), Cint,
                (Ptr{Cvoid}, Ptr{VAL}, Ptr{VAL}, Cint, Cint, Cdouble, Cint, Ref{Cint}, Ref{Cint},
                 Ptr{Cdouble}, Ref{Cint}),
                dev.handle, b, x0, inner, flexible, param.relativeTol, param.maxOuterIter, iter, flag,
                resvec, nres))
    return x0, param, Int(iter[]), resvec[1:nres[]]
end

# ---- src/Multigrid/SolveFuncs.jl:43-63: r -> z (one cycle from z = 0) -----------------------------
function getMultigridPreconditioner(param::MGparam{VAL,Int64}, dev::DeviceHierarchy, B::Array) where {VAL}
    check(ccall((:mgb200_adjust_nrhs, libmgb200), Cint, (Ptr{Cvoid}, Cint), dev.handle, size(B, 2)))
    z = zeros(VAL, size(B))
    return function (r::Array{VAL})
        z .= 0.0
        check(ccall((:mgb200_cycle, libmgb200), Cint, (Ptr{Cvoid}, Ptr{VAL}, Ptr{VAL}), dev.handle, r, z))
        return z
    end
end

end # module
