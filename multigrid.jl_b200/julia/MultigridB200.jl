# ccall shim that re-points the solve phase of Multigrid.jl at libmgb200.so.
#
# UNEXECUTED: this repository's build image has no Julia, so the file has never been run; every ccall signature below was
# written against include/mgb200.h and mirrors the ctypes binding multigrid.jl_b200/device.py, which IS exercised by the
# test-suite (tests/test_boundary.py feeds the same 1-based Int64 arrays through the ABI from a C++ caller).
#
# It is the binding a maintainer adds next to src/Multigrid/SolveFuncs.jl.  The host-side setup (MGsetup, SA_AMGsetup,
# getMGparam, ...) stays untouched; the front ends keep the reference's signatures and return tuples:
#     solveMG(param,b,x,verbose)                      -> (x, param, iter)           SolveFuncs.jl:3
#     solveCG_MG(AT,param,b,x0,verbose)               -> (x, param, iter)           SolveFuncs.jl:103
#     solveGMRES_MG(AT,param,b,x0,flexible,inner,v)   -> (x, param, iter, resvec)   SolveFuncs.jl:120
#     solveBiCGSTAB_MG(AT,param,b,x0,verbose)         -> (x, param, iter, nprec)    SolveFuncs.jl:85
#     getMultigridPreconditioner(param,B,verbose)     -> r -> z                     SolveFuncs.jl:43
# The device handle of a hierarchy lives in a registry keyed by the MGparam object (MGparam is the reference's struct and
# has no field for it); uploadHierarchy(param) is called at the end of MGsetup / SA_AMGsetup / replaceMatrixInHierarchy /
# transposeHierarchy, or lazily by the first solve.
#
# The library is located like the reference's own native libs (src/Multigrid/parRelax.jl:3, Vanka.jl:7): deps/builds/<name>.
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
    multi::Bool          # handle of the single-process multi-GPU entry (mgb200_multi_*)
end

# param -> its device hierarchy (identity keyed: two params with equal fields are different hierarchies)
const _devices = IdDict{Any,DeviceHierarchy}()

destroy!(d::DeviceHierarchy) = (d.handle == C_NULL && return;
    d.multi ? ccall((:mgb200_multi_destroy, libmgb200), Cint, (Ptr{Cvoid},), d.handle) :
              ccall((:mgb200_destroy, libmgb200), Cint, (Ptr{Cvoid},), d.handle);
    d.handle = C_NULL; nothing)

"clear!(param) of the reference (MGdef.jl:179-189) plus the device copy"
function clearDevice!(param::MGparam)
    haskey(_devices, param) && (destroy!(_devices[param]); delete!(_devices, param))
    return nothing
end

cycleArgs(param, levels) = (Int64[param.relaxPre(l) for l = 1:levels], Int64[param.relaxPost(l) for l = 1:levels],
                            param.relaxType == "Jac-GMRES" ? Cint(1) : Cint(0))

"""
    uploadHierarchy(param; device=0) -> DeviceHierarchy

Hands the raw CSC arrays of param.As / Ps / Rs (adjoint storage, 1-based Int64: MGdef.jl:75-77)
and param.relaxPrecs to the device.  Julia owns the arrays; the library copies during the call.
"""
function uploadHierarchy(param::MGparam{VAL,Int64}; device::Integer=0) where {VAL}
    hierarchyExists(param) || error("You have to do a setup first.")
    clearDevice!(param)
    levels = length(param.As)
    nrhs = length(param.memCycle) > 0 ? size(param.memCycle[1].x, 2) : 1
    pre, post, relaxKind = cycleArgs(param, levels)
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
                    d, 1))                                        # index_base = 1: the arrays as Julia holds them
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
    dev = DeviceHierarchy(h[], false)
    finalizer(destroy!, dev)
    _devices[param] = dev
    return dev
end

"""
    uploadHierarchyMulti(param, devices; replicateBelow=200_000) -> DeviceHierarchy

ONE handle that drives several GPUs from this Julia task (mgb200_multi_*): the global arrays are handed over unchanged,
the library slices the z-slabs of getOriginalBoundingBoxCells with NumCells = [1,1,G] (DDIndices.jl:41-47) and runs every
device on its own host thread inside each call.  Geometric hierarchies (param.Meshes), one right-hand side.
"""
function uploadHierarchyMulti(param::MGparam{VAL,Int64}, devices::Vector{<:Integer}; replicateBelow::Integer=200_000) where {VAL}
    hierarchyExists(param) || error("You have to do a setup first.")
    clearDevice!(param)
    G = length(devices); levels = length(param.As)
    pre, post, relaxKind = cycleArgs(param, levels)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:mgb200_multi_create, libmgb200), Cint,
                (Ref{Ptr{Cvoid}}, Cint, Ptr{Cint}, Cint, Cint, Cint, Cchar, Cint, Ptr{Int64}, Ptr{Int64}),
                h, G, Cint.(devices), valtype_code(VAL), levels, 1, Cchar(param.cycleType), relaxKind, pre, post))
    # slab g owns node planes g*c .. (g+1)*c - 1 (0-based), the last slab also the final plane
    offsets(ncells3, plane, n) = (c = div(ncells3, G); Int64[[g * c * plane for g = 0:G-1]; n])
    dist = G > 1
    for l = 1:levels-1
        AT, PT, RT = param.As[l], param.Ps[l], param.Rs[l]
        n, nc = size(AT, 2), size(param.As[l+1], 2)
        ro = cro = Ptr{Int64}(C_NULL); keep = nothing
        if dist && length(param.Meshes) > l && length(param.Meshes[l].n) == 3
            cf, cc = param.Meshes[l].n, param.Meshes[l+1].n
            if n >= replicateBelow && div(cf[3], G) >= 2 && div(cc[3], G) >= 1 && iseven(cf[3])
                keep = (offsets(cf[3], (cf[1] + 1) * (cf[2] + 1), n), offsets(cc[3], (cc[1] + 1) * (cc[2] + 1), nc))
                ro, cro = pointer(keep[1]), pointer(keep[2])
            end
        end
        keep === nothing && (dist = false)          # everything coarser is replicated too
        if length(param.Meshes) > l
            nf = Int64.(param.Meshes[l].n .+ 1); ncn = Int64.(param.Meshes[l+1].n .+ 1)
            check(ccall((:mgb200_multi_set_level_grid, libmgb200), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Ptr{Int64}),
                        h[], l, length(nf), nf, ncn))
        end
        d = convert(Vector{VAL}, param.relaxPrecs[l])
        GC.@preserve keep check(ccall((:mgb200_multi_upload_level, libmgb200), Cint,
                    (Ptr{Cvoid}, Cint, Int64, Int64, Ptr{Int64}, Ptr{Int64},
                     Ptr{Int64}, Ptr{Int64}, Ptr{VAL}, Ptr{Int64}, Ptr{Int64}, Ptr{real(VAL)},
                     Ptr{Int64}, Ptr{Int64}, Ptr{real(VAL)}, Ptr{VAL}, Cint),
                    h[], l, n, nc, ro, cro, AT.colptr, AT.rowval, AT.nzval, PT.colptr, PT.rowval, PT.nzval,
                    RT.colptr, RT.rowval, RT.nzval, d, 1))
    end
    AL = param.As[end]
    check(ccall((:mgb200_multi_upload_coarsest, libmgb200), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{VAL}, Cint),
                h[], size(AL, 2), AL.colptr, AL.rowval, AL.nzval, 1))
    dev = DeviceHierarchy(h[], true)
    finalizer(destroy!, dev)
    _devices[param] = dev
    return dev
end

# the handle of a hierarchy, uploaded on first use; mutable MGparam fields that steer the cycle are re-sent every call
function device(param::MGparam, b)
    dev = get(_devices, param, nothing)
    dev === nothing && (dev = uploadHierarchy(param))
    if !dev.multi
        levels = length(param.As)
        pre, post, _ = cycleArgs(param, levels)
        check(ccall((:mgb200_adjust_nrhs, libmgb200), Cint, (Ptr{Cvoid}, Cint), dev.handle, size(b, 2)))
        check(ccall((:mgb200_set_cycle, libmgb200), Cint, (Ptr{Cvoid}, Cchar, Ptr{Int64}, Ptr{Int64}),
                    dev.handle, Cchar(param.cycleType), pre, post))
    end
    return dev
end

# getAfun(AT, ...) is built from the AT of EVERY call (SolveFuncs.jl:65-82): upload it when it is not the hierarchy's own
# fine matrix, release a matrix left behind by an earlier call otherwise
function krylovMatrix(dev::DeviceHierarchy, AT::SparseMatrixCSC{VAL,Int64}, param::MGparam) where {VAL}
    dev.multi && (AT === param.As[1] || error("the multi-GPU handle multiplies with As[1]"); return)
    if AT === param.As[1]
        check(ccall((:mgb200_set_krylov_matrix, libmgb200), Cint,
                    (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Cvoid}, Cint), dev.handle, 0, C_NULL, C_NULL, C_NULL, 1))
    else
        check(ccall((:mgb200_set_krylov_matrix, libmgb200), Cint,
                    (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{VAL}, Cint),
                    dev.handle, size(AT, 2), AT.colptr, AT.rowval, AT.nzval, 1))
    end
end

# ---- src/Multigrid/SolveFuncs.jl:3-39 -----------------------------------------------------------
function solveMG(param::MGparam{VAL,Int64}, b::Array{VAL}, x::Array{VAL}, verbose::Bool) where {VAL}
    dev = device(param, b)
    iter = Ref{Cint}(0)
    resvec = zeros(param.maxOuterIter + 1)
    sym = dev.multi ? :mgb200_multi_solveMG : :mgb200_solveMG
    check(ccall((sym, libmgb200), Cint,
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
function solveCG_MG(AT::SparseMatrixCSC{VAL,Int64}, param::MGparam{VAL,Int64}, b::Array{VAL}, x0::Array{VAL},
                    verbose::Bool=false) where {VAL}
    dev = device(param, b)
    krylovMatrix(dev, AT, param)
    iter = Ref{Cint}(0); flag = Ref{Cint}(0)
    resvec = zeros(max(param.maxOuterIter, 1) * size(b, 2))
    sym = dev.multi ? :mgb200_multi_solveCG : :mgb200_solveCG
    check(ccall((sym, libmgb200), Cint,
                (Ptr{Cvoid}, Ptr{VAL}, Ptr{VAL}, Cdouble, Cint, Ref{Cint}, Ref{Cint}, Ptr{Cdouble}),
                dev.handle, b, x0, param.relativeTol, param.maxOuterIter, iter, flag, resvec))
    return x0, param, Int(iter[])
end

# ---- src/Multigrid/SolveFuncs.jl:85-99 ------------------------------------------------------------
function solveBiCGSTAB_MG(AT::SparseMatrixCSC{VAL,Int64}, param::MGparam{VAL,Int64}, b::Array{VAL}, x0::Array{VAL},
                          verbose::Bool=false) where {VAL}
    dev = device(param, b)
    dev.multi && error("solveBiCGSTAB_MG is not bound for the multi-GPU handle")
    krylovMatrix(dev, AT, param)
    iter = Ref{Cint}(0); flag = Ref{Cint}(0); nprec = Ref{Cint}(0)
    resvec = zeros((param.maxOuterIter + 1) * size(b, 2))
    check(ccall((:mgb200_solveBiCGSTAB, libmgb200), Cint,
                (Ptr{Cvoid}, Ptr{VAL}, Ptr{VAL}, Cdouble, Cint, Ref{Cint}, Ref{Cint}, Ptr{Cdouble}, Ref{Cint}),
                dev.handle, b, x0, param.relativeTol, param.maxOuterIter, iter, flag, resvec, nprec))
    return x0, param, Int(iter[]), Int(nprec[])
end

# ---- src/Multigrid/SolveFuncs.jl:120-132 ----------------------------------------------------------
function solveGMRES_MG(AT::SparseMatrixCSC{VAL,Int64}, param::MGparam{VAL,Int64}, b::Array{VAL}, x0::Array{VAL},
                       flexible::Bool, inner::Int64, verbose::Bool=false) where {VAL}
    dev = device(param, b)
    krylovMatrix(dev, AT, param)
    iter = Ref{Cint}(0); flag = Ref{Cint}(0); nres = Ref{Cint}(0)
    resvec = zeros(max(inner * param.maxOuterIter, 1))
    sym = dev.multi ? :mgb200_multi_solveFGMRES : :mgb200_solveFGMRES
    check(ccall((sym, libmgb200), Cint,
                (Ptr{Cvoid}, Ptr{VAL}, Ptr{VAL}, Cint, Cint, Cdouble, Cint, Ref{Cint}, Ref{Cint},
                 Ptr{Cdouble}, Ref{Cint}),
                dev.handle, b, x0, inner, flexible, param.relativeTol, param.maxOuterIter, iter, flag,
                resvec, nres))
    return x0, param, Int(iter[]), resvec[1:nres[]]
end

# ---- src/Multigrid/SolveFuncs.jl:43-63: r -> z (one cycle from z = 0) -----------------------------
function getMultigridPreconditioner(param::MGparam{VAL,Int64}, B::Array, verbose::Bool=false) where {VAL}
    hierarchyExists(param) || println("You have to do a setup first.")     # the reference prints here (SolveFuncs.jl:46-48)
    dev = device(param, B)
    z = zeros(VAL, size(B))
    sym = dev.multi ? :mgb200_multi_precondition : :mgb200_precondition
    return function (r::Array{VAL})
        # z .= 0 is a flag on the device, not a copy; like the reference the closure returns its own (aliased) buffer
        check(ccall((sym, libmgb200), Cint, (Ptr{Cvoid}, Ptr{VAL}, Ptr{VAL}), dev.handle, r, z))
        return z
    end
end

# ---- mixed precision (SolveFuncs.jl:52-60: VAL != eltype(B)) --------------------------------------
# A Float32 / ComplexF32 hierarchy under Float64 / ComplexF64 Krylov vectors: the outer handle holds only the Krylov
# matrix AT (double precision) and the Krylov vectors, its preconditioner is one cycle of the single-precision hierarchy
# on a rounded copy of the residual.  solveCG_MG(AT64, param32, b64, x64) dispatches on eltype(b), like the reference.
const _mixed = IdDict{Any,Tuple{DeviceHierarchy,Any}}()
function mixedDevice(param::MGparam{VAL,Int64}, AT::SparseMatrixCSC{VALD,Int64}, b) where {VAL<:Union{Float32,ComplexF32},VALD<:Union{Float64,ComplexF64}}
    inner = device(param, b)
    if haskey(_mixed, param) && _mixed[param][2] === AT
        outer = _mixed[param][1]
    else
        haskey(_mixed, param) && destroy!(_mixed[param][1])
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:mgb200_create_mixed, libmgb200), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cvoid}), h, inner.handle))
        check(ccall((:mgb200_set_krylov_matrix, libmgb200), Cint,
                    (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{VALD}, Cint),
                    h[], size(AT, 2), AT.colptr, AT.rowval, AT.nzval, 1))
        outer = DeviceHierarchy(h[], false)
        finalizer(destroy!, outer)
        _mixed[param] = (outer, AT)            # `inner` must outlive it: both hang on param
    end
    check(ccall((:mgb200_adjust_nrhs, libmgb200), Cint, (Ptr{Cvoid}, Cint), outer.handle, size(b, 2)))
    return outer
end
function solveCG_MG(AT::SparseMatrixCSC{VALD,Int64}, param::MGparam{VAL,Int64}, b::Array{VALD}, x0::Array{VALD},
                    verbose::Bool=false) where {VAL<:Union{Float32,ComplexF32},VALD<:Union{Float64,ComplexF64}}
    outer = mixedDevice(param, AT, b)
    iter = Ref{Cint}(0); flag = Ref{Cint}(0)
    resvec = zeros(max(param.maxOuterIter, 1) * size(b, 2))
    check(ccall((:mgb200_solveCG, libmgb200), Cint,
                (Ptr{Cvoid}, Ptr{VALD}, Ptr{VALD}, Cdouble, Cint, Ref{Cint}, Ref{Cint}, Ptr{Cdouble}),
                outer.handle, b, x0, param.relativeTol, param.maxOuterIter, iter, flag, resvec))
    return x0, param, Int(iter[])
end

# replaceMatrixInHierarchy (MGsetup.jl:226-270) with a resident device hierarchy: only the new As[1] travels; relaxation
# weights, Galerkin products and the coarsest factorisation are redone on the device (csrc/galerkin.cuh), then
# param.As[2:end] and param.relaxPrecs are refreshed from the device (values only: the sparsity is that of the first
# setup).  When the device path does not apply (another sparsity, row-partitioned hierarchy) the reference's host
# function runs and the device hierarchy is dropped for re-upload.
function replaceMatrixInHierarchy(param::MGparam{VAL,Int64}, AT::SparseMatrixCSC{VAL,Int64}, verbose::Bool=false) where {VAL}
    dev = get(_devices, param, nothing)
    if dev === nothing || dev.multi || !(param.relaxType in ("Jac", "Jac-GMRES", "SPAI"))
        clearDevice!(param)
        return Multigrid.replaceMatrixInHierarchy(param, AT, verbose)
    end
    omega = isa(param.relaxParam, Array) ? Float64.(real.(param.relaxParam)) : fill(Float64(real(param.relaxParam)), param.levels)
    done = Ref{Cint}(0)
    check(ccall((:mgb200_replace_matrix, libmgb200), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{VAL}, Cint, Cint, Ptr{Float64}, Ref{Cint}),
                dev.handle, size(AT, 2), AT.colptr, AT.rowval, AT.nzval, 1, param.relaxType == "SPAI" ? 1 : 0, omega, done))
    if done[] == 0
        clearDevice!(param)
        return Multigrid.replaceMatrixInHierarchy(param, AT, verbose)
    end
    param.As[1] = AT
    for l = 1:(param.levels - 1)
        check(ccall((:mgb200_download_relax_prec, libmgb200), Cint, (Ptr{Cvoid}, Cint, Ptr{VAL}), dev.handle, l, param.relaxPrecs[l]))
        check(ccall((:mgb200_download_values, libmgb200), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{VAL}, Int64),
                    dev.handle, l + 1, 0, param.As[l+1].nzval, nnz(param.As[l+1])))
    end
    param.doTranspose = 0
    return
end

end # module
