# DDFB200.jl -- Julia glue over libddf_b200 (include/ddf_b200.h).
#
# UNEXECUTED: there is no Julia in the build image, so this file has never been run.
# It shows the binding a DDF.jl maintainer would add (the precedent is
# examples/wave-gpu.jl:22-34,121-126: swap the array type inside Fun/Op and overload
# mul!).  The same C symbols are exercised by the Python ctypes binding
# (ddf.jl_b200/_lib.py) in the GPU tests.
module DDFB200

using DDF
using Libdl
using LinearAlgebra
using SparseArrays

const lib = "libddf_b200"
# ccall needs a constant (symbol, library) pair; entry points chosen at run time go through dlsym
const libhandle = Ref{Ptr{Cvoid}}(C_NULL)
function fptr(sym::Symbol)
    libhandle[] == C_NULL && (libhandle[] = Libdl.dlopen(lib))
    return Libdl.dlsym(libhandle[], sym)
end

struct DDFB200Error <: Exception
    msg::String
end
last_error() = unsafe_string(ccall((:ddf_last_error, lib), Cstring, ()))
check(status::Cint) = status == 0 ? nothing : throw(DDFB200Error(last_error()))

# ---------------------------------------------------------------- handles
mutable struct Ctx
    h::Ptr{Cvoid}
    function Ctx(device::Integer=0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ddf_ctx_create, lib), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r))
        c = new(r[])
        finalizer(x -> ccall((:ddf_ctx_destroy, lib), Cint, (Ptr{Cvoid},), x.h), c)
        return c
    end
end

mutable struct DeviceManifold{D}
    h::Ptr{Cvoid}
    ctx::Ctx
    host::Manifold{D}          # the reference's own object: metadata, sizes
end

"""
Upload `simplices[D]`, `coords`, `weights` of a reference Manifold and assemble the
faces/boundaries on the device -- replaces the loop at src/Manifolds.jl:361-368.
"""
function DeviceManifold(ctx::Ctx, mfd::Manifold{D,C,S}) where {D,C,S}
    sD = get_simplices(mfd, D).op                      # SparseMatrixCSC{One,Int}
    coords = get_coords(mfd).vec                       # Vector{SVector{C,Float64}} == C x V column-major
    weights = get_weights(mfd).vec
    r = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve sD coords weights begin
        check(ccall((:ddf_mesh_create, lib), Cint,
                    (Ptr{Cvoid}, Cint, Cint, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64},
                     Ref{Ptr{Cvoid}}),
                    ctx.h, D, C, size(sD, 1), size(sD, 2), sD.colptr, sD.rowval,
                    reinterpret(Float64, coords), weights, r))
    end
    check(ccall((:ddf_mesh_assemble, lib), Cint, (Ptr{Cvoid},), r[]))
    m = DeviceManifold{D}(r[], ctx, mfd)
    finalizer(x -> ccall((:ddf_mesh_destroy, lib), Cint, (Ptr{Cvoid},), x.h), m)
    return m
end

"Bit-exact export so that Manifold._boundaries[R] can still be filled (src/Manifolds.jl:443-450)."
function get_boundaries(m::DeviceManifold, R::Int)
    n = Ref{Int64}(0); nf = Ref{Int64}(0)
    check(ccall((:ddf_mesh_nsimplices, lib), Cint, (Ptr{Cvoid}, Cint, Ref{Int64}), m.h, R, n))
    check(ccall((:ddf_mesh_nsimplices, lib), Cint, (Ptr{Cvoid}, Cint, Ref{Int64}), m.h, R - 1, nf))
    colptr = Vector{Int64}(undef, n[] + 1); rowval = Vector{Int64}(undef, (R + 1) * n[])
    nzval = Vector{Int8}(undef, (R + 1) * n[])
    check(ccall((:ddf_mesh_get_boundary_csc, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}, Ptr{Int64}, Ptr{Int8}),
                m.h, R, colptr, rowval, nzval))
    return SparseOp{R - 1,R}(SparseMatrixCSC{Int8,Int}(nf[], n[], colptr, rowval, nzval))
end

# ---------------------------------------------------------------- cochains
mutable struct DeviceVector <: AbstractVector{Float64}   # goes inside Fun.values (src/Funs.jl:18)
    h::Ptr{Cvoid}
    n::Int
end
function DeviceVector(ctx::Ctx, x::Vector{Float64})
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ddf_vec_create, lib), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}), ctx.h, length(x), r))
    v = own(DeviceVector(r[], length(x)))
    GC.@preserve x check(ccall((:ddf_vec_upload, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), v.h, x))
    return v
end
own(v::DeviceVector) = finalizer(y -> ccall((:ddf_vec_destroy, lib), Cint, (Ptr{Cvoid},), y.h), v)
Base.size(v::DeviceVector) = (v.n,)
function Base.Array(v::DeviceVector)
    x = Vector{Float64}(undef, v.n)
    check(ccall((:ddf_vec_download, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), v.h, x))
    return x
end
function LinearAlgebra.norm(v::DeviceVector, p::Real=2)
    out = Ref{Float64}(0)
    kind = p == Inf ? 0 : Int(p)
    check(ccall((:ddf_vec_norm, lib), Cint, (Ptr{Cvoid}, Cint, Ref{Float64}), v.h, kind, out))
    return out[]
end

# ---------------------------------------------------------------- operators
mutable struct DeviceOp <: AbstractMatrix{Float64}       # goes inside Op.values (src/Ops.jl:20)
    h::Ptr{Cvoid}
    m::DeviceManifold
end
function _op(sym::Symbol, m::DeviceManifold, P::PrimalDual, R::Int)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall(fptr(sym), Cint, (Ptr{Cvoid}, Cint, Cint, Ref{Ptr{Cvoid}}), m.h, P == Pr ? 1 : 0, R, r))
    return own(DeviceOp(r[], m))
end
# every handle the library hands out is released by a finalizer
own(op::DeviceOp) = finalizer(x -> ccall((:ddf_op_destroy, lib), Cint, (Ptr{Cvoid},), x.h), op)
function Base.size(A::DeviceOp)
    a = Ref{Int64}(0); b = Ref{Int64}(0)
    check(ccall((:ddf_op_shape, lib), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), A.h, a, b))
    return (Int(a[]), Int(b[]))
end

# same call surface as src/ManifoldOps.jl:17-146, dispatching on the device manifold
DDF.deriv(::Val{P}, ::Val{R}, m::DeviceManifold{D}) where {P,R,D} =
    Op{D,P,R + (P == Pr ? 1 : -1),P,R}(m.host, _op(:ddf_op_deriv, m, P, R))
DDF.boundary(::Val{P}, ::Val{R}, m::DeviceManifold{D}) where {P,R,D} =
    Op{D,P,R - (P == Pr ? 1 : -1),P,R}(m.host, _op(:ddf_op_boundary, m, P, R))
DDF.hodge(::Val{P}, ::Val{R}, m::DeviceManifold{D}) where {P,R,D} =
    Op{D,!P,R,P,R}(m.host, _op(:ddf_op_hodge, m, P, R))
DDF.invhodge(::Val{P}, ::Val{R}, m::DeviceManifold{D}) where {P,R,D} =
    Op{D,!P,R,P,R}(m.host, _op(:ddf_op_invhodge, m, P, R))
DDF.coderiv(::Val{P}, ::Val{R}, m::DeviceManifold{D}) where {P,R,D} =
    Op{D,P,R - (P == Pr ? 1 : -1),P,R}(m.host, _op(:ddf_op_coderiv, m, P, R))
DDF.laplace(::Val{P}, ::Val{R}, m::DeviceManifold{D}) where {P,R,D} =
    Op{D,P,R,P,R}(m.host, _op(:ddf_op_laplace, m, P, R))
DDF.isboundary(::Val{Pr}, ::Val{R}, m::DeviceManifold{D}) where {R,D} =
    Op{D,Pr,R,Pr,R}(m.host, _op(:ddf_op_isboundary, m, Pr, R))

# Op.values algebra used by src/Ops.jl:178-276 (dnz is the identity on DeviceOp: Ops.jl:40)
function Base.:*(A::DeviceOp, B::DeviceOp)                       # Ops.jl:228-232
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ddf_op_compose, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), A.h, B.h, r))
    return own(DeviceOp(r[], A.m))
end
function Base.adjoint(A::DeviceOp)                               # Ops.jl:258-260
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ddf_op_adjoint, lib), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), A.h, r))
    return own(DeviceOp(r[], A.m))
end
function Base.:*(a::Number, A::DeviceOp)                         # Ops.jl:198-200
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ddf_op_scale, lib), Cint, (Cdouble, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), Float64(a), A.h, r))
    return own(DeviceOp(r[], A.m))
end
function Base.:+(A::DeviceOp, B::DeviceOp)                       # Ops.jl:186-190
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ddf_op_add, lib), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Ref{Ptr{Cvoid}}), A.h, 1.0, B.h, r))
    return own(DeviceOp(r[], A.m))
end
Base.:-(A::DeviceOp, B::DeviceOp) = A + (-1.0) * B
function Base.:/(A::DeviceOp, a::Number)                         # Ops.jl:210-212: values ./ a, ONE rounding per entry
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ddf_op_div, lib), Cint, (Ptr{Cvoid}, Cdouble, Ref{Ptr{Cvoid}}), A.h, Float64(a), r))
    return own(DeviceOp(r[], A.m))
end

# dot(f, g) of src/ManifoldOps.jl:154-164 (Fun{D,Pr,D}: 1 ./ vol_D, Fun{D,Dl,0}: 1 ./ dualvol_0), one fused pass
function wdot(m::DeviceManifold, P::Bool, R::Int, f::DeviceVector, g::DeviceVector)
    out = Ref{Float64}(0.0)
    check(ccall((:ddf_vec_wdot, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Float64}),
                m.h, P ? 1 : 0, R, f.h, g.h, out))
    return out[]
end

# get_lookup(mfd, Ri, Rj).op (src/Manifolds.jl:455-498) as an index-only CSC found on the device
function lookup(m::DeviceManifold, Ri::Int, Rj::Int)
    nnz = Ref{Int64}(0)
    check(ccall((:ddf_mesh_get_lookup_csc, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ref{Int64}, Ptr{Int64}, Ptr{Int64}),
                m.h, Ri, Rj, nnz, C_NULL, C_NULL))
    nj = Ref{Int64}(0)
    check(ccall((:ddf_mesh_nsimplices, lib), Cint, (Ptr{Cvoid}, Cint, Ref{Int64}), m.h, Rj, nj))
    ni = Ref{Int64}(0)
    check(ccall((:ddf_mesh_nsimplices, lib), Cint, (Ptr{Cvoid}, Cint, Ref{Int64}), m.h, Ri, ni))
    colptr = Vector{Int64}(undef, nj[] + 1); rowval = Vector{Int64}(undef, nnz[])
    check(ccall((:ddf_mesh_get_lookup_csc, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ref{Int64}, Ptr{Int64}, Ptr{Int64}),
                m.h, Ri, Rj, nnz, colptr, rowval))
    return SparseMatrixCSC{Bool,Int}(ni[], nj[], colptr, rowval, fill(true, nnz[]))
end

# Op*Fun (src/Ops.jl:273-276) ends in `A.values * f.values.vec`:
function Base.:*(A::DeviceOp, x::DeviceVector)
    n, _ = size(A)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ddf_vec_create, lib), Cint, (Ptr{Cvoid}, Int64, Ref{Ptr{Cvoid}}), A.m.ctx.h, n, r))
    y = own(DeviceVector(r[], n))
    check(ccall((:ddf_op_apply, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), A.h, x.h, y.h))
    return y
end
# the ODE right-hand side of examples/wave.jl:100  f!(dU,U,p,t) = mul!(dU, F, U)
function LinearAlgebra.mul!(y::DeviceVector, A::DeviceOp, x::DeviceVector)
    check(ccall((:ddf_op_apply, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), A.h, x.h, y.h))
    return y
end

# `.values` of an assembled operator (src/Ops.jl:25: Op(...) runs dnz/chop): the device evaluates the
# expression with SparseArrays' spmatmul order and the chop after every product / sum
function SparseArrays.sparse(A::DeviceOp)
    nnz = Ref{Int64}(0)
    check(ccall((:ddf_op_to_csc, lib), Cint, (Ptr{Cvoid}, Ref{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                A.h, nnz, C_NULL, C_NULL, C_NULL))
    nr, nc = size(A)
    colptr = Vector{Int64}(undef, nc + 1); rowval = Vector{Int64}(undef, nnz[]); nzval = Vector{Float64}(undef, nnz[])
    check(ccall((:ddf_op_to_csc, lib), Cint, (Ptr{Cvoid}, Ref{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                A.h, nnz, colptr, rowval, nzval))
    return SparseMatrixCSC{Float64,Int}(nr, nc, colptr, rowval, nzval)
end

# The operator held the way the reference holds every Op -- its assembled, chopped sparse matrix (src/Ops.jl:25) -- but
# on the device: one row kernel per `A * f`, and the matrix a Krylov solver is handed in examples/poisson.jl:203
function assemble(A::DeviceOp)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ddf_op_assemble, lib), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), A.h, r))
    return own(DeviceOp(r[], A.m))
end

# storage of the rows of deriv(Pr,k): :bits32 / :bits64 (one word of index differences per row), :packed
# (prefix-packed rows) or :explicit (index table)
function deriv_format(m::DeviceManifold, k::Int)
    f = Ref{Cint}(0)
    check(ccall((:ddf_mesh_deriv_format, lib), Cint, (Ptr{Cvoid}, Cint, Ref{Cint}), m.h, k, f))
    return (:explicit, :packed, :bits32, :bits64)[f[] + 1]
end

# OrdinaryDiffEq's stage broadcasts on the state vector (`@.. tmp = uprev + dt*(a31*k1 + a32*k2)`,
# examples/wave.jl:134) in one pass
function lincomb!(out::DeviceVector, base::DeviceVector, dt::Float64, coefs::Vector{Float64}, ks::Vector{DeviceVector})
    hs = Ptr{Cvoid}[k.h for k in ks]
    GC.@preserve coefs hs check(ccall((:ddf_vec_lincomb, lib), Cint,
                                      (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Ptr{Float64}, Ptr{Ptr{Cvoid}}),
                                      out.h, base.h, dt, length(ks), coefs, hs))
    return out
end

# the whole fixed-step Tsit5 loop of examples/wave.jl:134 on the device (U = [u; udot])
function tsit5!(m::DeviceManifold, u::DeviceVector, udot::DeviceVector, dt::Float64, nsteps::Int)
    check(ccall((:ddf_wave_tsit5, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint),
                m.h, u.h, udot.h, dt, nsteps))
    return u, udot
end

# examples/poisson.jl:203-227: bicgstabl_iterator!(x, A', b'; Pl = Identity()) with IterativeSolvers' defaults
function bicgstabl!(x::DeviceVector, A::DeviceOp, b::DeviceVector; l::Int=2, reltol::Float64=sqrt(eps(Float64)),
                    abstol::Float64=0.0, max_mv_products::Int=size(A, 2))
    mv = Ref{Cint}(0); res = Ref{Float64}(0); conv = Ref{Cint}(0)
    check(ccall((:ddf_solve_bicgstabl, lib), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cdouble, Cdouble, Cint, Ref{Cint}, Ref{Float64}, Ref{Cint}),
                A.h, b.h, x.h, l, reltol, abstol, max_mv_products, mv, res, conv))
    return x, (mv_products=Int(mv[]), residual=res[], converged=conv[] != 0)
end

# fused Poisson relaxation sweep (north_star): u += theta (1-b).((rho - L u) ./ diag(L))
function jacobi!(m::DeviceManifold, u::DeviceVector, rho::DeviceVector, theta::Float64, nsweeps::Int)
    check(ccall((:ddf_poisson_jacobi, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint),
                m.h, u.h, rho.h, theta, nsweeps))
    return u
end

# fused step (north_star): v += dt (1-b).(L u); u += dt (1-b).v
function leapfrog!(m::DeviceManifold, u::DeviceVector, v::DeviceVector, dt::Float64, nsteps::Int)
    check(ccall((:ddf_wave_leapfrog, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint),
                m.h, u.h, v.h, dt, nsteps))
    return u, v
end

# ---------------------------------------------------------------------------------- one process per GPU
# The host process (MPI.jl / Distributed) broadcasts the 128-byte id of rank 0 itself.
function comm_unique_id()
    id = zeros(UInt8, 128)
    GC.@preserve id check(ccall((:ddf_comm_unique_id, lib), Cint, (Ptr{UInt8},), id))
    return id
end
comm_init!(ctx::Ptr{Cvoid}, nranks::Int, rank::Int, id::Vector{UInt8}) =
    GC.@preserve id check(ccall((:ddf_comm_init, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx, nranks, rank, id))

# A mesh cut by vertex ranges (this rank's closed star of its owned vertices, renumbered monotonically): owned local
# vertices are own_lo:own_hi-1 (0-based), peers[k] gets send_idx[send_ptr[k]+1:send_ptr[k+1]] and fills
# recv_idx[recv_ptr[k]+1:recv_ptr[k+1]] (0-based local ids, both sides in ascending global id).
function set_halo_lists!(m::DeviceManifold, own_lo::Int, own_hi::Int, peers::Vector{Int32}, send_ptr::Vector{Int64},
                         send_idx::Vector{Int64}, recv_ptr::Vector{Int64}, recv_idx::Vector{Int64})
    GC.@preserve peers send_ptr send_idx recv_ptr recv_idx check(ccall(
        (:ddf_mesh_set_halo_lists, lib), Cint,
        (Ptr{Cvoid}, Int64, Int64, Cint, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
        m.h, own_lo, own_hi, length(peers), peers, send_ptr, send_idx, recv_ptr, recv_idx))
    return m
end

# generator slab along the slowest axis: `plane` vertices per plane, ghost planes below / above the owned ones
set_halo!(m::DeviceManifold, plane::Int, ghost_lo::Int, ghost_hi::Int) =
    check(ccall((:ddf_mesh_set_halo, lib), Cint, (Ptr{Cvoid}, Int64, Int64, Int64), m.h, plane, ghost_lo, ghost_hi))

# the fused step on a partitioned mesh: owned rows + ghost exchange per step (peer-memory halo on slabs, NCCL lists otherwise)
function leapfrog_dist!(m::DeviceManifold, u::DeviceVector, v::DeviceVector, dt::Float64, nsteps::Int)
    check(ccall((:ddf_wave_leapfrog_dist, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint),
                m.h, u.h, v.h, dt, nsteps))
    return u, v
end

end # module
