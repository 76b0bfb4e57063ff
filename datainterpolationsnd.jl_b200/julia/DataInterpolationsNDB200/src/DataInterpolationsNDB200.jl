"""
    DataInterpolationsNDB200

Julia host side of the B200 evaluation engine: the same exported names, constructors and keyword
arguments as DataInterpolationsND.jl (reference `src/DataInterpolationsND.jl:102-104`), with the
batched evaluation path (`eval_unstructured[!]`, `eval_grid[!]`, `interp(out, t...)`) `ccall`ed into
`libdind.so` (C ABI in `include/dind.h`) instead of dispatched to a KernelAbstractions backend.

NOTE: this file could not be executed where it was written (no Julia in the build image); the same
C ABI is exercised by the Python binding in `datainterpolationsnd.jl_b200/__init__.py`, which is
what the parity tests run.  See INTEGRATION.md.
"""
module DataInterpolationsNDB200

using Libdl

export NDInterpolation, LinearInterpolationDimension, ConstantInterpolationDimension,
    BSplineInterpolationDimension, NURBSWeights, EmptyCache,
    eval_unstructured, eval_unstructured!, eval_grid, eval_grid!, eval_unstructured_gradient!, build_caches!,
    set_u!, shard_range, comm_unique_id, init_comm!, allgather_out!, eval_unstructured_allgather!

# a const String: `ccall((:sym, libdind), ...)` needs a constant library expression
const libdind = get(ENV, "LIBDIND", joinpath(@__DIR__, "..", "..", "..", "libdind.so"))

const DIND_F32, DIND_F64 = Cint(0), Cint(1)
const DIND_LINEAR, DIND_CONSTANT, DIND_BSPLINE = Cint(0), Cint(1), Cint(2)

dind_dtype(::Type{Float32}) = DIND_F32
dind_dtype(::Type{Float64}) = DIND_F64
dind_dtype(T) = throw(ArgumentError("libdind supports Float32 and Float64, got $T"))

# Every validation failure of the reference is an `@assert`; a negative status maps to the same.
function check(rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:dind_last_error, libdind), Cstring, ()))
    throw(AssertionError("[dind $rc] $msg"))
end

# Arrays are handed to the C ABI as raw pointers.  Host `Array`s are copied by the library; DEVICE arrays are
# BORROWED (zero-copy) — the analogue of `Adapt.adapt_structure` handing a device array to the reference's structs
# (reference src/DataInterpolationsND.jl:58-64, src/interpolation_dimensions.jl:29-35).  The CUDA.jl method
# `dind_pointer(x::CuArray) = reinterpret(Ptr{Cvoid}, pointer(x))` lives in ext/DataInterpolationsNDB200CUDAExt.jl
# (a package extension, so this package does not depend on CUDA.jl).
dind_pointer(x::Array) = Ptr{Cvoid}(pointer(x))
dind_pointer(x::AbstractArray) = throw(ArgumentError("libdind takes Array (host) or CuArray (device, with CUDA.jl loaded), got $(typeof(x))"))
"`x` in the element type `T` the handle computes in, without a copy when it already is (device arrays are never copied here)."
as_eltype(::Type{T}, x::AbstractArray{T}) where {T} = x
as_eltype(::Type{T}, x::Array) where {T} = convert(Array{T}, x)

abstract type AbstractInterpolationDimension end
abstract type AbstractInterpolationCache end
struct EmptyCache <: AbstractInterpolationCache end
struct NURBSWeights{W <: AbstractArray} <: AbstractInterpolationCache
    weights::W
end

# ---- dimensions: same fields as the reference structs (src/interpolation_dimensions.jl) ----------
# The reference fills `idx_eval` (and `basis_function_eval`) in the constructors (set_eval_idx!,
# src/interpolation_utils.jl:143-161; set_basis_function_eval!, src/spline_utils.jl:164-191) and its tests read them
# as FIELDS (test/test_interface.jl:53,59).  Here the hot path never needs them materialised (the kernels search and
# recompute in registers), so they are computed on the device on first access — `dim.idx_eval`,
# `dim.basis_function_eval` — through `getproperty`, and cached in the struct.
mutable struct LazyCaches
    idx_eval::Union{Nothing, Vector{Int}}
    basis_function_eval::Any
end
LazyCaches() = LazyCaches(nothing, nothing)

struct LinearInterpolationDimension{tT, teT} <: AbstractInterpolationDimension
    t::tT
    t_eval::teT
    _lazy::LazyCaches
end
LinearInterpolationDimension(t; t_eval = similar(t, 0)) = (validate_t(t); LinearInterpolationDimension(t, t_eval, LazyCaches()))

struct ConstantInterpolationDimension{tT, teT} <: AbstractInterpolationDimension
    t::tT
    left::Bool      # stored, never read by evaluation — exactly like the reference (dead field)
    t_eval::teT
    _lazy::LazyCaches
end
ConstantInterpolationDimension(t; left = true, t_eval = similar(t, 0)) =
    (validate_t(t); ConstantInterpolationDimension(t, left, t_eval, LazyCaches()))

struct BSplineInterpolationDimension{tT, teT, mT} <: AbstractInterpolationDimension
    t::tT
    knots_all::tT
    t_eval::teT
    degree::Int
    max_derivative_order_eval::Int
    multiplicities::mT
    _lazy::LazyCaches
end
function BSplineInterpolationDimension(
        t, degree; t_eval = similar(t, 0), max_derivative_order_eval::Integer = 0,
        multiplicities::Union{AbstractVector{<:Integer}, Nothing} = nothing
    )
    validate_t(t)
    if isnothing(multiplicities)                       # interpolation_dimensions.jl:168-173
        multiplicities = fill(1, length(t))
        multiplicities[[1, end]] .= degree + 1
    end
    @assert length(multiplicities) == length(t) "There must be the same amount of knots (points in t) as multiplicities."
    @assert all(m -> 1 ≤ m ≤ degree + 1, multiplicities) "All multiplicities must be between 1 and degree + 1."
    knots_all = similar(t, sum(multiplicities))        # expand_knot_vector_kernel (spline_utils.jl:1-20), host side
    k = 0
    for (ti, m) in zip(t, multiplicities), _ in 1:m
        knots_all[k += 1] = ti
    end
    return BSplineInterpolationDimension(t, knots_all, t_eval, Int(degree), Int(max_derivative_order_eval), multiplicities, LazyCaches())
end

function validate_t(t)                                  # interpolation_utils.jl:34-37
    @assert t isa AbstractVector{<:Number} "t must be an AbstractVector with number like elements."
    @assert all(>(0), diff(t)) "The elements of t must be sorted and unique."
end

Base.length(d::AbstractInterpolationDimension) = length(d.t)

# `dim.idx_eval` / `dim.basis_function_eval`: the reference's field names, computed on the device on first access by a
# private one-dimensional handle (dind_get_idx_eval / dind_get_basis_function_eval) and cached.
function Base.getproperty(d::AbstractInterpolationDimension, s::Symbol)
    if s === :idx_eval
        lazy = getfield(d, :_lazy)
        lazy.idx_eval === nothing && (lazy.idx_eval = compute_idx_eval(d))
        return lazy.idx_eval
    elseif s === :basis_function_eval && d isa BSplineInterpolationDimension
        lazy = getfield(d, :_lazy)
        lazy.basis_function_eval === nothing && (lazy.basis_function_eval = compute_basis_function_eval(d))
        return lazy.basis_function_eval
    end
    return getfield(d, s)
end
Base.propertynames(d::LinearInterpolationDimension) = (:t, :t_eval, :idx_eval)
Base.propertynames(d::ConstantInterpolationDimension) = (:t, :left, :t_eval, :idx_eval)
Base.propertynames(d::BSplineInterpolationDimension) =
    (:t, :knots_all, :t_eval, :idx_eval, :degree, :max_derivative_order_eval, :basis_function_eval, :multiplicities)

# a one-dimensional handle that only knows this dimension and its t_eval
function with_solo_handle(f, d::AbstractInterpolationDimension, ::Type{T}) where {T}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:dind_create, libdind), Cint, (Ptr{Ptr{Cvoid}}, Cint, Cint, Cint), h, 1, dind_dtype(T), 0))
    try
        set_dim_and_t_eval!(h[], 0, d, T)
        return f(h[])
    finally
        ccall((:dind_destroy, libdind), Cint, (Ptr{Cvoid},), h[])
    end
end
compute_type(d::AbstractInterpolationDimension) = promote_type(eltype(getfield(d, :t)), eltype(getfield(d, :t_eval))) == Float32 ? Float32 : Float64
function compute_idx_eval(d::AbstractInterpolationDimension)
    out = Vector{Int64}(undef, length(getfield(d, :t_eval)))
    isempty(out) && return out
    with_solo_handle(d, compute_type(d)) do h
        check(ccall((:dind_get_idx_eval, libdind), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}), h, 0, out))
    end
    return out
end
function compute_basis_function_eval(d::BSplineInterpolationDimension)
    T = compute_type(d)
    out = Array{T, 3}(undef, length(getfield(d, :t_eval)), getfield(d, :degree) + 1, getfield(d, :max_derivative_order_eval) + 1)
    isempty(out) && return out
    with_solo_handle(d, T) do h
        check(ccall((:dind_get_basis_function_eval, libdind), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), h, 0, out))
    end
    return out
end

# dind_set_dim + dind_set_t_eval of dimension `dim` at (0-based) position `d0` of handle `h`
function set_dim_and_t_eval!(h::Ptr{Cvoid}, d0::Integer, dim::AbstractInterpolationDimension, ::Type{T}) where {T}
    t = convert(Vector{T}, collect(getfield(dim, :t)))   # knots are a few hundred values: always staged through the host
    mult = dim isa BSplineInterpolationDimension ? convert(Vector{Int64}, getfield(dim, :multiplicities)) : Int64[]
    GC.@preserve t mult check(ccall((:dind_set_dim, libdind), Cint,
        (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Int64, Cint, Ptr{Int64}, Cint),
        h, d0, kind(dim), t, length(t), degree(dim),
        isempty(mult) ? C_NULL : pointer(mult), dim isa ConstantInterpolationDimension ? getfield(dim, :left) : true))
    te = as_eltype(T, getfield(dim, :t_eval))            # Array: copied by the library; CuArray: borrowed (zero-copy)
    GC.@preserve te check(ccall((:dind_set_t_eval, libdind), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64, Cint, Cuint),
        h, d0, isempty(te) ? C_NULL : dind_pointer(te), length(te), max_deriv(dim), 0))
    return te    # the caller keeps it alive while the handle borrows it
end
kind(::LinearInterpolationDimension) = DIND_LINEAR
kind(::ConstantInterpolationDimension) = DIND_CONSTANT
kind(::BSplineInterpolationDimension) = DIND_BSPLINE
degree(d::AbstractInterpolationDimension) = 0
degree(d::BSplineInterpolationDimension) = d.degree
max_deriv(d::AbstractInterpolationDimension) = 0
max_deriv(d::BSplineInterpolationDimension) = d.max_derivative_order_eval

# ---- the interpolation object -----------------------------------------------------------------------
mutable struct NDInterpolation{N_in, N_out, T, ID <: AbstractInterpolationDimension, gType, uType}
    u::uType
    interp_dims::NTuple{N_in, ID}
    cache::gType
    handle::Ptr{Cvoid}
    keep::Vector{Any}      # arrays the handle borrows (device t_eval / weights / u): alive as long as the handle
end

function NDInterpolation(u::AbstractArray{T}, interp_dims, cache::AbstractInterpolationCache; device::Integer = 0) where {T}
    interp_dims isa AbstractInterpolationDimension && (interp_dims = (interp_dims,))
    N_in = length(interp_dims)
    N_out = ndims(u) - N_in
    @assert N_out ≥ 0 "The number of dimensions of u must be at least the number of interpolation dimensions."
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:dind_create, libdind), Cint, (Ptr{Ptr{Cvoid}}, Cint, Cint, Cint), h, N_in, dind_dtype(T), device))
    itp = NDInterpolation{N_in, N_out, T, eltype(interp_dims), typeof(cache), typeof(u)}(u, Tuple(interp_dims), cache, h[], Any[])
    finalizer(x -> ccall((:dind_destroy, libdind), Cint, (Ptr{Cvoid},), x.handle), itp)
    for (d, dim) in enumerate(interp_dims)
        push!(itp.keep, set_dim_and_t_eval!(itp.handle, d - 1, dim, T))
    end
    if cache isa NURBSWeights
        w = as_eltype(T, cache.weights)
        dims = collect(Int64, size(w))
        GC.@preserve w dims check(ccall((:dind_set_weights, libdind), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Cint, Cuint), itp.handle, dind_pointer(w), dims, length(dims), 0))
        push!(itp.keep, w)
    end
    set_u!(itp, u)
    return itp
end
NDInterpolation(u, interp_dims; cache = EmptyCache(), kwargs...) = NDInterpolation(u, interp_dims, cache; kwargs...)

"""
    set_u!(itp, u)

Swap in a new `u` (re-evaluation with the cached t_eval indices/weights kept).  A host `Array` is copied; a device
array (`CuArray`) is borrowed: keep it unchanged until it is replaced — after changing it IN PLACE call `set_u!`
again (the library keeps derived copies: component-fastest `u`, `u .* weights`).
"""
function set_u!(itp::NDInterpolation{N_in, N_out, T}, u::AbstractArray{T}) where {N_in, N_out, T}
    dims = collect(Int64, size(u))
    GC.@preserve u dims check(ccall((:dind_set_u, libdind), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Cint, Cuint), itp.handle, dind_pointer(u), dims, length(dims), 0))
    itp.u = u
    return itp
end

get_output_size(itp::NDInterpolation{N_in}) where {N_in} = size(itp.u)[(N_in + 1):end]

orders_arg(derivative_orders::NTuple{N, <:Integer}) where {N} = Cint[derivative_orders...]

# ---- batched evaluation (src/interpolation_parallel.jl:14-103) -----------------------------------------
function eval_unstructured!(out::AbstractArray{T}, itp::NDInterpolation{N_in, N_out, T};
        derivative_orders::NTuple{N_in, <:Integer} = ntuple(_ -> 0, N_in)) where {N_in, N_out, T}
    @assert size(out, 1) == length(first(itp.interp_dims).t_eval) "The t_eval of all interpolation dimensions must have the same length as the first dimension of out."
    @assert size(out)[2:end] == get_output_size(itp) "The size of the last N_out dimensions of out must be the same as the output size of the interpolation."
    o = orders_arg(derivative_orders)
    GC.@preserve out o check(ccall((:dind_eval_unstructured, libdind), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cint}, Cuint), itp.handle, dind_pointer(out), o, 0))
    return out
end
function eval_unstructured(itp::NDInterpolation; kwargs...)
    out = similar(itp.u, (length(first(itp.interp_dims).t_eval), get_output_size(itp)...))
    return eval_unstructured!(out, itp; kwargs...)
end

function eval_grid!(out::AbstractArray{T}, itp::NDInterpolation{N_in, N_out, T};
        derivative_orders::NTuple{N_in, <:Integer} = ntuple(_ -> 0, N_in)) where {N_in, N_out, T}
    @assert all(i -> size(out, i) == length(itp.interp_dims[i].t_eval), 1:N_in) "For the first N_in dimensions of out the length must match the t_eval of the corresponding interpolation dimension."
    @assert size(out)[(N_in + 1):end] == get_output_size(itp) "The size of the last N_out dimensions of out must be the same as the output size of the interpolation."
    o = orders_arg(derivative_orders)
    GC.@preserve out o check(ccall((:dind_eval_grid, libdind), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cint}, Cuint), itp.handle, dind_pointer(out), o, 0))
    return out
end
function eval_grid(itp::NDInterpolation{N_in}; kwargs...) where {N_in}
    out = similar(itp.u, (map(d -> length(d.t_eval), itp.interp_dims)..., get_output_size(itp)...))
    return eval_grid!(out, itp; kwargs...)
end

# ---- beyond the reference: value + all first partial derivatives in one pass -----------------------------
"""
    eval_unstructured_gradient!(out, itp)

`out[k, .., 1]` = value, `out[k, .., 1 + d]` = ∂/∂t_d at the cached unstructured points; equivalent to
`N_in + 1` calls of `eval_unstructured!` with unit `derivative_orders`, in one kernel pass.
"""
function eval_unstructured_gradient!(out::AbstractArray{T}, itp::NDInterpolation{N_in, N_out, T}) where {N_in, N_out, T}
    @assert size(out) == (length(first(itp.interp_dims).t_eval), get_output_size(itp)..., N_in + 1)
    GC.@preserve out check(ccall((:dind_eval_unstructured_gradient, libdind), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cuint), itp.handle, dind_pointer(out), 0))
    return out
end

# ---- explicit cache construction (dind_build_caches): the constructors' set_eval_idx! step, timeable --------
function build_caches!(itp::NDInterpolation{N_in}; grid::Bool = false,
        derivative_orders::NTuple{N_in, <:Integer} = ntuple(_ -> 0, N_in), flags::Integer = 0) where {N_in}
    o = orders_arg(derivative_orders)
    GC.@preserve o check(ccall((:dind_build_caches, libdind), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cint}, Cuint), itp.handle, grid ? 1 : 0, o, flags))
    return itp
end

# ---- multi-GPU: one Julia process per GPU (MPI.jl / Distributed), u and knots replicated, points sharded -------------
"[lo, hi) of 0:n-1 owned by `rank` (0-based) of `world` — dind_shard_range; shards differ by at most one point."
function shard_range(n::Integer, rank::Integer, world::Integer)
    lo, hi = Ref{Int64}(0), Ref{Int64}(0)
    check(ccall((:dind_shard_range, libdind), Cint, (Int64, Cint, Cint, Ptr{Int64}, Ptr{Int64}), n, rank, world, lo, hi))
    return lo[], hi[]
end
"128 bytes made on rank 0 and sent to every rank by the host's own means (e.g. `MPI.Bcast!`)."
function comm_unique_id()
    id = Vector{UInt8}(undef, 128)
    check(ccall((:dind_comm_get_unique_id, libdind), Cint, (Ptr{UInt8},), id))
    return id
end
"Collective: give `itp`'s handle an NCCL communicator over `world` processes (one GPU each)."
function init_comm!(itp::NDInterpolation, unique_id::Vector{UInt8}, rank::Integer, world::Integer)
    check(ccall((:dind_comm_init, libdind), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Cint, Cint), itp.handle, unique_id, rank, world))
    return itp
end
"""
    allgather_out!(out_global, itp, out_local; n_total, unit = 1)

All-gather the outputs of a sharded evaluation (device arrays): `out_local` is this rank's `(unit, n_local, out...)`
block, `out_global` receives `(unit, n_total, out...)`.  `eval_unstructured!` shards: `unit = 1`;
`eval_grid!` slabs along the last interpolation axis: `unit = prod(size(out)[1:N_in-1])`.
"""
function allgather_out!(out_global::AbstractArray{T}, itp::NDInterpolation{N_in, N_out, T}, out_local::AbstractArray{T};
        n_total::Integer, unit::Integer = 1) where {N_in, N_out, T}
    n_comp = length(out_global) ÷ (n_total * unit)
    GC.@preserve out_global out_local check(ccall((:dind_allgather_out, libdind), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, Cuint),
        itp.handle, dind_pointer(out_local), dind_pointer(out_global), n_total, unit, n_comp, 0))
    return out_global
end

"""
    eval_unstructured_allgather!(out_global, itp; n_total, derivative_orders, n_chunks = 0)

Collective form of `eval_unstructured!` for a point-sharded evaluation: `itp` holds this rank's shard of the points
(`shard_range`), the shard is evaluated straight into its rows of the device array `out_global` `(n_total, out...)` and
the other ranks' rows arrive over NCCL while the next chunk of the shard is computed (dind_eval_unstructured_allgather).
"""
function eval_unstructured_allgather!(out_global::AbstractArray{T}, itp::NDInterpolation{N_in, N_out, T};
        n_total::Integer, derivative_orders::NTuple{N_in, <:Integer} = ntuple(_ -> 0, N_in), n_chunks::Integer = 0) where {N_in, N_out, T}
    @assert size(out_global) == (n_total, get_output_size(itp)...) "out_global must be (n_total, out...)"
    o = orders_arg(derivative_orders)
    GC.@preserve out_global o check(ccall((:dind_eval_unstructured_allgather, libdind), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cint}, Cint, Cuint), itp.handle, dind_pointer(out_global), n_total, o, n_chunks, 0))
    return out_global
end

# ---- single point calls (src/DataInterpolationsND.jl:73-100) through dind_eval_points with n = 1 --------
function (itp::NDInterpolation{N_in, N_out, T})(out::AbstractArray{T}, t::Tuple{Vararg{Number, N_in}};
        derivative_orders::NTuple{N_in, <:Integer} = ntuple(_ -> 0, N_in)) where {N_in, N_out, T}
    @assert size(out) == get_output_size(itp) "The size of out must match the size of the last N_out dimensions of u."
    pts = [T[ti] for ti in t]
    ptrs = Ptr{Cvoid}[pointer(p) for p in pts]
    o = orders_arg(derivative_orders)
    GC.@preserve out pts ptrs o check(ccall((:dind_eval_points, libdind), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Int64, Ptr{Cint}, Cuint), itp.handle, out, ptrs, 1, o, 0))
    return out
end
(itp::NDInterpolation)(out::AbstractArray, t_args::Vararg{Number}; kwargs...) = itp(out, t_args; kwargs...)
function (itp::NDInterpolation{N_in, N_out, T})(t::Tuple{Vararg{Number, N_in}}; kwargs...) where {N_in, N_out, T}
    out = zeros(T, get_output_size(itp))
    itp(out, t; kwargs...)
    return N_out == 0 ? out[] : out
end
(itp::NDInterpolation)(t_args::Vararg{Number}; kwargs...) = itp(t_args; kwargs...)

end # module
