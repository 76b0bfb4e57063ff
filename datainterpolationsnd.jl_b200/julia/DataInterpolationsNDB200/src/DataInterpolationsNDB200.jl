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
    eval_unstructured, eval_unstructured!, eval_grid, eval_grid!, eval_unstructured_gradient!, build_caches!

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

abstract type AbstractInterpolationDimension end
abstract type AbstractInterpolationCache end
struct EmptyCache <: AbstractInterpolationCache end
struct NURBSWeights{W <: AbstractArray} <: AbstractInterpolationCache
    weights::W
end

# ---- dimensions: same fields as the reference structs (src/interpolation_dimensions.jl) ----------
struct LinearInterpolationDimension{tT, teT} <: AbstractInterpolationDimension
    t::tT
    t_eval::teT
end
LinearInterpolationDimension(t; t_eval = similar(t, 0)) = (validate_t(t); LinearInterpolationDimension(t, t_eval))

struct ConstantInterpolationDimension{tT, teT} <: AbstractInterpolationDimension
    t::tT
    left::Bool      # stored, never read by evaluation — exactly like the reference (dead field)
    t_eval::teT
end
ConstantInterpolationDimension(t; left = true, t_eval = similar(t, 0)) =
    (validate_t(t); ConstantInterpolationDimension(t, left, t_eval))

struct BSplineInterpolationDimension{tT, teT, mT} <: AbstractInterpolationDimension
    t::tT
    knots_all::tT
    t_eval::teT
    degree::Int
    max_derivative_order_eval::Int
    multiplicities::mT
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
    return BSplineInterpolationDimension(t, knots_all, t_eval, Int(degree), Int(max_derivative_order_eval), multiplicities)
end

function validate_t(t)                                  # interpolation_utils.jl:34-37
    @assert t isa AbstractVector{<:Number} "t must be an AbstractVector with number like elements."
    @assert all(>(0), diff(t)) "The elements of t must be sorted and unique."
end

Base.length(d::AbstractInterpolationDimension) = length(d.t)
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
end

function NDInterpolation(u::AbstractArray{T}, interp_dims, cache::AbstractInterpolationCache; device::Integer = 0) where {T}
    interp_dims isa AbstractInterpolationDimension && (interp_dims = (interp_dims,))
    N_in = length(interp_dims)
    N_out = ndims(u) - N_in
    @assert N_out ≥ 0 "The number of dimensions of u must be at least the number of interpolation dimensions."
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:dind_create, libdind), Cint, (Ptr{Ptr{Cvoid}}, Cint, Cint, Cint), h, N_in, dind_dtype(T), device))
    itp = NDInterpolation{N_in, N_out, T, eltype(interp_dims), typeof(cache), typeof(u)}(u, Tuple(interp_dims), cache, h[])
    finalizer(x -> ccall((:dind_destroy, libdind), Cint, (Ptr{Cvoid},), x.handle), itp)
    for (d, dim) in enumerate(interp_dims)
        t = convert(Vector{T}, dim.t)
        mult = dim isa BSplineInterpolationDimension ? convert(Vector{Int64}, dim.multiplicities) : Int64[]
        GC.@preserve t mult check(ccall((:dind_set_dim, libdind), Cint,
            (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Int64, Cint, Ptr{Int64}, Cint),
            itp.handle, d - 1, kind(dim), t, length(t), degree(dim),
            isempty(mult) ? C_NULL : pointer(mult), dim isa ConstantInterpolationDimension ? dim.left : true))
        te = convert(Vector{T}, dim.t_eval)      # a CuArray would be passed as a device pointer instead (zero-copy)
        GC.@preserve te check(ccall((:dind_set_t_eval, libdind), Cint,
            (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64, Cint, Cuint),
            itp.handle, d - 1, isempty(te) ? C_NULL : pointer(te), length(te), max_deriv(dim), 0))
    end
    if cache isa NURBSWeights
        w = convert(Array{T}, cache.weights)
        dims = collect(Int64, size(w))
        GC.@preserve w dims check(ccall((:dind_set_weights, libdind), Cint,
            (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Cint, Cuint), itp.handle, w, dims, length(dims), 0))
    end
    set_u!(itp, u)
    return itp
end
NDInterpolation(u, interp_dims; cache = EmptyCache(), kwargs...) = NDInterpolation(u, interp_dims, cache; kwargs...)

"Swap in a new `u` (re-evaluation with the cached t_eval indices/weights kept)."
function set_u!(itp::NDInterpolation{N_in, N_out, T}, u::AbstractArray{T}) where {N_in, N_out, T}
    dims = collect(Int64, size(u))
    GC.@preserve u dims check(ccall((:dind_set_u, libdind), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int64}, Cint, Cuint), itp.handle, u, dims, length(dims), 0))
    itp.u = u
    return itp
end

get_output_size(itp::NDInterpolation{N_in}) where {N_in} = size(itp.u)[(N_in + 1):end]

# idx_eval / basis_function_eval are computed on the device on demand (fields of the reference structs)
function idx_eval(itp::NDInterpolation, d::Integer)
    out = Vector{Int64}(undef, length(itp.interp_dims[d].t_eval))
    check(ccall((:dind_get_idx_eval, libdind), Cint, (Ptr{Cvoid}, Cint, Ptr{Int64}), itp.handle, d - 1, out))
    return out
end
function basis_function_eval(itp::NDInterpolation{N_in, N_out, T}, d::Integer) where {N_in, N_out, T}
    dim = itp.interp_dims[d]
    out = Array{T, 3}(undef, length(dim.t_eval), dim.degree + 1, dim.max_derivative_order_eval + 1)
    check(ccall((:dind_get_basis_function_eval, libdind), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), itp.handle, d - 1, out))
    return out
end

orders_arg(derivative_orders::NTuple{N, <:Integer}) where {N} = Cint[derivative_orders...]

# ---- batched evaluation (src/interpolation_parallel.jl:14-103) -----------------------------------------
function eval_unstructured!(out::AbstractArray{T}, itp::NDInterpolation{N_in, N_out, T};
        derivative_orders::NTuple{N_in, <:Integer} = ntuple(_ -> 0, N_in)) where {N_in, N_out, T}
    @assert size(out, 1) == length(first(itp.interp_dims).t_eval) "The t_eval of all interpolation dimensions must have the same length as the first dimension of out."
    @assert size(out)[2:end] == get_output_size(itp) "The size of the last N_out dimensions of out must be the same as the output size of the interpolation."
    o = orders_arg(derivative_orders)
    GC.@preserve out o check(ccall((:dind_eval_unstructured, libdind), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cint}, Cuint), itp.handle, out, o, 0))
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
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cint}, Cuint), itp.handle, out, o, 0))
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
        (Ptr{Cvoid}, Ptr{Cvoid}, Cuint), itp.handle, out, 0))
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
