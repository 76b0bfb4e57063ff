# Package extension (loaded when CUDA.jl is): device arrays go to libdind as device pointers — zero-copy, the role
# `Adapt.adapt_structure` plays for the reference's structs (reference src/DataInterpolationsND.jl:58-64,
# src/interpolation_dimensions.jl:29-35, :78-85, :149-160).  NOTE: never executed where it was written (no Julia).
module DataInterpolationsNDB200CUDAExt

using DataInterpolationsNDB200
using CUDA
import Adapt

const DND = DataInterpolationsNDB200

DND.dind_pointer(x::CuArray) = reinterpret(Ptr{Cvoid}, pointer(x))
# device arrays are never converted on the way in: a CuArray{Float32} u needs Float32 knots / t_eval
DND.as_eltype(::Type{T}, x::CuArray{T}) where {T} = x
DND.as_eltype(::Type{T}, x::CuArray) where {T} = CuArray{T}(x)

# adapt_structure(to, x) rebuilds the structs around adapted arrays, like the reference; the new NDInterpolation
# owns a fresh handle that borrows the adapted (device) arrays.
Adapt.adapt_structure(to, d::DND.LinearInterpolationDimension) =
    DND.LinearInterpolationDimension(Adapt.adapt(to, d.t); t_eval = Adapt.adapt(to, d.t_eval))
Adapt.adapt_structure(to, d::DND.ConstantInterpolationDimension) =
    DND.ConstantInterpolationDimension(Adapt.adapt(to, d.t); left = d.left, t_eval = Adapt.adapt(to, d.t_eval))
Adapt.adapt_structure(to, d::DND.BSplineInterpolationDimension) =
    DND.BSplineInterpolationDimension(Adapt.adapt(to, d.t), d.degree; t_eval = Adapt.adapt(to, d.t_eval),
        max_derivative_order_eval = d.max_derivative_order_eval, multiplicities = d.multiplicities)
Adapt.adapt_structure(to, c::DND.NURBSWeights) = DND.NURBSWeights(Adapt.adapt(to, c.weights))
Adapt.adapt_structure(to, ::DND.EmptyCache) = DND.EmptyCache()
Adapt.adapt_structure(to, itp::DND.NDInterpolation) =
    DND.NDInterpolation(Adapt.adapt(to, itp.u), map(d -> Adapt.adapt_structure(to, d), itp.interp_dims);
        cache = Adapt.adapt_structure(to, itp.cache))

end # module
