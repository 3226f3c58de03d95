# AGPBlackwell.jl -- the thin Julia shim a maintainer adds on the reference side so that AbstractGPs.jl's
# dense hot path runs on libagp.so (hand-written sm_100a CUDA behind the C ABI of include/agp.h).
#
# Julia is NOT available in the build image, so this file is shipped as source and has never been
# executed (INTEGRATION.md lists what a maintainer should check first); every `ccall` below is mirrored 1:1 by the ctypes binding
# abstractgps.jl_b200/_cabi.py, which IS exercised by the test-suite -- the ABI is what is tested.
#
# Seams used (SURVEY.md s1, s8b): ordinary multiple dispatch on
#   FiniteGP{<:GP{<:Union{ZeroMean,ConstMean,CustomMean},<:SupportedKernel}, <:Inputs, <:Diagonal}
# for logpdf / posterior / rand / elbo, and a device-resident factor type `DeviceCholesky` that slots
# into PosteriorGP.data.C with methods for the operator set of src/util/common_covmat_ops.jl.
# Anything else (other kernels, dense Sigma_y, ...) falls through to the stock reference methods.
module AGPBlackwell

using AbstractGPs, KernelFunctions, LinearAlgebra, FillArrays
import AbstractGPs: posterior, mean_and_var, elbo, approx_log_evidence, FiniteGP, PosteriorGP, VFE, DTC, GP
import AbstractGPs: Xt_invA_X, Xt_invA_Y, diag_Xt_invA_X, tr_Xt_invA_X
import Distributions: logpdf
import Random

const libagp = get(ENV, "AGP_LIB", "libagp.so")

# ---- POD structs of include/agp.h ------------------------------------------------------------
struct AgpKernel; family::Int32; transform::Int32; variance::Float64; scale::Float64; linear_c::Float64; ard::Ptr{Cvoid}; end
struct AgpMean;   kind::Int32; c::Float64; v::Ptr{Cvoid}; end
struct AgpNoise;  kind::Int32; s::Float64; v::Ptr{Cvoid}; end
const AGP_F32, AGP_F64 = Int32(0), Int32(1)
const AGP_POINT_MAJOR, AGP_FEATURE_MAJOR = Int32(0), Int32(1)
agp_dtype(::Type{Float32}) = AGP_F32
agp_dtype(::Type{Float64}) = AGP_F64

mutable struct Ctx
    h::Ptr{Cvoid}
    lock::ReentrantLock          # a ctx is not re-entrant (agp.h "Conventions")
end
const CTX = Ref{Union{Nothing,Ctx}}(nothing)
function ctx()
    if CTX[] === nothing
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:agp_init, libagp), Int32, (Ptr{Ptr{Cvoid}}, Int32, Ptr{Cvoid}), h, 0, C_NULL)
        rc == 0 || error("agp_init failed ($rc): no CUDA device? (there is no CPU fallback)")
        CTX[] = Ctx(h[], ReentrantLock())
    end
    return CTX[]
end

function check(c::Ctx, rc::Int32)
    rc == 0 && return
    msg = unsafe_string(ccall((:agp_last_error, libagp), Cstring, (Ptr{Cvoid},), c.h))
    rc == 1 && throw(PosDefException(ccall((:agp_last_info, libagp), Int64, (Ptr{Cvoid},), c.h)))  # cholesky(.) behaviour
    rc == 2 && throw(DimensionMismatch(msg))
    error("libagp status $rc: $msg")
end

# ---- kernel / mean / noise translation ----------------------------------------------------------
const Stationary = Union{SqExponentialKernel,Matern12Kernel,Matern32Kernel,Matern52Kernel}
family(::SqExponentialKernel) = Int32(0); family(::Matern12Kernel) = Int32(1)
family(::Matern32Kernel) = Int32(2);      family(::Matern52Kernel) = Int32(3); family(::LinearKernel) = Int32(4)

# Kernels the engine implements: a base kernel wrapped in any nesting of ScaledKernel / TransformedKernel{Scale|ARD}.
# Everything else (sums, products, periodic, ...) is NOT claimed: the methods below `invoke` the stock reference method.
supported(::Union{Stationary,LinearKernel}) = true
supported(k::ScaledKernel) = supported(k.kernel)
supported(k::TransformedKernel{<:Any,<:Union{ScaleTransform,ARDTransform}}) = supported(k.kernel)
supported(::Kernel) = false
supported(::Union{AbstractGPs.ZeroMean,AbstractGPs.ConstMean,AbstractGPs.CustomMean}) = true
supported(::AbstractGPs.MeanFunction) = false
supported(f::GP) = supported(f.kernel) && supported(f.mean)
supported(::Any) = false

# flattened description: (family, sigma_f^2, linear c, per-dimension input scaling w) with w === nothing (identity),
# a scalar (ScaleTransform) or a vector (ARDTransform).  (k o t1) o t2 evaluates k(t1(t2(x))): diagonal scalings commute,
# so nested transforms multiply (the Python mirror's `_chain`).
flat(k::Union{Stationary,LinearKernel}) = (family(k), 1.0, k isa LinearKernel ? Float64(only(k.c)) : 0.0, nothing)
function flat(k::ScaledKernel)
    fam, var, c, w = flat(k.kernel)
    (fam, var * Float64(only(k.σ²)), c * 1.0, w)      # sigma^2 * (x'y + c) keeps c inside: the engine applies variance to both
end
combine(::Nothing, t) = t
combine(w, t) = w .* t
function flat(k::TransformedKernel{<:Any,<:ScaleTransform})
    fam, var, c, w = flat(k.kernel)
    (fam, var, c, combine(w, Float64(only(k.transform.s))))
end
function flat(k::TransformedKernel{<:Any,<:ARDTransform})
    fam, var, c, w = flat(k.kernel)
    (fam, var, c, combine(w, Float64.(k.transform.v)))
end
# returns (AgpKernel, keepalive)
function kernel_spec(k, T)
    fam, var, c, w = flat(k)
    w === nothing && return (AgpKernel(fam, 0, var, 1.0, c, C_NULL), nothing)
    w isa Real && return (AgpKernel(fam, 1, var, Float64(w), c, C_NULL), nothing)
    v = convert(Vector{T}, w)
    return (AgpKernel(fam, 2, var, 1.0, c, pointer(v)), v)
end
mean_spec(::AbstractGPs.ZeroMean, x, T) = (AgpMean(0, 0.0, C_NULL), nothing)
mean_spec(m::AbstractGPs.ConstMean, x, T) = (AgpMean(1, Float64(m.c), C_NULL), nothing)
function mean_spec(m::AbstractGPs.CustomMean, x, T)      # arbitrary closure: evaluated host-side
    v = convert(Vector{T}, AbstractGPs.mean_vector(m, x)); (AgpMean(2, 0.0, pointer(v)), v)
end
noise_spec(Σ::Diagonal{<:Any,<:Fill}, T) = (AgpNoise(0, Float64(Σ.diag.value), C_NULL), nothing)
function noise_spec(Σ::Diagonal, T); v = convert(Vector{T}, Σ.diag); (AgpNoise(1, 0.0, pointer(v)), v); end

points(x::ColVecs{T}) where {T} = (x.X, AGP_POINT_MAJOR, size(x.X, 1))
points(x::RowVecs{T}) where {T} = (x.X, AGP_FEATURE_MAJOR, size(x.X, 2))
points(x::AbstractVector{T}) where {T<:Real} = (x, AGP_POINT_MAJOR, 1)
# a second input collection (inducing / test points) in the SAME storage order as the first: the ABI takes one layout flag
same_layout(z::ColVecs, layout) = layout == AGP_POINT_MAJOR ? z.X : permutedims(z.X)
same_layout(z::RowVecs, layout) = layout == AGP_FEATURE_MAJOR ? z.X : permutedims(z.X)
same_layout(z::AbstractVector{<:Real}, layout) = z

# ---- device factor (boundary #2) --------------------------------------------------------------
mutable struct DeviceCholesky{T}
    h::Ptr{Cvoid}
    n::Int
    function DeviceCholesky{T}(h, n) where {T}
        C = new{T}(h, n)
        finalizer(c -> ccall((:agp_post_free, libagp), Int32, (Ptr{Cvoid},), c.h), C)
    end
end
function Base.getproperty(C::DeviceCholesky{T}, s::Symbol) where {T}
    s === :U || return getfield(C, s)
    U = Matrix{T}(undef, C.n, C.n)
    check(ctx(), ccall((:agp_post_factor_export, libagp), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), C.h, U))
    return UpperTriangular(U)
end
function solve_lower(C::DeviceCholesky{T}, B::AbstractVecOrMat) where {T}      # C.U' \ B
    Bm = convert(Matrix{T}, reshape(B, C.n, :)); V = similar(Bm)
    check(ctx(), ccall((:agp_post_solve_lower, libagp), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}), C.h, Bm, size(Bm, 2), V))
    return B isa AbstractVector ? vec(V) : V
end
Xt_invA_X(A::DeviceCholesky, x::AbstractVector) = sum(abs2, solve_lower(A, x))
Xt_invA_X(A::DeviceCholesky, X::AbstractMatrix) = (V = solve_lower(A, X); Symmetric(V'V))
Xt_invA_Y(X::AbstractVecOrMat, A::DeviceCholesky, Y::AbstractVecOrMat) = solve_lower(A, X)' * solve_lower(A, Y)
diag_Xt_invA_X(A::DeviceCholesky, X::AbstractVecOrMat) = AbstractGPs.diag_At_A(solve_lower(A, X))
tr_Xt_invA_X(A::DeviceCholesky, X::AbstractVecOrMat) = sum(abs2, solve_lower(A, X))

# ---- fused fit: logpdf + posterior from ONE Gram and ONE factorisation ---------------------------
const DevInputs{T} = Union{Vector{T},ColVecs{T},RowVecs{T}}
const DevFiniteGP{T} = FiniteGP{<:GP,<:DevInputs{T},<:Diagonal}

function fit(fx::DevFiniteGP{T}, Y::AbstractVecOrMat; want_post::Bool=true) where {T<:Union{Float32,Float64}}
    c = ctx()
    X, layout, D = points(fx.x)
    Ym = convert(Matrix{T}, reshape(Y, length(fx), :))
    ks, k1 = kernel_spec(fx.f.kernel, T); ms, k2 = mean_spec(fx.f.mean, fx.x, T); ns, k3 = noise_spec(fx.Σy, T)
    lp = Vector{T}(undef, size(Ym, 2)); α = Vector{T}(undef, length(fx)); post = Ref{Ptr{Cvoid}}(C_NULL)
    lock(c.lock) do
        GC.@preserve X Ym k1 k2 k3 begin
            check(c, ccall((:agp_fit, libagp), Int32,
                (Ptr{Cvoid}, Int32, Ref{AgpKernel}, Ref{AgpMean}, Ref{AgpNoise}, Int32, Ptr{Cvoid}, Int64, Int32,
                 Ptr{Cvoid}, Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}),
                c.h, agp_dtype(T), ks, ms, ns, layout, X, length(fx), D, Ym, size(Ym, 2), lp,
                want_post ? pointer(α) : C_NULL, want_post ? post : C_NULL))
        end
    end
    lpv = Y isa AbstractVector ? lp[1] : lp
    want_post || return lpv, nothing
    δ = Ym[:, 1] - AbstractGPs.mean(fx)
    return lpv, PosteriorGP(fx.f, (α=α, C=DeviceCholesky{T}(post[], length(fx)), x=fx.x, δ=δ))
end

# replaces src/finite_gp_projection.jl:306-311 and src/exact_gpr_posterior.jl:29-35 for the device types
# priors the engine does not implement fall through to the reference's own methods (Julia dispatch, not a CPU fallback
# inside the engine)
logpdf(fx::DevFiniteGP{T}, Y::AbstractVecOrMat{<:Real}) where {T} =
    supported(fx.f) ? fit(fx, Y; want_post=false)[1] : invoke(logpdf, Tuple{FiniteGP,typeof(Y)}, fx, Y)
posterior(fx::DevFiniteGP{T}, y::AbstractVector{<:Real}) where {T} =
    supported(fx.f) ? fit(fx, y)[2] : invoke(posterior, Tuple{FiniteGP,AbstractVector{<:Real}}, fx, y)

# replaces src/exact_gpr_posterior.jl:85-90 (+ src/finite_gp_projection.jl:154-158) with the fused cross-Gram path
const DevPosterior = PosteriorGP{<:GP,<:NamedTuple{(:α, :C, :x, :δ),<:Tuple{Any,DeviceCholesky,Any,Any}}}
function mean_and_var(fx::FiniteGP{<:DevPosterior,<:DevInputs{T},<:Diagonal}) where {T}
    p = fx.f; c = ctx(); Xs, layout, D = points(fx.x); M = length(fx)
    ms, k2 = mean_spec(p.prior.mean, fx.x, T); ns, k3 = noise_spec(fx.Σy, T)
    μ = Vector{T}(undef, M); v = Vector{T}(undef, M)
    lock(c.lock) do
        GC.@preserve Xs k2 k3 check(c, ccall((:agp_post_mean_var, libagp), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int64, Ref{AgpMean}, Ref{AgpNoise}, Ptr{Cvoid}, Ptr{Cvoid}),
            p.data.C.h, layout, Xs, M, ms, ns, μ, v))
    end
    return μ, v
end

# FiniteGP over a device posterior: logpdf / rand / sequential conditioning stay on the device
# (src/finite_gp_projection.jl:306-318, :233-237 and src/exact_gpr_posterior.jl:46-56 for f = PosteriorGP)
const DevPostFiniteGP{T} = FiniteGP{<:DevPosterior,<:DevInputs{T},<:Diagonal}
function logpdf(fx::DevPostFiniteGP{T}, Y::AbstractVecOrMat{<:Real}) where {T}
    p = fx.f; c = ctx(); Xs, layout, D = points(fx.x); M = length(fx)
    Ym = convert(Matrix{T}, reshape(Y, M, :)); lp = Vector{T}(undef, size(Ym, 2))
    ms, k2 = mean_spec(p.prior.mean, fx.x, T); ns, k3 = noise_spec(fx.Σy, T)
    lock(c.lock) do
        GC.@preserve Xs Ym k2 k3 check(c, ccall((:agp_post_logpdf, libagp), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int64, Ref{AgpMean}, Ref{AgpNoise}, Ptr{Cvoid}, Int32, Ptr{Cvoid}),
            p.data.C.h, layout, Xs, M, ms, ns, Ym, size(Ym, 2), lp))
    end
    return Y isa AbstractVector ? lp[1] : lp
end
function Random.rand(rng::Random.AbstractRNG, fx::DevPostFiniteGP{T}, S::Int) where {T}
    p = fx.f; c = ctx(); Xs, layout, D = points(fx.x); M = length(fx)
    Z = randn(rng, T, M, S); out = similar(Z)
    ms, k2 = mean_spec(p.prior.mean, fx.x, T); ns, k3 = noise_spec(fx.Σy, T)
    lock(c.lock) do
        GC.@preserve Xs k2 k3 check(c, ccall((:agp_post_rand, libagp), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int64, Ref{AgpMean}, Ref{AgpNoise}, Ptr{Cvoid}, Int32, Ptr{Cvoid}),
            p.data.C.h, layout, Xs, M, ms, ns, Z, S, out))
    end
    return out
end
function posterior(fx::DevPostFiniteGP{T}, y::AbstractVector{<:Real}) where {T}
    p = fx.f; c = ctx(); X2, layout, D = points(fx.x); N2 = length(fx); N1 = p.data.C.n
    yv = convert(Vector{T}, y); α = Vector{T}(undef, N1 + N2); post = Ref{Ptr{Cvoid}}(C_NULL)
    ms, k2 = mean_spec(p.prior.mean, fx.x, T); ns, k3 = noise_spec(fx.Σy, T)
    lock(c.lock) do
        GC.@preserve X2 yv k2 k3 check(c, ccall((:agp_post_extend, libagp), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ref{AgpMean}, Ref{AgpNoise}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}),
            p.data.C.h, layout, X2, N2, yv, ms, ns, α, post))
    end
    δ = vcat(p.data.δ, yv - AbstractGPs.mean_vector(p.prior.mean, fx.x))
    return PosteriorGP(p.prior, (α=α, C=DeviceCholesky{T}(post[], N1 + N2), x=vcat(p.data.x, fx.x), δ=δ))
end

# replaces rand(rng, fx, S) src/finite_gp_projection.jl:233-237: the normals come from the caller's rng
function Random.rand(rng::Random.AbstractRNG, fx::DevFiniteGP{T}, S::Int) where {T}
    supported(fx.f) || return invoke(Random.rand, Tuple{Random.AbstractRNG,FiniteGP,Int}, rng, fx, S)
    c = ctx(); X, layout, D = points(fx.x); Z = randn(rng, T, length(fx), S); out = similar(Z)
    ks, k1 = kernel_spec(fx.f.kernel, T); ms, k2 = mean_spec(fx.f.mean, fx.x, T); ns, k3 = noise_spec(fx.Σy, T)
    lock(c.lock) do
        GC.@preserve X k1 k2 k3 check(c, ccall((:agp_rand, libagp), Int32,
            (Ptr{Cvoid}, Int32, Ref{AgpKernel}, Ref{AgpMean}, Ref{AgpNoise}, Int32, Ptr{Cvoid}, Int64, Int32, Ptr{Cvoid}, Int32, Ptr{Cvoid}),
            c.h, agp_dtype(T), ks, ms, ns, layout, X, length(fx), D, Z, S, out))
    end
    return out
end

# replaces approx_log_evidence(::VFE / ::DTC, fx, y) src/sparse_approximations.jl:248-254, :282-286: ONE streamed pass
# returns both objectives (elbo, dtc)
function sparse_objectives(fz::FiniteGP, fx::DevFiniteGP{T}, y::AbstractVector{<:Real}) where {T}
    fz.f === fx.f || throw(ArgumentError("the inducing and the data FiniteGP must share the prior"))
    length(y) == length(fx) || throw(DimensionMismatch("length(fx) = $(length(fx)) but length(y) = $(length(y))"))
    c = ctx(); X, layout, D = points(fx.x); Z = same_layout(fz.x, layout)
    ks, k1 = kernel_spec(fx.f.kernel, T); ms, k2 = mean_spec(fx.f.mean, fx.x, T)
    ns, k3 = noise_spec(fx.Σy, T); js, k4 = noise_spec(fz.Σy, T)
    yv = convert(Vector{T}, y); out = Vector{T}(undef, 2)
    lock(c.lock) do
        GC.@preserve X Z yv out k1 k2 k3 k4 check(c, ccall((:agp_vfe_elbo, libagp), Int32,
            (Ptr{Cvoid}, Int32, Ref{AgpKernel}, Ref{AgpMean}, Ref{AgpNoise}, Int32, Ptr{Cvoid}, Int64, Int32, Ptr{Cvoid}, Int64,
             Ref{AgpNoise}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
            c.h, agp_dtype(T), ks, ms, ns, layout, X, length(fx), D, Z, length(fz), js, yv, pointer(out, 1), pointer(out, 2)))
    end
    return out[1], out[2]
end
approx_log_evidence(vfe::VFE, fx::DevFiniteGP{T}, y::AbstractVector{<:Real}) where {T} =
    supported(fx.f) ? sparse_objectives(vfe.fz, fx, y)[1] : invoke(approx_log_evidence, Tuple{VFE,FiniteGP,AbstractVector{<:Real}}, vfe, fx, y)
approx_log_evidence(dtc::DTC, fx::DevFiniteGP{T}, y::AbstractVector{<:Real}) where {T} =
    supported(fx.f) ? sparse_objectives(dtc.fz, fx, y)[2] : invoke(approx_log_evidence, Tuple{DTC,FiniteGP,AbstractVector{<:Real}}, dtc, fx, y)

# ---- posterior(::VFE, fx, y) src/sparse_approximations.jl:58-75 and its predictive mean_and_var :212-217 ------------
# The reference's ApproxPosteriorGP caches host matrices; the device posterior keeps chol(K_zz), chol(A A' + I) and
# m_eps in HBM behind an agp_vfe_post handle.
mutable struct DeviceApproxPosterior{T,Tprior,Tz} <: AbstractGPs.AbstractGP
    h::Ptr{Cvoid}
    prior::Tprior
    z::Tz
    function DeviceApproxPosterior{T}(h, prior, z) where {T}
        p = new{T,typeof(prior),typeof(z)}(h, prior, z)
        finalizer(q -> ccall((:agp_vfe_post_free, libagp), Int32, (Ptr{Cvoid},), q.h), p)
    end
end
function posterior(vfe::Union{VFE,DTC}, fx::DevFiniteGP{T}, y::AbstractVector{<:Real}) where {T}
    supported(fx.f) || return invoke(posterior, Tuple{typeof(vfe),FiniteGP,AbstractVector{<:Real}}, vfe, fx, y)
    fz = vfe.fz
    fz.f === fx.f || throw(ArgumentError("the inducing and the data FiniteGP must share the prior"))
    c = ctx(); X, layout, D = points(fx.x); Z = same_layout(fz.x, layout)
    ks, k1 = kernel_spec(fx.f.kernel, T); ms, k2 = mean_spec(fx.f.mean, fx.x, T)
    ns, k3 = noise_spec(fx.Σy, T); js, k4 = noise_spec(fz.Σy, T)
    yv = convert(Vector{T}, y); post = Ref{Ptr{Cvoid}}(C_NULL)
    lock(c.lock) do
        GC.@preserve X Z yv k1 k2 k3 k4 check(c, ccall((:agp_vfe_fit, libagp), Int32,
            (Ptr{Cvoid}, Int32, Ref{AgpKernel}, Ref{AgpMean}, Ref{AgpNoise}, Int32, Ptr{Cvoid}, Int64, Int32, Ptr{Cvoid}, Int64,
             Ref{AgpNoise}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}),
            c.h, agp_dtype(T), ks, ms, ns, layout, X, length(fx), D, Z, length(fz), js, yv, post))
    end
    return DeviceApproxPosterior{T}(post[], fx.f, fz.x)
end
function mean_and_var(fx::FiniteGP{<:DeviceApproxPosterior{T},<:DevInputs{T},<:Diagonal}) where {T}
    p = fx.f; c = ctx(); Xs, layout, D = points(fx.x); M = length(fx)
    μ = Vector{T}(undef, M); v = Vector{T}(undef, M)
    lock(c.lock) do
        GC.@preserve Xs check(c, ccall((:agp_vfe_mean_var, libagp), Int32,
            (Ptr{Cvoid}, Int32, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}), p.h, layout, Xs, M, μ, v))
    end
    # the handle applies Zero / Const prior means; a closure mean is evaluated host-side (src/mean_function.jl:52-55)
    p.prior.mean isa AbstractGPs.CustomMean && (μ .+= AbstractGPs.mean_vector(p.prior.mean, fx.x))
    return μ, v .+ diag(fx.Σy)
end
AbstractGPs.mean_and_var(p::DeviceApproxPosterior{T}, x::DevInputs{T}) where {T} = mean_and_var(p(x, zero(T)))
AbstractGPs.mean(fx::FiniteGP{<:DeviceApproxPosterior}) = mean_and_var(fx)[1]
AbstractGPs.var(fx::FiniteGP{<:DeviceApproxPosterior}) = mean_and_var(fx)[2]

# ---- reverse-mode rule for logpdf (test/finite_gp_projection.jl:152-178, examples/1-mauna-loa/script.jl:200-242) ------
# agp_post_logpdf_grad returns, from the factor of ONE fit, d logpdf / d (total sigma_f^2, total ScaleTransform factor,
# LinearKernel c, scalar noise, constant mean, total ARD weights) and the per-point noise gradient.  `logpdf_and_gradient`
# exposes them flat; the `rrule` maps them back onto the nested kernel structs (chain rule through the products the shim
# forms in `flat`: sigma_f^2 = prod of the ScaledKernel factors, w = prod of the transform scalings).  Inputs x are treated
# as constants (NoTangent); a CustomMean closure is not differentiated.
import ChainRulesCore
const CRC = ChainRulesCore

function logpdf_and_gradient(fx::DevFiniteGP{T}, y::AbstractVector{<:Real}) where {T}
    lp, post = fit(fx, y)
    D = points(fx.x)[3]; N = length(fx)
    g = Vector{Float64}(undef, 5 + D); nd = Vector{T}(undef, N)
    c = ctx()
    lock(c.lock) do
        GC.@preserve g nd check(c, ccall((:agp_post_logpdf_grad, libagp), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Cvoid}), post.data.C.h, g, nd))
    end
    return lp, (variance=g[1], scale=g[2], linear_c=g[3], noise=g[4], mean_c=g[5], ard=g[6:end], noise_diag=nd, y=-post.data.α)
end

# tangent of a (nested) kernel: `tot` carries the flattened totals the gradients refer to
kernel_tangent(k::Stationary, g, var, w) = CRC.NoTangent()
kernel_tangent(k::LinearKernel, g, var, w) = CRC.Tangent{typeof(k)}(; c=[g.linear_c])
function kernel_tangent(k::ScaledKernel, g, var, w)          # d/d sigma_i^2 = d/d var * var / sigma_i^2
    CRC.Tangent{typeof(k)}(; kernel=kernel_tangent(k.kernel, g, var, w), σ²=[g.variance * var / only(k.σ²)])
end
function kernel_tangent(k::TransformedKernel{<:Any,<:ScaleTransform}, g, var, w)
    s_i = only(k.transform.s)
    ds = w isa Real ? g.scale * w / s_i : sum(g.ard .* w) / s_i      # scalar layer under vector totals: sum over dimensions
    CRC.Tangent{typeof(k)}(; kernel=kernel_tangent(k.kernel, g, var, w), transform=CRC.Tangent{typeof(k.transform)}(; s=[ds]))
end
function kernel_tangent(k::TransformedKernel{<:Any,<:ARDTransform}, g, var, w)
    dv = g.ard .* w ./ k.transform.v                                    # d/d v_i[d] = d/d w[d] * w[d] / v_i[d]
    CRC.Tangent{typeof(k)}(; kernel=kernel_tangent(k.kernel, g, var, w), transform=CRC.Tangent{typeof(k.transform)}(; v=dv))
end
mean_tangent(m::AbstractGPs.ConstMean, g) = CRC.Tangent{typeof(m)}(; c=g.mean_c)
mean_tangent(m, g) = CRC.NoTangent()
noise_tangent(Σ::Diagonal{<:Any,<:Fill}, g) = CRC.Tangent{typeof(Σ)}(; diag=CRC.Tangent{typeof(Σ.diag)}(; value=g.noise))
noise_tangent(Σ::Diagonal, g) = CRC.Tangent{typeof(Σ)}(; diag=collect(g.noise_diag))

function CRC.rrule(::typeof(logpdf), fx::DevFiniteGP{T}, y::AbstractVector{<:Real}) where {T}
    supported(fx.f) || return nothing                                     # no rule: AD differentiates the stock method
    lp, g = logpdf_and_gradient(fx, y)
    _, var, _, w = flat(fx.f.kernel)
    w === nothing && (w = 1.0)
    function logpdf_pullback(Δ)
        d = CRC.unthunk(Δ)
        gs = map(v -> v isa Number ? d * v : d .* v, g)
        f̄ = CRC.Tangent{typeof(fx.f)}(; mean=mean_tangent(fx.f.mean, gs), kernel=kernel_tangent(fx.f.kernel, gs, var, w))
        f̄x = CRC.Tangent{typeof(fx)}(; f=f̄, x=CRC.NoTangent(), Σy=noise_tangent(fx.Σy, gs))
        return CRC.NoTangent(), f̄x, gs.y
    end
    return lp, logpdf_pullback
end

end # module
