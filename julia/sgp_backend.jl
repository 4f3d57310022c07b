# sgp_backend.jl -- libsgp.so behind the unchanged ReactiveMP rule surface of biaslab/GaussianProcessNode.
#
# UNTESTED SOURCE: Julia is not installed in the build image, so this file has never been executed (syntax and the third-party field names it
# touches -- KernelFunctions' ScaledKernel / TransformedKernel / ScaleTransform / ARDTransform, ReactiveMP's weightedmean_precision -- are
# written from memory of those packages).  The identical call sequence IS exercised, through ctypes, by gaussianprocessnode_b200/nodes.py
# (tests/test_nodes_gpu.py, tests/test_multisgp_nodes_gpu.py); this file is the reference-side binding a maintainer adds, kept consistent
# with include/sgp.h by hand.  Include it from GPnode/UniSGPnode.jl in place of the rule bodies it overrides (INTEGRATION.md).

const libsgp = "libsgp"                       # libsgp.so on LD_LIBRARY_PATH

mutable struct SGPHandle
    ptr::Ptr{Cvoid}
    function SGPHandle(device::Integer = 0)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:sgp_create, libsgp), Cint, (Ref{Ptr{Cvoid}}, Cint), ref, device)
        rc == 0 || error("sgp_create failed ($rc): no CUDA device / library built for another arch")
        h = new(ref[]); finalizer(x -> ccall((:sgp_destroy, libsgp), Cvoid, (Ptr{Cvoid},), x.ptr), h); h
    end
end
sgp_check(h, rc) = rc == 0 || error(unsafe_string(ccall((:sgp_last_error, libsgp), Cstring, (Ptr{Cvoid},), h.ptr)))

# One staging queue per interface: ReactiveMP may interleave the :v / :w rules and the energy node by node.
mutable struct SGPQueue
    x::Vector{Float64}; y::Vector{Float64}; v::Vector{Float64}; n::Int
end
SGPQueue() = SGPQueue(Float64[], Float64[], Float64[], 0)
function enqueue!(q::SGPQueue, x, μ_y, v_y)
    append!(q.x, x); push!(q.y, μ_y); push!(q.v, v_y); q.n += 1
end
function reset!(q::SGPQueue)
    empty!(q.x); empty!(q.y); empty!(q.v); q.n = 0
end

# meta: the reference's fields, same order (helper_functions/gp_helperfunction.jl:33-44), then the library handle and the shim's state
mutable struct UniSGPMeta{I,K}
    method; Xu::I; Ψ0::Matrix{Float64}; Ψ1_trans::Matrix{Float64}; Ψ2::Matrix{Float64}
    KuuL; kernel::K; Uv; counter::Int; N::Int
    h::SGPHandle
    qv::SGPQueue; qw::SGPQueue; qe::SGPQueue       # staging of the :v rule, the :w rule and the average energy
    wbar::Float64                                   # mean(q_w) of the current :v pass (the message itself is built by the N-th prod)
    θkey::Vector{Float64}; kuukey::Vector{Float64}; kuu_jitter::Float64
    resident::UInt64; swept::Bool                   # hash of the data the library holds; whether its statistics belong to them
end
UniSGPMeta(method, Xu, Ψ0, Ψ1_trans, Ψ2, KuuL, kernel, Uv, counter, N; device = 0, kuu_jitter = 0.0) =
    UniSGPMeta(method, Xu, Ψ0, Ψ1_trans, Ψ2, KuuL, kernel, Uv, counter, N, SGPHandle(device), SGPQueue(), SGPQueue(), SGPQueue(), 0.0,
               Float64[], Float64[], kuu_jitter, UInt64(0), false)

# (kind, σ², ℓ) of the KernelFunctions kernel the notebooks build: σ² * with_lengthscale(SEKernel() | Matern32Kernel() | Matern52Kernel(), ℓ)
# (experiments/regression_kin40k.ipynb:108)
function kernel_params(k, D::Int)
    σ2 = first(k.σ²); tk = k.kernel                 # ScaledKernel(TransformedKernel(base, transform), σ²)
    base = tk.kernel; tr = tk.transform
    ℓ = tr isa KernelFunctions.ScaleTransform ? fill(1 / first(tr.s), D) : 1 ./ tr.v      # with_lengthscale(k, ℓ) = k ∘ ScaleTransform(1/ℓ) | ARDTransform(1 ./ ℓ)
    kind = base isa KernelFunctions.Matern32Kernel ? 1 : base isa KernelFunctions.Matern52Kernel ? 2 : 0
    return kind, σ2, collect(Float64, ℓ)
end

function configure!(meta, θ)                     # kernel(θ) + Xu into the library when θ changed
    θ == meta.θkey && return
    Z = meta.Xu isa AbstractVector{<:Number} ? reshape(collect(Float64, meta.Xu), 1, :) : reduce(hcat, meta.Xu)     # D×M column-major
    D, M = size(Z)
    kind, σ2, ℓ = kernel_params(meta.kernel(θ), D)
    sgp_check(meta.h, ccall((:sgp_set_kernel, libsgp), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Ptr{Cdouble}), meta.h.ptr, kind, D, σ2, ℓ))
    sgp_check(meta.h, ccall((:sgp_set_inducing, libsgp), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), meta.h.ptr, M, Z))
    meta.θkey = copy(θ); meta.resident = UInt64(0); meta.swept = false
end

function ensure_kuu!(meta)                       # host side in the notebooks: fastcholesky!(Kuu).L (regression_kin40k.ipynb:183-184)
    meta.kuukey == meta.θkey && return
    M = length(meta.Xu); L = zeros(M, M)
    sgp_check(meta.h, ccall((:sgp_kuu_factor, libsgp), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cdouble}), meta.h.ptr, meta.kuu_jitter, L))
    meta.KuuL = LowerTriangular(L); meta.kuukey = copy(meta.θkey)
end

# flush a queue into the library (sgp_set_data); a pass that queued what is already resident costs nothing
function upload!(meta, q::SGPQueue)
    q.n == meta.N || error("UniSGP: $(q.n) queued points, meta.N = $(meta.N)")
    key = hash((meta.θkey, q.x, q.y, q.v))
    if key != meta.resident
        sgp_check(meta.h, ccall((:sgp_set_data, libsgp), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                                meta.h.ptr, q.n, q.x, q.y, any(!iszero, q.v) ? q.v : C_NULL, C_NULL))
        meta.resident = key; meta.swept = false
    end
    reset!(q)
end

# ---- :v rule: same signature as GPnode/UniSGPnode.jl:144-158 / :161-173 --------------------------------------
@rule UniSGP(:v, Marginalisation) (q_out::Any, q_in::PointMass, q_w::Any, q_θ::PointMass, meta::UniSGPMeta) = begin
    configure!(meta, mean(q_θ))
    μ_y, v_y = q_out isa PointMass ? (mean(q_out), 0.0) : mean_var(q_out)
    enqueue!(meta.qv, mean(q_in), μ_y, v_y)
    meta.wbar = mean(q_w)
    return BufferUniSGP(nothing, meta)            # the message is materialised by the N-th prod
end

# ---- prod: same method as GPnode/UniSGPnode.jl:62-73 -----------------------------------------------------------
function ReactiveMP.prod(::GenericProd, left::NormalDistributionsFamily, right::BufferUniSGP)
    meta = right.meta
    meta.counter += 1
    meta.counter == meta.N || return left        # first N-1 folds: the pending messages ride in the staging queue
    M = length(meta.Xu); w = meta.wbar
    upload!(meta, meta.qv)
    ψ0 = Ref(0.0); sy2 = Ref(0.0); ψ1 = vec(meta.Ψ1_trans)
    sgp_check(meta.h, ccall((:sgp_sweep_psi, libsgp), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}),
                            meta.h.ptr, ψ0, ψ1, meta.Ψ2, sy2))
    meta.Ψ0[1] = ψ0[]; meta.swept = true
    ξ0, Λ0 = weightedmean_precision(left)
    μ = zeros(M); Σ = zeros(M, M); Uv = zeros(M, M)
    sgp_check(meta.h, ccall((:sgp_posterior_v, libsgp), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                            meta.h.ptr, ξ0, Matrix(Λ0), w, μ, Σ, Uv))
    meta.Uv = UpperTriangular(Uv); meta.counter = 0
    return MvNormalWeightedMeanPrecision(ξ0 + w * ψ1, Λ0 + w * meta.Ψ2)    # mean_cov of it == (μ, Σ) just computed
end

# sum_n of the :w rule / energy ingredients; mu_v and Uv are the PREVIOUS sweep's, as the VMP schedule has them (UniSGPnode.jl:201, 212)
function w_terms!(meta, q::SGPQueue, q_v)
    ensure_kuu!(meta)
    upload!(meta, q)
    if !meta.swept
        sgp_check(meta.h, ccall((:sgp_sweep_psi, libsgp), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                                meta.h.ptr, C_NULL, C_NULL, C_NULL, C_NULL))
        meta.swept = true
    end
    s1 = Ref(0.0); s2 = Ref(0.0)
    sgp_check(meta.h, ccall((:sgp_w_terms, libsgp), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}, Ref{Cdouble}),
                            meta.h.ptr, mean(q_v), Matrix(meta.Uv), s1, s2))
    return s1[], s2[]
end

# Neutral element of the Gamma product for the first N-1 nodes: shape 1 and the smallest positive rate (a PROPER Gamma; the product of
# the N messages then carries (N-1) * floatmin(Float64) ~ 2e-308 * N extra rate).
const NEUTRAL_GAMMA = GammaShapeRate(1.0, floatmin(Float64))

# ---- :w rule: same signature as GPnode/UniSGPnode.jl:196-216 / :219-238 -----------------------------------------
@rule UniSGP(:w, Marginalisation) (q_out::Any, q_in::PointMass, q_v::MultivariateNormalDistributionsFamily, q_θ::PointMass, meta::UniSGPMeta) = begin
    configure!(meta, mean(q_θ))
    μ_y, v_y = q_out isa PointMass ? (mean(q_out), 0.0) : mean_var(q_out)
    enqueue!(meta.qw, mean(q_in), μ_y, v_y)
    meta.qw.n == meta.N || return NEUTRAL_GAMMA
    s1, s2 = w_terms!(meta, meta.qw, q_v)
    return GammaShapeRate(1.0 + meta.N / 2, 0.5 * (s1 + s2))          # = product of the N messages Γ(1.5, rate_n)
end

# ---- average energy: same signature as GPnode/UniSGPnode.jl:337-359 / :363-387 -------------------------------------
@average_energy UniSGP (q_out::Any, q_in::PointMass, q_v::MultivariateNormalDistributionsFamily, q_w::Any, q_θ::PointMass, meta::UniSGPMeta) = begin
    configure!(meta, mean(q_θ))
    μ_y, v_y = q_out isa PointMass ? (mean(q_out), 0.0) : mean_var(q_out)
    enqueue!(meta.qe, mean(q_in), μ_y, v_y)
    meta.qe.n == meta.N || return 0.0
    s1, s2 = w_terms!(meta, meta.qe, q_v)
    E_logw = q_w isa PointMass ? log(mean(q_w)) : mean(log, q_w)
    return 0.5 * mean(q_w) * (s1 + s2) + 0.5 * meta.N * (log(2π) - E_logw)      # sum over the N nodes of U_n
end

# ---- :out rule over a test set (GPnode/UniSGPnode.jl:96-104; regression_kin40k.ipynb:289-304) and the classification drivers' predict_new
# (classification_banana.ipynb:289-293): Xt D×Nt
function predict_new(Xt::Matrix{Float64}, qv, qw, θ, meta::UniSGPMeta; probit = false)
    configure!(meta, θ)
    Nt = size(Xt, 2); m = zeros(Nt)
    if !probit
        sgp_check(meta.h, ccall((:sgp_predict_mean, libsgp), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), meta.h.ptr, Nt, Xt, mean(qv), m))
        return [NormalMeanPrecision(mi, mean(qw)) for mi in m]
    end
    p = zeros(Nt); vf = Ref(0.0)
    sgp_check(meta.h, ccall((:sgp_predict_probit, libsgp), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ref{Cdouble}, Ptr{Cdouble}),
                            meta.h.ptr, Nt, Xt, mean(qv), mean(qw), m, vf, p))
    return [Bernoulli(pi) for pi in p]            # = @call_rule Probit(:out, Marginalisation)(m_in = NormalMeanPrecision(m_n, w̄))
end

# ---- theta step: same keyword signature as helper_functions/derivative_helper.jl:59-63 ---------------------------------------
# θ = [σ²_raw, ℓ_raw...] with softplus (regression_kin40k.ipynb:108); re-points the meta's library context at (x_data, y_data)
function grad_llh_new!(grad, θ; y_data, x_data, v, Uv, w, kernel, Xu, chunk_size = 4, meta::UniSGPMeta, jitter = 0.0)
    configure!(meta, θ)                                   # kernel(θ), Xu -> library
    X = x_data isa AbstractVector{<:Number} ? reshape(collect(Float64, x_data), 1, :) : reduce(hcat, x_data)      # D×N
    sgp_check(meta.h, ccall((:sgp_set_data, libsgp), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                            meta.h.ptr, size(X, 2), X, y_data, C_NULL, C_NULL))
    meta.resident = UInt64(0); meta.swept = false         # the node's cached view of the library no longer holds
    val = Ref(0.0); dσ2 = Ref(0.0); dℓ = zeros(size(X, 1))
    sgp_check(meta.h, ccall((:sgp_theta_objective, libsgp), Cint,
                            (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cdouble, Ref{Cdouble}, Ref{Cdouble}, Ptr{Cdouble}),
                            meta.h.ptr, v, Matrix(Uv), w, jitter, val, dσ2, dℓ))
    meta.kuukey = Float64[]                               # sgp_theta_objective refactored K_uu with ITS jitter
    σ′(t) = 1 / (1 + exp(-t))                             # softplus′
    grad[1] = dσ2[] * σ′(θ[1])
    grad[2:end] .= length(θ) == 2 ? sum(dℓ) * σ′(θ[2]) : dℓ .* σ′.(θ[2:end])
    return grad
end

# ---- one call per mini-batch: H2D of the batch, sweep, D2H of the statistics with a single host synchronisation -----------
# (sgp_sweep_psi_host; X D×N, y N; Ψ1 / Ψ2 best wrapped around sgp_pinned_alloc memory)
function sweep_psi_host!(meta::UniSGPMeta, X::Matrix{Float64}, y::Vector{Float64})
    ψ0 = Ref(0.0); sy2 = Ref(0.0)
    sgp_check(meta.h, ccall((:sgp_sweep_psi_host, libsgp), Cint,
                            (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}),
                            meta.h.ptr, size(X, 2), X, y, C_NULL, C_NULL, ψ0, vec(meta.Ψ1_trans), meta.Ψ2, sy2))
    meta.resident = UInt64(0); meta.swept = true
    return ψ0[], sy2[]
end

# the same with ALL statistics in one packed buffer (one device-to-host copy, half the bytes over the bus):
# buf = [lower triangle of Ψ2 column by column (LAPACK uplo = 'L' packed storage, M(M+1)/2) | Ψ1 (M) | Ψ0, Σw(ȳ²+v), Σw, n].
# `buf` is best a vector over sgp_pinned_alloc memory; Ψ2[i, j] = buf[i + (j - 1) * (2M - j) ÷ 2] for i ≥ j (1-based).
function sweep_psi_host_packed!(meta::UniSGPMeta, X::Matrix{Float64}, y::Vector{Float64}, buf::Vector{Float64})
    M = size(meta.Ψ2, 1); tri = M * (M + 1) ÷ 2
    @assert length(buf) == tri + M + 4
    sgp_check(meta.h, ccall((:sgp_sweep_psi_host_packed, libsgp), Cint,
                            (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                            meta.h.ptr, size(X, 2), X, y, C_NULL, C_NULL, C_NULL, C_NULL, buf, C_NULL))
    meta.resident = UInt64(0); meta.swept = true
    return buf[tri + M + 1], buf[tri + M + 2]          # Ψ0, Σw(ȳ²+v); Ψ1 = view(buf, tri+1:tri+M)
end

# ---- MultiSGP :in messages for a whole chain (GPnode/MultiSGPnode.jl:162-236, prod override :38-45) ------------------------------
# Xp: d×P×N cubature points (P per node), R: D×N columns W μ_y,n (row-major N×D for the library = this array's memory), Mv = reshape(μ_v, M, D),
# S = sum(create_blockmatrix(Σ_v + μ_v μ_v', D, M) .* W).  Returns f (P×N); with derivatives = true also ∇f (d×P×N) and ∇²f (d×d×P×N).
function in_logmessage(h::SGPHandle, Xp::Array{Float64,3}, R::Matrix{Float64}, Mv::Matrix{Float64}, S::Matrix{Float64}, trW::Float64; derivatives = false)
    d, P, N = size(Xp); D = size(Mv, 2)
    f = Matrix{Float64}(undef, P, N)
    g = derivatives ? Array{Float64,3}(undef, d, P, N) : nothing
    H = derivatives ? Array{Float64,4}(undef, d, d, P, N) : nothing
    sgp_check(h, ccall((:sgp_in_logmessage, libsgp), Cint,
                       (Ptr{Cvoid}, Int64, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                       h.ptr, N, P, Xp, D, R, Mv, S, trW, f, derivatives ? g : C_NULL, derivatives ? H : C_NULL))
    return derivatives ? (f, g, H) : f
end

# the prod override for all N (left_n, right_n) pairs: spherical-radial points of every left_n -> ONE library call -> moment matching
function prod_gaussian_logpdf_batch(h::SGPHandle, ms::Matrix{Float64}, Ps::Array{Float64,3}, R, Mv, S, trW)        # ms d×N, Ps d×d×N
    d, N = size(ms); c = sqrt(d + 1.0)
    Xp = Array{Float64,3}(undef, d, 2d + 1, N)
    for n in 1:N
        L = cholesky(Symmetric(Ps[:, :, n])).L
        Xp[:, 1, n] = ms[:, n]
        for j in 1:d
            Xp[:, 1 + j, n] = ms[:, n] + c * L[:, j]; Xp[:, 1 + d + j, n] = ms[:, n] - c * L[:, j]
        end
    end
    wts = vcat(1 / (d + 1), fill(0.5 / (d + 1), 2d))
    g = exp.(in_logmessage(h, Xp, R, Mv, S, trW)) .* wts                         # P×N
    out = Vector{Any}(undef, N)
    for n in 1:N
        Z = sum(g[:, n]); μ = Xp[:, :, n] * g[:, n] / Z
        if isnan(μ[1]); out[n] = MvNormalMeanCovariance(ms[:, n], Ps[:, :, n]); continue; end      # MultiSGPnode.jl:40-41
        Δ = Xp[:, :, n] .- μ
        out[n] = MvNormalMeanCovariance(μ, (Δ .* g[:, n]') * Δ' / Z)
    end
    return out
end
