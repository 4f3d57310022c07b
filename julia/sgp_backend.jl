# sgp_backend.jl -- libsgp.so behind the unchanged ReactiveMP rule surface of biaslab/GaussianProcessNode.
#
# SOURCE ONLY: Julia is not installed in the build image, so this file has never been executed.  The identical call sequence
# is exercised through ctypes by gaussianprocessnode_b200/nodes.py (tests/test_nodes_gpu.py, tests/test_multisgp_nodes_gpu.py).
# Include it from GPnode/UniSGPnode.jl in place of the rule bodies it overrides (INTEGRATION.md explains every step; the C
# ABI is include/sgp.h).  Pinned staging buffers: sgp_pinned_alloc + unsafe_wrap.  Theta step: sgp_theta_objective replaces
# neg_log_backwardmess_fast + ForwardDiff.gradient! (helper_functions/derivative_helper.jl:23-67).

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

# meta: the reference's fields (helper_functions/gp_helperfunction.jl:33-44) + handle + staging
mutable struct UniSGPMeta{I,K}
    method; Xu::I; Ψ0::Matrix{Float64}; Ψ1_trans::Matrix{Float64}; Ψ2::Matrix{Float64}
    KuuL; kernel::K; Uv; counter::Int; N::Int
    h::SGPHandle; xq::Vector{Float64}; yq::Vector{Float64}; vq::Vector{Float64}; θkey::Vector{Float64}
end

function configure!(meta, θ)                     # kernel(θ) + Xu into the library when θ changed
    θ == meta.θkey && return
    σ2, ℓ = kernel_params(meta.kernel, θ)        # e.g. softplus(θ[1]), softplus.(θ[2:end])  (regression_kin40k.ipynb:108)
    Z = reduce(hcat, meta.Xu)                    # Vector{Vector} -> D×M column-major = what sgp_set_inducing takes
    D, M = size(Z)
    sgp_check(meta.h, ccall((:sgp_set_kernel, libsgp), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Ptr{Cdouble}), meta.h.ptr, 0, D, σ2, ℓ))
    sgp_check(meta.h, ccall((:sgp_set_inducing, libsgp), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}), meta.h.ptr, M, Z))
    meta.θkey = copy(θ)
end

# ---- :v rule: same signature as GPnode/UniSGPnode.jl:144-158 / :161-173 --------------------------------------
@rule UniSGP(:v, Marginalisation) (q_out::Any, q_in::PointMass, q_w::Any, q_θ::PointMass, meta::UniSGPMeta) = begin
    configure!(meta, mean(q_θ))
    μ_y, v_y = q_out isa PointMass ? (mean(q_out), 0.0) : mean_var(q_out)
    append!(meta.xq, mean(q_in)); push!(meta.yq, μ_y); push!(meta.vq, v_y)
    return BufferUniSGP(mean(q_w), meta)         # carries w̄; the message is materialised by the N-th prod
end

# ---- prod: same method as GPnode/UniSGPnode.jl:62-73 -----------------------------------------------------------
function ReactiveMP.prod(::GenericProd, left::NormalDistributionsFamily, right::BufferUniSGP)
    meta = right.meta
    meta.counter += 1
    meta.counter == meta.N || return left        # first N-1 folds: the pending messages ride in the staging buffers
    M = length(meta.Xu); w = right.qv
    sgp_check(meta.h, ccall((:sgp_set_data, libsgp), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                            meta.h.ptr, meta.N, meta.xq, meta.yq, any(!iszero, meta.vq) ? meta.vq : C_NULL, C_NULL))
    ψ0 = Ref(0.0); sy2 = Ref(0.0); ψ1 = vec(meta.Ψ1_trans)
    sgp_check(meta.h, ccall((:sgp_sweep_psi, libsgp), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}),
                            meta.h.ptr, ψ0, ψ1, meta.Ψ2, sy2))
    ξ0, Λ0 = weightedmean_precision(left)
    μ = zeros(M); Σ = zeros(M, M); Uv = zeros(M, M)
    sgp_check(meta.h, ccall((:sgp_posterior_v, libsgp), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                            meta.h.ptr, ξ0, Λ0, w, μ, Σ, Uv))
    meta.Uv = UpperTriangular(Uv); meta.counter = 0
    empty!(meta.xq); empty!(meta.yq); empty!(meta.vq)
    return MvNormalWeightedMeanPrecision(ξ0 + w * ψ1, Λ0 + w * meta.Ψ2)    # mean_cov of it == (μ, Σ) just computed
end

# ---- :w rule: same signature as GPnode/UniSGPnode.jl:196-216 / :219-238 -----------------------------------------
@rule UniSGP(:w, Marginalisation) (q_out::Any, q_in::PointMass, q_v::MultivariateNormalDistributionsFamily, q_θ::PointMass, meta::UniSGPMeta) = begin
    meta.wcount += 1
    meta.wcount == meta.N || return GammaShapeRate(1.0, 0.0)          # neutral element of the Gamma product
    meta.wcount = 0
    s1 = Ref(0.0); s2 = Ref(0.0)
    sgp_check(meta.h, ccall((:sgp_w_terms, libsgp), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}, Ref{Cdouble}),
                            meta.h.ptr, mean(q_v), Matrix(meta.Uv), s1, s2))
    return GammaShapeRate(1.0 + meta.N / 2, 0.5 * (s1[] + s2[]))      # = product of the N messages Γ(1.5, rate_n)
end


# ---- theta step: same keyword signature as helper_functions/derivative_helper.jl:59-63 ---------------------------------------
function grad_llh_new!(grad, θ; y_data, x_data, v, Uv, w, kernel, Xu, chunk_size = 4, meta::UniSGPMeta, jitter = 0.0)
    configure!(meta, θ)                                   # kernel(θ), Xu -> library
    X = reduce(hcat, x_data)                              # D×N
    sgp_check(meta.h, ccall((:sgp_set_data, libsgp), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                            meta.h.ptr, size(X, 2), X, y_data, C_NULL, C_NULL))
    val = Ref(0.0); dσ2 = Ref(0.0); dℓ = zeros(size(X, 1))
    sgp_check(meta.h, ccall((:sgp_theta_objective, libsgp), Cint,
                            (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cdouble, Ref{Cdouble}, Ref{Cdouble}, Ptr{Cdouble}),
                            meta.h.ptr, v, Matrix(Uv), w, jitter, val, dσ2, dℓ))
    σ′(t) = 1 / (1 + exp(-t))                             # softplus′: θ = [σ²_raw, ℓ_raw...]  (regression_kin40k.ipynb:108)
    grad[1] = dσ2[] * σ′(θ[1])
    grad[2:end] .= dℓ .* σ′.(θ[2:end])
    return grad
end


# ---- one call per mini-batch: H2D of the batch, sweep, D2H of the statistics with a single host synchronisation -----------
# (sgp_sweep_psi_host; X D×N, y N; Ψ1 / Ψ2 best wrapped around sgp_pinned_alloc memory)
function sweep_psi_host!(meta::UniSGPMeta, X::Matrix{Float64}, y::Vector{Float64})
    ψ0 = Ref(0.0); sy2 = Ref(0.0)
    sgp_check(meta.h, ccall((:sgp_sweep_psi_host, libsgp), Cint,
                            (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}),
                            meta.h.ptr, size(X, 2), X, y, C_NULL, C_NULL, ψ0, vec(meta.Ψ1_trans), meta.Ψ2, sy2))
    return ψ0[], sy2[]
end

# ---- MultiSGP :in messages for a whole chain (GPnode/MultiSGPnode.jl:162-236, prod override :38-45) ------------------------------
# Xp: d×P×N cubature points (P per node), R: D×N columns W μ_y,n (row-major N×D for the library = this array's memory), Mv = reshape(μ_v, M, D),
# S = sum(create_blockmatrix(Σ_v + μ_v μ_v', D, M) .* W).  Returns f (P×N); with derivatives = true also ∇f (d×P×N) and ∇²f (d×d×P×N).
function in_logmessage(meta, Xp::Array{Float64,3}, R::Matrix{Float64}, Mv::Matrix{Float64}, S::Matrix{Float64}, trW::Float64; derivatives = false)
    d, P, N = size(Xp); D = size(Mv, 2)
    f = Matrix{Float64}(undef, P, N)
    g = derivatives ? Array{Float64,3}(undef, d, P, N) : nothing
    H = derivatives ? Array{Float64,4}(undef, d, d, P, N) : nothing
    sgp_check(meta.h, ccall((:sgp_in_logmessage, libsgp), Cint,
                            (Ptr{Cvoid}, Int64, Cint, Ptr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                            meta.h.ptr, N, P, Xp, D, R, Mv, S, trW, f, derivatives ? g : C_NULL, derivatives ? H : C_NULL))
    return derivatives ? (f, g, H) : f
end

# the prod override for all N (left_n, right_n) pairs: spherical-radial points of every left_n -> ONE library call -> moment matching
function prod_gaussian_logpdf_batch(meta, ms::Matrix{Float64}, Ps::Array{Float64,3}, R, Mv, S, trW)        # ms d×N, Ps d×d×N
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
    g = exp.(in_logmessage(meta, Xp, R, Mv, S, trW)) .* wts                      # P×N
    out = Vector{Any}(undef, N)
    for n in 1:N
        Z = sum(g[:, n]); μ = Xp[:, :, n] * g[:, n] / Z
        if isnan(μ[1]); out[n] = MvNormalMeanCovariance(ms[:, n], Ps[:, :, n]); continue; end      # MultiSGPnode.jl:40-41
        Δ = Xp[:, :, n] .- μ
        out[n] = MvNormalMeanCovariance(μ, (Δ .* g[:, n]') * Δ' / Z)
    end
    return out
end
