# GPUGRAPE.jl -- glue that plugs libqocgrape.so into QuOptimalControl.jl's solve(prob, alg) dispatch.
# `include` this file after src/solve.jl (it uses Problem, EnsembleProblem, init_ensemble, SolutionResult, ...).
# UNEXECUTED in this build environment (Julia is not installed); the identical C ABI is exercised from Python.
# See INTEGRATION.md.

using Optim, Parameters
const libqoc = "libqocgrape.so"            # on LD_LIBRARY_PATH, or an absolute path
const QOC_MAX_DEVICES = 16

struct QocDesc                              # mirrors `qoc_desc` in include/qocgrape.h, field for field
    sys_type::Cint; D::Cint; K::Cint; N::Cint; M::Cint; R::Cint
    T::Cdouble; gradient::Cint; convention::Cint; device::Cint
    expm_theta::Cdouble; flags::Cint
    n_devices::Cint; device_ids::NTuple{QOC_MAX_DEVICES,Cint}
end

Base.@kwdef struct GPUGRAPE{OPTS}
    n_slices::Int
    gradient::Symbol = :first_order        # :first_order (GRAPE) | :exact (ADGRAPE semantics)
    convention::Symbol = :inplace          # UnitaryGate sign: grad_func! (:inplace) | grad_func (:static)
    devices::Vector{Int} = [0]             # several ordinals: the ensemble members are sharded over these GPUs inside this process
    pure_state::Bool = true                # D > 16 pure-state transfers on sparse closed systems: vector sweep, same F and G
    penalty::Tuple{Float64,Float64} = (0.0, 0.0)   # weights of C3 (amplitudes) and C4 (variations), src/cost_functions.jl:29-39
    optimizer::Symbol = :optim             # :optim (Optim.LBFGS on the Julia side) | :native (qoc_minimize_lbfgs, Hager-Zhang)
    optim_options::OPTS = Optim.Options()
end

_code(::StateTransfer) = 0; _code(::UnitaryGate) = 1; _code(::CoherenceTransfer) = 2

_check(h, rc) = rc == 0 || error("libqocgrape: ", unsafe_string(ccall((:qoc_last_error, libqoc), Cstring, (Ptr{Cvoid},), h)))

# README's older problem names as keyword constructors onto Problem (src/problems.jl:19-28)
ClosedStateTransfer(; kw...)         = Problem(; sys_type = StateTransfer(), kw...)
UnitarySynthesis(; kw...)            = Problem(; sys_type = UnitaryGate(), kw...)
OpenSystemCoherenceTransfer(; kw...) = Problem(; sys_type = CoherenceTransfer(), kw...)

struct QocLbfgsOptions; max_iters::Cint; history::Cint; g_tol::Cdouble; f_tol::Cdouble; max_linesearch::Cint; linesearch::Cint; end
struct QocLbfgsResult; minimum::Cdouble; g_norm::Cdouble; iterations::Cint; f_calls::Cint; converged::Cint; end

function _qoc_solve(members, wts, guess, alg::GPUGRAPE)
    p1 = members[1]
    D = size(p1.A, 1); K = p1.n_controls; N = alg.n_slices; M = length(members)
    nd = length(alg.devices)
    ids = ntuple(i -> Cint(i <= nd ? alg.devices[i] : 0), QOC_MAX_DEVICES)
    desc = QocDesc(_code(p1.sys_type), D, K, N, M, 1, p1.T, alg.gradient == :exact ? 1 : 0,
                   alg.convention == :static ? 1 : 0, alg.devices[1], 0.0, alg.pure_state ? 0 : 1,   # flags: QOC_FLAG_NO_PURE_STATE = 1
                   nd > 1 ? nd : 0, ids)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    _check(C_NULL, ccall((:qoc_create, libqoc), Cint, (Ref{Ptr{Cvoid}}, Ref{QocDesc}), href, desc))
    h = href[]
    try
        # Julia arrays are column-major ComplexF64 already: concatenate, no transposes
        A  = ComplexF64[p.A[i]    for i in 1:D*D, p in members]                 # [D*D, M]
        B  = ComplexF64[p.B[c][i] for i in 1:D*D, c in 1:K, p in members]       # [D*D, K, M]
        Xi = ComplexF64[p.Xi[i]   for i in 1:D*D, p in members]
        Xt = ComplexF64[p.Xt[i]   for i in 1:D*D, p in members]
        w  = Float64.(wts)
        GC.@preserve A B Xi Xt w _check(h, ccall((:qoc_set_system, libqoc), Cint,
            (Ptr{Cvoid}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{ComplexF64}, Ptr{Float64}, Cint),
            h, A, B, Xi, Xt, w, 0))
        if alg.penalty != (0.0, 0.0)
            _check(h, ccall((:qoc_set_penalty, libqoc), Cint, (Ptr{Cvoid}, Cdouble, Cdouble), h, alg.penalty[1], alg.penalty[2]))
        end
        if alg.optimizer == :native                                              # the optimiser loop inside the library
            xout = similar(guess, Float64); res = Ref(QocLbfgsResult(0, 0, 0, 0, 0))
            opt = Ref(QocLbfgsOptions(0, 0, 0.0, -1.0, 0, 0))                  # defaults = Optim.LBFGS(): m = 10, g_tol 1e-8, Hager-Zhang
            GC.@preserve guess xout _check(h, ccall((:qoc_minimize_lbfgs, libqoc), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ref{QocLbfgsOptions}, Ptr{Float64}, Ref{QocLbfgsResult}), h, guess, opt, xout, res))
            return (minimum = res[].minimum, minimizer = xout, iterations = res[].iterations, f_calls = res[].f_calls)
        end
        Fbuf = Ref{Float64}(0.0)
        topt = (F, G, x) -> begin            # same closure contract as src/solve.jl:75-100 / :164-196
            gptr = G === nothing ? Ptr{Float64}(C_NULL) : pointer(G)
            GC.@preserve x G _check(h, ccall((:qoc_eval, libqoc), Cint,
                (Ptr{Cvoid}, Ptr{Float64}, Ref{Float64}, Ptr{Float64}), h, x, Fbuf, gptr))
            F === nothing ? nothing : Fbuf[]
        end
        return Optim.optimize(Optim.only_fg!(topt), guess, Optim.LBFGS(), alg.optim_options)   # src/solve.jl:138
    finally
        ccall((:qoc_destroy, libqoc), Cint, (Ptr{Cvoid},), h)
    end
end

function solve(prob::Problem, alg::GPUGRAPE)
    res = _qoc_solve([prob], [1.0], prob.guess, alg)
    SolutionResult(res, res.minimum, res.minimizer, prob, alg)                   # src/solve.jl:139
end

function solve(ens::EnsembleProblem, alg::GPUGRAPE)
    members = init_ensemble(ens)                                                  # src/tools.jl:42-53
    res = _qoc_solve(members, ens.wts, members[1].guess, alg)                     # devices = 0:7 shards the members over 8 GPUs
    EnsembleSolutionResult(res, res.minimum, res.minimizer, ens, alg)             # src/solve.jl:245
end

# Gradient-free callers (dCRAB, src/dCRAB.jl:13-89): a handle created with R = n candidates and qoc_eval(h, X, F, C_NULL)
# evaluates a whole Nelder-Mead simplex / candidate set in one call (value only, no backward sweep):
#   user_func_batch(X::Array{Float64,3}) = (ccall((:qoc_eval, libqoc), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
#                                                 h, X, Fvec, C_NULL); Fvec)           # X is K x N x R, Fvec has R entries
