# GAMCSamplerB200.jl -- the `ccall` front-end a GAMCSampler.jl maintainer would add to route the GAMC hot path
# through libgamc_b200.so (include/gamc_b200.h).  It keeps the package's own names and keyword arguments
# (reference src/GAMCSampler.jl:27-48) and only marshals pointers: `Matrix{Float64}` / `Vector{Float64}` are passed
# as-is (column-major FP64 is the ABI's layout), no CUDA.jl, no kernels, no CPU fallback.
#
# NOT EXECUTED IN THE BUILD IMAGE: Julia is not installed there (`julia: command not found`).  The executable twin of
# this file is gamcsampler.jl_b200/frontend.py (ctypes), which the tests and bench.py drive through the same symbols.
# Written for Julia >= 1.6 (the reference itself is Julia 0.6: `Void`, `chol`, `tic/toc` no longer exist).
module GAMCSamplerB200

export GAMC, GAMCTuner, VanillaMCTuner, AcceptanceRateMCTuner, BasicMCRange, LogisticModel, BasicMCJob,
       run!, output, acceptance,
       am_only_update!, smmala_only_update!, mod_update!, rand_update!, cos_update!, rand_exp_decay_update!,
       rand_lin_decay_update!, rand_pow_decay_update!, rand_quad_decay_update!,
       exp_decay, lin_decay, pow_decay, quad_decay

const libgamc = get(ENV, "GAMC_LIB", joinpath(@__DIR__, "..", "libgamc_b200.so"))

# ---- status codes -> Julia exceptions (reference: AssertionError, PosDefException, SingularException) -------------
const GAMC_OK, GAMC_INVALID_ARG, GAMC_NONFINITE_INIT, GAMC_NOT_POSDEF = Cint(0), Cint(1), Cint(2), Cint(3)
struct GamcError <: Exception
  status::Cint
  msg::String
end
function check(h::Ptr{Cvoid}, st::Cint)
  st == GAMC_OK && return nothing
  msg = h == C_NULL ? "" : unsafe_string(ccall((:gamc_last_error_string, libgamc), Cstring, (Ptr{Cvoid},), h))
  st == GAMC_INVALID_ARG && throw(AssertionError(msg))
  st == GAMC_NONFINITE_INIT && throw(AssertionError(msg))
  st == GAMC_NOT_POSDEF && throw(LinearAlgebra.PosDefException(0))
  throw(GamcError(st, msg))
end
import LinearAlgebra

# ---- decay functions (src/stats/schedule.jl:1-7) -------------------------------------------------------------------
exp_decay(i::Integer, tot::Integer, a::Real=10., b::Real=0.) = (1-b)*exp(-a*i/tot)+b
lin_decay(i::Integer, tot::Integer, a::Real=1e-3, b::Real=0.) = (1-b)*(1/(1+a*i))+b
pow_decay(i::Integer, tot::Integer, a::Real=1e-3, b::Real=0.) = (1-b)*(a^(i/tot))+b
quad_decay(i::Integer, tot::Integer, a::Real=1e-3, b::Real=0.) = (1-b)*(1/(1+a*abs2(i)))+b

# ---- schedules (src/samplers/GAMC.jl:264-337): the device path needs (kind, params), so the `update!` functions
#      return a descriptor instead of mutating sstate; `GAMC(update=rand_exp_decay_update!(10.))` replaces the
#      reference's closure `(sstate, pstate, i, tot) -> rand_exp_decay_update!(sstate, pstate, i, tot, 10.)`.
struct Schedule
  kind::Cint
  params::NTuple{3, Float64}
end
sched(kind, p...) = Schedule(Cint(kind), ntuple(i -> i <= length(p) ? Float64(p[i]) : NaN, 3))
am_only_update!() = sched(0)
smmala_only_update!() = sched(1)
mod_update!(n::Integer=10) = sched(2, n)
rand_update!(p::Real=0.5) = sched(3, p)
cos_update!(a::Real=10., b::Real=0.2, c::Real=0.2) = sched(4, a, b, c)
rand_exp_decay_update!(a::Real=10., b::Real=0.) = sched(5, a, b)
rand_lin_decay_update!(a::Real=1e-2, b::Real=0.) = sched(6, a, b)
rand_pow_decay_update!(a::Real=1e-3, b::Real=0.) = sched(7, a, b)
rand_quad_decay_update!(a::Real=1e-5, b::Real=0.) = sched(8, a, b)

# ---- sampler / tuners / range (src/samplers/GAMC.jl:126-151, src/tuners/GAMCTuner.jl:15-23) ------------------------
struct GAMC
  update!::Schedule
  transform::Nothing            # softabs etc.: SURVEY.md 8(f) "next" row
  driftstep::Float64
  minorscale::Float64
  c::Float64
  t0::Int
  function GAMC(update!, transform, driftstep, minorscale, c, t0)
    @assert driftstep > 0 "Drift step is not positive"
    @assert 0 < minorscale "Constant minorscale must be positive, got $minorscale"
    @assert 0 <= c "Constant c must be non-negative"
    @assert t0 > 0 "t0 is not positive"
    new(update!, transform, driftstep, minorscale, c, t0)
  end
end
GAMC(; update::Schedule=rand_update!(), transform=nothing, driftstep::Real=1., minorscale::Real=1., c::Real=0.05, t0::Integer=3) =
  GAMC(update, transform, driftstep, minorscale, c, t0)
Base.show(io::IO, s::GAMC) = print(io, "GAMC sampler: drift step = ", s.driftstep, ", scaling of stabilizing covariance = ",
                                   s.minorscale, ", weight of stabilizing mixture component = ", s.c)

Base.@kwdef struct VanillaMCTuner
  period::Int = 100
  verbose::Bool = false
end
struct AcceptanceRateMCTuner
  targetrate::Float64
  score_k::Float64              # score = x -> logistic_rate_score(x, score_k)
  period::Int
  verbose::Bool
end
AcceptanceRateMCTuner(targetrate::Real; score_k::Real=7., period::Integer=100, verbose::Bool=false) =
  AcceptanceRateMCTuner(targetrate, score_k, period, verbose)
struct GAMCTuner
  smmalatuner::VanillaMCTuner
  amtuner::VanillaMCTuner
  totaltuner::Union{VanillaMCTuner, AcceptanceRateMCTuner}
  verbose::Bool
end
GAMCTuner(s::VanillaMCTuner, a::VanillaMCTuner, t) = GAMCTuner(s, a, t, t.verbose || s.verbose || a.verbose)
Base.show(io::IO, ::GAMCTuner) = print(io, "GAMCTuner")

Base.@kwdef struct BasicMCRange
  nsteps::Int
  burnin::Int = 0
end

# ---- C structs (must match include/gamc_b200.h field for field) ----------------------------------------------------
struct CConfig
  driftstep::Cdouble; minorscale::Cdouble; c::Cdouble
  t0::Int32; schedule::Int32
  sched_params::NTuple{3, Cdouble}
  total_tuner_kind::Int32
  targetrate::Cdouble; score_k::Cdouble
  period::Int32; verbose_total::Int32; verbose_smmala::Int32; verbose_am::Int32
  nsteps::Int64; burnin::Int64
  am_mode::Int32; sampler_kind::Int32
  reuse_tensor::Int32; reserved_::Int32
end
struct CState
  count::Int64; updatetensorcount::Int64; step::Cdouble; sqrttunestep::Cdouble
  total_accepted::Int64; total_proposed::Int64; total_totproposed::Int64; total_rate::Cdouble
  smmala_accepted::Int64; smmala_proposed::Int64; smmala_totproposed::Int64
  am_accepted::Int64; am_proposed::Int64; am_totproposed::Int64
  smmalafrequency::Cdouble; amfrequency::Cdouble; logtarget::Cdouble; ratio::Cdouble
  presentupdatetensor::Int32; pastupdatetensor::Int32; error_flag::Int32; error_pivot::Int32
  error_iteration::Int64; chain_saved::Int64
end

function cconfig(s::GAMC, t::GAMCTuner, r::BasicMCRange; am_mode::Integer=1)
  tt = t.totaltuner
  isacc = tt isa AcceptanceRateMCTuner
  CConfig(s.driftstep, s.minorscale, s.c, s.t0, s.update!.kind, s.update!.params, isacc ? 1 : 0,
          isacc ? tt.targetrate : NaN, isacc ? tt.score_k : 7., tt.period, tt.verbose, t.smmalatuner.verbose, t.amtuner.verbose,
          r.nsteps, r.burnin, am_mode, 0, 0, 0)
end

# ---- model + job (examples/logistic_regression/src/GAMC/simulations.jl:39-48,71-77) --------------------------------
struct LogisticModel        # likelihood_model([Hyperparameter(:λ), Data(:X), Data(:y), p]) with the closed-form callbacks :25-37
  X::Matrix{Float64}
  y::Vector{Float64}
  λ::Float64
end

mutable struct BasicMCJob
  handle::Ptr{Cvoid}
  p::Int
  range::BasicMCRange
  sampler::GAMC
  tuner::GAMCTuner
end

function BasicMCJob(model::LogisticModel, sampler::GAMC, range::BasicMCRange, v0::Dict; tuner::GAMCTuner=GAMCTuner(VanillaMCTuner(), VanillaMCTuner(), VanillaMCTuner()),
                    device::Integer=0, am_mode::Integer=1)
  href = Ref{Ptr{Cvoid}}(C_NULL)
  check(C_NULL, ccall((:gamc_create, libgamc), Cint, (Ref{Ptr{Cvoid}}, Int32), href, device))
  h = href[]
  N, p = size(model.X)
  check(h, ccall((:gamc_target_logistic, libgamc), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Int64, Int32, Cdouble),
                 h, model.X, N, model.y, N, p, model.λ))                      # replaces the callbacks + Data(:X), Data(:y)
  cfg = Ref(cconfig(sampler, tuner, range; am_mode=am_mode))
  θ0 = convert(Vector{Float64}, v0[:p])
  check(h, ccall((:gamc_sampler_init, libgamc), Cint, (Ptr{Cvoid}, Ref{CConfig}, Ptr{Cdouble}), h, cfg, θ0))   # initialize! + sampler_state
  job = BasicMCJob(h, p, range, sampler, tuner)
  finalizer(j -> ccall((:gamc_destroy, libgamc), Cint, (Ptr{Cvoid},), j.handle), job)
  job
end

# run(job): nsteps x iterate!.  Host-drawn tape = the reference's in-line randn/rand draws in fixed per-iteration slots.
function run!(job::BasicMCJob; rng=Base.GLOBAL_RNG)
  n, p = job.range.nsteps, job.p
  u_sched, u_mix, u_acc, z = rand(rng, n), rand(rng, n), rand(rng, n), randn(rng, p, n)
  check(job.handle, ccall((:gamc_tape_set, libgamc), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                          job.handle, n, u_sched, u_mix, z, u_acc))
  check(job.handle, ccall((:gamc_run, libgamc), Cint, (Ptr{Cvoid}, Int64), job.handle, n))
  check(job.handle, ccall((:gamc_sync, libgamc), Cint, (Ptr{Cvoid},), job.handle))
  job
end

struct Chain
  value::Matrix{Float64}             # p x nsamples
  diagnosticvalues::Matrix{Bool}     # 1 x nsamples (:accept)
end
function output(job::BasicMCJob)
  ns = max(0, job.range.nsteps - job.range.burnin)
  value = Matrix{Float64}(undef, job.p, ns)
  acc = Vector{UInt8}(undef, ns)
  check(job.handle, ccall((:gamc_chain_get, libgamc), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cdouble}, Ptr{UInt8}), job.handle, 0, ns, value, acc))
  Chain(value, reshape(acc .!= 0, 1, ns))
end
acceptance(c::Chain) = sum(c.diagnosticvalues)/length(c.diagnosticvalues)

# job.sstate.tune.totaltune.step and job.sstate.updatetensorcount (simulations.jl:85-86)
function state(job::BasicMCJob)
  st = Ref{CState}()
  check(job.handle, ccall((:gamc_get_state, libgamc), Cint, (Ptr{Cvoid}, Ref{CState}), job.handle, st))
  st[]
end

end # module
