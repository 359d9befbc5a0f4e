# GAMCSamplerB200.jl -- the `ccall` front-end a GAMCSampler.jl maintainer would add to route the GAMC hot path
# through libgamc_b200.so (include/gamc_b200.h).  It keeps the package's own names and keyword arguments
# (reference src/GAMCSampler.jl:27-48) and only marshals pointers: `Matrix{Float64}` / `Vector{Float64}` are passed
# as-is (column-major FP64 is the ABI's layout), no CUDA.jl, no kernels, no CPU fallback.
#
# Every entry point of include/gamc_b200.h is bound here (tests/test_host_logic.py checks the list against the header).
#
# NOT EXECUTED IN THE BUILD IMAGE: Julia is not installed there (`julia: command not found`).  The executable twin of
# this file is gamcsampler.jl_b200/frontend.py (ctypes), which the tests and bench.py drive through the same symbols.
# Written for Julia >= 1.6 (the reference itself is Julia 0.6: `Void`, `chol`, `tic/toc` no longer exist).
module GAMCSamplerB200

import LinearAlgebra

export GAMC, MALA, GAMCTuner, VanillaMCTuner, AcceptanceRateMCTuner, BasicMCRange, LogisticModel, GenericModel, DeviceModel,
       MvTModel, RVModel, BasicMCJob, BatchedMCJob, run!, output, acceptance, softabs, Softabs,
       am_only_update!, smmala_only_update!, mod_update!, rand_update!, cos_update!, rand_exp_decay_update!,
       rand_lin_decay_update!, rand_pow_decay_update!, rand_quad_decay_update!,
       exp_decay, lin_decay, pow_decay, quad_decay,
       eval_logtarget, eval_upto_tensor, eval_partials, state, set_state!, sstate_array, monitor, tune_log, tape_get,
       comm_unique_id, comm_init!, comm_peer_path, comm_set_timeout!, probe_fp64_peak, profile!, profile_get, launch_count

const libgamc = get(ENV, "GAMC_LIB", joinpath(@__DIR__, "..", "libgamc_b200.so"))
const Handle = Ptr{Cvoid}

# ---- status codes -> Julia exceptions (reference: AssertionError, PosDefException, SingularException) -------------
const GAMC_OK, GAMC_INVALID_ARG, GAMC_NONFINITE_INIT, GAMC_NOT_POSDEF = Cint(0), Cint(1), Cint(2), Cint(3)
struct GamcError <: Exception
  status::Cint
  msg::String
end
function check(h::Handle, st::Cint)
  st == GAMC_OK && return nothing
  msg = h == C_NULL ? "" : unsafe_string(ccall((:gamc_last_error_string, libgamc), Cstring, (Handle,), h))
  st == GAMC_INVALID_ARG && throw(AssertionError(msg))
  st == GAMC_NONFINITE_INIT && throw(AssertionError(msg))
  st == GAMC_NOT_POSDEF && throw(LinearAlgebra.PosDefException(0))
  throw(GamcError(st, msg))
end
version() = unsafe_string(ccall((:gamc_version, libgamc), Cstring, ()))

# ---- decay functions (src/stats/schedule.jl:1-7) -------------------------------------------------------------------
exp_decay(i::Integer, tot::Integer, a::Real=10., b::Real=0.) = (1-b)*exp(-a*i/tot)+b
lin_decay(i::Integer, tot::Integer, a::Real=1e-3, b::Real=0.) = (1-b)*(1/(1+a*i))+b
pow_decay(i::Integer, tot::Integer, a::Real=1e-3, b::Real=0.) = (1-b)*(a^(i/tot))+b
quad_decay(i::Integer, tot::Integer, a::Real=1e-3, b::Real=0.) = (1-b)*(1/(1+a*abs2(i)))+b

# ---- schedules (src/samplers/GAMC.jl:264-337).  `update!` is either one of the nine schedules with its parameters bound
#      (`GAMC(update=rand_exp_decay_update!(10.))` replaces the reference's closure
#      `(sstate, pstate, i, tot) -> rand_exp_decay_update!(sstate, pstate, i, tot, 10.)`) or ANY function
#      `(i, tot, u) -> Bool` (true = SMMALA step), u being the iteration's schedule uniform from the tape in place of the
#      reference's in-line `rand(Bernoulli(.))` -- installed with gamc_set_schedule_callback.
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

# ---- metric transform (Klara softabs; examples/t_distribution/src/GAMC/simulations.jl:39) ---------------------------
struct Softabs
  alpha::Float64
end
softabs(alpha::Real=1000.) = Softabs(Float64(alpha))
function softabs_host(H::Matrix{Float64}, alpha::Float64)
  F = LinearAlgebra.eigen(LinearAlgebra.Symmetric((H + H') / 2))
  s = map(l -> abs(alpha * l) < 1e-8 ? 1 / alpha : l * coth(alpha * l), F.values)
  F.vectors * LinearAlgebra.Diagonal(s) * F.vectors'
end

# ---- sampler / tuners / range (src/samplers/GAMC.jl:126-151, src/tuners/GAMCTuner.jl:15-23) ------------------------
struct GAMC
  update!::Union{Schedule, Function}
  transform::Union{Nothing, Softabs, Function}
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
GAMC(; update=rand_update!(), transform=nothing, driftstep::Real=1., minorscale::Real=1., c::Real=0.05, t0::Integer=3) =
  GAMC(update, transform, driftstep, minorscale, c, t0)
Base.show(io::IO, s::GAMC) = print(io, "GAMC sampler: drift step = ", s.driftstep, ", scaling of stabilizing covariance = ",
                                   s.minorscale, ", weight of stabilizing mixture component = ", s.c)

# Klara's MALA(driftstep): the baseline of the reference's efficiency tables (examples/logistic_regression/src/MALA/simulations.jl:37)
struct MALA
  driftstep::Float64
  MALA(driftstep::Real=1.) = (@assert driftstep > 0 "Drift step is not positive"; new(driftstep))
end

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
GAMCTuner() = GAMCTuner(VanillaMCTuner(), VanillaMCTuner(), VanillaMCTuner())
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

function cconfig(s::GAMC, t::GAMCTuner, r::BasicMCRange; am_mode::Integer=1, reuse_tensor::Bool=false)
  tt = t.totaltuner
  isacc = tt isa AcceptanceRateMCTuner
  kind, params = s.update! isa Schedule ? (s.update!.kind, s.update!.params) : (Cint(1), (NaN, NaN, NaN))   # a function: callback installed separately
  CConfig(s.driftstep, s.minorscale, s.c, s.t0, kind, params, isacc ? 1 : 0,
          isacc ? tt.targetrate : NaN, isacc ? tt.score_k : 7., tt.period, tt.verbose, t.smmalatuner.verbose, t.amtuner.verbose,
          r.nsteps, r.burnin, am_mode, 0, reuse_tensor ? 1 : 0, 0)
end
function cconfig(s::MALA, t, r::BasicMCRange)
  tt = t isa GAMCTuner ? t.totaltuner : t
  isacc = tt isa AcceptanceRateMCTuner
  CConfig(s.driftstep, 1., 0., 1, 1, (NaN, NaN, NaN), isacc ? 1 : 0, isacc ? tt.targetrate : NaN, isacc ? tt.score_k : 7., tt.period,
          tt.verbose, 0, 0, r.nsteps, r.burnin, 0, 1, 0, 0)
end

# ---- models ----------------------------------------------------------------------------------------------------------
struct LogisticModel        # likelihood_model([Hyperparameter(:λ), Data(:X), Data(:y), p]) with the closed-form callbacks, simulations.jl:25-48
  X::Matrix{Float64}
  y::Vector{Float64}
  λ::Float64
end
# any model given by its Klara-style closures, evaluated on the HOST (gamc_target_callback): uptotensorlogtarget(θ) -> (lt, grad, G)
struct GenericModel
  p::Int
  uptotensorlogtarget::Function
  logtarget::Union{Nothing, Function}
end
GenericModel(p, f) = GenericModel(p, f, nothing)
# a model that lives on the DEVICE (gamc_target_device_callback): enqueue!(dθ::Ptr{Cdouble}, p, level, dout::Ptr{Cdouble}, stream::Ptr{Cvoid})
# must only enqueue work on `stream` (e.g. CUDA.jl kernels launched with `stream=CuStream(stream)`) that writes
# [logtarget, grad(p), pad, G(p x p column-major)] at dout (G at offset (p + 2) & ~1)
struct DeviceModel
  p::Int
  enqueue!::Function
end
Base.@kwdef struct MvTModel  # examples/t_distribution/src/GAMC/simulations.jl:17-33
  n::Int = 20
  ν::Float64 = 30.
  c::Float64 = 0.9
end
struct RVModel               # examples/radial_velocity/src/stat_model.jl:18-63
  times::Vector{Float64}
  obs::Vector{Float64}
  sigma_obs::Vector{Float64}
  nplanets::Int
end

# ---- single-chain job (examples/logistic_regression/src/GAMC/simulations.jl:39-48,71-77) ---------------------------
mutable struct BasicMCJob
  handle::Handle
  p::Int
  range::BasicMCRange
  sampler::Union{GAMC, MALA}
  tuner
  keep::Vector{Any}          # closures and their @cfunction pointers must outlive the handle
end

# host-evaluated generic target: the C callback receives the job's model through `user`
function target_trampoline(θp::Ptr{Cdouble}, p::Int32, level::Int32, ltp::Ptr{Cdouble}, gp::Ptr{Cdouble}, Gp::Ptr{Cdouble}, user::Ptr{Cvoid})::Int32
  try
    model, transform = unsafe_pointer_to_objref(user)::Tuple{GenericModel, Any}
    θ = copy(unsafe_wrap(Array, θp, Int(p)))
    if level == 0 && model.logtarget !== nothing
      unsafe_store!(ltp, Float64(model.logtarget(θ)))
      return Int32(0)
    end
    lt, grad, G = model.uptotensorlogtarget(θ)
    unsafe_store!(ltp, Float64(lt))
    level >= 1 && copyto!(unsafe_wrap(Array, gp, Int(p)), grad)
    if level >= 2
      G = Matrix{Float64}(G)
      transform isa Softabs && (G = softabs_host(G, transform.alpha))            # src/samplers/iterate/GAMC.jl:22-24,39-41
      transform isa Function && (G = transform(G))
      copyto!(unsafe_wrap(Array, Gp, Int(p) * Int(p)), vec(G))
    end
    return Int32(0)
  catch
    return Int32(1)
  end
end
function enqueue_trampoline(dθ::Ptr{Cdouble}, p::Int32, level::Int32, dout::Ptr{Cdouble}, stream::Ptr{Cvoid}, user::Ptr{Cvoid})::Int32
  try
    (unsafe_pointer_to_objref(user)::DeviceModel).enqueue!(dθ, Int(p), Int(level), dout, stream)
    return Int32(0)
  catch
    return Int32(1)
  end
end
function schedule_trampoline(i::Int64, tot::Int64, u::Cdouble, user::Ptr{Cvoid})::Int32
  f = unsafe_pointer_to_objref(user)::Base.RefValue{Function}
  try
    return Int32(f[](i, tot, u) ? 1 : 0)
  catch
    return Int32(1)
  end
end

function create(device::Integer)
  href = Ref{Handle}(C_NULL)
  check(C_NULL, ccall((:gamc_create, libgamc), Cint, (Ref{Handle}, Int32), href, device))
  href[]
end

function BasicMCJob(model, sampler::Union{GAMC, MALA}, range::BasicMCRange, v0::Dict; tuner=GAMCTuner(), device::Integer=0, am_mode::Integer=1,
                    reuse_tensor::Bool=false, comm=nothing, stream::Ptr{Cvoid}=C_NULL)
  h = create(device)
  keep = Any[]
  stream != C_NULL && check(h, ccall((:gamc_set_stream, libgamc), Cint, (Handle, Ptr{Cvoid}), h, stream))
  if model isa LogisticModel
    sampler isa GAMC && sampler.transform !== nothing &&
      error("transform= is supported for GenericModel / DeviceModel and for BatchedMCJob; the built-in logistic target takes transform=nothing")
    N, p = size(model.X)
    check(h, ccall((:gamc_target_logistic, libgamc), Cint, (Handle, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Int64, Int32, Cdouble),
                   h, model.X, N, model.y, N, p, model.λ))                      # replaces the callbacks + Data(:X), Data(:y)
  elseif model isa GenericModel
    p = model.p
    box = (model, sampler isa GAMC ? sampler.transform : nothing)
    push!(keep, box)
    cb = @cfunction(target_trampoline, Int32, (Ptr{Cdouble}, Int32, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}))
    check(h, ccall((:gamc_target_callback, libgamc), Cint, (Handle, Int32, Ptr{Cvoid}, Any), h, p, cb, box))
  elseif model isa DeviceModel
    p = model.p
    push!(keep, model)
    cb = @cfunction(enqueue_trampoline, Int32, (Ptr{Cdouble}, Int32, Int32, Ptr{Cdouble}, Ptr{Cvoid}, Ptr{Cvoid}))
    check(h, ccall((:gamc_target_device_callback, libgamc), Cint, (Handle, Int32, Ptr{Cvoid}, Any), h, p, cb, model))
  else
    error("unsupported model type $(typeof(model))")
  end
  comm === nothing || comm_init!(h, comm...)                                     # (uid, nranks, rank): this process holds one row shard
  if sampler isa GAMC && sampler.update! isa Function
    fref = Ref{Function}(sampler.update!)
    push!(keep, fref)
    cb = @cfunction(schedule_trampoline, Int32, (Int64, Int64, Cdouble, Ptr{Cvoid}))
    check(h, ccall((:gamc_set_schedule_callback, libgamc), Cint, (Handle, Ptr{Cvoid}, Any), h, cb, fref))
  else
    check(h, ccall((:gamc_set_schedule_callback, libgamc), Cint, (Handle, Ptr{Cvoid}, Ptr{Cvoid}), h, C_NULL, C_NULL))
  end
  cfg = Ref(sampler isa MALA ? cconfig(sampler, tuner, range) : cconfig(sampler, tuner, range; am_mode=am_mode, reuse_tensor=reuse_tensor))
  θ0 = convert(Vector{Float64}, v0[:p])
  check(h, ccall((:gamc_sampler_init, libgamc), Cint, (Handle, Ref{CConfig}, Ptr{Cdouble}), h, cfg, θ0))   # initialize! + sampler_state
  job = BasicMCJob(h, p, range, sampler, tuner, keep)
  finalizer(j -> ccall((:gamc_destroy, libgamc), Cint, (Handle,), j.handle), job)
  job
end

# run(job): nsteps x iterate!.  Host-drawn tape = the reference's in-line randn/rand draws in fixed per-iteration slots;
# `device_tape=seed` draws it on the device instead (gamc_tape_generate, Philox), with no host traffic.
function run!(job::BasicMCJob; rng=Base.GLOBAL_RNG, device_tape::Union{Nothing, Integer}=nothing, n::Integer=job.range.nsteps)
  p, h = job.p, job.handle
  if device_tape === nothing
    u_sched, u_mix, u_acc, z = rand(rng, n), rand(rng, n), rand(rng, n), randn(rng, p, n)
    check(h, ccall((:gamc_tape_set, libgamc), Cint, (Handle, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), h, n, u_sched, u_mix, z, u_acc))
  else
    check(h, ccall((:gamc_tape_generate, libgamc), Cint, (Handle, Int64, UInt64), h, n, device_tape))
  end
  check(h, ccall((:gamc_run, libgamc), Cint, (Handle, Int64), h, n))
  check(h, ccall((:gamc_sync, libgamc), Cint, (Handle,), h))
  job.tuner isa GAMCTuner && job.tuner.verbose && print_burnin_report(job)
  job
end

# one call with host tape in and host chain out (gamc_iterate)
function iterate!(job::BasicMCJob, u_sched::Vector{Float64}, u_mix::Vector{Float64}, z::Matrix{Float64}, u_acc::Vector{Float64})
  n = length(u_sched)
  value = Matrix{Float64}(undef, job.p, n); acc = Vector{UInt8}(undef, n); nsaved = Ref{Int64}(0)
  check(job.handle, ccall((:gamc_iterate, libgamc), Cint,
                          (Handle, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{UInt8}, Ref{Int64}),
                          job.handle, n, u_sched, u_mix, z, u_acc, value, acc, nsaved))
  value[:, 1:nsaved[]], acc[1:nsaved[]] .!= 0
end

function peek_kinds(job::BasicMCJob, n::Integer)
  out = Vector{UInt8}(undef, n)
  check(job.handle, ccall((:gamc_peek_kinds, libgamc), Cint, (Handle, Int64, Ptr{UInt8}), job.handle, n, out))
  out
end

function tape_get(job::BasicMCJob, first::Integer, n::Integer)
  us, um, ua, z = Vector{Float64}(undef, n), Vector{Float64}(undef, n), Vector{Float64}(undef, n), Matrix{Float64}(undef, job.p, n)
  check(job.handle, ccall((:gamc_tape_get, libgamc), Cint, (Handle, Int64, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                          job.handle, first, n, us, um, z, ua))
  us, um, z, ua
end

struct Chain
  value::Matrix{Float64}             # p x nsamples
  diagnosticvalues::Matrix{Bool}     # 1 x nsamples (:accept)
  logtarget::Vector{Float64}         # outopts[:monitor] containing :logtarget
end
function output(job::BasicMCJob)
  ns = max(0, job.range.nsteps - job.range.burnin)
  value = Matrix{Float64}(undef, job.p, ns)
  acc = Vector{UInt8}(undef, ns)
  lt = Vector{Float64}(undef, ns)
  check(job.handle, ccall((:gamc_chain_get, libgamc), Cint, (Handle, Int64, Int64, Ptr{Cdouble}, Ptr{UInt8}), job.handle, 0, ns, value, acc))
  check(job.handle, ccall((:gamc_chain_get_logtarget, libgamc), Cint, (Handle, Int64, Int64, Ptr{Cdouble}), job.handle, 0, ns, lt))
  Chain(value, reshape(acc .!= 0, 1, ns), lt)
end
acceptance(c::Chain) = sum(c.diagnosticvalues)/length(c.diagnosticvalues)

# job.sstate.tune.totaltune.step and job.sstate.updatetensorcount (simulations.jl:85-86)
function state(job::BasicMCJob)
  st = Ref{CState}()
  check(job.handle, ccall((:gamc_get_state, libgamc), Cint, (Handle, Ref{CState}), job.handle, st))
  st[]
end
# fields of MuvGAMCState / pstate (gamc_array_id): :value, :gradlogtarget, :tensorlogtarget, :lastmean, :secondlastmean,
# :oldinvtensor, :cholinvtensor, :oldfirstterm, :proposedvalue, and the monitored extras of src/samplers/iterate/GAMC.jl:74-90
const ARRAY_IDS = Dict(:value => 0, :gradlogtarget => 1, :tensorlogtarget => 2, :lastmean => 3, :secondlastmean => 4, :oldinvtensor => 5,
                       :cholinvtensor => 6, :oldfirstterm => 7, :proposedvalue => 8, :gradloglikelihood => 9, :gradlogprior => 10,
                       :tensorloglikelihood => 11, :tensorlogprior => 12)
function sstate_array(job::BasicMCJob, name::Symbol)
  id = ARRAY_IDS[name]
  ismat = id in (2, 5, 6, 11, 12)
  out = ismat ? Matrix{Float64}(undef, job.p, job.p) : Vector{Float64}(undef, job.p)
  check(job.handle, ccall((:gamc_get_array, libgamc), Cint, (Handle, Int32, Ptr{Cdouble}), job.handle, id, out))
  out
end
# (pstate.loglikelihood, pstate.logprior) of the current point
function monitor(job::BasicMCJob)
  ll, lp = Ref{Cdouble}(0), Ref{Cdouble}(0)
  check(job.handle, ccall((:gamc_get_monitor, libgamc), Cint, (Handle, Ref{Cdouble}, Ref{Cdouble}), job.handle, ll, lp))
  ll[], lp[]
end
# restart a chain from saved host state (gamc_set_state); pass the tape of the REMAINING iterations to run!
function set_state!(job::BasicMCJob, st::CState, θ::Vector{Float64}, lastmean::Vector{Float64}, secondlastmean::Vector{Float64},
                    oldinvtensor::Union{Nothing, Matrix{Float64}}=nothing)
  check(job.handle, ccall((:gamc_set_state, libgamc), Cint, (Handle, Ref{CState}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                          job.handle, Ref(st), θ, lastmean, secondlastmean, oldinvtensor === nothing ? C_NULL : oldinvtensor))
  job
end
# records of the burn-in report (16 x nrecords), and the report itself in the reference's format (iterate/GAMC.jl:218-268)
function tune_log(job::BasicMCJob)
  n = Ref{Int64}(0)
  check(job.handle, ccall((:gamc_tune_log_get, libgamc), Cint, (Handle, Ptr{Cdouble}, Int64, Ref{Int64}), job.handle, C_NULL, 0, n))
  out = Matrix{Float64}(undef, 16, n[])
  n[] > 0 && check(job.handle, ccall((:gamc_tune_log_get, libgamc), Cint, (Handle, Ptr{Cdouble}, Int64, Ref{Int64}), job.handle, out, n[], n))
  out
end
function print_burnin_report(job::BasicMCJob)
  t = job.tuner
  for r in eachcol(tune_log(job))
    println("Burnin iteration ", Int(r[1]), " out of ", job.range.burnin, "...")
    t.totaltuner.verbose && println("  Total : ", Int(r[2]), "/", Int(r[3]), " (", round(100r[4], digits=2), "%) acceptance rate")
    t.smmalatuner.verbose && println("  SMMALA: ", Int(r[5]), "/", Int(r[6]), " (", round(100r[7], digits=2), "%) acceptance rate, ", Int(r[6]), "/", Int(r[3]),
                                     " (", round(100r[8], digits=2), "%) running frequency")
    t.amtuner.verbose && println("  AM    : ", Int(r[9]), "/", Int(r[10]), " (", round(100r[11], digits=2), "%) acceptance rate, ", Int(r[10]), "/", Int(r[3]),
                                 " (", round(100r[12], digits=2), "%) running frequency")
  end
end

# ---- Klara closures on their own: parameter.logtarget!(pstate) / parameter.uptotensorlogtarget!(pstate) -----------
function eval_logtarget(job::BasicMCJob, θ::Vector{Float64})
  lt = Ref{Cdouble}(0)
  check(job.handle, ccall((:gamc_eval_logtarget, libgamc), Cint, (Handle, Ptr{Cdouble}, Ref{Cdouble}), job.handle, θ, lt))
  lt[]
end
function eval_upto_tensor(job::BasicMCJob, θ::Vector{Float64})
  lt = Ref{Cdouble}(0); g = Vector{Float64}(undef, job.p); G = Matrix{Float64}(undef, job.p, job.p)
  check(job.handle, ccall((:gamc_eval_upto_tensor, libgamc), Cint, (Handle, Ptr{Cdouble}, Ref{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), job.handle, θ, lt, g, G))
  lt[], g, G
end
function eval_partials(job::BasicMCJob, θ::Vector{Float64})     # this rank's likelihood-only sums, before the all-reduce
  ll = Ref{Cdouble}(0); g = Vector{Float64}(undef, job.p); G = Matrix{Float64}(undef, job.p, job.p)
  check(job.handle, ccall((:gamc_eval_partials, libgamc), Cint, (Handle, Ptr{Cdouble}, Ref{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), job.handle, θ, ll, g, G))
  ll[], g, G
end

# ---- data already on the device / generated on the device -----------------------------------------------------------
target_logistic_device!(h::Handle, dX::Ptr{Cdouble}, ldX::Integer, dy::Ptr{Cdouble}, N::Integer, p::Integer, λ::Real) =
  check(h, ccall((:gamc_target_logistic_device, libgamc), Cint, (Handle, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Int64, Int32, Cdouble), h, dX, ldX, dy, N, p, λ))
target_logistic_synth!(h::Handle, seed::Integer, row0::Integer, nrows::Integer, p::Integer, θtrue::Vector{Float64}, λ::Real) =
  check(h, ccall((:gamc_target_logistic_synth, libgamc), Cint, (Handle, UInt64, Int64, Int64, Int32, Ptr{Cdouble}, Cdouble), h, seed, row0, nrows, p, θtrue, λ))
function target_read_rows(h::Handle, p::Integer, r0::Integer, n::Integer)
  X = Matrix{Float64}(undef, n, p); y = Vector{Float64}(undef, n)
  check(h, ccall((:gamc_target_read_rows, libgamc), Cint, (Handle, Int64, Int64, Ptr{Cdouble}, Ptr{Cdouble}), h, r0, n, X, y))
  X, y
end

# ---- multi-GPU: one process per GPU, observations row-sharded (SURVEY.md 8e) ----------------------------------------
function comm_unique_id()
  id = Vector{UInt8}(undef, 128)
  check(C_NULL, ccall((:gamc_comm_unique_id, libgamc), Cint, (Ptr{UInt8},), id))
  id                                   # broadcast these 128 bytes with Distributed / MPI, then comm_init! on every rank
end
comm_init!(h::Handle, uid::Vector{UInt8}, nranks::Integer, rank::Integer) =
  check(h, ccall((:gamc_comm_init, libgamc), Cint, (Handle, Ptr{UInt8}, Int32, Int32), h, uid, nranks, rank))
function comm_peer_path(h::Handle)
  a = Ref{Int32}(0)
  check(h, ccall((:gamc_comm_peer_path, libgamc), Cint, (Handle, Ref{Int32}), h, a))
  a[]                                  # bit 0: AM scalar over NVLink mailboxes, bit 1: SMMALA payload summed inside k_land
end
comm_set_timeout!(h::Handle, seconds::Real) = check(h, ccall((:gamc_comm_set_timeout, libgamc), Cint, (Handle, Cdouble), h, seconds))
# metric route of the logistic targets set up afterwards: 0 = auto, 1 = FP64 DMMA, 2 = exact int8 digit products on tcgen05
set_metric_route!(h::Handle, route::Integer; slices::Integer = 0, smax::Integer = 0) =
  check(h, ccall((:gamc_set_metric_route, libgamc), Cint, (Handle, Int32, Int32, Int32), h, route, slices, smax))

function metric_route(h::Handle)
  r = Ref{Int32}(0); k = Ref{Int32}(0); s = Ref{Int32}(0)
  check(h, ccall((:gamc_get_metric_route, libgamc), Cint, (Handle, Ref{Int32}, Ref{Int32}, Ref{Int32}), h, r, k, s))
  (r[], k[], s[])                      # route (1 = FP64 DMMA, 2 = int8 digit products), digit planes, smax
end

# ---- batched chains: the examples' `while i <= nchains ... run(job)` loop as one persistent kernel -------------------
mutable struct BatchedMCJob
  handle::Handle
  p::Int
  nchains::Int
  range::BasicMCRange
  bigd::Bool                 # t target of dimension n > 32: the structured large-dimension path (gamc_bigd_*)
end
# diagonal / off-diagonal of the tridiagonal inverse scale matrix inv((ν-2)Σ/ν), Σ_ij = c^|i-j|
function mvt_tridiagonal(m::MvTModel)
  s = m.ν / ((m.ν - 2) * (1 - m.c^2))
  td = fill(s * (1 + m.c^2), m.n); td[1] = td[end] = s
  m.n == 1 && (td[1] = m.ν / (m.ν - 2))
  td, fill(-s * m.c, max(m.n - 1, 0))
end
function mvt_constants(m::MvTModel)
  Σt = [(m.ν - 2) * m.c^abs(i - j) / m.ν for i in 1:m.n, j in 1:m.n]
  Sinv = inv(Σt); Sinv = (Sinv + Sinv') / 2
  logconst = first(LinearAlgebra.logabsdet(Σt))
  lc = lgamma_((m.ν + m.n) / 2) - lgamma_(m.ν / 2) - m.n / 2 * log(m.ν * pi) - logconst / 2
  Sinv, lc
end
lgamma_(x) = first(Base.Math.lgamma_r(x))     # SpecialFunctions.loggamma where available
function BatchedMCJob(model, sampler::GAMC, range::BasicMCRange, v0s::Matrix{Float64}; tuner::GAMCTuner=GAMCTuner(), device::Integer=0,
                      chain_offset::Integer=0, max_factors::Integer=0)
  sampler.update! isa Schedule || error("the batched kernel decides the branch on the device: pass one of the nine schedules")
  h = create(device)
  if model isa MvTModel && model.n > 32
    td, te = mvt_tridiagonal(model)
    check(h, ccall((:gamc_bigd_target_mvt, libgamc), Cint, (Handle, Int32, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}), h, model.n, model.ν, td, te))
    check(h, ccall((:gamc_bigd_set_chain_offset, libgamc), Cint, (Handle, Int64), h, chain_offset))
    cfg = Ref(cconfig(sampler, tuner, range))
    tr = sampler.transform
    check(h, ccall((:gamc_bigd_init, libgamc), Cint, (Handle, Ref{CConfig}, Int32, Ptr{Cdouble}, Int32, Cdouble, Int32),
                   h, cfg, size(v0s, 2), v0s, tr isa Softabs ? 1 : 0, tr isa Softabs ? tr.alpha : 0., max_factors))
    job = BatchedMCJob(h, model.n, size(v0s, 2), range, true)
    finalizer(j -> ccall((:gamc_destroy, libgamc), Cint, (Handle,), j.handle), job)
    return job
  end
  if model isa RVModel
    p = 5 * model.nplanets + 1
    check(h, ccall((:gamc_batched_target_rv, libgamc), Cint, (Handle, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Int32),
                   h, model.times, model.obs, model.sigma_obs, length(model.times), model.nplanets))
  elseif model isa LogisticModel
    N, p = size(model.X)
    check(h, ccall((:gamc_batched_target_logistic, libgamc), Cint, (Handle, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Int64, Int32, Cdouble),
                   h, model.X, N, model.y, N, p, model.λ))
  else
    Sinv, lc = mvt_constants(model); p = model.n
    check(h, ccall((:gamc_batched_target_mvt, libgamc), Cint, (Handle, Int32, Cdouble, Ptr{Cdouble}, Cdouble), h, p, model.ν, Sinv, lc))
  end
  nchains = size(v0s, 2)                       # v0s: p x nchains (column c = initial value of chain c)
  @assert size(v0s, 1) == p
  check(h, ccall((:gamc_batched_set_chain_offset, libgamc), Cint, (Handle, Int64), h, chain_offset))
  cfg = Ref(cconfig(sampler, tuner, range))
  tr = sampler.transform
  check(h, ccall((:gamc_batched_init, libgamc), Cint, (Handle, Ref{CConfig}, Int32, Ptr{Cdouble}, Int32, Cdouble),
                 h, cfg, nchains, v0s, tr isa Softabs ? 1 : 0, tr isa Softabs ? tr.alpha : 0.))
  job = BatchedMCJob(h, p, nchains, range, false)
  finalizer(j -> ccall((:gamc_destroy, libgamc), Cint, (Handle,), j.handle), job)
  job
end
function run!(job::BatchedMCJob; seed::Integer=0, n::Integer=job.range.nsteps)     # device Philox streams keyed by (seed, chain, iteration)
  if job.bigd
    check(job.handle, ccall((:gamc_bigd_run, libgamc), Cint, (Handle, Int64, UInt64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                            job.handle, n, seed, C_NULL, C_NULL, C_NULL, C_NULL))
    check(job.handle, ccall((:gamc_sync, libgamc), Cint, (Handle,), job.handle))
    return job
  end
  check(job.handle, ccall((:gamc_batched_run, libgamc), Cint, (Handle, Int64, UInt64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                          job.handle, n, seed, C_NULL, C_NULL, C_NULL, C_NULL))
  check(job.handle, ccall((:gamc_sync, libgamc), Cint, (Handle,), job.handle))
  job
end
function output(job::BatchedMCJob)
  cap = max(0, job.range.nsteps - job.range.burnin)
  value = Array{Float64}(undef, job.p, cap, job.nchains)          # [p][cap][chain] in column-major = the ABI's [chain][cap][p]
  acc = Matrix{UInt8}(undef, cap, job.nchains)
  if job.bigd
    check(job.handle, ccall((:gamc_bigd_chain_get, libgamc), Cint, (Handle, Int32, Int32, Ptr{Cdouble}, Ptr{UInt8}), job.handle, 0, job.nchains, value, acc))
  else
    check(job.handle, ccall((:gamc_batched_chain_get, libgamc), Cint, (Handle, Int32, Int32, Ptr{Cdouble}, Ptr{UInt8}), job.handle, 0, job.nchains, value, acc))
  end
  [Chain(value[:, :, c], reshape(acc[:, c] .!= 0, 1, cap), Float64[]) for c in 1:job.nchains]
end
function state(job::BatchedMCJob, chain::Integer)
  st = Ref{CState}(); θ = Vector{Float64}(undef, job.p)
  if job.bigd
    check(job.handle, ccall((:gamc_bigd_get_state, libgamc), Cint, (Handle, Int32, Ref{CState}, Ptr{Cdouble}), job.handle, chain - 1, st, θ))
    return st[], θ
  end
  check(job.handle, ccall((:gamc_batched_get_state, libgamc), Cint, (Handle, Int32, Ref{CState}, Ptr{Cdouble}), job.handle, chain - 1, st, θ))
  st[], θ
end

# ---- measurement hooks ----------------------------------------------------------------------------------------------
const KERNEL_FAMILIES = (:loglik, :loglik_grad, :metric, :reduce, :linalg, :allreduce, :land, :smmala_pre, :smmala_post, :am_pre, :am_post, :handover)
profile!(h::Handle, on::Bool) = check(h, ccall((:gamc_profile_enable, libgamc), Cint, (Handle, Int32), h, on ? 1 : 0))
profile_reset!(h::Handle) = check(h, ccall((:gamc_profile_reset, libgamc), Cint, (Handle,), h))
function profile_get(h::Handle)
  out = Dict{Symbol, Tuple{Int64, Float64}}()
  for (k, name) in enumerate(KERNEL_FAMILIES)
    n = Ref{Int64}(0); ms = Ref{Cdouble}(0)
    check(h, ccall((:gamc_profile_get, libgamc), Cint, (Handle, Int32, Ref{Int64}, Ref{Cdouble}), h, k - 1, n, ms))
    out[name] = (n[], ms[])
  end
  out
end
function launch_count(h::Handle)
  n = Ref{Int64}(0)
  check(h, ccall((:gamc_launch_count, libgamc), Cint, (Handle, Ref{Int64}), h, n))
  n[]
end
function probe_fp64_peak(h::Handle)
  a = Ref{Cdouble}(0); b = Ref{Cdouble}(0)
  check(h, ccall((:gamc_probe_fp64_peak, libgamc), Cint, (Handle, Ref{Cdouble}, Ref{Cdouble}), h, a, b))
  a[], b[]                             # (DMMA, DFMA) TFLOP/s of this device
end

end # module
