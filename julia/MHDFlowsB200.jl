# MHDFlowsB200.jl -- Julia host shim over libmhdflows_b200.so (pure `ccall`; no CUDA.jl, no kernels here).
#
# Same exported names and keyword arguments as MHDFlows.jl for the hot path, so a user script changes
# `using MHDFlows` to `using MHDFlowsB200` (reference: src/MHDFlows.jl:65-88, src/pgen.jl:64-127,
# src/utils/IC.jl:41-109, src/integrator.jl:31-198, src/utils/UserInterface.jl:65-86,
# src/DiagnosticWrapper.jl:14-105).  NOT EXECUTED in the build environment (Julia is not installed there);
# the Python ctypes mirror mhdflows_jl_b200/ is the tested twin of this file -- keep them line-for-line
# (tests/test_abi.py checks that every entry point bound here exists in the header with the same argument count).
module MHDFlowsB200

export Problem, SetUpProblemIC!, stepforward!, TimeIntegrator!, getCFL!, ProbDiagnostic, Diagnostic,
       increment!, CPU, GPU, nothingfunction, spectralline, h_k_sum, h_m_sum,
       N97ForceDriving!, GetN97vars_And_function, SetUpN97!, setforcing!,
       A99ForceDriving!, GetA99vars_And_function, SetUpFk, A99GPU, DivVCorrection!, DivBCorrection!, setvpfield!,
       savefile, Restart!, readMHDFlows, DivFreeSpectraMap, SetUpRandomPhaseIC!,
       NDForceDriving!, GetNDvars_And_function, SetUpND!, stepper_stats, ScaleDecomposition, VectorPotential, CF, SFC

const lib = get(ENV, "MHDFLOWS_B200_LIB", "libmhdflows_b200.so")

struct CPU end
struct GPU; device::Int; end
GPU() = GPU(0)
nothingfunction(args...) = nothing

# mirrors `mhdf_config` in include/mhdflows_b200.h field by field
struct MhdfConfig
  nx::Cint; ny::Cint; nz::Cint
  Lx::Cdouble; Ly::Cdouble; Lz::Cdouble
  nu::Cdouble; eta::Cdouble
  n_nu::Cint
  dt::Cdouble
  physics::Cint; stepper::Cint; dtype::Cint; device::Cint
  rank::Cint; nranks::Cint
  nccl_id::Ptr{Cvoid}
  vp::Cint
  nd::Cint
end

const MHDF_HD, MHDF_MHD, MHDF_EMHD = 0, 1, 2
const MHDF_FRESH, MHDF_STALE = 0, 1

mutable struct Clock{T}          # FourierFlows.Clock, read through the library
  h::Ptr{Cvoid}
end
function _clock(c::Clock)
  t = Ref{Cdouble}(); dt = Ref{Cdouble}(); s = Ref{Clonglong}()
  ccall((:mhdf_get_clock, lib), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Ref{Cdouble}, Ref{Clonglong}), c.h, t, dt, s)
  (t = t[], dt = dt[], step = Int(s[]))
end
Base.getproperty(c::Clock, s::Symbol) = s === :h ? getfield(c, :h) : getproperty(_clock(c), s)
function Base.setproperty!(c::Clock, s::Symbol, v)
  h = getfield(c, :h)
  if s === :dt
    check(h, ccall((:mhdf_set_dt, lib), Cint, (Ptr{Cvoid}, Cdouble), h, v))
  elseif s === :t
    check(h, ccall((:mhdf_set_clock, lib), Cint, (Ptr{Cvoid}, Cdouble, Clonglong), h, v, _clock(c).step))
  elseif s === :step
    check(h, ccall((:mhdf_set_clock, lib), Cint, (Ptr{Cvoid}, Cdouble, Clonglong), h, _clock(c).t, v))
  end
end

struct Flag; b::Bool; e::Bool; vp::Bool; c::Bool; s::Bool; end
struct Grid{T}; nx::Int; ny::Int; nz::Int; Lx::Float64; Ly::Float64; Lz::Float64; dx::Float64; dy::Float64; dz::Float64; end
struct Params; ν::Float64; η::Float64; nν::Int; nη::Int; end

mutable struct Vars; usr_vars::Any; end

mutable struct MHDFlowsProblem{T}
  h::Ptr{Cvoid}
  clock::Clock{T}
  grid::Grid{T}
  params::Params
  flag::Flag
  Nl::Int
  usr_func::Vector{Any}
  calcF::Any            # an arbitrary closure (host-callback path) or nothing
  vars::Vars            # vars.usr_vars as in the reference (datastructure.jl:5-35); the fields are read with realfield()
end

function check(h, code)
  code == 0 && return nothing
  msg = unsafe_string(ccall((:mhdf_last_error, lib), Cstring, (Ptr{Cvoid},), h))
  error("mhdflows_b200 error $code: $msg")       # the reference reports through error(...), pgen.jl:99,104
end

"""
    Problem(dev; nx, ny, nz, Lx, Ly, Lz, dt, ν, nν, η, nη, B_field, EMHD, stepper, T, ...)   (pgen.jl:64-127)
"""
function Problem(dev; nx = 64, ny = nx, nz = nx, Lx = 2π, Ly = Lx, Lz = Lx, cₛ = 0.0, dt = 0.0,
                 ν = 0.0, nν = 0, η = 0.0, nη = 0, B_field = false, EMHD = false, Compressibility = false,
                 Shear = false, VP_method = false, Dye_Module = false, stepper = "RK4", calcF = nothingfunction,
                 T = Float32, aliased_fraction = 1/3, usr_vars = [], usr_params = [], usr_func = [])
  dev isa CPU && error("this build is the B200 path only: Problem(GPU(); ...)")
  cₛ == 0.0 && Compressibility && error("You should define cₛ")
  Shear && error("Shear haven't fully implemented yet!")
  (Compressibility || Dye_Module) && error("outside the B200 hot path")
  VP_method && EMHD && error("VP_method: the EMHD equation has no volume-penalisation terms")
  builtin = (calcF === nothingfunction || calcF === N97ForceDriving! || calcF === A99ForceDriving! ||
             calcF === A99GPU.A99ForceDriving! || calcF === NDForceDriving!)
  stepper in ("RK4", "LSRK54", "HM89") || error("stepper must be \"RK4\", \"LSRK54\" or (EMHD) \"HM89\" on the B200 path")
  stepper == "HM89" && !EMHD && error("stepper \"HM89\" exists for EMHD problems only (Problems.jl:124-126)")
  stepper == "HM89" && calcF !== nothingfunction && error("HM89 with a forcing function is not supported (HM89.jl:182-196)")
  physics = EMHD ? MHDF_EMHD : (B_field ? MHDF_MHD : MHDF_HD)
  cfg = MhdfConfig(nx, ny, nz, Lx, Ly, Lz, ν, η, nν, dt, physics, stepper == "RK4" ? 0 : (stepper == "LSRK54" ? 1 : 2),
                   T === Float32 ? 0 : 1, dev.device, 0, 1, C_NULL, VP_method ? 1 : 0,
                   calcF === NDForceDriving! ? 1 : 0)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  code = ccall((:mhdf_create, lib), Cint, (Ref{MhdfConfig}, Ref{Ptr{Cvoid}}), cfg, h)
  code == 0 || check(C_NULL, code)
  prob = MHDFlowsProblem{T}(h[], Clock{T}(h[]), Grid{T}(nx, ny, nz, Lx, Ly, Lz, Lx/nx, Ly/ny, Lz/nz),
                            Params(ν, η, nν, 0), Flag(B_field, EMHD, VP_method, false, false),
                            physics == MHDF_MHD ? 6 : 3, isempty(usr_func) ? Any[nothingfunction] : collect(Any, usr_func),
                            builtin ? nothing : calcF, Vars(usr_vars))
  builtin || installcalcF!(prob)
  finalizer(p -> ccall((:mhdf_destroy, lib), Cint, (Ptr{Cvoid},), p.h), prob)
  return prob
end

fieldid(prob, s::Symbol) = prob.flag.e ? Dict(:bx=>0, :by=>1, :bz=>2)[s] :
                           Dict(:ux=>0, :uy=>1, :uz=>2, :bx=>3, :by=>4, :bz=>5)[s]

# Field arguments may live on the host (Array) or on the device (CuArray): the library copies with cudaMemcpyDefault
# (include/mhdflows_b200.h), so a device array is passed by its device pointer and never bounces through the host.
_fieldarg(::Type{T}, A::Array) where {T} = Array{T,3}(A)
_fieldarg(::Type{T}, A) where {T} = (eltype(A) === T || error("device fields must already have the problem's element type"); A)
_fieldptr(A::Array) = Ptr{Cvoid}(pointer(A))
_fieldptr(A) = Ptr{Cvoid}(UInt(pointer(A)))          # CuArray: CuPtr -> raw address (unified addressing)

# ---- arbitrary calcF! closures (pgen.jl:231-234) through the library's host callback -----------------------------------------
# The library calls back at the beginning of every right-hand-side evaluation; `sol` of that evaluation is read with which = 2
# (MHDF_STAGE), what calcF! adds to a zero N is uploaded and added to the right-hand side.  The device-resident forcings of the
# reference (N97, A99, negative damping) do not take this path.
function _forcing_trampoline(user::Ptr{Cvoid}, t::Cdouble)::Cint
  prob = unsafe_pointer_to_objref(user)
  T = typeof(prob).parameters[1]; g = prob.grid
  try
    S = Array{Complex{T},4}(undef, g.nx ÷ 2 + 1, g.ny, g.nz, prob.Nl)
    for i in 1:prob.Nl
      A = Array{Complex{T},3}(undef, g.nx ÷ 2 + 1, g.ny, g.nz)
      check(prob.h, ccall((:mhdf_get_spectral, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}), prob.h, i - 1, 2, A))
      S[:, :, :, i] .= A
    end
    N = zero(S)
    prob.calcF(N, S, T(t), prob.clock, prob.vars, prob.params, prob.grid)
    for i in 1:prob.Nl
      A = N[:, :, :, i]
      check(prob.h, ccall((:mhdf_set_forcing_spectral, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), prob.h, i - 1, any(!iszero, A) ? pointer(A) : C_NULL))
    end
    return Cint(0)
  catch err
    @error "calcF! failed" exception = err
    return Cint(-1)
  end
end
function installcalcF!(prob)
  fn = @cfunction(_forcing_trampoline, Cint, (Ptr{Cvoid}, Cdouble))
  check(prob.h, ccall((:mhdf_set_forcing_callback, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), prob.h, fn, pointer_from_objref(prob)))
end

"SetUpProblemIC!(prob; ux, uy, uz, bx, by, bz)   (utils/IC.jl:41-109)"
function SetUpProblemIC!(prob; ux = [], uy = [], uz = [], bx = [], by = [], bz = [])
  T = typeof(prob).parameters[1]
  function put(s, A)
    A == [] && return nothing
    B = _fieldarg(T, A)
    GC.@preserve B check(prob.h, ccall((:mhdf_set_real, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), prob.h, fieldid(prob, s), _fieldptr(B)))
  end
  if !prob.flag.e; put(:ux, ux); put(:uy, uy); put(:uz, uz); end
  if prob.flag.b;  put(:bx, bx); put(:by, by); put(:bz, bz); end
  return nothing
end

"vars.ux etc.: real-space view; `stale = true` reproduces the reference's vars (c2r of the last stage input)"
function realfield(prob, s::Symbol; stale = true)
  T = typeof(prob).parameters[1]; g = prob.grid
  A = Array{T,3}(undef, g.nx, g.ny, g.nz)
  check(prob.h, ccall((:mhdf_get_real, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}), prob.h, fieldid(prob, s), stale ? 1 : 0, A))
  A
end
"prob.sol[:, :, :, i] (dealiased modes read back as zero)"
function sol(prob, i::Int)
  T = typeof(prob).parameters[1]; g = prob.grid
  A = Array{Complex{T},3}(undef, g.nx ÷ 2 + 1, g.ny, g.nz)
  check(prob.h, ccall((:mhdf_get_spectral, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}), prob.h, i - 1, 0, A))
  A
end

"constant real-space forcing of one field: the calcF! hook for time-independent forcings (pgen.jl:231-234)"
function setforcing!(prob, s::Symbol, F)
  T = typeof(prob).parameters[1]
  check(prob.h, ccall((:mhdf_set_forcing, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), prob.h, fieldid(prob, s), Array{T,3}(F)))
end
"N97 Taylor-Green forcing (pgen/TaylorGreenDynamo.jl:12-40): pass N97ForceDriving! as calcF, then SetUpN97!(prob; F0, kf)"
N97ForceDriving!(args...) = error("N97ForceDriving! is applied inside the library")
GetN97vars_And_function(dev, nx, ny, nz; T = Float32) = ([], N97ForceDriving!)
function SetUpN97!(prob; F0 = 1, kf = 2)
  g = prob.grid
  x = reshape([-g.Lx/2 + (i-1)*g.dx for i in 1:g.nx], (g.nx, 1, 1))
  y = reshape([-g.Ly/2 + (i-1)*g.dy for i in 1:g.ny], (1, g.ny, 1))
  z = reshape([-g.Lz/2 + (i-1)*g.dz for i in 1:g.nz], (1, 1, g.nz))
  setforcing!(prob, :ux, @. F0 *  sin(kf*x) * cos(kf*y) * cos(kf*z))
  setforcing!(prob, :uy, @. F0 * -cos(kf*x) * sin(kf*y) * cos(kf*z))
  nothing
end

# mirrors `mhdf_a99` in include/mhdflows_b200.h
struct MhdfA99
  variant::Cint
  amp::Cdouble; kf::Cdouble; sigma2::Cdouble; b::Cdouble
  seed::Culonglong; call::Culonglong
end
"A99_vars (pgen/A99ForceDriving.jl:5-16): A, b are the user knobs; the spectral tables are evaluated per mode on the device"
mutable struct A99_vars{T}
  A::T; b::T; σ²::T; kf::T
  Fk_A::Float64       # normalisation inside the Fk table (SetUpFk)
  seed::UInt64
  variant::Int
end
A99ForceDriving!(args...) = error("A99ForceDriving! is applied inside the library")
"GetA99vars_And_function(dev, nx, ny, nz; T)   (pgen/A99ForceDriving.jl:18-31)"
GetA99vars_And_function(dev, nx, ny, nz; T = Float32, C = false, seed = 0) =
  C ? error("A99ForceDriving_Compressible! is outside the B200 hot path") :
      (A99_vars{T}(1, 1, 1, 1, NaN, seed, 1), A99ForceDriving!)
function _a99_integral(g, kf, σ²)
  tot = 0.0
  for iz in 0:g.nz-1, iy in 0:g.ny-1, ix in 0:g.nx÷2
    ky = (iy < g.ny ÷ 2 ? iy : iy - g.ny) * 2π / g.Ly
    kz = (iz < g.nz ÷ 2 ? iz : iz - g.nz) * 2π / g.Lz
    k = sqrt((ix * 2π / g.Lx)^2 + ky^2 + kz^2)
    tot += exp(-(k - kf)^2 / σ²) / (k + 1)^2
  end
  tot
end
function _push_a99(prob, uv::A99_vars)
  calls = Ref{Culonglong}(0)
  check(prob.h, ccall((:mhdf_forcing_a99_calls, lib), Cint, (Ptr{Cvoid}, Ref{Culonglong}), prob.h, calls))
  a = MhdfA99(uv.variant, Float64(uv.A) * uv.Fk_A, uv.kf, uv.σ², uv.b, uv.seed, calls[])
  check(prob.h, ccall((:mhdf_set_forcing_a99, lib), Cint, (Ptr{Cvoid}, Ref{MhdfA99}), prob.h, a))
end
"SetUpFk(prob; kf, P, σ²)   (pgen/A99ForceDriving.jl:93-127)"
function SetUpFk(prob; kf = 2, P = 1, σ² = 1)
  g = prob.grid
  uv = prob.vars.usr_vars::A99_vars
  uv.Fk_A = sqrt(P * 3 * (g.Lx/g.dx) * (g.Ly/g.dy) * (g.Lz/g.dz) / _a99_integral(g, kf, σ²) * (1/g.dx/g.dy/g.dz))
  uv.kf = kf; uv.σ² = σ²
  _push_a99(prob, uv)
end
"module A99GPU (pgen/A99ForceDriving_GPU.jl): own basis vectors, real Φ, clipped gᵢ, Im N = 0 on the kr = 0 plane"
module A99GPU
  import ..A99_vars, .._a99_integral, .._push_a99
  A99ForceDriving!(args...) = error("A99GPU.A99ForceDriving! is applied inside the library")
  GetA99vars_And_function(dev, nx, ny, nz; T = Float32, seed = 0) =
    (A99_vars{T}(1, 1, 1, 1, NaN, seed, 2), A99ForceDriving!, SetUpFk!)
  function SetUpFk!(prob; kf = 2.0, P = 1.0, σ = 1.0, b = 1.0)
    g = prob.grid
    uv = prob.vars.usr_vars::A99_vars
    A = sqrt(P * 3 * (g.Lx/g.dx) * (g.Ly/g.dy) * (g.Lz/g.dz) / _a99_integral(g, kf, σ^2) * (1/g.dx/g.dy/g.dz))
    uv.A = A; uv.σ² = σ^2; uv.b = b
    uv.kf = b                  # sic (pgen/A99ForceDriving_GPU.jl:45)
    uv.Fk_A = Float64(uv.A)    # the kernel multiplies A twice (:89, :106)
    _push_a99(prob, uv)
  end
end

"ND_vars / NDForceDriving! / SetUpND! (pgen/NegativeDamping.jl:8-51): negative-damping forcing F_i = f_i u_i, applied inside the library"
mutable struct ND_vars; P::Float64; end
NDForceDriving!(args...) = error("NDForceDriving! is applied inside the library")
GetNDvars_And_function(dev, nx, ny, nz; T = Float32) = (ND_vars(0.0), NDForceDriving!)
function SetUpND!(prob, P, fx, fy, fz)
  T = typeof(prob).parameters[1]
  a, b, c = _fieldarg(T, fx), _fieldarg(T, fy), _fieldarg(T, fz)
  GC.@preserve a b c check(prob.h, ccall((:mhdf_set_forcing_nd, lib), Cint, (Ptr{Cvoid}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                                         prob.h, P, _fieldptr(a), _fieldptr(b), _fieldptr(c)))
  prob.vars.usr_vars.P = P
  nothing
end

"params.χ, U₀x … B₀z of a VP_method problem: setvpfield!(prob, :χ, mask)   (datastructure.jl:80-81,94-95; IC.jl:93-106)"
function setvpfield!(prob, s::Symbol, A)
  T = typeof(prob).parameters[1]
  which = Dict(:χ=>0, :U₀x=>1, :U₀y=>2, :U₀z=>3, :B₀x=>4, :B₀y=>5, :B₀z=>6)[s]
  check(prob.h, ccall((:mhdf_set_vp_field, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), prob.h, which, Array{T,3}(A)))
end

"DivVCorrection!(prob) / DivBCorrection!(prob)   (Solver/VPSolver.jl:61-137)"
DivVCorrection!(prob) = check(prob.h, ccall((:mhdf_div_correction, lib), Cint, (Ptr{Cvoid}, Cint), prob.h, 0))
DivBCorrection!(prob) = check(prob.h, ccall((:mhdf_div_correction, lib), Cint, (Ptr{Cvoid}, Cint), prob.h, 1))

# ---- on-device analysis of the state (utils/MHDAnalysis.jl:54-82, 129-174; utils/TurbStatTool.jl:67-120) ----------------------
# The reference applies these to arrays a script hands over (usually prob.vars.*); here they act on the problem's own fields:
# which = 1 the stale vars.* (default), 0 the true state.
_group(g) = g === :u ? Cint(0) : Cint(1)
function _three(prob::MHDFlowsProblem{T}, f) where T
  out = Array{T}(undef, prob.grid.nx, prob.grid.ny, prob.grid.nz, 3)
  check(prob.h, f(out))
  (out[:, :, :, 1], out[:, :, :, 2], out[:, :, :, 3])
end
"ScaleDecomposition(prob; group = :b, kf = [1, 5])   (MHDAnalysis.jl:54-82): the components restricted to kf[1] <= |k| <= kf[2]"
ScaleDecomposition(prob; group = :b, kf = [1, 5], which = 1) = _three(prob, out ->
  ccall((:mhdf_scale_decomposition, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Cdouble, Ptr{Cvoid}), prob.h, _group(group), which, minimum(kf), maximum(kf), out))
"VectorPotential(prob)   (MHDAnalysis.jl:129-174): a with curl a = b in the Coulomb gauge"
VectorPotential(prob; which = 1) = _three(prob, out ->
  ccall((:mhdf_vector_potential, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), prob.h, which, out))
"CF(prob; group)   (TurbStatTool.jl:67): fftshift(real(ifft(abs.(fft(V)).^2))) of the three components"
function CF(prob; group = :b, which = 1)
  c = _three(prob, out -> ccall((:mhdf_correlation, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}), prob.h, _group(group), which, out))
  sh = (prob.grid.nx ÷ 2, prob.grid.ny ÷ 2, prob.grid.nz ÷ 2)
  map(x -> circshift(x, sh), c)                       # fftshift
end
"SFC(prob; group)   (TurbStatTool.jl:72): 2 (mean(V) - CF(V)) per component, as written"
function SFC(prob::MHDFlowsProblem{T}; group = :b, which = 1) where T
  f0 = (group === :u || prob.flag.e) ? 0 : 3
  n3 = prob.grid.nx * prob.grid.ny * prob.grid.nz
  g = prob.grid
  m = map(1:3) do i                                    # mean(V) = V^(k = 0) / N^3
    A = Array{Complex{T},3}(undef, g.nx ÷ 2 + 1, g.ny, g.nz)
    check(prob.h, ccall((:mhdf_get_spectral, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}), prob.h, f0 + i - 1, which, A))
    T(real(A[1, 1, 1]) / n3)
  end
  c = CF(prob; group = group, which = which)
  ntuple(i -> 2 .* (m[i] .- c[i]), 3)
end

"stepforward!(prob) == stepforward!(prob.sol, prob.clock, prob.timestepper, prob.eqn, prob.vars, prob.params, prob.grid)"
stepforward!(prob, n::Int = 1) = check(prob.h, ccall((:mhdf_step, lib), Cint, (Ptr{Cvoid}, Cint), prob.h, n))

"HM89TimeStepper: (fixed-point iterations of the last step, its last error norm max |Bⁿ - B¹|)   (timestepper/HM89.jl:61-84)"
function stepper_stats(prob)
  it = Ref{Clonglong}(); ε = Ref{Cdouble}()
  check(prob.h, ccall((:mhdf_stepper_stats, lib), Cint, (Ptr{Cvoid}, Ref{Clonglong}, Ref{Cdouble}), prob.h, it, ε))
  (it[], ε[])
end

"getCFL!(prob, t_diff; Coef)   (integrator.jl:158-198)"
function getCFL!(prob, t_diff; Coef = 0.3)
  dt = Ref{Cdouble}()
  check(prob.h, ccall((:mhdf_cfl_dt, lib), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Ref{Cdouble}), prob.h, Coef, t_diff, dt))
  dt[]
end

"ProbDiagnostic(prob)   (utils/UserInterface.jl:65-86)"
function ProbDiagnostic(prob)
  KE = Ref{Cdouble}(); ME = Ref{Cdouble}()
  code = ccall((:mhdf_energy, lib), Cint, (Ptr{Cvoid}, Cint, Ref{Cdouble}, Ref{Cdouble}), prob.h, MHDF_STALE, KE, ME)
  code == -4 && error("detected NaN! Quit the simulation right now.")
  check(prob.h, code)
  ke, me = round(KE[], sigdigits = 3), round(ME[], sigdigits = 3)
  prob.flag.e ? me : (prob.flag.b ? (ke, me) : ke)
end

# Diagnostic (DiagnosticWrapper.jl:14-105)
mutable struct Diagnostic{T}
  calc::Function; prob; data::Vector{T}; t::Vector{Float64}; steps::Vector{Int}; freq::Int; i::Int
end
function Diagnostic(calc, prob; freq = 1, nsteps = 100, ndata = ceil(Int, (nsteps + 1) / freq))
  v = calc(prob)
  d = Diagnostic{typeof(v)}(calc, prob, Vector{typeof(v)}(undef, ndata), zeros(ndata), zeros(Int, ndata), freq, 1)
  d.data[1] = v; d.t[1] = prob.clock.t; d.steps[1] = prob.clock.step
  d
end
function update!(d::Diagnostic, i)
  if i > length(d.steps)
    n = length(d.steps); resize!(d.data, 2n); resize!(d.t, 2n); resize!(d.steps, 2n)
  end
  d.data[i] = d.calc(d.prob); d.t[i] = d.prob.clock.t; d.steps[i] = d.prob.clock.step; d.i = i
  nothing
end
increment!(d::Diagnostic) = (d.prob.clock.step % d.freq == 0 && update!(d, d.i + 1); nothing)
increment!(ds::AbstractVector) = (foreach(increment!, ds); nothing)

# ---- checkpoints in the reference's on-disk format (integrator.jl:208-288, utils/IC.jl:245-257) ----------------------------
# HDF5.jl is the reference's own dependency (Project.toml); it is loaded lazily so that the shim itself needs no package.
_h5() = (isdefined(Main, :HDF5) ? getfield(Main, :HDF5) : Base.require(Base.PkgId(Base.UUID("f67ccb44-e63f-5c2f-98bd-6dc0ccc4ba2f"), "HDF5")))
const _UNAMES = ((:ux, "i_velocity"), (:uy, "j_velocity"), (:uz, "k_velocity"))
const _BNAMES = ((:bx, "i_mag_field"), (:by, "j_mag_field"), (:bz, "k_mag_field"))
"savefile(prob, file_number; file_path_and_name)   (integrator.jl:259-288): the stale vars and the time"
function savefile(prob, file_number; file_path_and_name = "")
  H = _h5()
  name = file_path_and_name * "_t_" * lpad(string(file_number), 4, "0") * ".h5"
  H.h5open(name, "w") do fw
    if !prob.flag.e; for (s, ds) in _UNAMES; write(fw, ds, realfield(prob, s)); end; end
    if prob.flag.b;  for (s, ds) in _BNAMES; write(fw, ds, realfield(prob, s)); end; end
    write(fw, "time", prob.clock.t)
  end
  name
end
"readMHDFlows(FileName)   (utils/IC.jl:245-257)"
function readMHDFlows(FileName)
  H = _h5()
  H.h5open(FileName, "r") do f
    Dict(k => read(f, k) for k in keys(f))
  end
end
"Restart!(prob, file_path_and_name)   (integrator.jl:208-257): fields into vars / sol, clock.t restored"
function Restart!(prob, file_path_and_name)
  d = readMHDFlows(file_path_and_name)
  if !prob.flag.e; SetUpProblemIC!(prob; ux = d["i_velocity"], uy = d["j_velocity"], uz = d["k_velocity"]); end
  if prob.flag.b;  SetUpProblemIC!(prob; bx = d["i_mag_field"], by = d["j_mag_field"], bz = d["k_mag_field"]); end
  prob.clock.t = d["time"]
  nothing
end

"SetUpRandomPhaseIC!(prob; seed_u, seed_b, k_peak, P, k0): DivFreeSpectraMap + SetUpProblemIC! without leaving the device"
function SetUpRandomPhaseIC!(prob; seed_u = nothing, seed_b = nothing, k_peak = 0.0, P = 1, k0 = -5/3/2)
  rp(g, seed) = check(prob.h, ccall((:mhdf_set_random_phase, lib), Cint, (Ptr{Cvoid}, Cint, Culonglong, Cdouble, Cdouble, Cdouble),
                                    prob.h, g, seed, k0, P, k_peak))
  seed_u !== nothing && !prob.flag.e && rp(0, seed_u)
  seed_b !== nothing && prob.flag.b && rp(1, seed_b)
  nothing
end
"DivFreeSpectraMap(Nx, Ny, Nz; Lx, P, k0, b, T, k_peak, seed) -> Fx, Fy, Fz   (utils/IC.jl:122-179), built on the device"
function DivFreeSpectraMap(Nx::Int, Ny::Int, Nz::Int; Lx = 2π, dev = GPU(), P = 1, k0 = -5/3/2, b = 1, T = Float32, k_peak = 0.0, seed = 0)
  p = Problem(dev; nx = Nx, ny = Ny, nz = Nz, Lx = Lx, T = T)
  SetUpRandomPhaseIC!(p; seed_u = seed, k_peak = k_peak, P = P, k0 = k0)
  F = (realfield(p, :ux; stale = false), realfield(p, :uy; stale = false), realfield(p, :uz; stale = false))
  finalize(p)
  F
end

"TimeIntegrator!(prob, t₀, N₀; usr_dt, CFL_Coef, diags, save, save_loc, filename, file_number, dump_dt, ...)   (integrator.jl:31-156)"
function TimeIntegrator!(prob, t₀::Number, N₀::Int; usr_dt = 0.0, CFL_Coef = 0.25, CFL_function = nothingfunction,
                         diags = [], dynamic_dashboard = true, loop_number = 100, save = false,
                         save_loc = "", filename = "", file_number = 0, dump_dt = 0, kwargs...)
  file_path_and_name = ""
  if save                                           # integrator.jl:44-51
    (length(save_loc) == 0 || length(filename) == 0 || dump_dt == 0) &&
      error("Save Function Turned ON but save_loc/filename/dump_dt is not declared!\n")
    file_path_and_name = save_loc * filename
    savefile(prob, file_number; file_path_and_name = file_path_and_name)
    file_number += 1
  end
  t_next_save = prob.clock.t + dump_dt              # integrator.jl:75
  updateCFL! = CFL_function === nothingfunction ? getCFL! : CFL_function
  p = prob.params
  vi = prob.flag.b ? (prob.flag.e ? p.η : max(p.ν, p.η)) : p.ν
  nv = prob.flag.b ? (prob.flag.e ? p.nη : max(p.nν, p.nη)) : p.nν
  dl = min(prob.grid.dx, prob.grid.dy, prob.grid.dz)
  t_diff = nv > 1 ? CFL_Coef * dl^nv / vi : CFL_Coef * dl^2 / vi
  prob.clock.step = 0
  usr_dt != 0.0 && (prob.clock.dt = usr_dt)
  if prob.flag.vp                                   # integrator.jl:85-88
    DivVCorrection!(prob); prob.flag.b && DivBCorrection!(prob)
  end
  time = @elapsed while (N₀ >= prob.clock.step) && (t₀ >= prob.clock.t)
    usr_dt == 0.0 && updateCFL!(prob, t_diff; Coef = CFL_Coef)
    stepforward!(prob)
    increment!(diags)
    if prob.flag.vp                                 # integrator.jl:118-122
      DivVCorrection!(prob); prob.flag.b && DivBCorrection!(prob)
    end
    for foo! in prob.usr_func; foo!(prob); end
    if save && prob.clock.t >= t_next_save          # integrator.jl:136-141
      ProbDiagnostic(prob)
      savefile(prob, file_number; file_path_and_name = file_path_and_name)
      t_next_save += dump_dt
      file_number += 1
    end
  end
  n = prob.grid.nx * prob.grid.ny * prob.grid.nz
  print("Total CPU/GPU time run = $(round(time, digits = 3)) s, zone update per second = $(round(prob.clock.step * n / time, digits = 3)) \n")
  nothing
end

"spectralline of a state field (utils/MHDAnalysis.jl:237-255)"
function spectralline(prob, s::Symbol)
  g = prob.grid
  krmax = round(Int, sqrt((g.nx ÷ 2 * 2π / g.Lx)^2 + (g.ny ÷ 2 * 2π / g.Ly)^2 + (g.nz ÷ 2 * 2π / g.Lz)^2) + 1)
  Pk = zeros(Cdouble, krmax)
  check(prob.h, ccall((:mhdf_spectrum, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cint), prob.h, fieldid(prob, s), Pk, krmax))
  Pk, [Pk[r] > 0 ? r : 0 for r in 1:krmax]
end

function _hel(prob)
  a = Ref{Cdouble}(); b = Ref{Cdouble}(); c = Ref{Cdouble}()
  check(prob.h, ccall((:mhdf_helicity, lib), Cint, (Ptr{Cvoid}, Ref{Cdouble}, Ref{Cdouble}, Ref{Cdouble}), prob.h, a, b, c))
  a[], b[], c[]
end
"sum(h_k(ux,uy,uz)) and sum(h_m(bx,by,bz)) of the current state (utils/MHDAnalysis.jl:94-117)"
h_k_sum(prob) = _hel(prob)[1]
h_m_sum(prob) = _hel(prob)[2]

end # module
