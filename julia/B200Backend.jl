# B200Backend.jl -- the reference-side binding of libjudi_b200.so.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI (the build image has no Julia).  It is the `ccall` twin of
# JUDI's ten `devito_interface` methods (src/TimeModeling/Modeling/python_interface.jl:30-198) and of
# the `ac.J_adjoint` call in `_multi_src_fg` (src/TimeModeling/Modeling/misfit_fg.jl:69-73).
# Include it from src/JUDI.jl after python_interface.jl and select it with JUDI.set_backend(:b200)
# (or ENV["JUDI_BACKEND"]="b200"); everything above `devito_interface` stays untouched.
module B200Backend

using ..JUDI: Geometry, JUDIOptions, PhysicalParameter, AbstractModel, time_resample,
              _maybe_pad_t0, setup_grid, origin, spacing, nbl, _params

const LIB = get(ENV, "JUDI_B200_LIB", "libjudi_b200.so")
const CTX = Ref{Ptr{Cvoid}}(C_NULL)

struct JudiParam           # judi_param
    kind::Cint; value::Cfloat; data::Ptr{Cfloat}
end
none() = JudiParam(0, 0f0, C_NULL)
param(::Nothing) = none()
param(x::Number) = JudiParam(1, Float32(x), C_NULL)
param(x::Array{Float32}) = JudiParam(2, 0f0, pointer(x))        # Julia arrays are x-fastest: no copy
param(x::PhysicalParameter) = param(x.data)

struct JudiModelDesc       # judi_model_desc
    ndim::Cint; shape::NTuple{3,Cint}; spacing::NTuple{3,Cfloat}; origin::NTuple{3,Cfloat}
    nbl::Cint; space_order::Cint; fs::Cint; dt::Cfloat
    m::JudiParam; rho::JudiParam; b::JudiParam; epsilon::JudiParam; delta::JudiParam
    theta::JudiParam; phi::JudiParam; dm::JudiParam; params_on_device::Cint
end

struct JudiJopts           # judi_jopts
    is_residual::Cint; checkpointing::Cint; n_checkpoints::Cint; t_sub::Cint; ic::Cint
    born_fwd::Cint; nlind::Cint; illum::Cint; snapshot_region::Cint
    ws::Ptr{Cfloat}        # extended-source weights on the padded grid, or C_NULL
    nfreq::Cint; dft_sub::Cint; freq_list::Ptr{Cfloat}   # on-the-fly DFT (options.frequencies)
end

lasterr() = unsafe_string(ccall((:judi_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 || error("judi_b200: " * lasterr())      # `retry` in time_modeling_serial.jl:95 still applies

function ctx()
    if CTX[] == C_NULL
        dev = parse(Int, get(ENV, "JUDI_B200_DEVICE", "0"))   # one Julia worker per GPU
        CTX[] = ccall((:judi_ctx_create, LIB), Ptr{Cvoid}, (Cint,), dev)
        CTX[] == C_NULL && error("judi_b200: " * lasterr())
    end
    CTX[]
end

# Page-lock arrays that outlive many shots (model parameters, observed data, gradient buffers): the
# library then copies them at PCIe speed (measured on a C3 shot: Model 12 -> 3 ms, gradient
# read-back 67 -> 5 ms, 529 -> 471 ms per fwi shot).  Arrays are unpinned by their finalizer.
const PINNED = IdDict{Any,Bool}()
function pin!(a::Array{Float32})
    haskey(PINNED, a) && return a
    check(ccall((:judi_host_register, LIB), Cint, (Ptr{Cvoid}, Csize_t), a, sizeof(a)))
    PINNED[a] = true
    finalizer(x -> ccall((:judi_host_unregister, LIB), Cint, (Ptr{Cvoid},), x), a)
    return a
end
pin!(a) = a

pad3(t, fill) = ntuple(i -> i <= length(t) ? t[i] : fill, 3)
const IC = Dict("as" => 0, "isic" => 1, "fwi" => 2)

# replaces devito_model (src/TimeModeling/Utils/auxiliaryFunctions.jl:26-37)
function b200_model(model::AbstractModel, options::JUDIOptions, dm)
    p = Dict(_params(model))
    foreach(v -> pin!(v isa PhysicalParameter ? v.data : v), values(p))
    get_p(k) = haskey(p, k) ? param(p[k]) : none()
    n = size(model)
    desc = JudiModelDesc(length(n), Cint.(pad3(n, 1)), Cfloat.(pad3(spacing(model), 1)),
                         Cfloat.(pad3(origin(model), 0)), nbl(model), options.space_order,
                         options.free_surface, something(options.dt_comp, 0f0),
                         get_p(:m), get_p(:rho), none(), get_p(:epsilon), get_p(:delta),
                         get_p(:theta), get_p(:phi), param(dm), 0)
    h = GC.@preserve p dm ccall((:judi_model_create, LIB), Ptr{Cvoid},
                                (Ptr{Cvoid}, Ref{JudiModelDesc}), ctx(), desc)
    h == C_NULL && error("judi_b200: " * lasterr())
    if haskey(p, :qp)   # visco-acoustic model (Model(n, d, o, m, rho, qp)); f0 = options.f0 as the reference passes it
        qp = param(p[:qp])
        GC.@preserve p check(ccall((:judi_model_set_visco, LIB), Cint, (Ptr{Cvoid}, Ref{JudiParam}, Cfloat, Cint),
                                   h, qp, options.f0, 0))
    end
    return h
end
critical_dt(h) = ccall((:judi_model_critical_dt, LIB), Cfloat, (Ptr{Cvoid},), h)
function padded_shape(h, ndim)
    out = zeros(Cint, ndim); check(ccall((:judi_model_padded_shape, LIB), Cint, (Ptr{Cvoid}, Ptr{Cint}), h, out)); Tuple(out)
end
destroy(h) = ccall((:judi_model_destroy, LIB), Cvoid, (Ptr{Cvoid},), h)

# d_obs = Pr*F*Ps'*q   (python_interface.jl:30-43 -> interface.py:17)
function b200_interface(h, srcGeometry::Geometry, srcData::Array, recGeometry::Geometry, ::Nothing,
                        ::Nothing, options::JUDIOptions, illum::Bool, fw::Bool, n)
    dtComp = critical_dt(h)
    qIn = _maybe_pad_t0(time_resample(srcData, srcGeometry, dtComp), srcGeometry, recGeometry, dtComp)
    sc = Float32.(setup_grid(srcGeometry, n)); rc = Float32.(setup_grid(recGeometry, n))   # hcat(x[,y],z)
    nt = size(qIn, 1); rec = zeros(Float32, nt, size(rc, 1))
    I = illum ? zeros(Float32, padded_shape(h, length(n))) : nothing
    check(ccall((:judi_forward_rec, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Cint),
                h, fw, size(sc, 1), sc, qIn, nt, size(rc, 1), rc, rec, illum ? I : C_NULL, 0))
    return illum ? (rec, I) : rec
end

# d_lin = J*dm         (python_interface.jl:83-99 -> interface.py:208)
# fw=false (the Jacobian of F'): the adjoint scheme is the forward scheme in reversed time for every
# propagator but the visco-acoustic one, so J(F')*dm = reverse(J(F; reverse(q))*dm) -- reverse qIn
# before the call and the record after it (judi.jl_b200/interface.py:born_rec does the same).
function b200_born(h, srcGeometry, srcData, recGeometry, options::JUDIOptions, illum::Bool, n; fw::Bool=true)
    dtComp = critical_dt(h)
    qIn = _maybe_pad_t0(time_resample(srcData, srcGeometry, dtComp), srcGeometry, recGeometry, dtComp)
    fw || (qIn = reverse(qIn, dims=1))
    sc = Float32.(setup_grid(srcGeometry, n)); rc = Float32.(setup_grid(recGeometry, n))
    nt = size(qIn, 1); rec = zeros(Float32, nt, size(rc, 1))
    check(ccall((:judi_born_rec, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cint, Ptr{Cfloat}, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Cint),
                h, size(sc, 1), sc, qIn, nt, size(rc, 1), rc, IC[options.IC], rec, C_NULL, 0))
    return fw ? rec : reverse(rec, dims=1)
end

# (f, g) of one shot   (misfit_fg.jl:69-73 / python_interface.jl:102-121 -> interface.py:279)
function b200_J_adjoint(h, src_coords, qIn, rec_coords, dIn, options::JUDIOptions, n;
                        is_residual=false, born_fwd=false, nlind=false, illum=false, misfit=nothing,
                        ws=nothing)
    nt = size(qIn, 1)
    freqs = Float32.(options.frequencies)                     # python_interface.jl:112
    g = pin!(zeros(Float32, padded_shape(h, length(n)))); f = Ref{Cfloat}(0)
    Iu = illum ? similar(g) : nothing; Iv = illum ? similar(g) : nothing
    o = JudiJopts(is_residual, options.optimal_checkpointing, 0, options.subsampling_factor,
                  IC[options.IC], born_fwd, nlind, illum, options.sum_padding ? 0 : 1,
                  ws === nothing ? Ptr{Cfloat}(C_NULL) : pointer(ws),
                  length(freqs), options.dft_subsampling_factor[1],
                  isempty(freqs) ? Ptr{Cfloat}(C_NULL) : pointer(freqs))
    # custom misfit: judi_misfit_fn; the default mse (losses.jl:15-19) stays on the device (C_NULL)
    cb = misfit === nothing ? C_NULL :
        @cfunction($((ps, po, pr, nt_, nr_, _) -> begin
            dsyn = unsafe_wrap(Array, ps, (Int(nt_), Int(nr_))); dobs = unsafe_wrap(Array, po, (Int(nt_), Int(nr_)))
            fv, r = misfit(dsyn, dobs); copyto!(unsafe_wrap(Array, pr, (Int(nt_), Int(nr_))), r); Cfloat(fv)
        end), Cfloat, (Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cint, Ptr{Cvoid}))
    GC.@preserve freqs ws check(ccall((:judi_J_adjoint, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ref{JudiJopts},
                 Ptr{Cvoid}, Ptr{Cvoid}, Ref{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Cint),
                h, size(src_coords, 1), src_coords, qIn, nt, size(rec_coords, 1), rec_coords, dIn, o,
                cb, C_NULL, f, g, illum ? Iu : C_NULL, illum ? Iv : C_NULL, 0))
    return illum ? (f[], g, Iu, Iv) : (f[], g)
end

# (f, g) summed over several shots of this worker in ONE call (multi_src_fg!, propagation.jl:93-107)
struct JudiShot            # judi_shot
    nsrc::Cint; src_coords::Ptr{Cfloat}; wavelet::Ptr{Cfloat}; nt::Cint
    nrec::Cint; rec_coords::Ptr{Cfloat}; recin::Ptr{Cfloat}
end
function b200_fg_multishot(h, shots::Vector{<:NTuple{4,Array{Float32}}}, options::JUDIOptions, n;
                           born_fwd=false, nlind=false)
    g = pin!(zeros(Float32, padded_shape(h, length(n)))); f = Ref{Cfloat}(0)
    o = JudiJopts(false, options.optimal_checkpointing, 0, options.subsampling_factor, IC[options.IC],
                  born_fwd, nlind, false, options.sum_padding ? 0 : 1, Ptr{Cfloat}(C_NULL), 0, 1, Ptr{Cfloat}(C_NULL))
    cs = [JudiShot(size(sc, 1), pointer(sc), pointer(q), size(q, 1), size(rc, 1), pointer(rc), pointer(d))
          for (sc, q, rc, d) in shots]
    GC.@preserve shots cs check(ccall((:judi_fg_multishot, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{JudiShot}, Ref{JudiJopts}, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Cfloat}, Ptr{Cfloat}, Cint),
                h, length(cs), cs, o, C_NULL, C_NULL, f, g, 0))
    return f[], g
end

# ---- the sum over workers behind the C-ABI (replaces reduce! / run_and_reduce, distributed.jl:60-95,
# propagation.jl:23-50).  One Julia worker per GPU: worker 1 draws the NCCL id, the 128 bytes travel
# over Distributed (`id = remotecall_fetch(b200_comm_id, 1)`), every worker calls b200_comm_init.
b200_comm_id() = (id = zeros(UInt8, 128); check(ccall((:judi_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id)); id)
b200_comm_init(id::Vector{UInt8}, nranks::Int, rank::Int) =
    check(ccall((:judi_comm_init, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx(), nranks, rank, id))
# multi_src_fg! of this worker's shots + the reduction over all workers: every worker gets the
# physical-grid gradient and the objective of ALL shots (propagation.jl:99-106, misfit_fg.jl:77)
function b200_fg_multishot_reduced(h, shots::Vector{<:NTuple{4,Array{Float32}}}, options::JUDIOptions, n;
                                   born_fwd=false, nlind=false)
    g = pin!(zeros(Float32, n)); f = Ref{Cfloat}(0)
    o = JudiJopts(false, options.optimal_checkpointing, 0, options.subsampling_factor, IC[options.IC],
                  born_fwd, nlind, false, options.sum_padding ? 0 : 1, Ptr{Cfloat}(C_NULL), 0, 1, Ptr{Cfloat}(C_NULL))
    cs = [JudiShot(size(sc, 1), pointer(sc), pointer(q), size(q, 1), size(rc, 1), pointer(rc), pointer(d))
          for (sc, q, rc, d) in shots]
    GC.@preserve shots cs check(ccall((:judi_fg_multishot_reduced, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{JudiShot}, Ref{JudiJopts}, Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ref{Cfloat}, Ptr{Cfloat}, Cint),
                h, length(cs), cs, o, C_NULL, C_NULL, options.sum_padding, f, g, 0))
    return f[], g
end

# ---- extended and wavefield sources (python_interface.jl:58-81, 124-198 -> interface.py:50-204, 244-276)
# `weights` arrive on the physical grid; JUDI zero-pads them (python_interface.jl:129)
function b200_forward_rec_w(h, wpad::Array{Float32}, qIn, rec_coords, fw::Bool)
    nt = size(qIn, 1); rec = zeros(Float32, nt, size(rec_coords, 1))
    check(ccall((:judi_forward_rec_w, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Cint),
                h, fw, wpad, qIn, nt, size(rec_coords, 1), rec_coords, rec, C_NULL, 0))
    return rec
end
function b200_adjoint_w(h, rec_coords, dIn, qIn, fw::Bool, n)
    w = zeros(Float32, padded_shape(h, length(n)))
    check(ccall((:judi_adjoint_w, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Cint),
                h, fw, size(rec_coords, 1), rec_coords, dIn, qIn, size(qIn, 1), w, C_NULL, 0))
    return w
end
function b200_born_rec_w(h, wpad::Array{Float32}, qIn, rec_coords, ic::String)
    nt = size(qIn, 1); rec = zeros(Float32, nt, size(rec_coords, 1))
    check(ccall((:judi_born_rec_w, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cint, Ptr{Cfloat}, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Cint),
                h, wpad, qIn, nt, size(rec_coords, 1), rec_coords, IC[ic], rec, C_NULL, 0))
    return rec
end
# u_src: (padded grid..., nt) column-major = time slowest, as JUDI's judiWavefield data permuted
function b200_forward_wf_src(h, u_src::Array{Float32}, rec_coords, fw::Bool)
    nt = size(u_src)[end]; rec = zeros(Float32, nt, size(rec_coords, 1))
    check(ccall((:judi_forward_wf_src, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Cint),
                h, fw, u_src, nt, size(rec_coords, 1), rec_coords, rec, C_NULL, 0))
    return rec
end
function b200_forward_wf_src_norec(h, u_src::Array{Float32}, fw::Bool)
    u = similar(u_src)
    check(ccall((:judi_forward_wf_src_norec, LIB), Cint,
                (Ptr{Cvoid}, Cint, Ptr{Cfloat}, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Cint),
                h, fw, u_src, size(u_src)[end], u, C_NULL, 0))
    return u
end

# ---- the steps around the propagators, on the device (SURVEY 8f rank 2) ------------------------
# remove_padding(gradient, modelPy.padsizes; true_adjoint) (auxiliaryFunctions.jl:79-89)
function b200_remove_padding(h, g::Array{Float32}, n; true_adjoint::Bool=false)
    out = zeros(Float32, n)
    check(ccall((:judi_remove_padding, LIB), Cint, (Ptr{Cvoid}, Ptr{Cfloat}, Cint, Ptr{Cfloat}, Cint),
                h, g, true_adjoint, out, 0))
    return out
end
# time_resample(data, t_in, t_new) (time_utilities.jl:32-48)
function b200_time_resample(data::Matrix{Float32}, t_in::StepRangeLen, t_new::StepRangeLen, nt_out::Int)
    out = zeros(Float32, nt_out, size(data, 2))
    check(ccall((:judi_time_resample, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cfloat}, Cint, Cint, Cfloat, Cfloat, Ptr{Cfloat}, Cint, Cfloat, Cfloat, Cint),
                ctx(), data, size(data, 1), size(data, 2), first(t_in), step(t_in), out, nt_out,
                first(t_new), step(t_new), 0))
    return out
end

end # module
