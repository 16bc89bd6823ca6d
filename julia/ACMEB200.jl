# ACMEB200.jl -- the binding a maintainer of HSU-ANT/ACME.jl would add to drive the
# B200 library (include/acmeb200.h) from the reference's own API.
#
# NOT RUNNABLE IN THIS REPOSITORY'S IMAGE (no Julia toolchain); it documents the intended
# integration.  It keeps `@circuit` / `DiscreteModel` / `run!` untouched: the model is derived by
# ACME exactly as before, then `BatchRunner(model; batch, params)` uploads the derived matrices and
# `run!(runner, Y, U)` replaces the per-sample loop of src/ACME.jl:650-715.
module ACMEB200

using ACME
using ACME: DiscreteModel, nx, nu, ny, nn, np, nq

const libacmeb200 = get(ENV, "ACMEB200_LIB", "libacmeb200.so")

# ---- mirrors of the C structs (include/acmeb200.h)
struct CArray;  ptr::Ptr{Float64}; stride::Int64; end
struct CElem;   kind::Int32; q_offset::Int32; param_offset::Int32; nparam::Int32; end
struct CCache
    n_points::Int32; n_columns::Int32
    cut_dim::Ptr{Int32}; cut_val::Ptr{Float64}; ps_idx::Ptr{Int32}; ps::Ptr{Float64}; zs::Ptr{Float64}
end
struct CSubDesc
    nn::Int32; nq::Int32; np::Int32; nelem::Int32
    dq::CArray; eq::CArray; fqprev::CArray; pexp::CArray; q0::CArray; fq::CArray; init_z::CArray
    elems::Ptr{CElem}; params::CArray; nparams::Int32; reserved::Int32; cache::CCache
end
struct CModelDesc
    abi_version::Int32; nx::Int32; nu::Int32; ny::Int32; nsub::Int32; solver::Int32; maxiter::Int32
    cache_capacity::Int32; tol::Float64
    a::CArray; b::CArray; c::CArray; x0::CArray; dy::CArray; ey::CArray; fy::CArray; y0::CArray
    subs::Ptr{CSubDesc}
end

const ELEM_DIODE, ELEM_BJT, ELEM_POT, ELEM_MOSFET, ELEM_OPAMP_TANH, ELEM_JA = Int32.(1:6)

check(rc) = rc == 0 || error(unsafe_string(ccall((:acmeb200_last_error, libacmeb200), Cstring, ())))

"""
Element table of one sub-problem: walks the closures ACME built
(`model.nonlinear_eq_funcs[i]` captures `circ_nl_func::CircuitNLFunc`, src/ACME.jl:177;
each entry of `.fs` captures `q_indices` and `nleqfunc`, src/circuit.jl:76-80) and classifies
every element closure by its captured variables (src/elements.jl: diode `(is, η)` :236-244,
bjt `βf…` :309-401, potentiometer `(r,)` :20-30, mosfet `vt, polarity` :448-479,
tanh op-amp `(gain, scale)` :537-546, Jiles-Atherton `Ms…` :104-129).
"""
function element_table(nleq)
    cnl = getfield(nleq, :circ_nl_func)
    elems = CElem[]; params = Float64[]
    for f in cnl.fs
        qidx = getfield(f, :q_indices); law = getfield(f, :nleqfunc)
        names = fieldnames(typeof(law))
        kind, p = if names == (:is, :η)
            ELEM_DIODE, Float64[law.is, law.η]
        elseif :βf in names
            ELEM_BJT, Float64[law.ise, law.isc, law.ηe, law.ηc, law.βf, law.βr, law.ile, law.ilc,
                              law.ηel, law.ηcl, law.vaf, law.var, law.ikf, law.ikr]
        elseif names == (:r,)
            ELEM_POT, Float64[law.r]
        elseif :polarity in names && :vt in names
            vt = collect(Float64, law.vt); α = collect(Float64, law.α)
            ELEM_MOSFET, Float64[law.polarity, law.λ, length(vt), length(α),
                                 vcat(vt, zeros(4 - length(vt)))..., vcat(α, zeros(4 - length(α)))...]
        elseif names == (:gain, :scale)
            ELEM_OPAMP_TANH, Float64[law.gain, law.scale]
        elseif :Ms in names
            ELEM_JA, Float64[law.Ms, law.a, law.α, law.c, law.k]
        else
            error("unsupported non-linear element closure $(typeof(law))")
        end
        push!(elems, CElem(kind, first(qidx) - 1, length(params), length(p)))
        append!(params, p)
    end
    return elems, params
end

mutable struct BatchRunner
    model::DiscreteModel
    batch::Int
    handle::Ptr{Cvoid}
end

shared(a::Array{Float64}) = CArray(pointer(a), 0)
perinst(a::Array{Float64}, len) = CArray(pointer(a), len)

"""
    BatchRunner(model; batch=1, params=nothing)

`params[i]` is an `nparams × batch` matrix of element parameters for sub-problem `i`
(parameter sweeps); `nothing` shares the model's own parameters.
"""
function BatchRunner(model::DiscreteModel; batch::Integer=1, params=nothing, solver::Integer=2)
    nsub = length(model.solvers)
    keep = Any[]                      # everything the descriptor points into
    subs = Vector{CSubDesc}(undef, nsub)
    for i in 1:nsub
        elems, p = element_table(model.nonlinear_eq_funcs[i])
        pm = params === nothing || params[i] === nothing ? p : Matrix{Float64}(params[i])
        base = model.solvers[i]      # HomotopySolver{CachingSolver{SimpleSolver}} (src/solvers.jl:247, 319)
        while hasproperty(base, :basesolver); base = base.basesolver; end
        init_z = copy(base.last_z)   # origin of the SimpleSolver (src/solvers.jl:155)
        push!(keep, elems, pm, init_z)
        subs[i] = CSubDesc(nn(model, i), nq(model, i), np(model, i), length(elems),
            shared(model.dqs[i]), shared(model.eqs[i]), shared(model.fqprevs[i]), shared(model.pexps[i]),
            shared(model.q0s[i]), shared(model.fqs[i]), shared(init_z), pointer(elems),
            pm isa Matrix ? perinst(pm, size(pm, 1)) : shared(pm), length(p), 0,
            CCache(0, 0, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL))
    end
    desc = Ref(CModelDesc(1, nx(model), nu(model), ny(model), nsub, solver, 0, 0, 0.0,
        shared(model.a), shared(model.b), shared(model.c), shared(model.x0),
        shared(model.dy), shared(model.ey), shared(model.fy), shared(model.y0), pointer(subs)))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep subs model begin
        check(ccall((:acmeb200_model_create, libacmeb200), Cint,
                    (Ref{CModelDesc}, Int64, Int64, Ref{Ptr{Cvoid}}), desc, 0, batch, h))
    end
    r = BatchRunner(model, batch, h[])
    finalizer(r -> ccall((:acmeb200_model_destroy, libacmeb200), Cvoid, (Ptr{Cvoid},), r.handle), r)
    return r
end

"""
    run!(runner::BatchRunner, Y::Array{Float64,3}, U::Array{Float64,3})

`size(U) == (nu, N, B)`, `size(Y) == (ny, N, B)`: instance `b` sees exactly the `nu × N` matrix the
reference's `run!` takes (src/ACME.jl:658-664).  Raises the reference's own messages.
"""
function ACME.run!(r::BatchRunner, Y::Array{Float64,3}, U::Array{Float64,3})
    m = r.model
    size(U, 1) == nu(m) || throw(DimensionMismatch("input matrix has $(size(U,1)) rows, but model has $(nu(m)) inputs"))
    size(Y, 1) == ny(m) || throw(DimensionMismatch("output matrix has $(size(Y,1)) rows, but model has $(ny(m)) outputs"))
    size(U, 2) == size(Y, 2) || throw(DimensionMismatch("input matrix has $(size(U,2)) columns, output matrix has $(size(Y,2)) columns"))
    N = size(U, 2)
    GC.@preserve U Y check(ccall((:acmeb200_run, libacmeb200), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, UInt32, Ptr{Cvoid}),
        r.handle, U, nu(m) * N, Y, ny(m) * N, N, 0, C_NULL))
    status = Vector{UInt32}(undef, r.batch)
    check(ccall((:acmeb200_get_status, libacmeb200), Cint, (Ptr{Cvoid}, Ptr{UInt32}, Ptr{Int64}), r.handle, status, C_NULL))
    any(s -> s & 0x2 != 0, status) && error("Failed to converge while solving non-linear equation, got non-finite result.")
    any(s -> s & 0x1 != 0, status) && @warn "Failed to converge while solving non-linear equation."
    return Y
end

ACME.run!(r::BatchRunner, U::Array{Float64,3}) = ACME.run!(r, Array{Float64,3}(undef, ny(r.model), size(U, 2), r.batch), U)

"""
    run_samplemajor!(runner::BatchRunner, Y::Array{Float64,3}, U::Array{Float64,3})

Same computation on sample-major streams (`ACMEB200_SAMPLE_MAJOR`, thread-per-instance kernels only):
`size(U) == (nu, B, N)`, `size(Y) == (ny, B, N)`, i.e. `U[:, :, n]` is one time step of the whole batch.
Bit-identical results; the device tiles then move as contiguous 256-byte rows.
"""
function run_samplemajor!(r::BatchRunner, Y::Array{Float64,3}, U::Array{Float64,3})
    m = r.model
    size(U, 1) == nu(m) || throw(DimensionMismatch("input matrix has $(size(U,1)) rows, but model has $(nu(m)) inputs"))
    size(Y, 1) == ny(m) || throw(DimensionMismatch("output matrix has $(size(Y,1)) rows, but model has $(ny(m)) outputs"))
    size(U, 3) == size(Y, 3) || throw(DimensionMismatch("input matrix has $(size(U,3)) columns, output matrix has $(size(Y,3)) columns"))
    (size(U, 2) == r.batch && size(Y, 2) == r.batch) || throw(DimensionMismatch("streams must hold $(r.batch) instances"))
    GC.@preserve U Y check(ccall((:acmeb200_run, libacmeb200), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, UInt32, Ptr{Cvoid}),
        r.handle, U, nu(m) * r.batch, Y, ny(m) * r.batch, size(U, 3), 0x4, C_NULL))
    return Y
end

"""
    cache_sizes(runner; sub=1) -> (stored::Vector{Int32}, capacity::Int)

Solutions the learning `CachingSolver` of sub-problem `sub` has stored so far, per instance
(`num_ps`, src/solvers.jl:321-323).  The device store is the reference's (same stored solutions, k-d tree rebuilt on
the same schedule); `capacity` is its physical size (a full store stops accepting solutions, flag 2 of `cache_info`).
"""
function cache_sizes(r::BatchRunner; sub::Integer=1)
    n = Vector{Int32}(undef, r.batch)
    cap = Ref{Int32}(0)
    check(ccall((:acmeb200_get_cache_sizes, libacmeb200), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}),
                r.handle, sub - 1, n, cap))
    return n, Int(cap[])
end

"""
    cache_info(runner; sub=1) -> Matrix{Int32} (8 x batch)

Rows: `num_ps`, `new_count`, `new_count_limit` (src/solvers.jl:321-325), the capacity the reference's doubling arrays
would have, points in the current tree, flags (1 frozen, 2 capacity reached, 4 search heap overflow), 2 reserved.
"""
function cache_info(r::BatchRunner; sub::Integer=1)
    a = Matrix{Int32}(undef, 8, r.batch)
    check(ccall((:acmeb200_get_cache_info, libacmeb200), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}), r.handle, sub - 1, a))
    return a
end

"""
    frozen_cache(solver::ACME.CachingSolver) -> CCache (+ the arrays it points into)

Freezes what a `CachingSolver` of the reference has learnt (`ps_tree.ps[:, 1:num_ps]`, `zs[:, 1:num_ps]`,
src/solvers.jl:321-323) into the read-only cache of a sub-problem descriptor.  The tree is rebuilt by the library's
host-side `KDTree` constructor (`acmeb200_kdtree_build`, src/kdtree.jl:11-73 -- the code the device runs when a learning
store rebuilds its tree) over the stored solutions alone: the reference's own tree also sorts the zero-filled spare
columns of its doubled arrays into place (kdtree.jl:37), which makes (p = 0, z = 0) a start point.
"""
function frozen_cache(s)
    n = s.num_ps
    ps = Matrix{Float64}(s.ps_tree.ps[:, 1:n]); zs = Matrix{Float64}(s.zs[:, 1:n])
    cut_dim = Vector{Int32}(undef, max(n - 1, 1)); cut_val = Vector{Float64}(undef, max(n - 1, 1)); ps_idx = Vector{Int32}(undef, n)
    check(ccall((:acmeb200_kdtree_build, libacmeb200), Cint, (Int32, Int32, Int32, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Int32}),
                size(ps, 1), n, n, ps, cut_dim, cut_val, ps_idx))
    keep = (ps, zs, cut_dim, cut_val, ps_idx)
    return CCache(n, n, pointer(cut_dim), pointer(cut_val), pointer(ps_idx), pointer(ps), pointer(zs)), keep
end

"""
    solver_state(runner) -> Vector{UInt8};  solver_state!(runner, blob)

Everything mutable of the device model -- `x`, the extrapolation origins `last_p, last_z, last_LU, last_Jp`
(src/solvers.jl:155-158), the learnt solution stores and their trees (src/solvers.jl:321-325), status, statistics -- as
one blob: what `deepcopy(model)` carries besides the matrices.  Restores into a runner of the same model, batch and
kernel, on any device: checkpointing, moving a batch between GPUs, warm starts.
"""
function solver_state(r::BatchRunner)
    n = ccall((:acmeb200_solver_state_size, libacmeb200), Int64, (Ptr{Cvoid},), r.handle)
    n > 0 || check(Cint(n))
    blob = Vector{UInt8}(undef, n)
    check(ccall((:acmeb200_get_solver_state, libacmeb200), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int64), r.handle, blob, n))
    return blob
end
solver_state!(r::BatchRunner, blob::Vector{UInt8}) =
    check(ccall((:acmeb200_set_solver_state, libacmeb200), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int64), r.handle, blob, length(blob)))

"""
    extrapolation_origin(runner; sub=1) -> (p::Matrix, z::Matrix)      # get_extrapolation_origin, src/solvers.jl:199
"""
function extrapolation_origin(r::BatchRunner; sub::Integer=1)
    p = Matrix{Float64}(undef, np(r.model, sub), r.batch); z = Matrix{Float64}(undef, nn(r.model, sub), r.batch)
    check(ccall((:acmeb200_get_extrapolation_origin, libacmeb200), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}), r.handle, sub - 1, p, z))
    return p, z
end

"""
    eval_jq(runner, q::Matrix; sub=1) -> Array{Float64,3}       # Jq of CircuitNLFunc (src/circuit.jl:10-17), nn x nq x batch

Element Jacobians of the whole batch at one `q` (nq x batch) per instance, each with its own element
parameters: what a batched `linearize` (src/ACME.jl:520-546) needs at the steady state.
"""
function eval_jq(r::BatchRunner, q::Matrix{Float64}; sub::Integer=1)
    qh = permutedims(q)                                                # [nq][batch], instance index fastest
    out = Array{Float64,3}(undef, r.batch, nn(r.model, sub), size(q, 1))
    check(ccall((:acmeb200_eval_jq, libacmeb200), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}), r.handle, sub - 1, qh, out))
    return permutedims(out, (2, 3, 1))
end

# ---- devices: a model lives on the device that is current when it is created
device_count() = (n = Ref{Int32}(0); check(ccall((:acmeb200_device_count, libacmeb200), Cint, (Ref{Int32},), n)); Int(n[]))
set_device(d::Integer) = check(ccall((:acmeb200_set_device, libacmeb200), Cint, (Int32,), d))

"""
    MultiGpuRunner(desc_builder, batch; n_gpus=0);  run!(r::MultiGpuRunner, Y, U)

The batch over all the GPUs of the box from one Julia process (`acmeb200_multi_create/run`): contiguous instance shards,
one device model per GPU, host arrays for the whole batch, every shard's host-buffer pipeline on its own thread inside
the library (which never calls back into Julia, so a plain `ccall` suffices).  `desc_builder()` returns the same
`Ref{CModelDesc}` (and the arrays it points into) that `BatchRunner` assembles.
"""
mutable struct MultiGpuRunner
    model::DiscreteModel
    batch::Int
    handle::Ptr{Cvoid}
end
function MultiGpuRunner(model::DiscreteModel, desc::Ref{CModelDesc}, keep, batch::Integer; n_gpus::Integer=0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep check(ccall((:acmeb200_multi_create, libacmeb200), Cint, (Ref{CModelDesc}, Int64, Int32, Ref{Ptr{Cvoid}}), desc, batch, n_gpus, h))
    r = MultiGpuRunner(model, batch, h[])
    finalizer(r -> ccall((:acmeb200_multi_destroy, libacmeb200), Cvoid, (Ptr{Cvoid},), r.handle), r)
    return r
end
function ACME.run!(r::MultiGpuRunner, Y::Array{Float64,3}, U::Array{Float64,3})
    m = r.model; N = size(U, 2)
    size(U, 1) == nu(m) || throw(DimensionMismatch("input matrix has $(size(U,1)) rows, but model has $(nu(m)) inputs"))
    size(Y) == (ny(m), N, r.batch) || throw(DimensionMismatch("output must be $(ny(m)) x $N x $(r.batch)"))
    GC.@preserve U Y check(ccall((:acmeb200_multi_run, libacmeb200), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Int64, Int64, UInt32), r.handle, U, nu(m) * N, Y, ny(m) * N, N, 0))
    return Y
end

end # module
