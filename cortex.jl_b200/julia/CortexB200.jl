# CortexB200.jl — the `ccall` glue a Cortex.jl maintainer adds to put the B200 engine behind the existing API.
#
# NOT executed in this repository's CI: the build image has no `julia` (see DESIGN.md §1). It is a 1:1, mechanically
# checkable mirror of include/cortex_b200.h; the same ABI is exercised by the Python ctypes frontend in the tests.
#
# Usage (on a machine with Julia, Cortex.jl, BipartiteFactorGraphs.jl and libcortex_b200.so):
#     using Cortex, BipartiteFactorGraphs, CortexB200
#     engine = CortexB200.B200InferenceEngine(graph; rules = Dict(:likelihood => (CortexB200.RULE_GAUSS_OBS, [1.0]),
#                                                               :transition => (CortexB200.RULE_GAUSS_RW,  [1.0])),
#                                             family = CortexB200.FAMILY_GAUSS_CANON, value_dim = 2)
#     CortexB200.set_value!(engine, Cortex.get_connection_message_to_factor(...)-equivalent signal id, value)
#     Cortex.update_marginals!(engine, variable_ids)          # dispatches to cxb_update_marginals
module CortexB200

using Cortex

const LIB = get(ENV, "CORTEX_B200_LIB", "libcortex_b200.so")

# ---- enums of include/cortex_b200.h ---------------------------------------------------------------------------
const OK, ERR_NOT_PENDING, ERR_NO_RULE, ERR_OUT_OF_CONTRACT, ERR_BAD_ARG, ERR_UNSUPPORTED_ENGINE, ERR_CUDA, ERR_STATE = 0:7
const F32, F64 = 0, 1
const KIND_UNSPECIFIED, KIND_M2F, KIND_M2V, KIND_PRODUCT, KIND_MARGINAL, KIND_JOINT = 0:5
const FAMILY_GAUSS_CANON, FAMILY_CATEGORICAL, FAMILY_GAUSS_MV, FAMILY_BETA, FAMILY_SUM = 0:4
const RULE_NONE, RULE_GAUSS_OBS, RULE_GAUSS_RW, RULE_CAT_TABLE, RULE_POTTS, RULE_HMM_EMIT,
      RULE_GAUSS_MV_OBS, RULE_GAUSS_MV_RW, RULE_BETA_BERNOULLI, RULE_SCALE2 = 0:9
const RESOLVER_NONE, RESOLVER_DEFAULT_BP, RESOLVER_MEAN_FIELD = 0:2

struct UpdateStats
    levels::Int64
    updates::Int64
    updates_by_kind::NTuple{6, Int64}
    final_marginals::Int64
    final_linked::Int64
    kernel_launches::Int64
end

# ---- error mapping (SURVEY §8b): status -> the exception the reference throws on that path ------------------------
function check(h::Ptr{Cvoid}, status::Int32)
    status == OK && return nothing
    msg = unsafe_string(ccall((:cxb_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    status == ERR_NOT_PENDING && throw(ArgumentError(msg))                      # src/signal.jl:399-405
    status == ERR_NO_RULE && error(msg)                                         # src/inference_engine.jl:358-360
    status == ERR_UNSUPPORTED_ENGINE && throw(Cortex.UnsupportedModelEngineError(nothing, nothing))
    error("cortex_b200 status $status: $msg")
end

"""
    B200InferenceEngine(model_engine; rules, family, value_dim, dtype = F32, device = 0,
                        dependency_resolver = RESOLVER_DEFAULT_BP)

Walks the 7 backend generics of the model engine once (src/model_engine.jl:329-391) and hands flat arrays to
`cxb_graph_build`; registers one rule kernel per `Factor.functional_form`; resolves dependencies on the host side of the
library exactly as `DefaultDependencyResolver` does (src/dependencies.jl).
"""
mutable struct B200InferenceEngine{M}
    model_engine::M
    handle::Ptr{Cvoid}
    id_offset::Int              # Julia ids are 1-based, the library's 0-based
    type_of_form::Dict{Any, Int32}
end

function B200InferenceEngine(model_engine::M; rules::Dict, family::Integer, value_dim::Integer, dtype::Integer = F32,
                             device::Integer = 0, dependency_resolver::Integer = RESOLVER_DEFAULT_BP) where {M}
    Cortex.throw_if_engine_unsupported(model_engine)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:cxb_create, LIB), Int32, (Int32, Int32, Int32, Int32, Ref{Ptr{Cvoid}}), device, dtype, value_dim, family, href)
    st == OK || error("cxb_create failed with status $st (a CUDA device is required; there is no CPU fallback)")
    h = href[]
    vids = collect(Int, Cortex.get_variable_ids(model_engine))
    fids = collect(Int, Cortex.get_factor_ids(model_engine))
    n_ids = maximum(vcat(vids, fids); init = 0)
    is_factor = zeros(UInt8, n_ids); ftype = zeros(Int32, n_ids)
    type_of_form = Dict{Any, Int32}()
    for f in fids
        is_factor[f] = 1
        form = Cortex.get_factor_functional_form(Cortex.get_factor(model_engine, f))
        ftype[f] = get!(type_of_form, form, Int32(length(type_of_form)))
    end
    ev = Int64[]; ef = Int64[]
    for f in fids, v in Cortex.get_connected_variable_ids(model_engine, f)
        push!(ev, v - 1); push!(ef, f - 1)
    end
    check(h, ccall((:cxb_graph_build, LIB), Int32,
                   (Ptr{Cvoid}, Int64, Ptr{UInt8}, Ptr{Int32}, Int64, Ptr{Int64}, Ptr{Int64}),
                   h, n_ids, is_factor, ftype, length(ev), ev, ef))
    for (form, (kind, params)) in rules
        haskey(type_of_form, form) || continue
        p = convert(Vector{Float64}, params)
        check(h, ccall((:cxb_register_rule, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Int64),
                       h, type_of_form[form], kind, p, length(p)))
    end
    check(h, ccall((:cxb_resolve_dependencies, LIB), Int32, (Ptr{Cvoid}, Int32), h, dependency_resolver))
    engine = B200InferenceEngine{M}(model_engine, h, 1, type_of_form)
    finalizer(e -> ccall((:cxb_destroy, LIB), Cvoid, (Ptr{Cvoid},), e.handle), engine)
    return engine
end

# ---- signal ids (get_variable_marginal / get_connection_message_to_* equivalents) ---------------------------------
marginal_id(e::B200InferenceEngine, v::Int) =
    ccall((:cxb_signal_id, LIB), Int64, (Ptr{Cvoid}, Int32, Int64, Int64), e.handle, KIND_MARGINAL, v - 1, -1)
message_to_variable_id(e::B200InferenceEngine, v::Int, f::Int) =
    ccall((:cxb_signal_id, LIB), Int64, (Ptr{Cvoid}, Int32, Int64, Int64), e.handle, KIND_M2V, v - 1, f - 1)
message_to_factor_id(e::B200InferenceEngine, v::Int, f::Int) =
    ccall((:cxb_signal_id, LIB), Int64, (Ptr{Cvoid}, Int32, Int64, Int64), e.handle, KIND_M2F, v - 1, f - 1)

# ---- data in / out: set_value! (src/signal.jl:232), get_value (:171) -----------------------------------------------
function set_values!(e::B200InferenceEngine, signal_ids::Vector{Int64}, values::Matrix{Float64})   # values: dim x n
    check(e.handle, ccall((:cxb_set_values, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Int64),
                          e.handle, length(signal_ids), signal_ids, values, size(values, 1)))
end
function get_values(e::B200InferenceEngine, signal_ids::Vector{Int64}, dim::Int)
    out = zeros(Float64, dim, length(signal_ids))
    check(e.handle, ccall((:cxb_get_values, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Int64),
                          e.handle, length(signal_ids), signal_ids, out, dim))
    return out
end
is_pending(e::B200InferenceEngine, sid::Int64) = ccall((:cxb_is_pending, LIB), Int32, (Ptr{Cvoid}, Int64), e.handle, sid) == 1
is_computed(e::B200InferenceEngine, sid::Int64) = ccall((:cxb_is_computed, LIB), Int32, (Ptr{Cvoid}, Int64), e.handle, sid) == 1
link_signal_to_variable!(e::B200InferenceEngine, v::Int, sid::Int64) =
    check(e.handle, ccall((:cxb_link_signal, LIB), Int32, (Ptr{Cvoid}, Int64, Int64), e.handle, v - 1, sid))

# ---- the scheduler entry points: more specific methods of the reference generics ---------------------------------
function Cortex.request_inference_for(e::B200InferenceEngine, variable_ids::Union{AbstractVector, Tuple})
    ids = Int64[v - 1 for v in variable_ids]
    check(e.handle, ccall((:cxb_request_inference, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}), e.handle, length(ids), ids))
    return ids
end

function scan_inference_request(e::B200InferenceEngine)
    n = ccall((:cxb_scan, LIB), Int64, (Ptr{Cvoid}, Ptr{Int64}, Int64), e.handle, C_NULL, 0)
    out = zeros(Int64, n)
    ccall((:cxb_scan, LIB), Int64, (Ptr{Cvoid}, Ptr{Int64}, Int64), e.handle, out, n)
    return out
end

function Cortex.update_marginals!(e::B200InferenceEngine, variable_ids::Union{AbstractVector, Tuple})
    ids = Int64[v - 1 for v in variable_ids]
    stats = Ref{UpdateStats}()
    check(e.handle, ccall((:cxb_update_marginals, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ref{UpdateStats}),
                          e.handle, length(ids), ids, stats))
    return nothing                                  # the reference returns nothing (src/inference_engine.jl:631)
end
Cortex.update_marginals!(e::B200InferenceEngine, variable_id) = Cortex.update_marginals!(e, (variable_id,))

# ---- structured model engines (closed-form plans) ------------------------------------------------------------------
mutable struct GaussianChainBatch
    handle::Ptr{Cvoid}
    n_chains::Int
    n_steps::Int
end
function GaussianChainBatch(n_chains::Int, n_steps::Int; dtype::Integer = F32, device::Integer = 0)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:cxb_chains_create, LIB), Int32, (Int32, Int32, Int64, Int64, Ref{Ptr{Cvoid}}), device, dtype, n_chains, n_steps, href)
    st == OK || error("cxb_chains_create failed with status $st")
    c = GaussianChainBatch(href[], n_chains, n_steps)
    finalizer(x -> ccall((:cxb_chains_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.handle), c)
    return c
end
set_noise!(c::GaussianChainBatch, q::Vector{Float64}, r::Vector{Float64}) =
    ccall((:cxb_chains_set_noise, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), c.handle, q, r)
set_observations!(c::GaussianChainBatch, y::Matrix{Float32}) =                       # y is [B, T] column-major == [T][B]
    ccall((:cxb_chains_set_observations, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), c.handle, y)
function update_marginals!(c::GaussianChainBatch)
    n = Ref{Int64}(0)
    st = ccall((:cxb_chains_update_marginals, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), c.handle, n)
    st == OK || error(unsafe_string(ccall((:cxb_chains_last_error, LIB), Cstring, (Ptr{Cvoid},), c.handle)))
    return n[]
end
function get_marginals(c::GaussianChainBatch)
    out = zeros(Float32, 2, c.n_chains, c.n_steps)
    ccall((:cxb_chains_get_marginals, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), c.handle, out)
    return out
end

end # module
