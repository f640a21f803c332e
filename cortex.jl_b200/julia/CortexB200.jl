# CortexB200.jl — puts the B200 engine behind Cortex.jl's own API (SURVEY §8b / §8f.3).
#
#     using Cortex, BipartiteFactorGraphs, CortexB200
#     model  = CortexB200.B200ModelEngine(graph;                       # any supported model engine, walked ONCE
#                  rules = Dict(:likelihood => (CortexB200.RULE_GAUSS_OBS, [1.0]), :transition => (CortexB200.RULE_GAUSS_RW, [1.0])),
#                  family = CortexB200.FAMILY_GAUSS_CANON, value_dim = 2)
#     engine = Cortex.InferenceEngine(model_engine = model, prepare_signals_metadata = false, resolve_dependencies = false)
#     CortexB200.set_message_to_factor!(model, y[t], likelihood[t], [obs, 0.0])        # set_value!(m2f(y_t, lik_t), ...)
#     Cortex.update_marginals!(engine, x)                                              # -> cxb_update_marginals
#     Cortex.get_value(Cortex.get_variable_marginal(Cortex.get_variable(engine, x[1])))  # fetched from the device
#
# `B200ModelEngine` is a model-engine BACKEND in the reference's sense (src/model_engine.jl:269-391): it implements the trait
# and the seven generics, LAZILY — the device owns every signal; `get_variable` / `get_connection` materialise a `Variable` /
# `Connection` VIEW whose signals are fresh host `InferenceSignal`s carrying the device's current value and variant (nothing is
# allocated per signal up front: a 10^8-signal graph stays a few flat arrays). Signal variants and dependencies are made by
# the library (`cxb_graph_build`, `cxb_resolve_dependencies`), hence `prepare_signals_metadata = false, resolve_dependencies =
# false` in the `InferenceEngine` constructor (src/inference_engine.jl:60-89). The hot path is a MORE SPECIFIC method
# `Cortex.update_marginals!(::InferenceEngine{<:B200ModelEngine}, ids)` (the reference's are at src/inference_engine.jl:555,
# 559), so every other `InferenceEngine` keeps running the reference loop.
#
# NOT executed in this repository: the build image and the GPU boxes have no `julia` (DESIGN.md §1). The same ABI, call for
# call, is exercised by the Python ctypes frontend (`cortex.jl_b200/model_engine.py::B200ModelEngine` mirrors this file and is
# tested on the GPU: tests/test_b200_model_engine.py).
module CortexB200

using Cortex
import Cortex: is_engine_supported, get_variable, get_factor, get_variable_ids, get_factor_ids, get_connection,
               get_connected_variable_ids, get_connected_factor_ids, update_marginals!, request_inference_for

const LIB = get(ENV, "CORTEX_B200_LIB", "libcortex_b200.so")

# ---- enums of include/cortex_b200.h ---------------------------------------------------------------------------
const OK, ERR_NOT_PENDING, ERR_NO_RULE, ERR_OUT_OF_CONTRACT, ERR_BAD_ARG, ERR_UNSUPPORTED_ENGINE, ERR_CUDA, ERR_STATE, ERR_INTERNAL = 0:8
const F32, F64 = 0, 1
const KIND_UNSPECIFIED, KIND_M2F, KIND_M2V, KIND_PRODUCT, KIND_MARGINAL, KIND_JOINT = 0:5
const FAMILY_GAUSS_CANON, FAMILY_CATEGORICAL, FAMILY_GAUSS_MV, FAMILY_BETA, FAMILY_SUM,
      FAMILY_GAUSS_MP, FAMILY_GAMMA, FAMILY_POINT = 0:7
const RULE_NONE, RULE_GAUSS_OBS, RULE_GAUSS_RW, RULE_CAT_TABLE, RULE_POTTS, RULE_HMM_EMIT,
      RULE_GAUSS_MV_OBS, RULE_GAUSS_MV_RW, RULE_BETA_BERNOULLI, RULE_SCALE2, RULE_NORMAL_MEAN_FIELD,
      RULE_NORMAL_STRUCTURED, RULE_PROGRAM = 0:12
# opcodes of RULE_PROGRAM (a user-defined rule as a stack program: params = [n_consts, consts..., code...]; see
# include/cortex_b200.h and `rule_program` below)
const OP_DEP, OP_CONST, OP_PARAM, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_EXP, OP_LOG, OP_SQRT, OP_STORE, OP_TSET, OP_TGET,
      OP_NDEPS = 1:15
const RESOLVER_NONE, RESOLVER_DEFAULT_BP, RESOLVER_MEAN_FIELD = 0:2
const SCHEDULE_AUTO, SCHEDULE_LEVEL, SCHEDULE_SEQUENTIAL, RAN_REPLAY, RAN_PLAN = 0:4
const DEP_INTERMEDIATE, DEP_WEAK, DEP_NO_LISTEN, DEP_NO_CHECK_COMPUTED = 1, 2, 16, 32

struct UpdateStats
    levels::Int64
    updates::Int64
    updates_by_kind::NTuple{6, Int64}
    final_marginals::Int64
    final_linked::Int64
    kernel_launches::Int64
end

"The level-synchronous schedule would differ from the reference's order (only with `schedule = SCHEDULE_LEVEL`)."
struct OutOfContractError <: Exception
    msg::String
end

# ---- error mapping (SURVEY §8b): status -> the exception the reference throws on that path ------------------------
function check(h::Ptr{Cvoid}, status::Integer)
    status == OK && return nothing
    msg = unsafe_string(ccall((:cxb_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    status == ERR_NOT_PENDING && throw(ArgumentError(msg))                      # src/signal.jl:399-405
    status == ERR_NO_RULE && error(msg)                                         # src/inference_engine.jl:358-360
    status == ERR_OUT_OF_CONTRACT && throw(OutOfContractError(msg))
    status == ERR_BAD_ARG && throw(ArgumentError(msg))
    error("cortex_b200 status $status: $msg")
end

"""
    B200ModelEngine(source; rules, family, value_dim, dtype = F32, device = 0, dependency_resolver = RESOLVER_DEFAULT_BP,
                    decode = identity, encode = identity)

A Cortex model-engine backend whose signals live on a B200. `source` is any supported model engine (e.g. a
`BipartiteFactorGraph`): its seven generics are walked once (src/model_engine.jl:329-391) and the graph handed to the library
as flat arrays; or use `B200ModelEngine(n_ids, is_factor, functional_forms, edge_variable, edge_factor; ...)` for graphs that
never exist as host objects. `rules` maps `Factor.functional_form` to a registered rule kernel `(RULE_*, params)` — the
counterpart of methods of `compute_message_to_variable!` dispatching on the functional form
(test/inference_engine_tests.jl:256-259). `decode` / `encode` convert between the `value_dim` numbers of a device value and
the Julia value the user wants to see (e.g. `v -> NormalCanonical(v[1], v[2])`).
"""
mutable struct B200ModelEngine
    handle::Ptr{Cvoid}
    variable_ids::Vector{Int}
    factor_ids::Vector{Int}
    forms::Dict{Int, Any}                  # factor id -> Factor.functional_form
    names::Dict{Int, Tuple{Symbol, Any}}   # variable id -> (name, index), kept from the source engine
    labels::Dict{Tuple{Int, Int}, Tuple{Symbol, Int}}  # (variable, factor) -> connection (label, index)
    factors_of::Dict{Int, Vector{Int}}     # ascending ids (the order contract of ext/BipartiteFactorGraphsExt/...:26-48)
    variables_of::Dict{Int, Vector{Int}}
    value_dim::Int
    dtype::Int
    decode::Any
    encode::Any
    trace::Bool
end

is_engine_supported(::B200ModelEngine) = Cortex.SupportedModelEngine()   # src/model_engine.jl:269-310

function B200ModelEngine(source; rules::AbstractDict, family::Integer, value_dim::Integer, dtype::Integer = F32, device::Integer = 0,
                         dependency_resolver::Integer = RESOLVER_DEFAULT_BP, decode = identity, encode = identity)
    Cortex.throw_if_engine_unsupported(source)
    vids = sort!(collect(Int, Cortex.get_variable_ids(source)))
    fids = sort!(collect(Int, Cortex.get_factor_ids(source)))
    forms = Dict{Int, Any}(f => Cortex.get_factor_functional_form(Cortex.get_factor(source, f)) for f in fids)
    names = Dict{Int, Tuple{Symbol, Any}}()
    for v in vids
        var = Cortex.get_variable(source, v)
        names[v] = (Cortex.get_variable_name(var), Cortex.get_variable_index(var))
    end
    ev = Int[]; ef = Int[]
    labels = Dict{Tuple{Int, Int}, Tuple{Symbol, Int}}()
    for f in fids, v in sort!(collect(Int, Cortex.get_connected_variable_ids(source, f)))
        push!(ev, v); push!(ef, f)
        c = Cortex.get_connection(source, v, f)
        labels[(v, f)] = (Cortex.get_connection_label(c), Cortex.get_connection_index(c))
    end
    n_ids = maximum(vcat(vids, fids); init = 0)
    is_factor = zeros(UInt8, n_ids)
    is_factor[fids] .= 1
    return B200ModelEngine(n_ids, is_factor, forms, ev, ef; rules, family, value_dim, dtype, device, dependency_resolver, decode, encode,
                           names, labels)
end

function B200ModelEngine(n_ids::Integer, is_factor::Vector{UInt8}, forms::AbstractDict, edge_variable::Vector{Int}, edge_factor::Vector{Int};
                         rules::AbstractDict, family::Integer, value_dim::Integer, dtype::Integer = F32, device::Integer = 0,
                         dependency_resolver::Integer = RESOLVER_DEFAULT_BP, decode = identity, encode = identity,
                         names = Dict{Int, Tuple{Symbol, Any}}(), labels = Dict{Tuple{Int, Int}, Tuple{Symbol, Int}}())
    href = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:cxb_create, LIB), Int32, (Int32, Int32, Int32, Int32, Ref{Ptr{Cvoid}}), device, dtype, value_dim, family, href)
    st == OK || error("cxb_create failed with status $st (a CUDA device is required; there is no CPU fallback)")
    h = href[]
    type_of_form = Dict{Any, Int32}()
    ftype = zeros(Int32, n_ids)
    for f in 1:n_ids
        is_factor[f] == 1 || continue
        ftype[f] = get!(type_of_form, forms[f], Int32(length(type_of_form)))
    end
    ev0 = Int64.(edge_variable .- 1); ef0 = Int64.(edge_factor .- 1)          # Julia ids are 1-based, the library's 0-based
    check(h, ccall((:cxb_graph_build, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{UInt8}, Ptr{Int32}, Int64, Ptr{Int64}, Ptr{Int64}),
                   h, n_ids, is_factor, ftype, length(ev0), ev0, ef0))
    for (form, (kind, params)) in rules
        haskey(type_of_form, form) || continue
        p = convert(Vector{Float64}, collect(params))
        check(h, ccall((:cxb_register_rule, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}, Int64), h, type_of_form[form], kind, p, length(p)))
    end
    check(h, ccall((:cxb_resolve_dependencies, LIB), Int32, (Ptr{Cvoid}, Int32), h, dependency_resolver))   # src/dependencies.jl:5-173
    factors_of = Dict{Int, Vector{Int}}(); variables_of = Dict{Int, Vector{Int}}()
    for (v, f) in zip(edge_variable, edge_factor)
        push!(get!(factors_of, v, Int[]), f); push!(get!(variables_of, f, Int[]), v)
    end
    foreach(sort!, values(factors_of)); foreach(sort!, values(variables_of))
    vids = [i for i in 1:n_ids if is_factor[i] == 0]; fids = [i for i in 1:n_ids if is_factor[i] == 1]
    m = B200ModelEngine(h, vids, fids, Dict{Int, Any}(forms), names, labels, factors_of, variables_of, value_dim, dtype, decode, encode, false)
    finalizer(e -> ccall((:cxb_destroy, LIB), Cvoid, (Ptr{Cvoid},), e.handle), m)
    return m
end

"""
    rule_program(consts::Vector{Float64}, code::Vector) -> (RULE_PROGRAM, params)

A user-defined message-to-variable rule for fixed-size values without rebuilding the library (the counterpart of a new method of
`compute_message_to_variable!`, src/inference_engine.jl:351-361). `consts[1]` is the default of the per-factor parameter; `code`
is the postfix program, e.g. the random-walk rule `(L, h) -> (L / (1 + q L), h / (1 + q L))`:

    rule_program([1.0, 1.0], [OP_CONST, 1, OP_PARAM, OP_DEP, 0, 0, OP_MUL, OP_ADD, OP_TSET, 0,
                              OP_DEP, 0, 0, OP_TGET, 0, OP_DIV, OP_STORE, 0, OP_DEP, 0, 1, OP_TGET, 0, OP_DIV, OP_STORE, 1])
"""
rule_program(consts::Vector{Float64}, code::Vector) = (RULE_PROGRAM, vcat(Float64(length(consts)), consts, Float64.(code)))

# ---- signals: dense ids on the device, views on the host ---------------------------------------------------------------
signal_id(m::B200ModelEngine, kind::Integer, v::Integer, f::Integer = 0) =
    ccall((:cxb_signal_id, LIB), Int64, (Ptr{Cvoid}, Int32, Int64, Int64), m.handle, kind, v - 1, f - 1)

function device_is_computed(m::B200ModelEngine, sid::Int64)
    r = ccall((:cxb_is_computed, LIB), Int32, (Ptr{Cvoid}, Int64), m.handle, sid)
    r < 0 && throw(ArgumentError("bad signal id"))
    return r == 1
end
function device_value(m::B200ModelEngine, sid::Int64)
    out = zeros(Float64, m.value_dim); ids = [sid]
    check(m.handle, ccall((:cxb_get_values, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Int64), m.handle, 1, ids, out, m.value_dim))
    return m.decode(out)
end
"A host `InferenceSignal` showing the device signal `sid`: its current value (`UndefValue()` if never computed) and variant."
function signal_view(m::B200ModelEngine, sid::Int64, variant)
    value = device_is_computed(m, sid) ? device_value(m, sid) : Cortex.UndefValue()
    return Cortex.Signal(Any, Cortex.InferenceSignalVariant, value, variant)      # src/signal.jl:94-104
end
function linked_signal_views(m::B200ModelEngine, v::Int)
    # linked signals are m2f messages (protocol B) or user signals; their ids come back from the library
    return Cortex.InferenceSignal[]   # the device keeps the links (cxb_link_signal); views are made on request with signal_view
end

# ---- the seven backend generics, src/model_engine.jl:329-391 ------------------------------------------------------------
get_variable_ids(m::B200ModelEngine) = m.variable_ids
get_factor_ids(m::B200ModelEngine) = m.factor_ids
get_connected_variable_ids(m::B200ModelEngine, factor_id::Int) = get(m.variables_of, factor_id, Int[])
get_connected_factor_ids(m::B200ModelEngine, variable_id::Int) = get(m.factors_of, variable_id, Int[])
function get_variable(m::B200ModelEngine, variable_id::Int)::Cortex.Variable
    sid = signal_id(m, KIND_MARGINAL, variable_id)
    sid < 0 && throw(ArgumentError("not a variable id: $variable_id"))
    name, index = get(m.names, variable_id, (:variable, variable_id))
    marginal = signal_view(m, sid, Cortex.InferenceSignalVariants.IndividualMarginal(variable_id))
    return Cortex.Variable(name = name, index = index, marginal = marginal, linked_signals = linked_signal_views(m, variable_id))
end
function get_factor(m::B200ModelEngine, factor_id::Int)::Cortex.Factor
    haskey(m.forms, factor_id) || throw(ArgumentError("not a factor id: $factor_id"))
    return Cortex.Factor(functional_form = m.forms[factor_id])
end
function get_connection(m::B200ModelEngine, variable_id::Int, factor_id::Int)::Cortex.Connection
    s2v = signal_id(m, KIND_M2V, variable_id, factor_id); s2f = signal_id(m, KIND_M2F, variable_id, factor_id)
    (s2v < 0 || s2f < 0) && throw(ArgumentError("no connection between variable $variable_id and factor $factor_id"))
    label, index = get(m.labels, (variable_id, factor_id), (:edge, 0))
    return Cortex.Connection(label = label, index = index,
        message_to_variable = signal_view(m, s2v, Cortex.InferenceSignalVariants.MessageToVariable(variable_id, factor_id)),
        message_to_factor = signal_view(m, s2f, Cortex.InferenceSignalVariants.MessageToFactor(variable_id, factor_id)))
end

# ---- data in: set_value!(signal, v) of src/signal.jl:232-253 addressed by what the signal IS -------------------------------
function set_device_values!(m::B200ModelEngine, sids::Vector{Int64}, values)
    flat = zeros(Float64, m.value_dim, length(sids))
    for (k, v) in enumerate(values)
        e = m.encode(v); flat[1:length(e), k] .= e
    end
    check(m.handle, ccall((:cxb_set_values, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Float64}, Int64), m.handle, length(sids), sids, flat, m.value_dim))
end
set_message_to_factor!(m::B200ModelEngine, v::Int, f::Int, value) = set_device_values!(m, [signal_id(m, KIND_M2F, v, f)], [value])
set_message_to_variable!(m::B200ModelEngine, v::Int, f::Int, value) = set_device_values!(m, [signal_id(m, KIND_M2V, v, f)], [value])
set_marginal!(m::B200ModelEngine, v::Int, value) = set_device_values!(m, [signal_id(m, KIND_MARGINAL, v)], [value])
"Bulk form: one host->device copy and one notification kernel for a whole vector of (variable, factor) observations."
set_messages_to_factor!(m::B200ModelEngine, vs, fs, values) =
    set_device_values!(m, Int64[signal_id(m, KIND_M2F, v, f) for (v, f) in zip(vs, fs)], values)
"link_signal_to_variable!(variable, m2f(variable, factor)), src/model_engine.jl:80-83 (protocol B of SURVEY Appendix B)."
link_message_to_factor!(m::B200ModelEngine, v::Int, f::Int) =
    check(m.handle, ccall((:cxb_link_signal, LIB), Int32, (Ptr{Cvoid}, Int64, Int64), m.handle, v - 1, signal_id(m, KIND_M2F, v, f)))
set_schedule!(m::B200ModelEngine, schedule::Integer) = check(m.handle, ccall((:cxb_set_schedule, LIB), Int32, (Ptr{Cvoid}, Int32), m.handle, schedule))
last_schedule(m::B200ModelEngine) = ccall((:cxb_last_schedule, LIB), Int32, (Ptr{Cvoid},), m.handle)

# custom wiring for user resolvers (test/inference_engine_tests.jl:597-621, 811-907): signals are addressed by id
create_signal!(m::B200ModelEngine) = ccall((:cxb_create_signal, LIB), Int64, (Ptr{Cvoid},), m.handle)
function add_dependency!(m::B200ModelEngine, signal::Int64, dependency::Int64; weak = false, listen = true, check_computed = true, intermediate = false)
    flags = (intermediate ? DEP_INTERMEDIATE : 0) | (weak ? DEP_WEAK : 0) | (listen ? 0 : DEP_NO_LISTEN) | (check_computed ? 0 : DEP_NO_CHECK_COMPUTED)
    check(m.handle, ccall((:cxb_add_dependency, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int32), m.handle, signal, dependency, flags))
end

# ---- warnings, src/inference_engine.jl:11-14 + src/dependencies.jl:40-43 ------------------------------------------------
function collect_warnings!(engine::Cortex.InferenceEngine{<:B200ModelEngine})
    m = Cortex.get_model_engine(engine)
    n = ccall((:cxb_get_warnings, LIB), Int64, (Ptr{Cvoid}, Ptr{Int64}, Int64), m.handle, C_NULL, 0)
    n <= 0 && return engine
    buf = zeros(Int64, n)
    ccall((:cxb_get_warnings, LIB), Int64, (Ptr{Cvoid}, Ptr{Int64}, Int64), m.handle, buf, n)
    for v in buf
        Cortex.add_warning!(engine, "Variable has no connected factors", Int(v) + 1)
    end
    return engine
end

# ---- the hot path: more specific than src/inference_engine.jl:555 and :559 ------------------------------------------------
update_marginals!(engine::Cortex.InferenceEngine{<:B200ModelEngine}, variable_id::Integer) = update_marginals!(engine, (variable_id,))
function update_marginals!(engine::Cortex.InferenceEngine{<:B200ModelEngine}, variable_ids::Union{AbstractVector, Tuple})
    m = Cortex.get_model_engine(engine)
    ids = Int64[Int64(v) - 1 for v in variable_ids]
    tracer = Cortex.get_trace(engine)
    if tracer !== nothing && !m.trace
        check(m.handle, ccall((:cxb_trace_enable, LIB), Int32, (Ptr{Cvoid}, Int32), m.handle, 1)); m.trace = true
    end
    stats = Ref(UpdateStats(0, 0, ntuple(_ -> 0, 6), 0, 0, 0))
    t0 = time_ns()
    st = ccall((:cxb_update_marginals, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ref{UpdateStats}), m.handle, length(ids), ids, stats)
    check(m.handle, st)
    tracer === nothing || push!(tracer.inference_requests, collect_trace(engine, variable_ids, time_ns() - t0))
    return nothing                                                       # as the reference (src/inference_engine.jl:631)
end

# request_inference_for + scan_inference_request, src/inference_engine.jl:294-323, 540-546 (the reference's DFS order)
function scan_pending(engine::Cortex.InferenceEngine{<:B200ModelEngine}, variable_ids)
    m = Cortex.get_model_engine(engine)
    ids = Int64[Int64(v) - 1 for v in variable_ids]
    check(m.handle, ccall((:cxb_request_inference, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}), m.handle, length(ids), ids))
    n = ccall((:cxb_scan_dfs, LIB), Int64, (Ptr{Cvoid}, Ptr{Int64}, Int64), m.handle, C_NULL, 0)
    buf = zeros(Int64, max(n, 1))
    n = ccall((:cxb_scan_dfs, LIB), Int64, (Ptr{Cvoid}, Ptr{Int64}, Int64), m.handle, buf, n)
    return [signal_view(m, sid, variant_of(m, sid)) for sid in buf[1:n]]
end

function variant_of(m::B200ModelEngine, sid::Int64)
    info = zeros(Int64, 5)
    check(m.handle, ccall((:cxb_signal_info, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}), m.handle, sid, info))
    kind, v, f = info[1], Int(info[2]) + 1, Int(info[3]) + 1
    V = Cortex.InferenceSignalVariants
    kind == KIND_M2F && return V.MessageToFactor(v, f)
    kind == KIND_M2V && return V.MessageToVariable(v, f)
    kind == KIND_MARGINAL && return V.IndividualMarginal(v)
    kind == KIND_PRODUCT && return V.ProductOfMessages(v, (Int(info[4]) + 1):(Int(info[5]) + 1), get_connected_factor_ids(m, v))
    kind == KIND_JOINT && return V.JointMarginal(f, Int[])
    return V.Unspecified()
end

# ---- tracer, src/inference_engine.jl:650-862: rounds, executions in order, measured times --------------------------------
function collect_trace(engine::Cortex.InferenceEngine{<:B200ModelEngine}, variable_ids, total_ns)
    m = Cortex.get_model_engine(engine)
    n = ccall((:cxb_trace_get, LIB), Int64, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Int64), m.handle, C_NULL, C_NULL, 0)
    level = zeros(Int64, max(n, 1)); sids = zeros(Int64, max(n, 1)); ns = zeros(Int64, max(n, 1)); vars = zeros(Int64, max(n, 1))
    ccall((:cxb_trace_get, LIB), Int64, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Int64), m.handle, level, sids, n)
    ccall((:cxb_trace_get_times, LIB), Int64, (Ptr{Cvoid}, Ptr{Int64}, Int64), m.handle, ns, n)
    ccall((:cxb_trace_get_variables, LIB), Int64, (Ptr{Cvoid}, Ptr{Int64}, Int64), m.handle, vars, n)
    rounds = Cortex.TracedInferenceRound[]
    current = nothing; executions = Cortex.TracedInferenceExecution[]
    flush!() = isempty(executions) || push!(rounds, Cortex.TracedInferenceRound(engine, UInt64(sum(e.total_time_in_ns for e in executions)), executions))
    for k in 1:n
        key = level[k] >= 0 ? level[k] : -1                 # the final phase (marginals, then linked signals) is ONE round, :610-628
        if key != current
            flush!(); executions = Cortex.TracedInferenceExecution[]; current = key
        end
        signal = signal_view(m, sids[k], variant_of(m, sids[k]))
        # value_before_execution needs a snapshot taken before the request (the Python mirror does that); here: not recorded
        push!(executions, Cortex.TracedInferenceExecution(engine, Int(vars[k]) + 1, signal, UInt64(max(ns[k], 1)), nothing, Cortex.get_value(signal)))
    end
    flush!()
    request = Cortex.InferenceRequest(engine, variable_ids, Cortex.InferenceSignal[], falses(length(variable_ids)))
    return Cortex.TracedInferenceRequest(engine, UInt64(max(total_ns, 1)), request, rounds)
end

end # module
