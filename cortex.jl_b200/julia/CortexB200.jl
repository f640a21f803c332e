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
const OK, ERR_NOT_PENDING, ERR_NO_RULE, ERR_OUT_OF_CONTRACT, ERR_BAD_ARG, ERR_UNSUPPORTED_ENGINE, ERR_CUDA, ERR_STATE, ERR_INTERNAL = 0:8
const F32, F64 = 0, 1
const KIND_UNSPECIFIED, KIND_M2F, KIND_M2V, KIND_PRODUCT, KIND_MARGINAL, KIND_JOINT = 0:5
const FAMILY_GAUSS_CANON, FAMILY_CATEGORICAL, FAMILY_GAUSS_MV, FAMILY_BETA, FAMILY_SUM,
      FAMILY_GAUSS_MP, FAMILY_GAMMA, FAMILY_POINT = 0:7
const RULE_NONE, RULE_GAUSS_OBS, RULE_GAUSS_RW, RULE_CAT_TABLE, RULE_POTTS, RULE_HMM_EMIT,
      RULE_GAUSS_MV_OBS, RULE_GAUSS_MV_RW, RULE_BETA_BERNOULLI, RULE_SCALE2, RULE_NORMAL_MEAN_FIELD,
      RULE_NORMAL_STRUCTURED = 0:11
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

# ---- custom wiring: what a user-defined AbstractDependencyResolver calls (test/inference_engine_tests.jl:597-621, 811-907) ----
const DEP_INTERMEDIATE, DEP_WEAK, DEP_NO_LISTEN, DEP_NO_CHECK_COMPUTED = 1, 2, 16, 32
create_signal!(e::B200InferenceEngine) = ccall((:cxb_create_signal, LIB), Int64, (Ptr{Cvoid},), e.handle)
function add_dependency!(e::B200InferenceEngine, signal::Int64, dependency::Int64; weak = false, listen = true,
                         check_computed = true, intermediate = false)
    flags = (intermediate ? DEP_INTERMEDIATE : 0) | (weak ? DEP_WEAK : 0) | (listen ? 0 : DEP_NO_LISTEN) |
            (check_computed ? 0 : DEP_NO_CHECK_COMPUTED)
    check(e.handle, ccall((:cxb_add_dependency, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int32), e.handle, signal, dependency, flags))
end
# set_variant!(signal, JointMarginal(factor_id, variable_ids)): the rule registered for the factor computes the signal
set_joint_marginal_variant!(e::B200InferenceEngine, signal::Int64, factor_id::Int) =
    check(e.handle, ccall((:cxb_set_signal_variant, LIB), Int32, (Ptr{Cvoid}, Int64, Int32, Int64, Int64),
                          e.handle, signal, KIND_JOINT, -1, factor_id - 1))
# delegate one id to a built-in resolver (Cortex.resolve_variable_dependencies!(DefaultDependencyResolver(), engine, id))
resolve_factor_dependencies!(e::B200InferenceEngine, resolver::Integer, factor_id::Int) =
    check(e.handle, ccall((:cxb_resolve_factor_dependencies, LIB), Int32, (Ptr{Cvoid}, Int32, Int64), e.handle, resolver, factor_id - 1))
resolve_variable_dependencies!(e::B200InferenceEngine, resolver::Integer, variable_id::Int) =
    check(e.handle, ccall((:cxb_resolve_variable_dependencies, LIB), Int32, (Ptr{Cvoid}, Int32, Int64), e.handle, resolver, variable_id - 1))

# value type of a variable's signals in a model that mixes types (VMP: NormalMeanPrecision / Gamma / observed Float64,
# test/runtests.jl:52-99 — in Julia the type travels with the value, the device needs it declared once)
function set_variable_families!(e::B200InferenceEngine, variable_ids::AbstractVector{<:Integer}, families::AbstractVector{<:Integer})
    ids = Int64[v - 1 for v in variable_ids]; fam = convert(Vector{Int32}, families)
    check(e.handle, ccall((:cxb_set_variable_families, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Int64}, Ptr{Int32}),
                          e.handle, length(ids), ids, fam))
end

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

# ---- structured engines of the other model families (same pattern: opaque handle, status codes, host arrays borrowed) ----

"Row shard of a Potts grid (BASELINE config 4). Neighbour shards live in other processes (one Julia process per GPU)."
mutable struct PottsGridShard
    handle::Ptr{Cvoid}
    rows::Int
    cols::Int
    K::Int
end
function PottsGridShard(rows::Int, cols::Int, K::Int, beta::Float64; dtype::Integer = F32, device::Integer = 0,
                        has_upper::Bool = false, has_lower::Bool = false)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:cxb_grid_create, LIB), Int32, (Int32, Int32, Int64, Int64, Int32, Float64, Int32, Int32, Ref{Ptr{Cvoid}}),
               device, dtype, rows, cols, K, beta, has_upper, has_lower, href)
    st == OK || error("cxb_grid_create failed with status $st (CUDA device required; there is no CPU fallback)")
    g = PottsGridShard(href[], rows, cols, K)
    finalizer(x -> ccall((:cxb_grid_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.handle), g)
    return g
end
grid_check(g::PottsGridShard, st::Int32) =
    st == OK || error(unsafe_string(ccall((:cxb_grid_last_error, LIB), Cstring, (Ptr{Cvoid},), g.handle)))
set_unary!(g::PottsGridShard, unary::Array{Float32, 3}) =            # K x cols x rows (column-major = [rows][cols][K] in C)
    grid_check(g, ccall((:cxb_grid_set_unary, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), g.handle, unary))
reset_messages!(g::PottsGridShard) = grid_check(g, ccall((:cxb_grid_reset_messages, LIB), Int32, (Ptr{Cvoid},), g.handle))
function sweep!(g::PottsGridShard)                                   # one synchronous sweep = protocol B of SURVEY Appendix B
    n = Ref{Int64}(0)
    grid_check(g, ccall((:cxb_grid_sweep, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), g.handle, n))
    return n[]
end
"128 bytes (two cudaIpcMemHandle_t) to hand to the row neighbours, e.g. with MPI.Sendrecv!."
function p2p_export(g::PottsGridShard)
    handles = Vector{UInt8}(undef, 128)
    grid_check(g, ccall((:cxb_grid_p2p_export, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}), g.handle, handles))
    return handles
end
"direction 0 = the shard above, 1 = the shard below; afterwards sweep! delivers the cut-edge messages itself (NVLink peer stores)."
p2p_connect!(g::PottsGridShard, direction::Integer, neighbour_handles::Vector{UInt8}) =
    grid_check(g, ccall((:cxb_grid_p2p_connect_ipc, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{UInt8}), g.handle, direction, neighbour_handles))
function get_marginals(g::PottsGridShard)
    out = Array{Float32, 3}(undef, g.K, g.cols, g.rows)
    grid_check(g, ccall((:cxb_grid_get_marginals, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), g.handle, out))
    return out
end

"Batch of discrete HMMs (BASELINE config 3): K = 64 register kernel, K >= 128 tcgen05 path, otherwise the generic kernel."
mutable struct HmmBatch
    handle::Ptr{Cvoid}
    B::Int
    T::Int
    K::Int
end
function HmmBatch(n_chains::Int, n_steps::Int, n_states::Int, n_symbols::Int; dtype::Integer = F32, device::Integer = 0)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:cxb_hmm_create, LIB), Int32, (Int32, Int32, Int64, Int64, Int32, Int32, Ref{Ptr{Cvoid}}),
               device, dtype, n_chains, n_steps, n_states, n_symbols, href)
    st == OK || error("cxb_hmm_create failed with status $st")
    m = HmmBatch(href[], n_chains, n_steps, n_states)
    finalizer(x -> ccall((:cxb_hmm_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.handle), m)
    return m
end
hmm_check(m::HmmBatch, st::Int32) = st == OK || error(unsafe_string(ccall((:cxb_hmm_last_error, LIB), Cstring, (Ptr{Cvoid},), m.handle)))
set_tables!(m::HmmBatch, transition::Matrix{Float64}, emission::Matrix{Float64}) =   # row-major on the C side: pass transposes
    hmm_check(m, ccall((:cxb_hmm_set_tables, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), m.handle,
                       permutedims(transition), permutedims(emission)))
set_observations!(m::HmmBatch, obs::Matrix{UInt8}) =                 # B x T (column-major = [T][B] in C)
    hmm_check(m, ccall((:cxb_hmm_set_observations, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}), m.handle, obs))
function update_marginals!(m::HmmBatch)
    n = Ref{Int64}(0)
    hmm_check(m, ccall((:cxb_hmm_update_marginals, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), m.handle, n))
    return n[]
end
function get_marginals(m::HmmBatch, t0::Int = 0, t1::Int = m.T)      # K x B x (t1 - t0)
    out = Array{Float32, 3}(undef, m.K, m.B, t1 - t0)
    hmm_check(m, ccall((:cxb_hmm_get_marginals, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}), m.handle, t0, t1, out))
    return out
end

"Arbitrary pairwise categorical graph, loopy BP by synchronous sweeps (BASELINE config 5). Ids are 0-based on the C side."
mutable struct PairwiseGraph
    handle::Ptr{Cvoid}
    n::Int
    K::Int
end
function PairwiseGraph(n_variables::Int, fac_u::Vector{Int64}, fac_v::Vector{Int64}, fac_table::Vector{Int32},
                       tables::Array{Float64, 3}; dtype::Integer = F32, device::Integer = 0)   # tables: K x K x n_tables, [x_hi, x_lo, t]
    K, n_tables = size(tables, 1), size(tables, 3)
    href = Ref{Ptr{Cvoid}}(C_NULL)
    st = ccall((:cxb_pairwise_create, LIB), Int32, (Int32, Int32, Int64, Int64, Int32, Int32, Ref{Ptr{Cvoid}}),
               device, dtype, n_variables, length(fac_u), K, n_tables, href)
    st == OK || error("cxb_pairwise_create failed with status $st")
    g = PairwiseGraph(href[], n_variables, K)
    finalizer(x -> ccall((:cxb_pairwise_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.handle), g)
    chk(s) = s == OK || error(unsafe_string(ccall((:cxb_pairwise_last_error, LIB), Cstring, (Ptr{Cvoid},), g.handle)))
    chk(ccall((:cxb_pairwise_set_graph, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int32}), g.handle, fac_u .- 1, fac_v .- 1, fac_table))
    chk(ccall((:cxb_pairwise_set_tables, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), g.handle, tables))
    return g
end
pw_check(g::PairwiseGraph, st::Int32) = st == OK || error(unsafe_string(ccall((:cxb_pairwise_last_error, LIB), Cstring, (Ptr{Cvoid},), g.handle)))
set_unary!(g::PairwiseGraph, unary::Matrix{Float32}) =               # K x n
    pw_check(g, ccall((:cxb_pairwise_set_unary, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), g.handle, unary))
reset_messages!(g::PairwiseGraph) = pw_check(g, ccall((:cxb_pairwise_reset_messages, LIB), Int32, (Ptr{Cvoid},), g.handle))
function get_marginals(g::PairwiseGraph)
    out = Matrix{Float32}(undef, g.K, g.n)
    pw_check(g, ccall((:cxb_pairwise_get_marginals, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), g.handle, out))
    return out
end
function sweep!(g::PairwiseGraph)
    n = Ref{Int64}(0)
    st = ccall((:cxb_pairwise_sweep, LIB), Int32, (Ptr{Cvoid}, Ref{Int64}), g.handle, n)
    st == OK || error(unsafe_string(ccall((:cxb_pairwise_last_error, LIB), Cstring, (Ptr{Cvoid},), g.handle)))
    return n[]
end

end # module
