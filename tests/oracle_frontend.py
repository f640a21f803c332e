"""Oracle-only pieces of the host frontend (TEST INFRASTRUCTURE, not shipped in the package).

The device engine runs registered rule kernels and table-driven traversals; the CPU oracle can additionally call back
into Python, which is how the reference's own tests (user-defined Julia rules, arbitrary traversal callbacks) are ported
literally. Both live here so that the product package holds nothing that only works on the oracle.
"""
from __future__ import annotations

import numpy as np

from tests._pkg import pkg as C

capi = C.capi


class CallbackProcessor(C.AbstractInferenceRequestProcessor):
    """User-defined Python rules with the reference signature ``(engine, variant, signal, dependencies) -> value``
    (src/inference_engine.jl:351-477), evaluated by the oracle through ``cxo_set_rule_callback``."""

    def __init__(self, value_dim: int = 1):
        self.value_dim = value_dim
        self.family = capi.FAMILY_SUM

    def _missing(self, name):
        raise C.NoRuleError(f"The function `{name}` is not implemented for the processor of type {type(self).__name__}")

    def compute_message_to_variable(self, engine, variant, signal, dependencies):
        self._missing("compute_message_to_variable!")

    def compute_message_to_factor(self, engine, variant, signal, dependencies):
        self._missing("compute_message_to_factor!")

    def compute_individual_marginal(self, engine, variant, signal, dependencies):
        self._missing("compute_individual_marginal!")

    def compute_product_of_messages(self, engine, variant, signal, dependencies):
        self._missing("compute_product_of_messages!")

    def compute_joint_marginal(self, engine, variant, signal, dependencies):
        self._missing("compute_joint_marginal!")

    def install(self, engine):
        """Called by InferenceEngine.__init__ (processors may hook into the engine)."""
        if not hasattr(engine.api, "set_rule_callback"):
            raise C.NoRuleError("Python rule callbacks are an oracle facility: the device engine runs registered rule kernels "
                                "(RuleProcessor)")
        processor, st, dim = self, engine.store, engine.store.value_dim

        def cb(_user, sid, kind, var, fac, ndeps, dep_ids, dep_values, out):
            try:
                signal = C.Signal(st, sid)
                deps = [C.Signal(st, dep_ids[i]) for i in range(ndeps)]
                variant = C.get_variant(signal)
                fn = {capi.KIND_M2V: processor.compute_message_to_variable,
                      capi.KIND_M2F: processor.compute_message_to_factor,
                      capi.KIND_MARGINAL: processor.compute_individual_marginal,
                      capi.KIND_PRODUCT: processor.compute_product_of_messages,
                      capi.KIND_JOINT: processor.compute_joint_marginal}.get(kind)
                if fn is None:
                    raise C.NoRuleError(f"Unprocessed signal variant: {variant}")  # src/inference_engine.jl:506
                val = np.atleast_1d(np.asarray(fn(engine, variant, signal, deps), dtype=np.float64)).ravel()
                for k in range(dim):
                    out[k] = val[k] if k < val.size else 0.0
                return capi.OK
            except Exception as e:  # noqa: BLE001 - surfaced to the caller of update_marginals
                engine._callback_error = e
                return capi.ERR_NO_RULE

        engine._callback_error = None
        engine._cb = capi.RULE_CB(cb)  # keep alive
        st.check(engine.api.set_rule_callback(st.h, engine._cb, None))


def process_dependencies_callback(f, signal, *, retry: bool = False) -> bool:
    """process_dependencies!(f, signal; retry) with an arbitrary Python callback (oracle: ``cxo_process_dependencies``)."""
    st = signal.store
    cb = capi.VISIT_CB(lambda _u, d: 1 if f(C.Signal(st, d)) else 0)
    r = st.api.process_dependencies(st.h, signal.sid, int(retry), cb, None)
    if r < 0:
        st.check(capi.ERR_BAD_ARG)
    return bool(r)
