"""Visual debugging of the (device-resident) signal graph: a GraphViz DOT dump of one signal's neighbourhood.

Counterpart of `GraphViz.load(::Cortex.Signal; max_depth, max_dependencies, max_listeners, ...)`
(ext/GraphVizExt/GraphVizExt.jl:292-339): same colour conventions — pending signals orange, computed signals green,
others white; weak dependencies dashed, intermediate grey, fresh blue, fresh + intermediate cadet blue; listener edges
solid black (listening) or dotted grey. Returns the DOT source as a string (the image has no GraphViz binding);
everything is read back through the C ABI, so it shows the state the kernels actually left on the device."""
from __future__ import annotations

from typing import Callable, List

from . import _capi as capi
from .inference_signal import (Signal, get_dependencies, get_dependency_props, get_listeners, get_value, get_variant, is_computed,
                               is_pending)


def _node(sig: Signal, name: str, title: str, variant_to_string_fn: Callable, show_value: bool, show_variant: bool) -> str:
    fill = "orange" if is_pending(sig) else ("palegreen" if is_computed(sig) else "white")
    rows = [title]
    if show_variant:
        rows.append(variant_to_string_fn(get_variant(sig)))
    if show_value:
        rows.append("value: " + (str(get_value(sig)) if is_computed(sig) else "UndefValue()"))
    label = "\\n".join(r.replace('"', "'") for r in rows)
    return f'  {name} [label="{label}", style="filled", fillcolor="{fill}", shape="box"];'


def signal_to_dot(signal: Signal, max_depth: int = 2, max_dependencies: int = 10, max_listeners: int = 10,
                  variant_to_string_fn: Callable = str, show_value: bool = True, show_variant: bool = True,
                  show_listeners: bool = True) -> str:
    lines: List[str] = ["digraph G {", '  rankdir="RL"']
    seen = {}

    def visit(sig: Signal, level: int) -> str:
        if sig.sid in seen:
            return seen[sig.sid]
        name = "main" if level == 0 else f"s{sig.sid}"
        seen[sig.sid] = name
        lines.append(_node(sig, name, "MainSignal" if level == 0 else f"Signal {sig.sid}", variant_to_string_fn, show_value, show_variant))
        if level >= max_depth:
            return name
        deps, props = get_dependencies(sig), get_dependency_props(sig)
        for k, (dep, nib) in enumerate(zip(deps, props)):
            if k >= max_dependencies:
                lines.append(f'  more_{name} [label="... {len(deps) - max_dependencies} more", shape="plaintext"];')
                lines.append(f"  more_{name} -> {name} [style=dotted];")
                break
            dn = visit(dep, level + 1)
            inter, weak, fresh = nib & capi.NIB_INTERMEDIATE, nib & capi.NIB_WEAK, nib & capi.NIB_FRESH
            color = "cadetblue" if (fresh and inter) else ("blue" if fresh else ("gray" if inter else "black"))
            style = "dashed" if weak else "solid"
            lines.append(f'  {dn} -> {name} [color="{color}", style="{style}"];')
        if show_listeners and level == 0:
            for k, lis in enumerate(get_listeners(sig)):
                if k >= max_listeners:
                    break
                ln = visit(lis, max_depth)  # listeners are shown without their own neighbourhood
                listening = any(d.sid == sig.sid for d in get_dependencies(lis))
                lines.append(f'  {name} -> {ln} [color="{"black" if listening else "gray"}", style="{"solid" if listening else "dotted"}"];')
        return name

    visit(signal, 0)
    lines.append("}")
    return "\n".join(lines)
