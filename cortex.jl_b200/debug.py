"""Visual debugging of the (device-resident) signal graph: a GraphViz DOT dump of one signal's neighbourhood.

Counterpart of `GraphViz.load(::Cortex.Signal; max_depth, max_dependencies, max_listeners, variant_to_string_fn,
show_value, show_variant, show_listeners)` (ext/GraphVizExt/GraphVizExt.jl:246-339) with the same conventions: pending
signals orange, computed signals green, others white; weak dependencies dashed, intermediate grey, fresh blue, fresh +
intermediate cadet blue; listener edges solid black (listening) or dotted grey (not listening); the same summaries when
`max_depth`, `max_dependencies` or `max_listeners` cut the picture ("N dependencies", "N more dependencies, k weak, ...",
"N more listeners, k active, ..."). Returns the DOT source as a string (the image has no GraphViz binding); everything is
read back through the C ABI, so it shows the state the kernels actually left on the device."""
from __future__ import annotations

from typing import Callable, List

from . import _capi as capi
from .inference_signal import (Signal, get_dependencies, get_dependency_props, get_listeners, get_listenmask, get_value, get_variant,
                               is_computed, is_pending)


def _esc(text: str) -> str:
    return text.replace('"', "'")


def _edge_style(weak: bool, intermediate: bool, fresh: bool):  # get_edge_style, GraphVizExt.jl:27-38
    color = "cadetblue" if (fresh and intermediate) else ("blue" if fresh else ("gray" if intermediate else "black"))
    return ("dashed" if weak else "solid"), color


def signal_to_dot(signal: Signal, max_depth: int = 2, max_dependencies: int = 10, max_listeners: int = 10,
                  variant_to_string_fn: Callable = str, show_value: bool = True, show_variant: bool = True,
                  show_listeners: bool = True) -> str:
    lines: List[str] = ["digraph G {", '  rankdir="RL"']
    counter = [0]

    def node(sig: Signal, name: str, title: str, level: int, depth_left: int) -> None:
        pending, computed = is_pending(sig), is_computed(sig)
        fill = "orange" if pending else ("palegreen" if computed else "white")  # get_node_attributes, :78-87
        rows = [title + (f" (depth {level})" if level > 0 else "")]
        if show_value:  # :41-49
            rows.append("Current value: " + (str(get_value(sig)) if computed else "UndefValue()") + (" (pending)" if pending else ""))
        if show_variant:  # :52-57
            rows.append("Variant: " + variant_to_string_fn(get_variant(sig)))
        deps, props = get_dependencies(sig), get_dependency_props(sig)
        edges: List[str] = []
        if not deps:
            rows.append("No dependencies")  # :398-399
        elif depth_left <= 0:
            rows.append(f"{len(deps)} dependencies")  # format_dependencies_summary, :67-76
            rows.append("Use `max_depth` to render more dependencies")
        else:
            shown = min(len(deps), max_dependencies)
            for i in range(shown):  # format_signal_dependencies, :447-496
                rows.append(f"- dependency {i + 1}")
                counter[0] += 1
                child = f"{name}dep{i + 1}_{counter[0]}"
                node(deps[i], child, "Dependency", level + 1, depth_left - 1)
                nib = props[i]
                style, color = _edge_style(bool(nib & capi.NIB_WEAK), bool(nib & capi.NIB_INTERMEDIATE), bool(nib & capi.NIB_FRESH))
                edges.append(f'  {child} -> {name} [style="{style}" color="{color}"];')
            if len(deps) > max_dependencies:  # calculate_dependency_stats / format_dependency_stats, :96-127
                rest = range(shown, len(deps))
                stats = [f"{len(deps) - shown} more dependencies"]
                for label, count in (("weak", sum(bool(props[i] & capi.NIB_WEAK) for i in rest)),
                                     ("intermediate", sum(bool(props[i] & capi.NIB_INTERMEDIATE) for i in rest)),
                                     ("fresh", sum(bool(props[i] & capi.NIB_FRESH) for i in rest)),
                                     ("pending", sum(is_pending(deps[i]) for i in rest))):
                    if count > 0:
                        stats.append(f"{count} {label}")
                rows.append("...")
                rows.append(", ".join(stats))
                rows.append("Use `max_dependencies` to show more dependencies")
        label = "\\n".join(_esc(r) for r in rows)
        lines.append(f'  {name} [label="{label}", style="filled", fillcolor="{fill}", shape="box"];')
        lines.extend(edges)

    node(signal, "main", "MainSignal", 0, max_depth)
    if show_listeners:  # listeners of the main signal only, without their own neighbourhood, :205-243
        listeners, mask = get_listeners(signal), get_listenmask(signal)
        shown = min(len(listeners), max_listeners)
        for k in range(shown):
            name = f"listener{k + 1}"
            node(listeners[k], name, "Listener", 0, 0)
            style, color = ("solid", "black") if mask[k] else ("dotted", "gray40")  # LISTENER_EDGE_STYLES, :22-24
            lines.append(f'  main -> {name} [style="{style}" color="{color}"];')
        if len(listeners) > max_listeners:  # :138-165
            active = sum(bool(mask[k]) for k in range(shown, len(listeners)))
            inactive = len(listeners) - shown - active
            stats = [f"{len(listeners) - shown} more listeners"]
            if active > 0:
                stats.append(f"{active} active")
            if inactive > 0:
                stats.append(f"{inactive} inactive")
            text = _esc(", ".join(stats) + "\\nUse `max_listeners` to show more listeners")
            lines.append(f'  more_listeners [label="{text}", shape="plaintext"];')
            lines.append('  main -> more_listeners [style="dotted" color="gray40"];')
    lines.append("}")
    return "\n".join(lines)
