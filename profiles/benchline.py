#!/usr/bin/env python
"""Condense bench.py output (stdin) to one short line per JSON record; pass other lines through (truncated)."""
import json
import sys

for line in sys.stdin:
    line = line.strip()
    if line.startswith("{"):
        try:
            d = json.loads(line)
        except Exception:
            print(line[:200])
            continue
        r = d.get("roofline", {})
        print(f"value={d.get('value'):.4g} ms/step={d.get('ms_per_step'):.4g} frac={r.get('frac')} kernel_ms={r.get('kernel_ms')} "
              f"e2e={d.get('e2e', {}).get('value'):.4g} launches={d.get('gpu_launches')} n_gpus={d.get('n_gpus')}")
    elif line:
        print(line[:200])
