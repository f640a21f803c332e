#!/bin/bash
# tuning sweep of the chains kernel (tile = time steps in flight, block = chains per CTA); prints kernel ms per variant
for dt in f32 f64; do
  for blk in 32 64 128 256; do
    for tile in 2 4 8 16; do
      out=$(CXB_CHAINS_TILE=$tile CXB_CHAINS_BLOCK=$blk timeout 120 python bench.py --dtype $dt --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1)
      [ -z "$out" ] && continue
      echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$dt tile=$tile block=$blk kernel_ms=%.4f frac=%.3f' % (d['roofline']['kernel_ms'], d['roofline']['frac']))" 2>/dev/null
    done
  done
done
