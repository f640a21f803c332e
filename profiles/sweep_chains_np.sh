#!/bin/bash
# non-pipelined variants of the chains kernel (CXB_CHAINS_PIPE=0)
for dt in f32 f64; do
  for blk in 64 128 256; do
    for tile in 4 8 16; do
      out=$(CXB_CHAINS_PIPE=0 CXB_CHAINS_TILE=$tile CXB_CHAINS_BLOCK=$blk timeout 120 python bench.py --dtype $dt --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1)
      [ -z "$out" ] && continue
      echo "$out" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$dt np tile=$tile block=$blk kernel_ms=%.4f frac=%.3f' % (d['roofline']['kernel_ms'], d['roofline']['frac']))" 2>/dev/null
    done
  done
done
