timeout 90 python -m pytest tests/test_device_parity.py tests/test_fullsize_parity.py -x -q -m gpu -k "tensor_core or hmm_sparse or hmm_kernels_against" 2>&1 | tail -3
run() { echo "== $*"; timeout 45 env "$@" python bench.py --workload hmm512 --hmm-chains 256 --hmm-steps 2000 --steps 2 --warmup 1 --no-cpu-baseline --others none 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step', d.get('ms_per_step'), 'value', d.get('value'))
"; }
run A=1
run CXB_HMM_TC_FLAGS=0
run CXB_HMM_TC_NT=64
