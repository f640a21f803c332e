run() { echo "== $*"; env "$@" python bench.py --workload hmm512 --hmm-chains 256 --hmm-steps 2000 --steps 2 --warmup 1 --no-cpu-baseline --others none 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('ms_per_step', d.get('ms_per_step'), 'value', d.get('value'))
"; }
run A=1
run CXB_HMM_TC_CLUSTER=8
run CXB_HMM_TC_CLUSTER=4
run CXB_HMM_TC_NT=64
run CXB_HMM_TC_NT=64 CXB_HMM_TC_CLUSTER=8
run CXB_HMM_TC_PIECES=2
run CXB_HMM_TC_PIECES=2 CXB_HMM_TC_CLUSTER=8
