# Round-end verification on one B200: GPU test-suite, smoke, one bench line per workload, the reference arm.
# usage (from the repo root, on the GPU box): bash profiles/run_all_benches.sh   (writes gpurun_out/bench_all.jsonl)
set -x
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/final_gpu_tests.log 2>&1; tail -3 gpurun_out/final_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
: > gpurun_out/bench_all.jsonl
timeout 300 python bench.py >> gpurun_out/bench_all.jsonl 2>gpurun_out/bench_default.err
timeout 300 python bench.py --dtype f64 --no-cpu-baseline >> gpurun_out/bench_all.jsonl 2>>gpurun_out/bench_default.err
timeout 300 python bench.py --workload potts_grid --steps 10 --warmup 3 >> gpurun_out/bench_all.jsonl 2>>gpurun_out/bench_default.err
timeout 300 python bench.py --workload hmm64 --steps 3 --warmup 3 >> gpurun_out/bench_all.jsonl 2>>gpurun_out/bench_default.err
timeout 300 python bench.py --workload hmm512 --steps 3 --warmup 3 >> gpurun_out/bench_all.jsonl 2>>gpurun_out/bench_default.err
timeout 300 python bench.py --workload powerlaw --steps 10 --warmup 3 >> gpurun_out/bench_all.jsonl 2>>gpurun_out/bench_default.err
timeout 300 python bench.py --workload chain1k --steps 5 --warmup 3 >> gpurun_out/bench_all.jsonl 2>>gpurun_out/bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>>gpurun_out/bench_default.err
while read l; do echo "$l" | python profiles/benchline.py; done < gpurun_out/bench_all.jsonl
cat gpurun_out/bench_reference.json | cut -c1-400
tail -5 gpurun_out/bench_default.err
