timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r02_gpu_tests_final.log 2>&1; tail -3 gpurun_out/r02_gpu_tests_final.log
date +%s > gpurun_out/bench_t0
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err; echo "bench rc=$?"
echo "bench seconds: $(( $(date +%s) - $(cat gpurun_out/bench_t0) ))"
tail -c 600 gpurun_out/r02_bench_final_n1.err
