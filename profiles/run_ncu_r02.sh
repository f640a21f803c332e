# Round-2 ncu evidence on one B200. Each profiled command is first run plain (must exit 0), then as a launch list
# (gpu__time_duration.sum, --clock-control none), then ONE --set full capture of its dominant kernel. Writes under gpurun_out/.
set -x
# (1) default bench workload: the Potts grid (2 warm-up steps + 1 step of 4 sweeps keeps the capture short)
CMD="python bench.py --steps 1 --warmup 3 --sweeps 4 --others none --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/r02_potts_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_potts_launches.csv $CMD > gpurun_out/r02_ncu_l_potts.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_potts_sweep -s 8 -c 1 -f -o gpurun_out/r02_prof_potts $CMD > gpurun_out/r02_ncu_f_potts.log 2>&1
# (2) the single entry point on a chain batch: set_values_prepared + update_marginals_prepared -> k_chain_plan
CMD="python bench.py --workload chains_engine --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/r02_chains_engine_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_chains_engine_launches.csv $CMD > gpurun_out/r02_ncu_l_ce.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain_plan -s 3 -c 1 -f -o gpurun_out/r02_prof_chain_plan $CMD > gpurun_out/r02_ncu_f_ce.log 2>&1
# (3) the memoised replay of a protocol-B sweep on the 1M-variable power-law graph
CMD="python bench.py --workload powerlaw_engine --pl-vars 1000000 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/r02_ple_plain.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ple_launches.csv $CMD > gpurun_out/r02_ncu_l_ple.log 2>&1
for f in r02_prof_potts r02_prof_chain_plan; do python profiles/extract_ncu.py gpurun_out/$f.ncu-rep > gpurun_out/$f.json 2>/dev/null; done
ls -la gpurun_out/ | tail -12
