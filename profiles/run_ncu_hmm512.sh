# ncu evidence for the K = 512 HMM tensor-core path (one B200): launch list of a short bench command, then one full capture
# of the paired step kernel. Run only after the same command has exited 0 without ncu. Writes under gpurun_out/.
set -x
TAG=${1:-r02}
CMD="python bench.py --workload hmm512 --hmm-steps 40 --steps 1 --warmup 3 --no-cpu-baseline --others none"
timeout 200 $CMD > gpurun_out/hmm512_plain_$TAG.log 2>&1 || exit 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_hmm512_launches.csv $CMD > gpurun_out/ncu_l_hmm512.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_hmm_tc_step_pair -s 60 -c 2 -f -o gpurun_out/${TAG}_prof_hmm512 $CMD > gpurun_out/ncu_f_hmm512.log 2>&1
tail -3 gpurun_out/ncu_f_hmm512.log
ls -la gpurun_out/
