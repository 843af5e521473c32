set -x
export ATLJ_NO_GRAPH=1
python tools/profile_step.py 80 300 5 > gpurun_out/r2_profile_step.txt 2>&1 || exit 1
cat gpurun_out/r2_profile_step.txt
ncu --set full --clock-control none --import-source on -k 'regex:pair_tile2|kdk_assign|scan_lookback|cell_scatter|cell_rankgather|tile2_table' -s 1806 -c 6 -f -o gpurun_out/r2_step_full python tools/profile_step.py 80 300 5 > gpurun_out/r2_ncu_full.log 2>&1
tail -3 gpurun_out/r2_ncu_full.log
unset ATLJ_NO_GRAPH
ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 3 --warmup 3 --melt 100 --no-cpu --no-other > gpurun_out/r2_ncu_launches.log 2>&1
tail -2 gpurun_out/r2_ncu_launches.log
ls -la gpurun_out
