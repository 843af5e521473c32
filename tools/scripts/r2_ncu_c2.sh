# ncu capture of the pair kernel at BASELINE config 2 (N = 32,000): where does one chunk's latency go?
export ATLJ_NO_GRAPH=1
python tools/profile_step.py 20 200 5 > gpurun_out/r2_c2_profile_step.txt 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k 'regex:pair_tile2|kdk_assign|scan_lookback|cell_scatter|cell_rankgather|tile2_table' -s 1206 -c 6 -f -o gpurun_out/r2_c2_full python tools/profile_step.py 20 200 5 > gpurun_out/r2_c2_ncu.log 2>&1
tail -2 gpurun_out/r2_c2_ncu.log
