# A/B of pair-kernel builds (ATLJ_LIB selects the library); prints one line per run
for lib in libatlj.so $(ls examples_b200 | grep '^libatlj_.*\.so$' | grep -v stats); do
  echo "== $lib"
  ATLJ_LIB=$PWD/examples_b200/$lib python tools/profile_step.py 80 300 40
  ATLJ_LIB=$PWD/examples_b200/$lib python tools/profile_step.py 80 300 40
done 2>&1 | tee gpurun_out/r2_ab_pair.txt
for lib in $(ls examples_b200 | grep '^libatlj_.*\.so$' | grep -v stats); do
  echo "== parity $lib"
  ATLJ_LIB=$PWD/examples_b200/$lib python -m pytest tests/test_force_parity.py tests/test_device_sim.py tests/test_slab.py -m gpu -x -q 2>&1 | tail -3
done 2>&1 | tee -a gpurun_out/r2_ab_pair.txt
