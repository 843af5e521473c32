# A/B of builds at BASELINE config 2 (N = 32,000) and at 2.05 M atoms; parity tests on every variant
for lib in libatlj.so $(ls examples_b200 | grep '^libatlj_.*\.so$' | grep -v stats); do
  echo "== $lib"
  ATLJ_LIB=$PWD/examples_b200/$lib python tools/profile_step.py 20 500 400
  ATLJ_LIB=$PWD/examples_b200/$lib python tools/profile_step.py 20 500 400
  ATLJ_LIB=$PWD/examples_b200/$lib python tools/profile_step.py 22 500 400
  ATLJ_LIB=$PWD/examples_b200/$lib python tools/profile_step.py 80 300 40
done 2>&1 | tee gpurun_out/r2_ab_c2.txt
for lib in $(ls examples_b200 | grep '^libatlj_.*\.so$' | grep -v stats); do
  echo "== parity $lib"
  ATLJ_LIB=$PWD/examples_b200/$lib python -m pytest tests -m gpu -x -q 2>&1 | tail -5
done 2>&1 | tee -a gpurun_out/r2_ab_c2.txt
