# programmatic dependent launch in graph replay: where does it stop paying? (ATLJ_PDL mask 46 vs 0)
for nc in 26 32 40 50 63; do
for pdl in 0 46; do
  echo "== nc=$nc ATLJ_PDL=$pdl"
  ATLJ_PDL=$pdl python tools/profile_step.py $nc 300 200 2>&1 | tail -1
done
done 2>&1 | tee gpurun_out/r2_ab_pdl2.txt
