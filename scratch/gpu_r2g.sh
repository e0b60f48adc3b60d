set -x
T=$1
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log | head -3
for F in auto 4; do
  if [ $F = auto ]; then unset ABM_VF_WARP_FOCAL; else export ABM_VF_WARP_FOCAL=$F; fi
  echo "== focal per CTA: $F"; timeout 300 python scratch/c5_tile_probe.py 2>&1 | tail -2
done
unset ABM_VF_WARP_FOCAL
timeout 300 python scratch/base_probe.py 100
timeout 300 ncu --set full --clock-control none --import-source on -k regex:base_step -s 25 -c 1 -f -o gpurun_out/${T}_prof_base python scratch/base_probe.py 30 > gpurun_out/${T}_ncu_base.log 2>&1
ncu -i gpurun_out/${T}_prof_base.ncu-rep --page raw --csv > gpurun_out/${T}_prof_base_raw.csv 2>/dev/null
ncu -i gpurun_out/${T}_prof_base.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${T}_prof_base_src.csv 2>/dev/null
ls -la gpurun_out/
