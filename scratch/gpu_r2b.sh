set -x
T=$1
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
for F in auto 1 2 4 8; do
  if [ $F = auto ]; then unset ABM_VF_WARP_FOCAL; else export ABM_VF_WARP_FOCAL=$F; fi
  echo "== focal per CTA: $F"; timeout 300 python scratch/c5_tile_probe.py 2>&1 | tail -3
done
unset ABM_VF_WARP_FOCAL
timeout 300 python scratch/other_configs_probe.py 2>&1 | grep -v "^   modes"
