set -x
T=$1
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log | head -3
for cfg in "4 256" "8 512" "4 512" "8 256"; do
  set -- $cfg; export ABM_VF_WARP_FOCAL=$1 ABM_VF_WARP_THREADS=$2
  echo "== focal per CTA $1, threads $2"; timeout 300 python scratch/c5_tile_probe.py 2>&1 | tail -2
done
unset ABM_VF_WARP_FOCAL ABM_VF_WARP_THREADS
timeout 300 python scratch/other_configs_probe.py 2>&1 | grep "C2"
ABM_VF_ONE_STEP_PER_LAUNCH=1 timeout 300 python scratch/other_configs_probe.py 2>&1 | grep "C2"
