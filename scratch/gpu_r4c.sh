set -x
T=r4c
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 150 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_base_gpu.py -m gpu -x -q -k "collision_phase_matches_oracle or full_step_in_lockstep or collision_lidar" > gpurun_out/${T}_racecheck_base.log 2>&1; echo rc=$?; tail -4 gpurun_out/${T}_racecheck_base.log
timeout 100 python scratch/base_probe.py 200 2>&1 | tail -3
