set -x
T=r4d
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_vf_gpu.py -m gpu -x -q -k "(matches_reference_fixture and (warp or onesided) and not line) or warp_kernel_on_a_sparse or (degenerate and (warp or onesided))" > gpurun_out/${T}_racecheck_vf_warp_onesided.log 2>&1; echo rc=$?; tail -6 gpurun_out/${T}_racecheck_vf_warp_onesided.log
timeout 100 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_vf_gpu.py -m gpu -x -q -k "small_runs_step_inside_one_launch" > gpurun_out/${T}_racecheck_vf_multistep.log 2>&1; echo rc=$?; tail -6 gpurun_out/${T}_racecheck_vf_multistep.log
