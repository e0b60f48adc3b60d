set -x
T=$1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_vf_gpu.py -m gpu -x -q -k "not full_size and not step_host and not long_run and not two_devices and not adapts_to_crowding and not packed_state and not never_reads_back" > gpurun_out/${T}_memcheck_vf.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/${T}_memcheck_vf.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_base_gpu.py tests/test_cs_gpu.py -m gpu -x -q -k "not full_loop and not fused_step_equals and not recorder" > gpurun_out/${T}_memcheck_base.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/${T}_memcheck_base.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_vf_gpu.py -m gpu -x -q -k "degenerate or (matches_reference_fixture and symmetric and not line)" > gpurun_out/${T}_racecheck_vf.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/${T}_racecheck_vf.log
