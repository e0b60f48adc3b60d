set -x
T=$1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_base_gpu.py -m gpu -x -q -k "resolution or lidar or radii" > gpurun_out/${T}_memcheck_base.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/${T}_memcheck_base.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_vf_gpu.py -m gpu -x -q -k "line_following or one_chunk or large_sparse" > gpurun_out/${T}_memcheck_vf.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/${T}_memcheck_vf.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_base_gpu.py -m gpu -x -q -k "heterogeneous_resolutions and one_cta" > gpurun_out/${T}_racecheck_base.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/${T}_racecheck_base.log
