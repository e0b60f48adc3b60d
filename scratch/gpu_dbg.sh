set -x
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_vf_gpu.py -m gpu -x -q -k "test_step_matches_reference_fixture and N20_R1200_walls and symmetric" 2>&1 | grep -v "^\s*$" | head -60
