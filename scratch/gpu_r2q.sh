set -x
T=$1
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log | head -3
timeout 300 python scratch/other_configs_probe.py 2>&1 | grep -v "   modes"
