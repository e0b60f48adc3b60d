set -x
T=$1
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log | head -3
timeout 300 python scratch/base_probe.py 100
timeout 300 python scratch/other_configs_probe.py 2>&1 | grep -v "^   modes" | head -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:base_step -s 25 -c 1 -f -o gpurun_out/${T}_prof_base python scratch/base_probe.py 30 > gpurun_out/${T}_ncu_base.log 2>&1
ncu -i gpurun_out/${T}_prof_base.ncu-rep --page raw --csv > gpurun_out/${T}_prof_base_raw.csv 2>/dev/null
ncu -i gpurun_out/${T}_prof_base.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${T}_prof_base_src.csv 2>/dev/null
