set -x
T=$1
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 300 python scratch/other_configs_probe.py 2>&1 | grep -v "^   modes"
ABM_VF_WARP_FOCAL=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:vf_step_warp -s 4 -c 1 -f -o gpurun_out/${T}_prof_warp python scratch/swarm_prof.py > gpurun_out/${T}_ncu_warp.log 2>&1
ncu -i gpurun_out/${T}_prof_warp.ncu-rep --page raw --csv > gpurun_out/${T}_prof_warp_raw.csv 2>/dev/null
ncu -i gpurun_out/${T}_prof_warp.ncu-rep --page source --csv > gpurun_out/${T}_prof_warp_src.csv 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s 60 -c 12 --csv --log-file gpurun_out/${T}_base_launches.csv python scratch/base_probe.py 30 > /dev/null 2>&1
cat gpurun_out/${T}_base_launches.csv | grep -v "^==" | cut -d, -f5,14,15 | tail -24
