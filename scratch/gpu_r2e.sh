set -x
T=$1
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log | head -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -1
timeout 300 python scratch/other_configs_probe.py 2>&1 | grep -v "^   modes"
ABM_BASE_SEPARATE_PHASES=1 timeout 300 python scratch/base_probe.py 100
timeout 300 python scratch/base_probe.py 100
for F in auto 4 8; do
  if [ $F = auto ]; then unset ABM_VF_WARP_FOCAL; else export ABM_VF_WARP_FOCAL=$F; fi
  echo "== focal per CTA: $F"; timeout 300 python scratch/c5_tile_probe.py 2>&1 | tail -2
done
unset ABM_VF_WARP_FOCAL
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -s 40 -c 4 --csv --log-file gpurun_out/${T}_base_launches.csv python scratch/base_probe.py 30 > /dev/null 2>&1
grep -v "^==" gpurun_out/${T}_base_launches.csv | cut -d, -f5,13,15 | tail -8
