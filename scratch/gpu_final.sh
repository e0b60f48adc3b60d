set -x
T=$1
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench.json | cut -c1-3000; tail -3 gpurun_out/${T}_bench.err
timeout 240 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; cut -c1-200 gpurun_out/${T}_bench_ref.json
ABM_BENCH_SWARM=0 ABM_BENCH_OTHER_CONFIGS=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_launches.log 2>&1
ABM_BENCH_SWARM=0 ABM_BENCH_OTHER_CONFIGS=0 timeout 250 ncu --set full --clock-control none --import-source on -k regex:vf_step_sym -s 3 -c 1 -f -o gpurun_out/${T}_prof_sym python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1
ncu -i gpurun_out/${T}_prof_sym.ncu-rep --page raw --csv > gpurun_out/${T}_prof_sym_raw.csv 2>/dev/null
# the swarm's per-rank kernel: 8192 focal agents against the 65 536-record table (tile 0 of 8, emulated exchange)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vf_step_warp -s 27 -c 1 -f -o gpurun_out/${T}_prof_warp python scratch/c5_tile_probe.py 8 > gpurun_out/${T}_ncu_warp.log 2>&1
ncu -i gpurun_out/${T}_prof_warp.ncu-rep --page raw --csv > gpurun_out/${T}_prof_warp_raw.csv 2>/dev/null
# configs[2]: the fused foraging step
timeout 300 ncu --set full --clock-control none --import-source on -k regex:base_step -s 25 -c 1 -f -o gpurun_out/${T}_prof_base python scratch/base_probe.py 30 > gpurun_out/${T}_ncu_base.log 2>&1
ncu -i gpurun_out/${T}_prof_base.ncu-rep --page raw --csv > gpurun_out/${T}_prof_base_raw.csv 2>/dev/null
python scratch/ncu_keys.py gpurun_out/${T}_prof_warp_raw.csv | head -12
