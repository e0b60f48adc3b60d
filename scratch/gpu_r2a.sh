set -x
T=$1
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
timeout 300 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cat gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
ABM_E2E_BATCHES=1 timeout 300 python bench.py --no-cpu-baseline --steps 30 > gpurun_out/${T}_bench_nb1.json 2>> gpurun_out/${T}_bench.err; python -c "import json;d=json.load(open('gpurun_out/${T}_bench_nb1.json'));print('NB=1 e2e', d['e2e'])"
timeout 300 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; cat gpurun_out/${T}_bench_ref.json | cut -c1-300
timeout 300 python scratch/other_configs_probe.py > gpurun_out/${T}_other.log 2>&1; cat gpurun_out/${T}_other.log
timeout 300 python scratch/c5_tile_probe.py > gpurun_out/${T}_c5tiles.log 2>&1; cat gpurun_out/${T}_c5tiles.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
