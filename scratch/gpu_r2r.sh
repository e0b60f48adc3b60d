set -x
T=$1
(time timeout 900 python -m pytest tests/test_vf_gpu.py -m gpu -q) > gpurun_out/${T}_pytest.log 2>&1; tail -6 gpurun_out/${T}_pytest.log | head -5
ABM_BENCH_SWARM=0 ABM_BENCH_OTHER_CONFIGS=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json')); print('ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'parity', d['parity']['ok'], d['parity']['field_bits_differ'], 'fp64 frac', d['fp64_pairs_fraction'])"
timeout 200 python scratch/dense_probe.py 2>&1 | tail -12
