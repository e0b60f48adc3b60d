export ABM_BENCH_SWARM=0 ABM_BENCH_OTHER_CONFIGS=0
timeout 600 python -m pytest tests/test_vf_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python bench.py --steps 60 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ms %.4f' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'], 'e2e %.4g' % d['e2e']['value'], 'parity ok', d['parity']['ok'], d['parity']['field_bits_differ'])"
timeout 100 python scratch/dense_probe.py 2>&1 | tail -8
