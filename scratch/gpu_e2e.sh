set -x
T=$1
export ABM_BENCH_SWARM=0 ABM_BENCH_OTHER_CONFIGS=0
timeout 300 python -m pytest tests/test_vf_gpu.py -m gpu -x -q -k "packed or host" 2>&1 | tail -3
for api in step_host; do for nb in 1 2 3; do
ABM_E2E_API=$api ABM_E2E_BATCHES=$nb timeout 200 python bench.py --steps 30 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$api', $nb, 'value %.4g e2e %.4g ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d['e2e'].get('host_enqueue_s'))"
done; done
