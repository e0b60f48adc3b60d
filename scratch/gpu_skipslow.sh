export ABM_BENCH_SWARM=0 ABM_BENCH_OTHER_CONFIGS=0
for v in 0 1; do
  if [ $v = 1 ]; then export ABM_VF_DEBUG_SKIP_SLOW=1; fi
  timeout 200 python bench.py --steps 40 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('skip_slow=$v ms %.4f' % d['ms_per_step'], 'parity ok', d['parity']['ok'])"
done
