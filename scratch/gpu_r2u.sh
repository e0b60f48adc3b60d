set -x
T=$1
for w in auto 0 1; do
if [ $w = auto ]; then unset ABM_VF_SYM_WIDE; else export ABM_VF_SYM_WIDE=$w; fi
ABM_BENCH_SWARM=0 ABM_BENCH_OTHER_CONFIGS=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench.err
python -c "
import json; d=json.load(open('gpurun_out/${T}_bench_$w.json')); print('wide=$w ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'vis', d['roofline']['visible_pair_fraction'], d['roofline']['kernel_launches_by_variant'])"
done
unset ABM_VF_SYM_WIDE
timeout 200 python scratch/dense_probe.py 2>&1 | grep "1370\|900px\|600px" 
