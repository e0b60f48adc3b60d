set -x
(time python -m pytest tests -m gpu -x -q) > gpurun_out/$1_pytest.log 2>&1; tail -3 gpurun_out/$1_pytest.log
python bench.py --no-cpu-baseline > gpurun_out/$1_bench.json 2> gpurun_out/$1_bench.err; python - <<PY
import json
d=json.load(open("gpurun_out/$1_bench.json"))
print(d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d.get("fp64_pairs_fraction"))
PY
