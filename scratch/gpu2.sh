set -x
T=$1
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/multigpu_check.py 4096 2>&1 | grep -v "^W\|^\*\*" | tail -20
timeout 300 $TR --master-port 29512 tests/multigpu_check.py 65536 2>&1 | grep -v "^W\|^\*\*" | tail -20
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 30 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; tail -5 gpurun_out/${T}_bench_n2.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_n2.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"])
print("parity", {k:v for k,v in d["parity"].items() if k not in ("oracle","tolerance")})
print("swarm", d["swarm"])
PY
