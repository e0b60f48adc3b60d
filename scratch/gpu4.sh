set -x
T=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/multigpu_check.py 65536 2>&1 | grep -v "^W\|^\*\*\|OMP_NUM" | tail -16
timeout 600 $TR --master-port 29513 bench.py --gpus 4 --steps 30 > gpurun_out/${T}_bench_n4.json 2> gpurun_out/${T}_bench_n4.err; tail -5 gpurun_out/${T}_bench_n4.err
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_n4.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"])
print("parity", {k:v for k,v in d["parity"].items() if k not in ("oracle","tolerance")})
print("swarm", d["swarm"])
PY
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 | cut -c1-200
