set -x
timeout 200 python -m pytest tests/test_facade_gpu.py tests/test_vf_gpu.py -q -m gpu -k "tiled or swarm or multi or two_gpus" 2>&1 | tail -5
timeout 200 python tests/multigpu_check.py 2>&1 | tail -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r1e_bench_n2.json 2> gpurun_out/r1e_bench_n2.err; cat gpurun_out/r1e_bench_n2.json | cut -c1-400
