set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/b10_gpu_tests.log 2>&1; tail -4 gpurun_out/b10_gpu_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_r1b_n2.json 2> gpurun_out/bench_r1b_n2.err; tail -c 600 gpurun_out/bench_r1b_n2.json
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 scripts/large_drain.py --size 65536 --blocks 1 --outlets 64 --out gpurun_out/cfg5_drain_65536_2gpu_64outlets.json > gpurun_out/cfg5.log 2>&1; tail -3 gpurun_out/cfg5.log
