set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_stripes.py -q -m gpu -k nvlink > gpurun_out/b15_stripes4.log 2>&1; tail -3 gpurun_out/b15_stripes4.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/bench_r1d_n4.json 2> gpurun_out/bench_r1d_n4.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_r1d_n4.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['ms_per_step'], d['gpu_launches'])"
