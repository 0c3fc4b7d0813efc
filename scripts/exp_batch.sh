set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "all_variants or golden" > gpurun_out/b2_parity.log 2>&1; tail -3 gpurun_out/b2_parity.log
for v in 10 11 13; do timeout 300 python scripts/profile_iterate.py --size 8192 --dtype f64 --iters 200 --variant $v 2>&1 | tail -1; done > gpurun_out/b2_f64.txt
cat gpurun_out/b2_f64.txt
timeout 300 python scripts/timeline.py --size 8192 --variant 13 --cta 70 --out gpurun_out/tl_v13.npy > gpurun_out/tl_v13.txt 2>&1
