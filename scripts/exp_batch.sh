set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "all_variants and float32" > gpurun_out/b11_parity.log 2>&1; tail -3 gpurun_out/b11_parity.log
for v in 7 14; do timeout 300 python scripts/profile_iterate.py --size 8192 --dtype f32 --iters 200 --add-mm 100 --variant $v 2>&1 | tail -1; done > gpurun_out/b11_f32.txt
cat gpurun_out/b11_f32.txt
