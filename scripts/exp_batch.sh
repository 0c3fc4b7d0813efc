set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/b4_gpu_tests.log 2>&1; tail -8 gpurun_out/b4_gpu_tests.log
