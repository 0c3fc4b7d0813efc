set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "adversarial or without_a_positive" > gpurun_out/b12_edge.log 2>&1; tail -15 gpurun_out/b12_edge.log
