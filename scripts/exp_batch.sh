set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/final_gpu_tests.log 2>&1; tail -3 gpurun_out/final_gpu_tests.log
python bench.py > gpurun_out/bench_r1c_n1.json 2> gpurun_out/bench_r1c_n1.err; tail -c 900 gpurun_out/bench_r1c_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 3040 -c 1100 --csv --log-file gpurun_out/r1c_bench_launches.csv python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3100 -c 2 -f -o gpurun_out/prof_r1c_bench_fused_f64_32768 python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
python bench.py --impl reference > gpurun_out/bench_r1c_ref.json 2> gpurun_out/bench_r1c_ref.err; tail -c 400 gpurun_out/bench_r1c_ref.json
python bench.py --size 8192 --dtype f32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1c_cfg2_8192_f32.json 2> gpurun_out/bench_r1c_cfg2.err; tail -c 400 gpurun_out/bench_r1c_cfg2_8192_f32.json
ls -la gpurun_out | tail -4
