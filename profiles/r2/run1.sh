mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_gpu_r2a.log 2>&1; echo "pytest rc=$?" > gpurun_out/run1_status.txt
tail -5 gpurun_out/pytest_gpu_r2a.log
python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?" >> gpurun_out/run1_status.txt
cat gpurun_out/bench_r2a.json | head -c 3000
tail -3 gpurun_out/bench_r2a.err
