mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention" -p no:cacheprovider > gpurun_out/pytest_fa2.log 2>&1; echo "fa pytest rc=$?" > gpurun_out/run3_status.txt
tail -8 gpurun_out/pytest_fa2.log
timeout 300 python profiles/flash_bench.py > gpurun_out/flash_v2.json 2> gpurun_out/flash_v2.err; echo "flash rc=$?" >> gpurun_out/run3_status.txt
cat gpurun_out/flash_v2.json; tail -3 gpurun_out/flash_v2.err
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/run3_status.txt
tail -15 gpurun_out/pytest_gpu_r2b.log
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?" >> gpurun_out/run3_status.txt
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2b.json"))
    print("bench", d["value"], d["e2e"]["value"], d["launches_per_step"], d["roofline"]["frac"], d["roofline"].get("frac_in_graph"), d["roofline_tensor"]["frac"], d["kernel_time_breakdown_ms"])
except Exception as e: print("bench failed", e)
PY
tail -3 gpurun_out/bench_r2b.err
