mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention" -p no:cacheprovider > gpurun_out/pytest_fa3.log 2>&1; echo "fa pytest rc=$?" > gpurun_out/run4_status.txt
tail -6 gpurun_out/pytest_fa3.log
timeout 300 python profiles/flash_bench.py > gpurun_out/flash_v3.json 2> gpurun_out/flash_v3.err; echo "flash rc=$?" >> gpurun_out/run4_status.txt
cat gpurun_out/flash_v3.json; tail -3 gpurun_out/flash_v3.err
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "gemm or conv or split" -p no:cacheprovider > gpurun_out/pytest_gemm_cl.log 2>&1; echo "gemm pytest rc=$?" >> gpurun_out/run4_status.txt
tail -12 gpurun_out/pytest_gemm_cl.log
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_determinism_gpu.py tests/test_device_stream_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_mod_cl.log 2>&1; echo "modules pytest rc=$?" >> gpurun_out/run4_status.txt
tail -12 gpurun_out/pytest_mod_cl.log
for cl in 1 0; do
L2D_SPLITK_CLUSTER=$cl timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_r2c_cl$cl.json 2> gpurun_out/bench_r2c_cl$cl.err; echo "bench cl$cl rc=$?" >> gpurun_out/run4_status.txt
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2c_cl$cl.json"))
    print("bench cl$cl", d["value"], d["e2e"]["value"], d["launches_per_step"], d["roofline"]["frac"], d["roofline"].get("frac_in_graph"), d["roofline_tensor"]["frac"], d["kernel_time_breakdown_ms"])
except Exception as e: print("bench failed", e)
PY
done
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2c.csv python profiles/run_step_for_ncu.py > gpurun_out/ncu_list_r2c.log 2>&1; echo "ncu rc=$?" >> gpurun_out/run4_status.txt
python profiles/summarize_launches.py gpurun_out/launches_r2c.csv 12 | head -45
cat gpurun_out/run4_status.txt
