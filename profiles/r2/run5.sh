mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention" -p no:cacheprovider > gpurun_out/pytest_fa4.log 2>&1; echo "fa pytest rc=$?" > gpurun_out/run5_status.txt
tail -4 gpurun_out/pytest_fa4.log
timeout 300 python profiles/flash_bench.py > gpurun_out/flash_v4.json 2> gpurun_out/flash_v4.err; echo "flash rc=$?" >> gpurun_out/run5_status.txt
cat gpurun_out/flash_v4.json; tail -3 gpurun_out/flash_v4.err
timeout 900 python -m pytest tests/test_taesd_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_taesd.log 2>&1; echo "taesd pytest rc=$?" >> gpurun_out/run5_status.txt
tail -25 gpurun_out/pytest_taesd.log
L2D_GEMM_LEAN=1 timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_modules_gpu.py -m gpu -q -k "gemm or conv or split or unet or temporal" -p no:cacheprovider > gpurun_out/pytest_lean.log 2>&1; echo "lean pytest rc=$?" >> gpurun_out/run5_status.txt
tail -6 gpurun_out/pytest_lean.log
for lean in 0 1; do for cl in 0 1; do
L2D_GEMM_LEAN=$lean L2D_SPLITK_CLUSTER=$cl timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_r2d_l${lean}c${cl}.json 2> gpurun_out/bench_r2d_l${lean}c${cl}.err; echo "bench l$lean c$cl rc=$?" >> gpurun_out/run5_status.txt
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_r2d_l${lean}c${cl}.json"))
    print("bench lean$lean cl$cl", round(d["value"],2), round(d["e2e"]["value"],2), d["launches_per_step"], round(d["roofline"]["frac"],3), round(d["roofline"].get("frac_in_graph",0),3), round(d["roofline_tensor"]["frac"],3), {k:(v["ms"],v.get("ms_in_graph")) for k,v in d["kernel_time_breakdown_ms"].items()})
except Exception as e: print("bench failed", e)
PY
done; done
cat gpurun_out/run5_status.txt
