mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention" -p no:cacheprovider > gpurun_out/pytest_fa8.log 2>&1; echo "fa pytest rc=$?" > gpurun_out/run9_status.txt
tail -6 gpurun_out/pytest_fa8.log
L2D_FLASH_OCC_DEBUG=1 timeout 300 python profiles/flash_bench.py > gpurun_out/flash_v8.json 2> gpurun_out/flash_v8.err; echo "flash rc=$?" >> gpurun_out/run9_status.txt
cat gpurun_out/flash_v8.json; grep occupancy gpurun_out/flash_v8.err
show() { python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2", round(d["value"],2), round(d["e2e"]["value"],2), d["launches_per_step"], round(d["roofline"]["frac"],3), round(d["roofline"].get("frac_in_graph",0),3), round(d["roofline_tensor"]["frac"],3), {k:(v["ms"],v.get("ms_in_graph")) for k,v in d["kernel_time_breakdown_ms"].items()}, (d.get("e2e_image") or {}).get("value"))
except Exception as e: print("$2 failed", e)
PY
}
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; show gpurun_out/bench_r2h.json cfg2; tail -2 gpurun_out/bench_r2h.err
cat gpurun_out/run9_status.txt
