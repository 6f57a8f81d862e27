mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_ops13.log 2>&1; echo "ops pytest rc=$?" > gpurun_out/run13_status.txt
tail -6 gpurun_out/pytest_ops13.log
timeout 300 python profiles/gemm_timeline.py > gpurun_out/gemm_timeline_hoist.txt 2>&1; cut -c1-260 gpurun_out/gemm_timeline_hoist.txt
show() { python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2", round(d["value"],2), round(d["e2e"]["value"],2), d["launches_per_step"], round(d["roofline"]["frac"],3), round(d["roofline"].get("frac_in_graph",0),3), round(d["roofline_tensor"]["frac"],3), {k:(v["ms"],v.get("ms_in_graph")) for k,v in d["kernel_time_breakdown_ms"].items()}, (d.get("e2e_image") or {}).get("value"))
except Exception as e: print("$2 failed", e)
PY
}
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; show gpurun_out/bench_r2j.json cfg2; tail -2 gpurun_out/bench_r2j.err
timeout 900 python -m pytest tests/test_modules_gpu.py tests/test_taesd_gpu.py tests/test_determinism_gpu.py tests/test_warmup_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_mod13.log 2>&1; echo "modules pytest rc=$?" >> gpurun_out/run13_status.txt
tail -6 gpurun_out/pytest_mod13.log
cat gpurun_out/run13_status.txt
