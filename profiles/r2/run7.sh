mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kv_attn_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_kv32.log 2>&1; echo "kv pytest rc=$?" > gpurun_out/run7_status.txt
tail -25 gpurun_out/pytest_kv32.log
timeout 300 python profiles/flash_bench.py > gpurun_out/flash_v6.json 2> gpurun_out/flash_v6.err; echo "flash rc=$?" >> gpurun_out/run7_status.txt
cat gpurun_out/flash_v6.json; grep occupancy gpurun_out/flash_v6.err
timeout 900 python -m pytest tests/test_image_stream_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_image2.log 2>&1; echo "image pytest rc=$?" >> gpurun_out/run7_status.txt
grep -a "parity\|image stream\|passed\|failed\|Error" gpurun_out/pytest_image2.log | tail -12
show() { python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2", round(d["value"],2), round(d["e2e"]["value"],2), d["launches_per_step"], round(d["roofline"]["frac"],3), round(d["roofline"].get("frac_in_graph",0),3), round(d["roofline_tensor"]["frac"],3), {k:(v["ms"],v.get("ms_in_graph")) for k,v in d["kernel_time_breakdown_ms"].items()}, (d.get("e2e_image") or {}).get("value"))
except Exception as e: print("$2 failed", e)
PY
}
timeout 600 python bench.py --config 4 --steps 30 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_cfg4_f.json 2> gpurun_out/bench_cfg4_f.err; show gpurun_out/bench_cfg4_f.json cfg4; tail -2 gpurun_out/bench_cfg4_f.err
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2f.log 2>&1; echo "full pytest rc=$?" >> gpurun_out/run7_status.txt
tail -12 gpurun_out/pytest_gpu_r2f.log
cat gpurun_out/run7_status.txt
