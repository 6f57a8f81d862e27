mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kv_attn_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_kv_slice.log 2>&1; echo "kv pytest rc=$?" > gpurun_out/run8_status.txt
tail -8 gpurun_out/pytest_kv_slice.log
L2D_K1_SLICE=1 timeout 900 python -m pytest tests/test_kv_attn_gpu.py -m gpu -q -p no:cacheprovider -k "full_size or golden" > gpurun_out/pytest_kv_slice16.log 2>&1; echo "kv slice16 pytest rc=$?" >> gpurun_out/run8_status.txt
tail -4 gpurun_out/pytest_kv_slice16.log
L2D_FLASH_OCC_DEBUG=1 timeout 300 python profiles/flash_bench.py > gpurun_out/flash_v7.json 2> gpurun_out/flash_v7.err; echo "flash rc=$?" >> gpurun_out/run8_status.txt
cat gpurun_out/flash_v7.json; grep occupancy gpurun_out/flash_v7.err
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k attention -p no:cacheprovider 2>&1 | tail -2
show() { python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2", round(d["value"],2), round(d["e2e"]["value"],2), d["launches_per_step"], round(d["roofline"]["frac"],3), round(d["roofline"].get("frac_in_graph",0),3), round(d["roofline_tensor"]["frac"],3), {k:(v["ms"],v.get("ms_in_graph")) for k,v in d["kernel_time_breakdown_ms"].items()}, (d.get("e2e_image") or {}).get("value"))
except Exception as e: print("$2 failed", e)
PY
}
timeout 600 python bench.py --config 4 --steps 30 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_cfg4_g.json 2> gpurun_out/bench_cfg4_g.err; show gpurun_out/bench_cfg4_g.json cfg4; tail -2 gpurun_out/bench_cfg4_g.err
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; show gpurun_out/bench_r2g.json cfg2
L2D_K1_SLICE=1 timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_r2g_slice.json 2> gpurun_out/bench_r2g_slice.err; show gpurun_out/bench_r2g_slice.json cfg2-slice16
cat gpurun_out/run8_status.txt
