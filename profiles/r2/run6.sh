mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_kv_attn_gpu.py -m gpu -q -k "attention or kv" -p no:cacheprovider > gpurun_out/pytest_fa5.log 2>&1; echo "fa+kv pytest rc=$?" > gpurun_out/run6_status.txt
tail -4 gpurun_out/pytest_fa5.log
timeout 300 python profiles/flash_bench.py > gpurun_out/flash_v5.json 2> gpurun_out/flash_v5.err; echo "flash rc=$?" >> gpurun_out/run6_status.txt
cat gpurun_out/flash_v5.json; tail -3 gpurun_out/flash_v5.err
timeout 900 python -m pytest tests/test_image_stream_gpu.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_image.log 2>&1; echo "image pytest rc=$?" >> gpurun_out/run6_status.txt
tail -12 gpurun_out/pytest_image.log
show() { python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2", round(d["value"],2), round(d["e2e"]["value"],2), d["launches_per_step"], round(d["roofline"]["frac"],3), round(d["roofline"].get("frac_in_graph",0),3), round(d["roofline_tensor"]["frac"],3), {k:(v["ms"],v.get("ms_in_graph")) for k,v in d["kernel_time_breakdown_ms"].items()}, d.get("e2e_image"), d.get("torch_fp16_eager"))
except Exception as e: print("$2 failed", e)
PY
}
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench rc=$?" >> gpurun_out/run6_status.txt
show gpurun_out/bench_r2e.json default; tail -3 gpurun_out/bench_r2e.err
L2D_LN_FOLD=0 timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_r2e_nofold.json 2> gpurun_out/bench_r2e_nofold.err
show gpurun_out/bench_r2e_nofold.json nofold
for c in 3 4; do timeout 600 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_cfg${c}_e.json 2> gpurun_out/bench_cfg${c}_e.err; show gpurun_out/bench_cfg${c}_e.json cfg$c; tail -2 gpurun_out/bench_cfg${c}_e.err; done
cat gpurun_out/run6_status.txt
