mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2", round(d["value"],2), round(d["e2e"]["value"],2), d.get("gpu_launches"), round(d["roofline"]["frac"],3), round(d["roofline"].get("frac_in_graph",0),3), {k:(v["ms"],v.get("ms_in_graph")) for k,v in d["kernel_time_breakdown_ms"].items() if k in ("gemm","spatial_attn","kv_attn","norm")})
except Exception as e: print("$2 failed", e)
PY
}
timeout 400 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "attention or flash" -x > gpurun_out/pytest_fa_s64.log 2>&1; echo "fa pytest rc=$?"; tail -3 gpurun_out/pytest_fa_s64.log
L2D_LIB_OVERRIDE=profiles/bin/libl2d_r20.so timeout 300 python profiles/flash_bench.py > gpurun_out/flash_r20.json 2> gpurun_out/flash_r20.err; echo "flash r20 rc=$?"; cat gpurun_out/flash_r20.json
timeout 300 python profiles/flash_bench.py > gpurun_out/flash_s64.json 2> gpurun_out/flash_s64.err; echo "flash s64 rc=$?"; cat gpurun_out/flash_s64.json
for rep in 1 2; do
L2D_LIB_OVERRIDE=profiles/bin/libl2d_r20.so timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_ab6_r20_$rep.json 2> gpurun_out/bench_ab6_r20_$rep.err; show gpurun_out/bench_ab6_r20_$rep.json r20-$rep
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_ab6_s64_$rep.json 2> gpurun_out/bench_ab6_s64_$rep.err; show gpurun_out/bench_ab6_s64_$rep.json s64-$rep
done
tail -2 gpurun_out/bench_ab6_s64_1.err
timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_determinism_gpu.py -q -m gpu -x > gpurun_out/pytest_mod_s64.log 2>&1; echo "modules+determinism pytest rc=$?"; tail -3 gpurun_out/pytest_mod_s64.log
