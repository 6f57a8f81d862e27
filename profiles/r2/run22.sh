mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2", round(d["value"],2), round(d["e2e"]["value"],2), d.get("gpu_launches"), round(d["roofline"]["frac"],3), round(d["roofline"].get("frac_in_graph",0),3), {k:(v["ms"],v.get("ms_in_graph")) for k,v in d["kernel_time_breakdown_ms"].items() if k in ("gemm","spatial_attn","kv_attn","norm")})
except Exception as e: print("$2 failed", e)
PY
}
timeout 400 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "attention or flash" -x > gpurun_out/pytest_fa_r128.log 2>&1; echo "fa pytest rc=$?"; tail -3 gpurun_out/pytest_fa_r128.log
L2D_LIB_OVERRIDE=profiles/bin/libl2d_r20.so timeout 300 python profiles/flash_bench.py > gpurun_out/flash_r20b.json 2> gpurun_out/flash_r20b.err; echo "flash r20 rc=$?"; cat gpurun_out/flash_r20b.json
L2D_LIB_OVERRIDE=profiles/bin/libl2d_r20.so L2D_FLASH_NQ=1 timeout 300 python profiles/flash_bench.py > gpurun_out/flash_r20_nq1.json 2> gpurun_out/flash_r20_nq1.err; echo "flash r20 (no NQ toggle in that lib) rc=$?"
timeout 300 python profiles/flash_bench.py > gpurun_out/flash_r128.json 2> gpurun_out/flash_r128.err; echo "flash r128 rc=$?"; cat gpurun_out/flash_r128.json
L2D_FLASH_NQ=1 timeout 300 python profiles/flash_bench.py > gpurun_out/flash_r128_nq1.json 2> gpurun_out/flash_r128_nq1.err; echo "flash r128 nq1 rc=$?"; cat gpurun_out/flash_r128_nq1.json
for rep in 1 2; do
L2D_LIB_OVERRIDE=profiles/bin/libl2d_r20.so timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_ab7_r20_$rep.json 2> gpurun_out/bench_ab7_r20_$rep.err; show gpurun_out/bench_ab7_r20_$rep.json r20-$rep
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_ab7_r128_$rep.json 2> gpurun_out/bench_ab7_r128_$rep.err; show gpurun_out/bench_ab7_r128_$rep.json r128-$rep
done
tail -2 gpurun_out/bench_ab7_r128_1.err
timeout 600 compute-sanitizer --tool memcheck python profiles/sanitize_ops.py > gpurun_out/sanitizer_memcheck_r2l.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_r2l.log
timeout 1500 python -m pytest tests -q -m gpu -x -s > gpurun_out/pytest_gpu_r2l.log 2>&1; echo "full pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r2l.log
cp gpurun_out/parity_report.json gpurun_out/parity_report_r2l.json 2>/dev/null
