mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2", round(d["value"],2), round(d["e2e"]["value"],2), d.get("gpu_launches"), round(d["roofline"]["frac"],3), round(d["roofline"].get("frac_in_graph",0),3), {k:(v["ms"],v.get("ms_in_graph")) for k,v in d["kernel_time_breakdown_ms"].items() if k in ("gemm","spatial_attn","kv_attn","norm")})
except Exception as e: print("$2 failed", e)
PY
}
timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "attention or flash" -x > gpurun_out/pytest_fa_pp2.log 2>&1; echo "fa pytest rc=$?"; tail -3 gpurun_out/pytest_fa_pp2.log
for pp in 0 1; do
L2D_FLASH_PINGPONG=$pp timeout 200 python profiles/flash_bench.py > gpurun_out/flash_rowregs_pp$pp.json 2> gpurun_out/flash_rowregs_pp$pp.err; echo "flash pp$pp rc=$?"; cat gpurun_out/flash_rowregs_pp$pp.json
done
for rep in 1 2; do
for pp in 0 1; do
L2D_FLASH_PINGPONG=$pp timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_ab8_pp$pp$rep.json 2> gpurun_out/bench_ab8_pp$pp$rep.err; show gpurun_out/bench_ab8_pp$pp$rep.json pingpong$pp-$rep
done
done
tail -2 gpurun_out/bench_ab8_pp11.err
