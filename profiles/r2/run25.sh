mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2", round(d["value"],2), round(d["e2e"]["value"],2), d.get("gpu_launches"), round(d["roofline"]["frac"],3), round(d["roofline"].get("frac_in_graph",0),3), {k:(v["ms"],v.get("ms_in_graph")) for k,v in d["kernel_time_breakdown_ms"].items() if k in ("gemm","spatial_attn","kv_attn","norm")})
except Exception as e: print("$2 failed", e)
PY
}
timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "attention or flash" -x > gpurun_out/pytest_fa_ptmem.log 2>&1; echo "fa pytest rc=$?"; tail -5 gpurun_out/pytest_fa_ptmem.log
timeout 200 python profiles/flash_bench.py > gpurun_out/flash_ptmem.json 2> gpurun_out/flash_ptmem.err; echo "flash ptmem rc=$?"; cat gpurun_out/flash_ptmem.json
L2D_LIB_OVERRIDE=profiles/bin/libl2d_r20.so timeout 200 python profiles/flash_bench.py > gpurun_out/flash_r20c.json 2> gpurun_out/flash_r20c.err; echo "flash r20 rc=$?"; cat gpurun_out/flash_r20c.json
timeout 200 python profiles/flash_timeline.py > gpurun_out/flash_timeline_ptmem.txt 2>&1; cat gpurun_out/flash_timeline_ptmem.txt
for rep in 1 2; do
L2D_LIB_OVERRIDE=profiles/bin/libl2d_r20.so timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_ab9_r20_$rep.json 2> gpurun_out/bench_ab9_r20_$rep.err; show gpurun_out/bench_ab9_r20_$rep.json r20-$rep
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_ab9_ptmem_$rep.json 2> gpurun_out/bench_ab9_ptmem_$rep.err; show gpurun_out/bench_ab9_ptmem_$rep.json ptmem-$rep
done
tail -2 gpurun_out/bench_ab9_ptmem_1.err
