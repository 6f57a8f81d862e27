mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2", round(d["value"],2), round(d["e2e"]["value"],2), d.get("gpu_launches"), round(d["roofline"]["frac"],3), round(d["roofline"].get("frac_in_graph",0),3), round(d.get("roofline_tensor",{}).get("frac",0),3), {k:(v["ms"],v.get("ms_in_graph")) for k,v in d["kernel_time_breakdown_ms"].items() if k in ("gemm","spatial_attn","kv_attn","norm")})
except Exception as e: print("$2 failed", e)
PY
}
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x > gpurun_out/pytest_ops_uniform.log 2>&1; echo "ops pytest rc=$?"; tail -3 gpurun_out/pytest_ops_uniform.log
timeout 200 python profiles/flash_bench.py > gpurun_out/flash_uniform.json 2> gpurun_out/flash_uniform.err; echo "flash uniform rc=$?"; cat gpurun_out/flash_uniform.json
timeout 200 python profiles/flash_timeline.py > gpurun_out/flash_timeline_uniform.txt 2>&1; head -6 gpurun_out/flash_timeline_uniform.txt
timeout 300 python profiles/gemm_timeline.py > gpurun_out/gemm_timeline_uniform.txt 2>&1; cat gpurun_out/gemm_timeline_uniform.txt | tail -5
timeout 600 python profiles/gemm_shapes_bench.py > gpurun_out/gemm_shapes_uniform.txt 2>&1; cat gpurun_out/gemm_shapes_uniform.txt
for rep in 1 2; do
L2D_LIB_OVERRIDE=profiles/bin/libl2d_ptmem.so timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_ab10_ptmem_$rep.json 2> gpurun_out/bench_ab10_ptmem_$rep.err; show gpurun_out/bench_ab10_ptmem_$rep.json ptmem-$rep
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_ab10_uniform_$rep.json 2> gpurun_out/bench_ab10_uniform_$rep.err; show gpurun_out/bench_ab10_uniform_$rep.json uniform-$rep
done
tail -2 gpurun_out/bench_ab10_uniform_1.err
