mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2", round(d["value"],2), round(d["e2e"]["value"],2), round(d["roofline"]["frac"],3), round(d["roofline"].get("frac_in_graph",0),3), {k:(v["ms"],v.get("ms_in_graph")) for k,v in d["kernel_time_breakdown_ms"].items() if k in ("gemm","spatial_attn","kv_attn","norm")})
except Exception as e: print("$2 failed", e)
PY
}
timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "attention or flash" -x > gpurun_out/pytest_fa_lazy.log 2>&1; echo "fa pytest rc=$?"; tail -3 gpurun_out/pytest_fa_lazy.log
L2D_LIB_OVERRIDE=profiles/bin/libl2d_r18.so timeout 300 python profiles/flash_bench.py > gpurun_out/flash_r18.json 2> gpurun_out/flash_r18.err; echo "flash r18 rc=$?"; cat gpurun_out/flash_r18.json
timeout 300 python profiles/flash_bench.py > gpurun_out/flash_lazy.json 2> gpurun_out/flash_lazy.err; echo "flash lazy rc=$?"; cat gpurun_out/flash_lazy.json
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x > gpurun_out/pytest_gemm_pipe1.log 2>&1; echo "gemm+ops pytest rc=$?"; tail -3 gpurun_out/pytest_gemm_pipe1.log
for rep in 1 2; do
L2D_LIB_OVERRIDE=profiles/bin/libl2d_r18.so timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_ab4_r18_$rep.json 2> gpurun_out/bench_ab4_r18_$rep.err; show gpurun_out/bench_ab4_r18_$rep.json r18-$rep
L2D_LIB_OVERRIDE=profiles/bin/libl2d_r18fa.so timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_ab4_fa_$rep.json 2> gpurun_out/bench_ab4_fa_$rep.err; show gpurun_out/bench_ab4_fa_$rep.json r18+fa-$rep
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_ab4_cur_$rep.json 2> gpurun_out/bench_ab4_cur_$rep.err; show gpurun_out/bench_ab4_cur_$rep.json r18+fa+pipe1-$rep
done
timeout 300 python profiles/gemm_timeline.py > gpurun_out/gemm_timeline_pipe1.txt 2>&1; cat gpurun_out/gemm_timeline_pipe1.txt | tail -6
tail -3 gpurun_out/bench_ab4_cur_1.err
