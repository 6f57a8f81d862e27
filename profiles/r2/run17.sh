mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2", round(d["value"],2), round(d["e2e"]["value"],2), round(d["roofline"]["frac"],3), round(d["roofline"].get("frac_in_graph",0),3), {k:(v["ms"],v.get("ms_in_graph")) for k,v in d["kernel_time_breakdown_ms"].items() if k in ("gemm","spatial_attn","kv_attn","norm")})
except Exception as e: print("$2 failed", e)
PY
}
for rep in 1 2; do
L2D_LIB_OVERRIDE=profiles/bin/libl2d_head2.so timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_ab3_head$rep.json 2> gpurun_out/bench_ab3_head$rep.err; show gpurun_out/bench_ab3_head$rep.json head$rep
for e in 0 1 2; do
L2D_K1_EAGER_PLANES=$e timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_ab3_cur_e$e$rep.json 2> gpurun_out/bench_ab3_cur_e$e$rep.err; show gpurun_out/bench_ab3_cur_e$e$rep.json cur_eager$e-$rep
done
done
tail -2 gpurun_out/bench_ab3_head1.err
