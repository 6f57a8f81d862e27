mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2", round(d["value"],2), round(d["e2e"]["value"],2), {k:(v["ms"],v.get("ms_in_graph")) for k,v in d["kernel_time_breakdown_ms"].items() if k in ("gemm","spatial_attn","kv_attn")})
except Exception as e: print("$2 failed", e)
PY
}
for rep in 1 2; do
for v in base2 d256 reord; do
L2D_LIB_OVERRIDE=profiles/bin/libl2d_$v.so timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_ab2_$v$rep.json 2> gpurun_out/bench_ab2_$v$rep.err; show gpurun_out/bench_ab2_$v$rep.json $v$rep
done
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_ab2_cur$rep.json 2> gpurun_out/bench_ab2_cur$rep.err; show gpurun_out/bench_ab2_cur$rep.json cur$rep
done
