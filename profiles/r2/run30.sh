mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2", round(d["value"],2), round(d["e2e"]["value"],2), d.get("launches_per_step"), {k:(v["ms"],v.get("ms_in_graph")) for k,v in d["kernel_time_breakdown_ms"].items() if k in ("gemm","norm")})
except Exception as e: print("$2 failed", e)
PY
}
run() { env $2 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_ab11_$1.json 2> gpurun_out/bench_ab11_$1.err; show gpurun_out/bench_ab11_$1.json $1; }
run base1 L2D_NOP=1
run cluster1 L2D_SPLITK_CLUSTER=1
run lnfold1 L2D_LN_FOLD=1
run base2 L2D_NOP=1
run cluster2 L2D_SPLITK_CLUSTER=1
run lnfold2 L2D_LN_FOLD=1
