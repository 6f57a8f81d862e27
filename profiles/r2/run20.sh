mkdir -p gpurun_out
show() { python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2", round(d["value"],2), round(d["e2e"]["value"],2), d.get("gpu_launches"), round(d["roofline"]["frac"],3), round(d["roofline"].get("frac_in_graph",0),3), {k:(v["ms"],v.get("ms_in_graph")) for k,v in d["kernel_time_breakdown_ms"].items() if k in ("gemm","spatial_attn","kv_attn","norm")})
except Exception as e: print("$2 failed", e)
PY
}
timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "groupnorm" -x > gpurun_out/pytest_gn_fused.log 2>&1; echo "gn pytest rc=$?"; tail -3 gpurun_out/pytest_gn_fused.log
timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_determinism_gpu.py -q -m gpu -x > gpurun_out/pytest_mod_gn.log 2>&1; echo "modules+determinism pytest rc=$?"; tail -3 gpurun_out/pytest_mod_gn.log
for rep in 1 2; do
for f in 0 1; do
L2D_GN_FUSED=$f timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_gnf$f$rep.json 2> gpurun_out/bench_gnf$f$rep.err; show gpurun_out/bench_gnf$f$rep.json gn_fused$f-$rep
done
done
tail -2 gpurun_out/bench_gnf11.err
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o gpurun_out/flash_lazy_full python profiles/flash_ncu_probe.py > gpurun_out/ncu_flash_lazy.log 2>&1; echo "ncu flash rc=$?"
tail -2 gpurun_out/ncu_flash_lazy.log
ncu -i gpurun_out/flash_lazy_full.ncu-rep --page raw --csv > gpurun_out/flash_lazy_raw.csv 2>/dev/null
ncu -i gpurun_out/flash_lazy_full.ncu-rep --page source --csv > gpurun_out/flash_lazy_source.csv 2>/dev/null
python profiles/ncu_key_metrics.py gpurun_out/flash_lazy_raw.csv | head -70
