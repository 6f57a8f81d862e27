mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck python profiles/sanitize_k1.py > gpurun_out/sanitizer_racecheck_k1.log 2>&1; echo "racecheck k1 rc=$?" > gpurun_out/run16_status.txt
tail -3 gpurun_out/sanitizer_racecheck_k1.log
for v in head cur; do
if [ $v = head ]; then export L2D_LIB_OVERRIDE=profiles/bin/libl2d_head.so; else unset L2D_LIB_OVERRIDE; fi
timeout 300 python profiles/flash_bench.py > gpurun_out/flash_ab_$v.json 2>/dev/null; cat gpurun_out/flash_ab_$v.json; echo
done
unset L2D_LIB_OVERRIDE
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu_r2g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/run16_status.txt
tail -5 gpurun_out/pytest_gpu_r2g.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/run16_status.txt; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2k.json")); print("cfg2", round(d["value"],2), round(d["e2e"]["value"],2), d["launches_per_step"], {k:(v["ms"],v.get("ms_in_graph")) for k,v in d["kernel_time_breakdown_ms"].items()}, (d.get("e2e_image") or {}).get("value"))
PY
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:gemm_f16|flash|kv_attn|groupnorm|layernorm|splitk' -s 9 -c 44 -f \
  -o gpurun_out/layer_l0_full python profiles/run_step_for_ncu.py > gpurun_out/ncu_full_l0.log 2>&1; echo "full l0 rc=$?" >> gpurun_out/run16_status.txt
ncu -i gpurun_out/layer_l0_full.ncu-rep --page raw --csv > gpurun_out/layer_l0_full_raw.csv 2>/dev/null
python profiles/summarize_ncu_full.py gpurun_out/layer_l0_full_raw.csv > gpurun_out/layer_l0_full_kernels.txt 2>&1; head -50 gpurun_out/layer_l0_full_kernels.txt
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:gemm_f16|flash|kv_attn|groupnorm|layernorm|splitk' -s 205 -c 48 -f \
  -o gpurun_out/layer_l2_full python profiles/run_step_for_ncu.py > gpurun_out/ncu_full_l2.log 2>&1; echo "full l2 rc=$?" >> gpurun_out/run16_status.txt
ncu -i gpurun_out/layer_l2_full.ncu-rep --page raw --csv > gpurun_out/layer_l2_full_raw.csv 2>/dev/null
python profiles/summarize_ncu_full.py gpurun_out/layer_l2_full_raw.csv > gpurun_out/layer_l2_full_kernels.txt 2>&1; head -52 gpurun_out/layer_l2_full_kernels.txt
rm -f gpurun_out/layer_l0_full.ncu-rep gpurun_out/layer_l2_full.ncu-rep
cat gpurun_out/run16_status.txt
