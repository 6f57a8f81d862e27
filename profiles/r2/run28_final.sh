#!/bin/bash
# Final round-2 measurement pass on the final code (everything lands in gpurun_out/; kept summaries are copied to profiles/r2/).
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_gpu_final2.log 2>&1; echo "pytest rc=$?" > gpurun_out/capture2_status.txt
tail -3 gpurun_out/pytest_gpu_final2.log
cp gpurun_out/parity_report.json gpurun_out/parity_report_final2.json 2>/dev/null
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_final2_cfg2.json 2> gpurun_out/bench_final2_cfg2.err; echo "bench cfg2 rc=$?" >> gpurun_out/capture2_status.txt
for c in 3 4; do timeout 600 python bench.py --config $c --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/bench_final2_cfg$c.json 2> gpurun_out/bench_final2_cfg$c.err; echo "bench cfg$c rc=$?" >> gpurun_out/capture2_status.txt; done
timeout 300 python bench.py --impl torch-fp16 --steps 30 --warmup 5 > gpurun_out/bench_final2_torch_fp16.json 2> gpurun_out/bench_final2_torch_fp16.err; echo "torch arm rc=$?" >> gpurun_out/capture2_status.txt
timeout 300 python profiles/flash_bench.py > gpurun_out/flash_final2.json 2>/dev/null
timeout 300 python profiles/gemm_shapes_bench.py > gpurun_out/gemm_shapes_final2.txt 2>&1
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_step_final2.csv python profiles/run_step_for_ncu.py > gpurun_out/ncu_list2.log 2>&1; echo "list rc=$?" >> gpurun_out/capture2_status.txt
timeout 300 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
  --clock-control none -k regex:kv_attn --csv --log-file gpurun_out/k1_traffic_final2.csv python profiles/run_step_for_ncu.py > gpurun_out/ncu_k1t2.log 2>&1; echo "k1t rc=$?" >> gpurun_out/capture2_status.txt
timeout 300 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o gpurun_out/flash_full3 python profiles/flash_ncu_probe.py > gpurun_out/ncu_flash3.log 2>&1
ncu -i gpurun_out/flash_full3.ncu-rep --page raw --csv > gpurun_out/flash_full3_raw.csv 2>/dev/null
python profiles/ncu_key_metrics.py gpurun_out/flash_full3_raw.csv > gpurun_out/flash_full_key_metrics_final2.txt 2>&1
rm -f gpurun_out/flash_full3.ncu-rep
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python profiles/sanitize_ops.py > gpurun_out/sanitizer_memcheck_final2.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/capture2_status.txt
tail -2 gpurun_out/sanitizer_memcheck_final2.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python profiles/sanitize_ops.py > gpurun_out/sanitizer_racecheck_final2.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/capture2_status.txt
tail -3 gpurun_out/sanitizer_racecheck_final2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final2.log 2>&1; echo "smoke rc=$?" >> gpurun_out/capture2_status.txt
cat gpurun_out/capture2_status.txt
python - <<'PY'
import json
for c in (2,3,4):
    try:
        d=json.load(open(f"gpurun_out/bench_final2_cfg{c}.json"))
        print(c, round(d["value"],2), round(d["e2e"]["value"],2), d["launches_per_step"], round(d["roofline"]["frac"],3), round(d["roofline"].get("frac_in_graph",0),3), round(d["roofline_tensor"]["frac"],3), (d.get("e2e_image") or {}).get("value"), (d.get("torch_fp16_eager") or {}).get("value"), d.get("vs_torch_fp16_eager"), (d.get("cpu_baseline") or {}).get("value"), {k:(v["ms"],v.get("ms_in_graph")) for k,v in d["kernel_time_breakdown_ms"].items() if k in ("gemm","spatial_attn","kv_attn","norm")})
    except Exception as e: print(c, "failed", e)
PY
cat gpurun_out/flash_final2.json
