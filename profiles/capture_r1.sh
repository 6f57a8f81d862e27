#!/bin/bash
# Round-1 measurement pass (run on the GPU box from the repo root; everything lands in gpurun_out/, the summaries
# that are kept are copied into profiles/r1/ afterwards -- see profiles/README.md):
#   bench line, ncu launch list of one step, K1 DRAM traffic for its 40 launches, ncu --set full of one level-0 layer
#   (K1 + tcgen05 GEMM + flash attention launches), family ablation, GPU test suite.
set -u
mkdir -p gpurun_out
timeout 400 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 200 python profiles/ablate_families.py 40 > gpurun_out/ablate_final.json 2>/dev/null; echo "ablate rc=$?"
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_final.csv python profiles/run_step_for_ncu.py > gpurun_out/ncu_list.log 2>&1; echo "list rc=$?"
timeout 200 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:kv_attn --csv --log-file gpurun_out/k1_traffic_final.csv python profiles/run_step_for_ncu.py > gpurun_out/ncu_k1t.log 2>&1; echo "k1t rc=$?"
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:kv_attn|gemm_f16|flash_attn' -s 10 -c 22 -f \
  -o gpurun_out/layer0_full python profiles/run_step_for_ncu.py > gpurun_out/ncu_full.log 2>&1; echo "full rc=$?"
timeout 420 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final.log
cat gpurun_out/bench_final.json
