mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o gpurun_out/gemm_full python profiles/gemm_ncu_probe.py > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?" > gpurun_out/run12_status.txt
tail -2 gpurun_out/ncu_gemm.log
ncu -i gpurun_out/gemm_full.ncu-rep --page raw --csv > gpurun_out/gemm_full_raw.csv 2>/dev/null
python profiles/ncu_key_metrics.py gpurun_out/gemm_full_raw.csv | head -90
ncu -i gpurun_out/gemm_full.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 > gpurun_out/gemm_full_source0.csv 2>/dev/null
python profiles/ncu_source_hot.py gpurun_out/gemm_full_source0.csv 45
cat gpurun_out/run12_status.txt
