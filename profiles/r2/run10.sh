mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -f -o gpurun_out/flash_full python profiles/flash_ncu_probe.py > gpurun_out/ncu_flash.log 2>&1; echo "ncu flash rc=$?" > gpurun_out/run10_status.txt
tail -3 gpurun_out/ncu_flash.log
ncu -i gpurun_out/flash_full.ncu-rep --page raw --csv > gpurun_out/flash_full_raw.csv 2>/dev/null
python profiles/ncu_key_metrics.py gpurun_out/flash_full_raw.csv | head -120
timeout 600 python profiles/gemm_shapes_bench.py > gpurun_out/gemm_shapes_r2.txt 2>&1; cat gpurun_out/gemm_shapes_r2.txt
timeout 300 python profiles/gemm_timeline.py > gpurun_out/gemm_timeline_r2.txt 2>&1; cat gpurun_out/gemm_timeline_r2.txt
cat gpurun_out/run10_status.txt
