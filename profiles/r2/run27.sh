mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "attention or flash" -x > gpurun_out/pytest_fa_poly.log 2>&1; echo "fa pytest rc=$?"; tail -3 gpurun_out/pytest_fa_poly.log
for pp in 1 0; do
L2D_FLASH_PINGPONG=$pp timeout 200 python profiles/flash_bench.py > gpurun_out/flash_poly_pp$pp.json 2> gpurun_out/flash_poly_pp$pp.err; echo "flash poly pp$pp rc=$?"; cat gpurun_out/flash_poly_pp$pp.json
done
L2D_LIB_OVERRIDE=profiles/bin/libl2d_nopoly.so timeout 200 python profiles/flash_bench.py > gpurun_out/flash_nopoly.json 2> gpurun_out/flash_nopoly.err; echo "flash nopoly rc=$?"; cat gpurun_out/flash_nopoly.json
L2D_LIB_OVERRIDE=profiles/bin/libl2d_nopoly.so L2D_FLASH_PINGPONG=0 timeout 200 python profiles/flash_bench.py > gpurun_out/flash_nopoly_pp0.json 2> gpurun_out/flash_nopoly_pp0.err; echo "flash nopoly pp0 rc=$?"; cat gpurun_out/flash_nopoly_pp0.json
timeout 200 python profiles/flash_timeline.py > gpurun_out/flash_timeline_poly.txt 2>&1; head -6 gpurun_out/flash_timeline_poly.txt
