mkdir -p gpurun_out
for pp in 0 1; do
L2D_FLASH_PINGPONG=$pp timeout 200 python profiles/flash_timeline.py > gpurun_out/flash_timeline_pp$pp.txt 2>&1; echo "timeline pp$pp rc=$?"; cat gpurun_out/flash_timeline_pp$pp.txt
done
