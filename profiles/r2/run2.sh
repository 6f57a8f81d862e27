mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "attention" -p no:cacheprovider > gpurun_out/pytest_fa.log 2>&1; echo "fa pytest rc=$?" > gpurun_out/run2_status.txt
tail -25 gpurun_out/pytest_fa.log
timeout 300 python profiles/flash_bench.py > gpurun_out/flash_new.json 2> gpurun_out/flash_new.err; echo "flash new rc=$?" >> gpurun_out/run2_status.txt
L2D_FLASH_LEGACY=1 timeout 300 python profiles/flash_bench.py > gpurun_out/flash_legacy.json 2> gpurun_out/flash_legacy.err
cat gpurun_out/flash_new.json gpurun_out/flash_legacy.json; tail -3 gpurun_out/flash_new.err
for c in 3 4; do L2D_FLASH_LEGACY=1 timeout 600 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg${c}_a.json 2> gpurun_out/bench_cfg${c}_a.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_cfg${c}_a.json"))
    print("cfg${c}", d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"].get("frac_in_graph"), d.get("torch_fp16_eager"))
except Exception as e: print("cfg${c} failed", e)
PY
done
