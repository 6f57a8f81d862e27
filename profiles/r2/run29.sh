mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_determinism_gpu.py -q -m gpu -k "groupnorm or determin or same or replay or pdl" -x > gpurun_out/pytest_gn_fix.log 2>&1; echo "gn+determinism pytest rc=$?"; tail -3 gpurun_out/pytest_gn_fix.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python profiles/sanitize_ops.py > gpurun_out/sanitizer_racecheck_final3.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/sanitizer_racecheck_final3.log
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-torch-baseline --no-image > gpurun_out/bench_final3_cfg2_short.json 2> gpurun_out/bench_final3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_final3_cfg2_short.json')); print(round(d['value'],2), round(d['e2e']['value'],2), d['launches_per_step'])"
