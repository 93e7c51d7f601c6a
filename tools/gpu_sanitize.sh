# on the box: compute-sanitizer over the round-2 kernels (gate, fused BN glue, loss) and one step of configs 3 / 4 / 2 at batch 1-2
export PYTHONWARNINGS=ignore
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q -x -k "gate or bn_apply or loss_kinds" --timeout 380 --timeout-method thread > gpurun_out/r2_sanitize_memcheck_kernels.log 2>&1; echo "memcheck kernels rc=$?"; tail -4 gpurun_out/r2_sanitize_memcheck_kernels.log | cut -c1-200
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -q -x -k "gate or bn_apply" --timeout 380 --timeout-method thread > gpurun_out/r2_sanitize_racecheck_kernels.log 2>&1; echo "racecheck kernels rc=$?"; tail -4 gpurun_out/r2_sanitize_racecheck_kernels.log | cut -c1-200
for c in 3 4 2; do
  timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/run_step.py --config $c --batch 2 --steps 2 > gpurun_out/r2_sanitize_memcheck_cfg$c.log 2>&1; echo "memcheck cfg$c rc=$?"; tail -3 gpurun_out/r2_sanitize_memcheck_cfg$c.log | cut -c1-200
done
