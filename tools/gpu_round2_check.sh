# on the box: MultiRes determinism diagnostic, full GPU test-suite (no -x: every failure in one run), cfg4 bench
timeout 300 python tools/diag_multires.py > gpurun_out/r2_diag_multires2.txt 2>&1; head -12 gpurun_out/r2_diag_multires2.txt | cut -c1-120
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread -s > gpurun_out/r2_gputest13.log 2>&1; tail -6 gpurun_out/r2_gputest13.log | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline --config 4 --ops-json gpurun_out/r2_ops_cfg4_i.json > gpurun_out/r2_bench_cfg4_i.json 2> gpurun_out/r2_bench_cfg4_i.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_cfg4_i.json') if l.startswith('{')][-1]); print('cfg4', d['value'], d['ms_per_step'])"
