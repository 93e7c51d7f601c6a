# on the box: the round's final evidence — GPU test-suite, smoke, bench lines of every config, reference arm
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread -s > gpurun_out/r2_gputest_final.log 2>&1; tail -4 gpurun_out/r2_gputest_final.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke_final.log 2>&1; tail -2 gpurun_out/r2_smoke_final.log
timeout 600 python bench.py --ops-json gpurun_out/r2_ops_cfg2_final.json > gpurun_out/r2_bench_cfg2_final.json 2> gpurun_out/r2_bench_cfg2_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
for c in 3 4 5; do timeout 400 python bench.py --no-cpu-baseline --config $c --ops-json gpurun_out/r2_ops_cfg${c}_final.json > gpurun_out/r2_bench_cfg${c}_final.json 2> gpurun_out/r2_bench_cfg${c}_final.err; done
timeout 300 python bench.py --no-cpu-baseline --mode predict > gpurun_out/r2_bench_predict_cfg2_final.json 2> gpurun_out/r2_bench_predict_cfg2_final.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2_bench_*final.json')) + ['gpurun_out/r2_bench_reference_arm.json']:
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('/')[-1], round(d['value'], 1), d['unit'], round(d.get('ms_per_step', 0), 3), 'e2e', round(d['e2e']['value'], 1), 'frac', d.get('roofline', {}).get('frac'), d.get('clocks', {}).get('sm_mhz'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
