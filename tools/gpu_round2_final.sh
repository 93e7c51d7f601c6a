# on the box: the round's final evidence — GPU test-suite and smoke on the final build, launch list of one cfg4 step
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --timeout-method thread -s > gpurun_out/r2_gputest_final.log 2>&1; tail -4 gpurun_out/r2_gputest_final.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke_final.log 2>&1; tail -2 gpurun_out/r2_smoke_final.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_cfg4.csv python tools/run_step.py --config 4 --steps 2 > gpurun_out/r2_ncu_cfg4.log 2>&1; tail -1 gpurun_out/r2_ncu_cfg4.log
timeout 300 python bench.py --no-cpu-baseline --config 4 --ops-json gpurun_out/r2_ops_cfg4_final.json > gpurun_out/r2_bench_cfg4_final.json 2> gpurun_out/r2_bench_cfg4_final.err
du -sh gpurun_out
