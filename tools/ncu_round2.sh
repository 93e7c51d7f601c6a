# round-2 ncu evidence (run on the box from the repo root): launch list of three cfg2 steps and of the gate / BN-glue kernels of one
# cfg3 step (time + DRAM bytes per launch).  `--set full` captures are summarised on the box by tools/ncu_summary.py (the .ncu-rep
# files are too large to come back): pass "full" as the first argument to add them.
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv python tools/run_step.py --config 2 --steps 3 > gpurun_out/r2_ncu_a.log 2>&1; tail -1 gpurun_out/r2_ncu_a.log
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gate_|bn_act2" -c 300 --csv --log-file gpurun_out/r2_ncu_gate_launches_final.csv python tools/run_step.py --config 3 --steps 1 > gpurun_out/r2_ncu_c.log 2>&1; tail -1 gpurun_out/r2_ncu_c.log
if [ "$1" == "full" ]; then
  timeout 600 ncu --set full --clock-control none -k regex:"conv_halo_kernel|wgrad_halo_kernel|conv_gemm_kernel" -c 53 -o /tmp/r2_prof_tensor_kernels -f python tools/run_step.py --config 2 --steps 1 > gpurun_out/r2_ncu_b.log 2>&1; tail -1 gpurun_out/r2_ncu_b.log
  python tools/ncu_summary.py /tmp/r2_prof_tensor_kernels.ncu-rep > gpurun_out/r2_prof_tensor_kernels.summary.txt 2>&1
  timeout 300 ncu --set full --clock-control none -k regex:"gate_out_kernel|gate_mid" -c 14 -o /tmp/r2_prof_gate_kernels -f python tools/run_step.py --config 3 --steps 1 > gpurun_out/r2_ncu_d.log 2>&1
  python tools/ncu_summary.py /tmp/r2_prof_gate_kernels.ncu-rep > gpurun_out/r2_prof_gate_kernels.summary.txt 2>&1
fi
rm -f gpurun_out/*.ncu-rep; du -sh gpurun_out
