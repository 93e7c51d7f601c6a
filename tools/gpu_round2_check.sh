# on the box: A/B of Adam beside backward (bucket size in MB; 0 = one launch after backward)
for mb in 8 0 8 0 4 16; do
  B2SEG_ADAM_OVERLAP_MB=$mb timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2_bench_ov$mb.json 2> gpurun_out/r2_bench_ov$mb.err; python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r2_bench_ov$mb.json') if l.startswith('{')][-1]); print('overlap $mb MB', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d.get('gpu_launches'), d['clocks'])"
done
