# on the box: A/B of the Adam grid cap (blocks per SM) with Adam beside backward
for bps in 24 2 4 24 2 8; do
  B2SEG_ADAM_BLOCKS_PER_SM=$bps timeout 300 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2_bench_bps$bps.json 2> gpurun_out/r2_bench_bps$bps.err; python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r2_bench_bps$bps.json') if l.startswith('{')][-1]); print('adam blocks/SM $bps', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), d['clocks']['sm_mhz'])"
done
