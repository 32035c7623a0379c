#!/bin/bash
# round-2 closing run on one B200: whole gpu suite, the default bench line, DRAM bytes of the fused pass at the
# benchmarked shape (single-pass ncu: no replay, so the 80 GB the kernel writes need no save / restore), launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout=600 2>&1 | grep -v "Warning\|warnings.warn\|Possible non" | tail -8 > gpurun_out/r2f_pytest.log
tail -3 gpurun_out/r2f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 4 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.log; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1]); r=d['roofline']; e=d['e2e']
print(d['value'], d['unit'], d['ms_per_step'], 'ms/step | frac', round(r['frac'],3), 'mix', r.get('mix_peak'), '| e2e', e['value'], e['fit_seconds'], '| cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], d['clocks'])
print(d.get('kernel_ms_per_step'))
P
CMD1M="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fused_tc_kernel -s 3 -c 3 --csv --log-file gpurun_out/r2f_fused_dram_1m.csv $CMD1M > gpurun_out/r2f_ncu_dram.log 2>&1; echo "dram rc=$?"; tail -4 gpurun_out/r2f_fused_dram_1m.csv | cut -c1-300
CMD="python bench.py --cells 100000 --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/r2f_plain_100k.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/r2f_ncu_launches.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/launches_r02.csv
