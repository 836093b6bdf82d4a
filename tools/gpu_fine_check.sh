#!/bin/bash
# GPU box: A/B the fine-sync modes at 1024 cycles, then per-kernel device times of the fine stage at the bench size.
timeout 200 python tools/fine_ab.py --cycles 1024 2>&1 | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('AB', j['mode1_ms'], j['mode0_ms'], j.get('field_mismatches'), j['mode0_stats']['fine_pass'], j['mode1_stats']['fine_pass'])"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_fine|k_fscan" --csv --log-file gpurun_out/launches_fine.csv python bench.py --device-only --steps 1 --warmup 3 > gpurun_out/ncu_bench_fine.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_fine.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows[-3:]: print(r[4][:24], float(r[-1])/1e6, "ms")
PY
