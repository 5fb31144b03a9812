#!/bin/bash
# Round-end check on one B200: GPU tests, smoke(), bench lines for every workload, per-kernel events, launch list.
R=${1:-r01}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > $O/${R}_bench_default.json 2>/dev/null
python bench.py --optimizer lazyadam --no-cpu-baseline > $O/${R}_bench_lazyadam.json 2>/dev/null
python bench.py --workload kuaishou --no-cpu-baseline --steps 10 > $O/${R}_bench_kuaishou_T250.json 2>/dev/null
python bench.py --workload small --no-cpu-baseline > $O/${R}_bench_small_batch500.json 2>/dev/null
python bench.py --no-cpu-baseline --profile-out $O/${R}_events_tc.json > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_ncu_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
for f in default lazyadam kuaishou_T250 small_batch500; do
  python -c "
import json; d=json.load(open('$O/${R}_bench_$f.json')); print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']))"
done
