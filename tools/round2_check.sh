#!/bin/bash
# Round-2 check on one B200: GPU tests, smoke(), default bench line, per-kernel events, launch list.
R=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${R}_pytest_gpu.log 2>&1; tail -5 $O/${R}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $O/${R}_bench_default.json 2> $O/${R}_bench_default.err; tail -3 $O/${R}_bench_default.err
timeout 600 python bench.py --no-cpu-baseline --profile-out $O/${R}_events_tc.json > $O/${R}_bench_events.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_ncu_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import json
d=json.load(open('$O/${R}_bench_default.json'))
print('default', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d.get('launches_per_step'))
print(d.get('top_kernels'))
PY
