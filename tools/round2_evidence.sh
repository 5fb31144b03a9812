#!/bin/bash
# Round-2 single-GPU evidence: bench lines of every workload that fits one GPU, standalone gather / scatter-add,
# ncu captures.  bash tools/round2_evidence.sh r02
R=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py > $O/${R}_bench_default.json 2> $O/${R}_bench_default.err; tail -2 $O/${R}_bench_default.err
timeout 600 python bench.py --no-cpu-baseline --no-extra --profile-out $O/${R}_events_tc.json > /dev/null 2>&1
timeout 600 python bench.py --optimizer lazyadam --no-cpu-baseline --no-extra > $O/${R}_bench_lazyadam.json 2>/dev/null
timeout 600 python bench.py --workload small --no-cpu-baseline --no-extra > $O/${R}_bench_small_batch500.json 2>/dev/null
timeout 900 python bench.py --workload synth50m --no-cpu-baseline --no-extra --steps 10 --windows 3 > $O/${R}_bench_synth50m_1gpu.json 2> $O/${R}_bench_synth50m_1gpu.err; tail -2 $O/${R}_bench_synth50m_1gpu.err
timeout 300 python tools/bench_gather.py > $O/${R}_bench_gather.json 2> $O/${R}_bench_gather.err; tail -2 $O/${R}_bench_gather.err
bash tools/profile_round.sh $R
timeout 900 python -m pytest tests -m gpu -q > $O/${R}_pytest_gpu.log 2>&1; tail -3 $O/${R}_pytest_gpu.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $O/${R}_memcheck_smoke.log 2>&1; tail -4 $O/${R}_memcheck_smoke.log
for f in default lazyadam small_batch500 synth50m_1gpu; do
  python -c "
import json; d=json.load(open('$O/${R}_bench_$f.json')); print('$f', round(d['value']), round(d['ms_per_step'],3), round(d['e2e']['value']), d.get('launches_per_step'))"
done
cat $O/${R}_bench_gather.json
