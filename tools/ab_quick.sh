#!/bin/bash
# Quick A/B on one B200: kernel-level + parity tests, then the default bench with graphs on / off and a per-kernel event profile.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_cli.py -m gpu -q > $O/ab_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed" $O/ab_pytest.log | tail -8
B="python bench.py --no-cpu-baseline --no-extra --windows 3"
timeout 300 $B > $O/ab_graph.json 2> $O/ab_graph.err; tail -2 $O/ab_graph.err
env ${AB_ENV:-CLSR_NO_TMA_STORE=1} CLSR_NO_GRAPH=1 timeout 300 $B --profile-out $O/ab_events_alt.json > $O/ab_alt.json 2> $O/ab_alt.err; tail -2 $O/ab_alt.err
CLSR_NO_GRAPH=1 timeout 300 $B --profile-out $O/ab_events.json > $O/ab_nograph.json 2> $O/ab_nograph.err; tail -2 $O/ab_nograph.err
python - <<PY
import json
for f in ("ab_graph","ab_alt","ab_nograph"):
    try:
        d=json.load(open("$O/%s.json"%f)); print(f, round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), d.get('launches_per_step'))
    except Exception as ex: print(f, 'failed', ex)
try:
    pk=json.load(open("$O/ab_events.json"))['per_kernel']
    print([(k, round(v['ms']*v['calls_per_step'],4)) for k,v in sorted(pk.items(), key=lambda kv:-kv[1]['ms']*kv[1]['calls_per_step'])[:24]])
    alt=json.load(open("$O/ab_events_alt.json"))['per_kernel']
    diff=[(k, round(v['ms']*v['calls_per_step'],4), round(alt[k]['ms']*alt[k]['calls_per_step'],4)) for k,v in pk.items() if k in alt and abs(v['ms']*v['calls_per_step']-alt[k]['ms']*alt[k]['calls_per_step'])>0.004]
    print('default vs alt:', sorted(diff, key=lambda x: x[1]-x[2]))
except Exception as ex: print('events failed', ex)
PY
