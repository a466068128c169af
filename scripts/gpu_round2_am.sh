#!/bin/bash
# round 2, state "am": the driver's own invocations: smoke(), default bench.py, reference arm
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out/r2am
mkdir -p $O
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; tail -n 6 $O/smoke.log
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2>&1 | grep real
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err ) 2>&1 | grep real
for f in $O/bench_*.json; do echo "$f $(grep -o '"ms_per_step": [0-9.]*' $f | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' $f | head -1) $(grep -o '"step_ms": {[^}]*}' $f)"; done
python - <<'P'
import json
d=json.load(open('gpurun_out/r2am/bench_default.json'))
print({k:d[k] for k in ('steps','warmup','warmup_done','gpu_launches','kernel_map_build_ms','clocks')})
print('cpu_baseline', d['cpu_baseline'])
print('roofline', {k:d['roofline'][k] for k in ('bound','achieved','peak','frac','traffic','tensor_pipe_active_pct','tensor_issued_tflops','share_of_step')}, d['roofline']['step_model'])
P
