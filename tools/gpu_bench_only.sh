#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python bench.py --dump-tuned $O/tuned.json ) > $O/bench.log 2>&1
echo "bench rc=$?" >> $O/bench.log
grep '^{"metric' $O/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'e2e_u8', d['e2e_u8']['value'], 'launches', d['gpu_launches'])
print('scan', d['scan_mode'] and d['scan_mode']['value'])
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','traffic','kernel','conv_family')})
print('kernels', d['roofline']['kernels_ms'])
print('clocks', d['clocks'])
"
