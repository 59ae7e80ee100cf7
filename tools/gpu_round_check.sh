#!/usr/bin/env bash
# Full check of the current tree on one B200: GPU parity suite, smoke, default bench (both arms).  Outputs in gpurun_out/.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -8 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke.log 2>&1
echo "smoke rc=$?" >> $O/smoke.log
tail -2 $O/smoke.log
( time timeout 900 python bench.py --dump-tuned $O/tuned.json ) > $O/bench.log 2>&1
echo "bench rc=$?" >> $O/bench.log
grep '^{"metric' $O/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'e2e_u8', d['e2e_u8']['value'])
print('scan', d['scan_mode'] and d['scan_mode']['value'], 'fusion', d['fusion'])
print('gpu_baseline', {k:(v['value'] if isinstance(v,dict) else v) for k,v in (d['gpu_baseline'] or {}).items() if k!='kind'})
print('cpu', d['cpu_baseline'])
print('alt', d['alt_modes'])
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','kernel','conv_family')})
print('kernels', d['roofline']['kernels_ms'])
"
tail -4 $O/bench.log | grep -E "real|rc="
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_ref.log 2>&1
tail -5 $O/bench_ref.log | cut -c1-400
