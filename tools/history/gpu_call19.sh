#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py tests/test_gpu_operator_surface.py -m gpu -q --tb=short -p no:cacheprovider -x ) > $O/pytest_k3.log 2>&1
echo "pytest rc=$?" >> $O/pytest_k3.log; tail -8 $O/pytest_k3.log
python - <<'P'
import torch, math
from diffmvs_b200 import ops, packing
import sys
for N in (6, 1):
  x = torch.rand(N, 48, 144, 200, 8, device="cuda")
  w = (torch.rand(1, 8, 3, 3, 3) - 0.5) / math.sqrt(216)
  pc = packing.pack_weight(w, torch.zeros(1)).to("cuda")
  for mode in (False, True):
      for _ in range(3): ops.conv3d_to1(x, pc, sigmoid_max=mode)
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      torch.cuda.synchronize(); e0.record()
      for _ in range(20): ops.conv3d_to1(x, pc, sigmoid_max=mode)
      e1.record(); torch.cuda.synchronize()
      print("conv3d_to1 %dx48x144x200 sigmoid_max=%s: %.3f ms" % (N, mode, e0.elapsed_time(e1) / 20))
P
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --no-fusion --no-scan-mode > $O/bench_c20.log 2>&1
grep '^{"metric' $O/bench_c20.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench:', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
