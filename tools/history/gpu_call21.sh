#!/usr/bin/env bash
mkdir -p gpurun_out
O=gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q --tb=short -p no:cacheprovider -x ) > $O/pytest_k4.log 2>&1
echo "pytest rc=$?" >> $O/pytest_k4.log; tail -8 $O/pytest_k4.log
python - <<'P'
import torch, math
from diffmvs_b200 import ops, packing
for (cin, cout, D, H, W) in ((32, 16, 12, 36, 50), (16, 8, 24, 72, 100)):
    x = torch.rand(1, D, H, W, cin, device="cuda")
    skip = torch.rand(1, 2 * D, 2 * H, 2 * W, cout, device="cuda")
    sd = {"d.conv.weight": torch.rand(cin, cout, 3, 3, 3) * 0.1, "d.bn.weight": torch.ones(cout), "d.bn.bias": torch.zeros(cout),
          "d.bn.running_mean": torch.zeros(cout), "d.bn.running_var": torch.ones(cout)}
    w, b = (t.cuda() for t in packing.pack_deconv3d_bn(sd, "d"))
    for _ in range(3): ops.deconv3d(x, w, b, skip)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): ops.deconv3d(x, w, b, skip)
    e1.record(); torch.cuda.synchronize()
    print("deconv3d %d->%d %dx%dx%d: %.3f ms" % (cin, cout, D, H, W, e0.elapsed_time(e1) / 20))
P
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-alt-modes --no-cpu-baseline --no-gpu-baseline --no-fusion --no-scan-mode > $O/bench_c21.log 2>&1
grep '^{"metric' $O/bench_c21.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench:', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
