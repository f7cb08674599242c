#!/bin/bash
# compute-sanitizer runs of the kernels: memcheck + racecheck + synccheck on smoke() (159 launches, every
# kernel family on small shapes) and on one training step of a 2-block model at 256x256 (P = 256: the fused-statistics
# hand-offs, the L = 16 fast attention, CTA-pair GEMM tiles).  Logs -> gpurun_out/<tag>_sanitize_*.log
TAG=${1:-r2}
export PYTHONUNBUFFERED=1
cat > /tmp/bf_san_step.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from bubbleformer_b200 import get_model
from bubbleformer_b200.losses import rel_l2_loss
from bubbleformer_b200.parallel import GradSink
from oracle.param_init import fluid_params
torch.manual_seed(0)
cfg = dict(input_fields=4, output_fields=4, patch_size=16, embed_dim=384, num_heads=6, processor_blocks=2, drop_path=0.2,
           attn_scale=True, feat_scale=True, num_fluid_params=9)
m = get_model("filmavit", time_window=5, **cfg).cuda().train()
sink = GradSink(m)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
x = torch.randn(B, 5, 4, 256, 256, device="cuda"); tgt = torch.randn_like(x); cond = fluid_params(B).cuda()
sink.begin_step()
loss = rel_l2_loss(m(x, cond), tgt)
loss.backward()
sink.finish()
torch.cuda.synchronize()
print("step ok, loss", float(loss), "finite grads", bool(torch.isfinite(sink.flat).all()))
PY
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python __graft_entry__.py smoke > gpurun_out/${TAG}_sanitize_${tool}_smoke.log 2>&1
  echo "$tool smoke rc=$?"; tail -4 gpurun_out/${TAG}_sanitize_${tool}_smoke.log
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/bf_san_step.py 16 > gpurun_out/${TAG}_sanitize_${tool}_step.log 2>&1
  echo "$tool step rc=$?"; tail -4 gpurun_out/${TAG}_sanitize_${tool}_step.log
done
