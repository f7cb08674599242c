"""Config 5 smoke/timing: film_avit_big (E=768, 12 heads, 12 blocks) at 1024x1024, bf16 fwd+bwd, per-GPU batch B."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bubbleformer_b200 import get_model
from bubbleformer_b200.losses import rel_l2_loss
from bubbleformer_b200.parallel import GradSink
from oracle.param_init import fluid_params
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
res = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cfg = dict(input_fields=4, output_fields=4, patch_size=16, embed_dim=768, num_heads=12, processor_blocks=12,
           drop_path=0.2, attn_scale=True, feat_scale=True, num_fluid_params=9)
torch.manual_seed(0)
m = get_model("filmavit", time_window=5, **cfg).cuda().train()
sink = GradSink(m)
x = torch.randn(B, 5, 4, res, res, device="cuda"); tgt = torch.randn_like(x); cond = fluid_params(B).cuda()
def step():
    sink.begin_step()
    loss = rel_l2_loss(m(x, cond), tgt)
    loss.backward()
    return loss
for _ in range(2):
    l = step()
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 3
for _ in range(n):
    l = step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
gn = float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters())))
flops = {1024: 15126.93e9, 512: 3745.49e9}.get(res, 0) * B
print(f"config5 B={B} {res}x{res}: {dt*1e3:.1f} ms/step, {B/dt:.2f} samples/s, {flops/dt/1e12:.0f} TFLOP/s algorithmic, "
      f"loss {float(l):.4f}, grad norm {gn:.4e}, finite={bool(torch.isfinite(l))}, "
      f"peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
