import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bubbleformer_b200 import get_model
from bubbleformer_b200.rollout import GraphedStep
from oracle.param_init import fluid_params
torch.manual_seed(3)
m = get_model("filmavit", input_fields=4, output_fields=4, time_window=5, patch_size=16, embed_dim=128, num_heads=2,
              processor_blocks=2, num_fluid_params=9).cuda().eval()
with torch.no_grad():
    for n, p in m.named_parameters():
        if "gamma" in n:
            p.copy_(0.05 * torch.randn_like(p))
x0 = torch.randn(1, 5, 4, 64, 64, device="cuda")
cond = fluid_params(1).cuda()
rel = lambda a, b: float((a - b).norm() / b.norm())
with torch.no_grad():
    e1 = m(x0, cond).clone(); e2 = m(x0, cond).clone()
    print("eager vs eager", rel(e1, e2))
    g = GraphedStep(m, x0, cond)
    g1 = g(x0, cond).clone(); g2 = g(x0, cond).clone()
    print("graph vs graph", rel(g1, g2), "graph vs eager", rel(g1, e1))
    x1 = torch.randn_like(x0)
    print("graph new input vs eager", rel(g(x1, cond).clone(), m(x1, cond)))
