"""Two forwards of the headline config at B clips (for ncu captures of the in-situ kernels).
usage: python tools/ncu_forward.py [B] [config] [bf16|fp16|strict]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afft_b200 import configs  # noqa: E402
from afft_b200.models import BaseModel  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
name = sys.argv[2] if len(sys.argv) > 2 else "ek100_sa_tsn"
precision = sys.argv[3] if len(sys.argv) > 3 else "bf16"
cfg, T, ncls, _ = configs.named_config(name)
torch.manual_seed(0)
dev = torch.device("cuda:0")
model = BaseModel(cfg, ncls, {}, precision=precision, max_batch=B).to(dev).eval()
feats = {m: torch.randn(B, T, d, 1, 1, 1, device=dev) for m, d in cfg["modal_dims"].items()}
kw = dict(mixup_fn=None, target=None, target_subclips=None, target_subclips_ignore_index=None)
with torch.no_grad():
    for _ in range(3):
        model(dict(feats), **kw)
torch.cuda.synchronize()
print("launches per forward:", model.future_predictor.last_launch_count())
