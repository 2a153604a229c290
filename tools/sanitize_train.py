"""One tiny training step (TrainState: direct gradients, attention dropout, dual conversion, native SGD) and one forward
with the TMA-staged epilogue - targets for compute-sanitizer.  usage: python tools/sanitize_train.py [config]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from afft_b200 import _capi, configs, synthetic, train as atrain  # noqa: E402
from afft_b200.models import BaseModel  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "egtea_sa"
cfg, T, ncls, _ = configs.named_config(name)
KW = dict(mixup_fn=None, target=None, target_subclips=None, target_subclips_ignore_index=None)
B, C = 4, list(ncls.values())[0]
model = BaseModel(cfg, ncls, {})
model.load_state_dict(synthetic.synthetic_state_dict(model, seed=0))
model = model.to("cuda:0").train()
feats = {m: t.reshape(B, T, -1, 1, 1, 1).cuda() for m, t in synthetic.synthetic_features(cfg["modal_dims"], B, T, seed=1).items()}
target = torch.zeros(B, 1, dtype=torch.long, device="cuda:0")
tsub = torch.zeros(B, T, dtype=torch.long, device="cuda:0")
state = atrain.TrainState(model.future_predictor, lr=1e-3, momentum=0.9, weight_decay=1e-6)
with state:
    for _ in range(2):
        state.zero()
        out, _ = model(dict(feats), **KW)
        atrain.reference_losses(out, target, tsub)["total"].backward()
        state.finish()
        state.step()
torch.cuda.synchronize()
model.eval()
_capi.check(_capi.lib().afft_set_gemm_epilogue(1))
big = {m: t.repeat(40, 1, 1, 1, 1, 1) for m, t in feats.items()}  # M > 128: the 2-CTA kernel with the TMA-staged epilogue
with torch.no_grad():
    model(big, **KW)
torch.cuda.synchronize()
_capi.check(_capi.lib().afft_set_gemm_epilogue(0))
print("done")
