"""Drop-in for reference models/base_model.py: same constructor, same call signature, same output dicts.

    model = BaseModel(cfg.model, num_classes, class_mappings)                # test.py:119, train.py:315
    outputs, outputs_target = model(feature_dict, mixup_fn=None, target=None,
                                    target_subclips=None, target_subclips_ignore_index=None)   # test.py:72-82
    outputs['logits/action']['all-fused'][:, 0, :]                            # test.py:86

Only the glue lives here (crop handling, the Identity backbones, (B, #clips, C, 1, 1, 1) -> (B, T, C)); the
arithmetic under ``self.future_predictor`` runs in libafft_b200.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn

from .future_prediction import cfg_get, cfg_items, instantiate

CLS_MAP_PREFIX = 'cls_map_'
PAST_LOGITS_PREFIX = 'past_'


class BaseModel(nn.Module):
    def __init__(self, model_cfg, num_classes: Dict[str, int],
                 class_mappings: Dict[Tuple[str, str], torch.FloatTensor], strict: bool = False,
                 max_batch: int = 64, precision: str = None):
        super().__init__()
        self.backbone = nn.ModuleDict()
        for mod, backbone_conf in cfg_items(cfg_get(cfg_get(model_cfg, "common"), "backbones")):
            bb = instantiate(backbone_conf)
            if not isinstance(bb, nn.Identity):
                raise NotImplementedError("only torch.nn.Identity backbones (pre-extracted features) are supported")
            self.backbone[mod] = bb
        self.future_predictor = instantiate(cfg_get(model_cfg, "CMFP"), model_cfg=model_cfg, num_classes=num_classes,
                                            strict=strict, max_batch=max_batch, precision=precision)
        for (src, dst), mapping in class_mappings.items():  # reference base_model.py:27-29
            self.register_buffer(f'{CLS_MAP_PREFIX}{src}_{dst}', mapping)

    @staticmethod
    def _to_sequence(data: torch.Tensor) -> torch.Tensor:
        """(B, #clips, C, T', H, W) -> (B, #clips*T', C): spatial mean, channels last (reference :40-46)."""
        if data.shape[-3:] == (1, 1, 1):
            return data.reshape(data.shape[0], data.shape[1], data.shape[2])  # mean over 1x1 is the identity
        feats = torch.mean(data, [-1, -2]).permute((0, 1, 3, 2))
        return torch.flatten(feats, 1, 2) if feats.ndim == 4 else feats

    def forward_singlecrop(self, data_dict, **kwargs):
        feats_past = {mod: self._to_sequence(self.backbone[mod](data)) for mod, data in data_dict.items()}
        target = kwargs['target']
        target_subclips = kwargs['target_subclips']
        target_subclips_ignore_index = kwargs['target_subclips_ignore_index']
        if kwargs['mixup_fn'] is not None:  # reference :53-56 (training only)
            feats_past, target, target_subclips, target_subclips_ignore_index = \
                kwargs['mixup_fn'](feats_past, target, target_subclips)
        outputs = self.future_predictor(feats_past)
        outputs_target = {'target': target, 'target_subclips': target_subclips,
                          'target_subclips_ignore_index': target_subclips_ignore_index}
        return outputs, outputs_target

    def forward(self, video_data, *args, **kwargs):
        """video_data: {mod: (B, #clips, C, T, H, W) or (B, #clips, #crops, C, T, H, W)} (reference :68-119)."""
        for key in ('mixup_fn', 'target', 'target_subclips', 'target_subclips_ignore_index'):
            kwargs.setdefault(key, None)
        crops = {}
        for mod, data in video_data.items():
            if data.ndim == 6:
                crops[mod] = [data]
            elif data.ndim == 7:
                crops[mod] = [data.squeeze(2)] if data.size(2) == 1 else list(torch.unbind(data, dim=2))
            else:
                raise NotImplementedError('Unsupported size %s' % (data.shape,))
        mods = sorted(crops)
        num_crops = max(len(crops[m]) for m in mods)
        per_crop = [{m: crops[m][i % len(crops[m])] for m in mods} for i in range(num_crops)]
        results = [self.forward_singlecrop(el, *args, **kwargs) for el in per_crop]
        output_targets = results[0][1]
        if num_crops == 1:
            return results[0][0], output_targets
        merged = {}
        for out_key in results[0][0]:
            if out_key == 'attentions':  # first crop only, as the reference
                merged[out_key] = results[0][0][out_key]
                continue
            merged[out_key] = {k: torch.mean(torch.stack([r[0][out_key][k] for r in results], dim=0), dim=0)
                               for k in results[0][0][out_key]}
        return merged, output_targets
