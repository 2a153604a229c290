"""Host-side mirror of the reference's ``models/`` package (same module, class and parameter names).

``BaseModel`` / ``CMFPEarly`` keep the reference's constructor and forward signatures
(reference models/base_model.py:15-17,68; models/future_prediction.py:228-291) and own their weights
as ordinary ``nn.Parameter``s under the reference's state-dict keys, so ``init_model`` /
``store_checkpoint`` (train.py:55-103,156-167) work unchanged.  The arithmetic of the whole path runs in
libafft_b200 (one ``afft_forward`` call per crop); there is no PyTorch fallback.
"""
from .base_model import BaseModel  # noqa: F401
