"""Owner of one native ``afft_handle``: builds the ``afft_config`` from the Python module tree, registers the
weights (reference state-dict names) and issues ``afft_forward``.

One engine per (module, device, T).  Weights are re-packed into the library's bf16 K-major storage whenever a
parameter's version counter changes (``load_state_dict`` / ``init_model`` after construction, optimizer steps).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _capi

# GPT-2 buffers present in checkpoints written with transformers 4.18 that are not weights
IGNORED_STATE_KEYS = (".attn.bias", ".attn.masked_bias")


class Engine:
    def __init__(self, *, fuser_kind: int, T: int, mod_names: List[str], mod_dims: List[int], dim: int,
                 fuser_depth: int, fuser_heads: int, modal_encoding: bool, frame_level_token: bool, cross_attn: bool,
                 norm_elementwise: bool, gpt_dim: int, gpt_layers: int, gpt_heads: int, cls_names: List[str],
                 cls_dims: List[int], precision: str, max_batch: int, device: torch.device, fp_output_len: int = 1,
                 stages: int = _capi.STAGE_ALL):
        if device.type != "cuda":
            raise _capi.AfftError("afft_b200 runs on CUDA devices only (sm_100a); there is no CPU path")
        self.lib = _capi.lib()
        self.device = device
        self.T, self.dim, self.max_batch = T, dim, max_batch
        self.mod_names, self.mod_dims = list(mod_names), list(mod_dims)
        self.cls_names, self.cls_dims = list(cls_names), list(cls_dims)
        self.fuser_kind, self.fuser_depth, self.fuser_heads = fuser_kind, fuser_depth, fuser_heads
        self.frame_level_token = frame_level_token
        self.fp_output_len = fp_output_len
        cfg = _capi.Config()
        cfg.fuser_kind, cfg.T, cfg.n_mod = fuser_kind, T, len(mod_names)
        for i, (n, d) in enumerate(zip(mod_names, mod_dims)):
            cfg.mod_name[i].value = n.encode()
            cfg.mod_dim[i] = d
        cfg.dim, cfg.fuser_depth, cfg.fuser_heads = dim, fuser_depth, fuser_heads
        cfg.modal_encoding, cfg.frame_level_token = int(modal_encoding), int(frame_level_token)
        cfg.cross_attn, cfg.norm_elementwise = int(cross_attn), int(norm_elementwise)
        cfg.gpt_dim, cfg.gpt_layers, cfg.gpt_heads = gpt_dim, gpt_layers, gpt_heads
        cfg.n_cls = len(cls_names)
        for i, (n, d) in enumerate(zip(cls_names, cls_dims)):
            cfg.cls_name[i].value = n.encode()
            cfg.cls_dim[i] = d
        self.precision = precision
        cfg.precision, cfg.max_batch = _capi.PRECISIONS[precision], max_batch
        cfg.fp_output_len = fp_output_len
        cfg.stages = stages
        self.stages = stages
        self.gpt_dim, self.gpt_layers, self.gpt_heads = gpt_dim, gpt_layers, gpt_heads
        cfg.device = device.index if device.index is not None else torch.cuda.current_device()
        self.cfg = cfg
        # The workspace is a PyTorch-allocator tensor owned by this object (caller-owned device memory); the
        # library carves its activation buffers out of it and never allocates or synchronises for them.
        need = C.c_size_t()
        _capi.check(self.lib.afft_workspace_bytes_for(C.byref(cfg), C.byref(need)))
        self.workspace = torch.empty(int(need.value) + 256, dtype=torch.uint8, device=device)
        base = (self.workspace.data_ptr() + 255) // 256 * 256
        h = C.c_void_p()
        _capi.check(self.lib.afft_create_in(C.byref(cfg), base, need.value, _capi.current_stream_ptr(device), C.byref(h)))
        self.handle = h
        self.max_ksplit = 4  # library default
        self._versions: Optional[tuple] = None

    def __deepcopy__(self, memo):
        return None  # a native handle belongs to one module instance on one device; copies build their own

    def close(self):
        if getattr(self, "handle", None):
            self.lib.afft_destroy(self.handle)
            self.handle = None
            self.workspace = None  # returned to PyTorch's allocator (stream-ordered: queued kernels keep it valid)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights ----
    def sync_weights(self, named_params: Dict[str, torch.Tensor]):
        """named_params: names relative to ``future_predictor.`` -> fp32 CUDA tensors."""
        versions = tuple((n, p.data_ptr(), p._version) for n, p in named_params.items())
        if versions == self._versions:
            return
        stream = _capi.current_stream_ptr(self.device)
        for name, p in named_params.items():
            if any(name.endswith(k) for k in IGNORED_STATE_KEYS):
                continue
            if p.device != self.device:
                raise _capi.AfftError(f"parameter {name} is on {p.device}, engine is on {self.device}")
            t = p.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            shape = (C.c_int64 * t.dim())(*t.shape)
            _capi.check(self.lib.afft_set_weight(self.handle, name.encode(), t.data_ptr(), t.dim(), shape, stream),
                        self.handle)
        buf = C.create_string_buffer(4096)
        n_missing = self.lib.afft_missing_weights(self.handle, buf, len(buf))
        if n_missing:
            raise _capi.AfftError(f"{n_missing} weights missing in the module tree: {buf.value.decode()}")
        self._versions = versions

    # ---- forward ----
    @property
    def n_slots(self) -> int:
        n = len(self.mod_names)
        if self.fuser_kind == _capi.FUSER_SA:
            return n + 1
        if self.fuser_kind == _capi.FUSER_TSA:
            return n + (1 if self.frame_level_token else 0)
        return n

    def forward(self, feats: List[torch.Tensor], want_attn: bool = True, want_gpt_attn: bool = False):
        """feats: per modality (fusion order) (B, T, C_m) fp32 contiguous CUDA tensors.
        Returns (orig_past (B,T,D), past_futures_buf (B,T+O,D), [logits_buf (B,T+O,ld)], attn or None), O = fp_output_len;
        with want_gpt_attn a fifth element: the GPT-2 attention probabilities (B, layers, heads, T, T)."""
        B = feats[0].shape[0]
        T, D, dev = self.T, self.dim, self.device
        io = _capi.IO()
        for i, f in enumerate(feats):
            if f.device != dev or f.dtype != torch.float32 or not f.is_contiguous():
                raise _capi.AfftError("features must be contiguous fp32 tensors on the engine's device")
            if tuple(f.shape) != (B, T, self.mod_dims[i]):
                raise _capi.AfftError(f"feature {self.mod_names[i]} has shape {tuple(f.shape)}, expected {(B, T, self.mod_dims[i])}")
            io.feat[i] = f.data_ptr()
        orig_past = torch.empty(B, T, D, device=dev, dtype=torch.float32)
        S = T + self.fp_output_len
        pf = torch.empty(B, S, D, device=dev, dtype=torch.float32)
        logits = []
        for k, c in enumerate(self.cls_dims):
            ld = (c + 3) // 4 * 4
            buf = torch.empty(B, S, ld, device=dev, dtype=torch.float32)
            logits.append(buf)
            io.logits[k] = buf.data_ptr()
            io.ld_logits[k] = ld
        io.orig_past, io.past_futures = orig_past.data_ptr(), pf.data_ptr()
        attn = None
        if want_attn and self.fuser_kind not in (_capi.FUSER_CA, _capi.FUSER_NONE):
            n, H = self.n_slots, self.fuser_heads
            if self.fuser_kind == _capi.FUSER_TSA:
                attn = torch.empty(B, self.fuser_depth, H, n * T, n * T, device=dev, dtype=torch.float32)
            else:
                attn = torch.empty(B, self.fuser_depth, T, H, n, n, device=dev, dtype=torch.float32)
            io.fuser_attn = attn.data_ptr()
        gpt_attn = None
        if want_gpt_attn:
            gpt_attn = torch.empty(B, self.gpt_layers, self.gpt_heads, T, T, device=dev, dtype=torch.float32)
            io.gpt_attn = gpt_attn.data_ptr()
        _capi.check(self.lib.afft_forward(self.handle, B, C.byref(io), _capi.current_stream_ptr(dev)), self.handle)
        return (orig_past, pf, logits, attn, gpt_attn) if want_gpt_attn else (orig_past, pf, logits, attn)

    def forward_fuser(self, feats: List[torch.Tensor], want_attn: bool = True):
        """AFFT_STAGE_FUSER handle: (fused (B, T, D), attention or None)."""
        B = feats[0].shape[0]
        T, D, dev = self.T, self.dim, self.device
        io = _capi.IO()
        for i, f in enumerate(feats):
            if f.device != dev or f.dtype != torch.float32 or not f.is_contiguous():
                raise _capi.AfftError("features must be contiguous fp32 tensors on the engine's device")
            if tuple(f.shape) != (B, T, self.mod_dims[i]):
                raise _capi.AfftError(f"feature {self.mod_names[i]} has shape {tuple(f.shape)}, expected {(B, T, self.mod_dims[i])}")
            io.feat[i] = f.data_ptr()
        fused = torch.empty(B, T, D, device=dev, dtype=torch.float32)
        io.orig_past = fused.data_ptr()
        attn = None
        if want_attn and self.fuser_kind != _capi.FUSER_CA:
            n, H = self.n_slots, self.fuser_heads
            shape = (B, self.fuser_depth, H, n * T, n * T) if self.fuser_kind == _capi.FUSER_TSA else (B, self.fuser_depth, T, H, n, n)
            attn = torch.empty(*shape, device=dev, dtype=torch.float32)
            io.fuser_attn = attn.data_ptr()
        _capi.check(self.lib.afft_forward(self.handle, B, C.byref(io), _capi.current_stream_ptr(dev)), self.handle)
        return fused, attn

    def forward_gpt(self, feats: torch.Tensor, want_attn: bool = False):
        """AFFT_STAGE_GPT handle: feats (B, T, G) fp32 -> (hidden states (B, T + O - 1, G), attention (B, layers, H, T, T) or None)."""
        B, T, G, dev, O = feats.shape[0], self.T, self.gpt_dim, self.device, self.fp_output_len
        if feats.device != dev or feats.dtype != torch.float32 or not feats.is_contiguous() or tuple(feats.shape) != (B, T, G):
            raise _capi.AfftError(f"predictor input must be a contiguous fp32 (B, {T}, {G}) tensor on the engine's device")
        io = _capi.IO()
        io.feat[0] = feats.data_ptr()
        prompt = torch.empty(B, T, G, device=dev, dtype=torch.float32)
        io.orig_past = prompt.data_ptr()
        gpt_attn = None
        if want_attn:
            gpt_attn = torch.empty(B, self.gpt_layers, self.gpt_heads, T, T, device=dev, dtype=torch.float32)
            io.gpt_attn = gpt_attn.data_ptr()
        new = None
        if O > 1:
            new = torch.empty(B, O - 1, G, device=dev, dtype=torch.float32)
            io.past_futures = new.data_ptr()
        _capi.check(self.lib.afft_forward(self.handle, B, C.byref(io), _capi.current_stream_ptr(dev)), self.handle)
        return (prompt if new is None else torch.cat([prompt, new], dim=1)), gpt_attn

    def forward_into(self, io: "_capi.IO", B: int):
        """Lowest-overhead call for benchmarking: caller pre-fills the io struct with persistent buffers."""
        _capi.check(self.lib.afft_forward(self.handle, B, C.byref(io), _capi.current_stream_ptr(self.device)),
                    self.handle)

    def profile_enable(self, on: bool = True):
        _capi.check(self.lib.afft_profile_enable(self.handle, int(on)), self.handle)

    def profile_read(self):
        """[(category, M, N, K, ms)] of the most recent forward (categories: 0 GEMM, 1 LN, 2 attention, 3 other)."""
        p = _capi.Profile()
        _capi.check(self.lib.afft_profile_read(self.handle, C.byref(p)), self.handle)
        return [(r.cat, r.M, r.N, r.K, r.ms) for r in p.recs[:p.n]]

    def set_max_ksplit(self, max_split: int):
        """Upper bound on the K splits of skinny GEMMs (1 = off); see ``afft_set_max_ksplit``."""
        _capi.check(self.lib.afft_set_max_ksplit(self.handle, int(max_split)), self.handle)
        self.max_ksplit = int(max_split)

    def launch_count(self) -> int:
        return int(self.lib.afft_last_launch_count(self.handle))

    def workspace_bytes(self) -> int:
        return int(self.lib.afft_workspace_bytes(self.handle))

    def weight_bytes(self) -> int:
        return int(self.lib.afft_weight_bytes(self.handle))
