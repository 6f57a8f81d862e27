"""`B200TinyVAE`: the object installed at `stream.vae`.

Call contract = what the reference pipeline uses of `diffusers.AutoencoderTiny`
(live2diff/utils/wrapper.py:468-470; live2diff/pipeline_stream_animation_depth.py:517-542, 565-571):

    vae.encode(x).latents            x [n,3,H,W] in [-1,1]   -> [n,4,H/8,W/8]     (retrieve_latents reads `.latents`)
    vae.decode(z, return_dict=False)[0]                      -> [n,3,H,W]
    vae.config.scaling_factor (1.0), vae.dtype, vae.device

plus `preprocess_u8` / `postprocess_u8`, the uint8 <-> [-1,1] conversions of `__call__` (:630) and
image_utils.postprocess_image, as single kernels.  Everything runs inside libl2d_b200.so (csrc/taesd.cu): every 3x3
convolution is a tcgen05 implicit GEMM.  Weights = `AutoencoderTiny.state_dict()` (same keys), so a real TAESD checkpoint
loads unchanged; `random_taesd_state_dict` gives seeded weights with the real shapes (no checkpoint can be downloaded here).
"""
from __future__ import annotations

import ctypes as C
import math
from collections import OrderedDict
from types import SimpleNamespace
from typing import Dict, Optional, Tuple

import torch

from ._lib import check, current_stream, lib, make_tensor_table

CH = 64
ENC_BLOCKS = (1, 3, 3, 3)
DEC_BLOCKS = (3, 3, 3, 1)


def taesd_param_spec() -> "OrderedDict[str, Tuple[int, ...]]":
    """name -> shape in AutoencoderTiny.state_dict() order (nn.Sequential indices incl. the ReLU / Upsample slots)."""
    s: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def conv(p, cin, cout, bias=True):
        s[p + ".weight"] = (cout, cin, 3, 3)
        if bias:
            s[p + ".bias"] = (cout,)

    def block(p):
        for i in (0, 2, 4):
            conv(f"{p}.conv.{i}", CH, CH)

    i = 0
    for stage, nb in enumerate(ENC_BLOCKS):
        conv(f"encoder.layers.{i}", 3 if stage == 0 else CH, CH, bias=stage == 0)
        i += 1
        for _ in range(nb):
            block(f"encoder.layers.{i}")
            i += 1
    conv(f"encoder.layers.{i}", CH, 4)
    conv("decoder.layers.0", 4, CH)
    i = 2
    for stage, nb in enumerate(DEC_BLOCKS):
        for _ in range(nb):
            block(f"decoder.layers.{i}")
            i += 1
        last = stage == len(DEC_BLOCKS) - 1
        if not last:
            i += 1
        conv(f"decoder.layers.{i}", CH, 3 if last else CH, bias=last)
        i += 1
    return s


def random_taesd_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded fp32 weights with the real TAESD shapes; conv weights ~ N(0, gain / fan_in) so activations stay O(1)
    through the 40 layers, biases small."""
    gen = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, shape in taesd_param_spec().items():
        if name.endswith(".weight"):
            fan_in = shape[1] * 9
            # the last conv of a residual block is scaled down so that 10 stacked blocks do not blow up
            gain = 0.5 if ".conv.4." in name else 1.4
            sd[name] = torch.randn(shape, generator=gen) * (gain / math.sqrt(fan_in))
        else:
            sd[name] = torch.randn(shape, generator=gen) * 0.05
    return sd


class _EncodeOutput:
    def __init__(self, latents):
        self.latents = latents


class B200TinyVAE:
    def __init__(self, state_dict: Dict[str, torch.Tensor], height: int, width: int, max_batch: int = 1,
                 device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("B200TinyVAE needs a CUDA device; live2diff_b200 has no CPU fallback")
        self.device = torch.device(device or "cuda")
        self.dtype = torch.float16
        self.config = SimpleNamespace(scaling_factor=1.0, latent_channels=4)
        self.height, self.width, self.max_batch = height, width, max_batch
        spec = taesd_param_spec()
        missing = [k for k in spec if k not in state_dict]
        if missing:
            raise KeyError(f"state_dict is missing {len(missing)} tensors, e.g. {missing[:3]}")
        named = {}
        for k, shape in spec.items():
            t = state_dict[k]
            if tuple(t.shape) != tuple(shape):
                raise ValueError(f"{k}: shape {tuple(t.shape)} != expected {tuple(shape)}")
            named[k] = t.detach().to(device=self.device, dtype=torch.float16).contiguous()
        arr, keep = make_tensor_table(named)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().l2d_taesd_create(C.byref(handle), arr, len(named), max_batch, height, width))
        self._handle = handle
        del named, arr, keep

    def __del__(self):
        if getattr(self, "_handle", None):
            lib().l2d_taesd_destroy(self._handle)
            self._handle = None

    def to(self, *args, **kwargs):
        return self

    @property
    def device_bytes(self) -> int:
        return lib().l2d_taesd_device_bytes(self._handle)

    def _check(self, t, c, h, w, name):
        if t.dim() != 4 or t.shape[1] != c or t.shape[2] != h or t.shape[3] != w or t.shape[0] > self.max_batch:
            raise ValueError(f"{name}: expected [n<={self.max_batch},{c},{h},{w}], got {tuple(t.shape)}")
        return t.to(device=self.device, dtype=torch.float16).contiguous()

    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        x = self._check(x, 3, self.height, self.width, "image")
        z = torch.empty(x.shape[0], 4, self.height // 8, self.width // 8, dtype=torch.float16, device=self.device)
        check(lib().l2d_taesd_encode(self._handle, x.data_ptr(), z.data_ptr(), x.shape[0], current_stream()))
        return _EncodeOutput(z) if return_dict else (z,)

    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True, clip: bool = False):
        z = self._check(z, 4, self.height // 8, self.width // 8, "latents")
        img = torch.empty(z.shape[0], 3, self.height, self.width, dtype=torch.float16, device=self.device)
        check(lib().l2d_taesd_decode(self._handle, z.data_ptr(), img.data_ptr(), z.shape[0], int(clip), current_stream()))
        return SimpleNamespace(sample=img) if return_dict else (img,)

    # ---- pre / post-processing of __call__ ------------------------------------------------------------------------
    @torch.no_grad()
    def preprocess_u8(self, frame_u8: torch.Tensor) -> torch.Tensor:
        """uint8 [n,H,W,3] (CUDA) -> fp16 [n,3,H,W] in [-1,1]."""
        if frame_u8.dtype != torch.uint8 or frame_u8.dim() != 4 or frame_u8.shape[3] != 3 or not frame_u8.is_cuda:
            raise ValueError("frame must be a CUDA uint8 tensor [n,H,W,3]")
        frame_u8 = frame_u8.contiguous()
        n, h, w, _ = frame_u8.shape
        out = torch.empty(n, 3, h, w, dtype=torch.float16, device=frame_u8.device)
        check(lib().l2d_image_u8_to_f16(frame_u8.data_ptr(), out.data_ptr(), n, h, w, current_stream()))
        return out

    @torch.no_grad()
    def postprocess_u8(self, img: torch.Tensor) -> torch.Tensor:
        """fp16 [n,3,H,W] in [-1,1] -> uint8 [n,H,W,3]."""
        img = img.to(dtype=torch.float16).contiguous()
        n, _, h, w = img.shape
        out = torch.empty(n, h, w, 3, dtype=torch.uint8, device=img.device)
        check(lib().l2d_image_f16_to_u8(img.data_ptr(), out.data_ptr(), n, h, w, current_stream()))
        return out
