"""f3: B200TinyVAE (csrc/taesd.cu) against the oracle restatement of diffusers AutoencoderTiny
(oracle/taesd_oracle.py -- parity UNPINNED: diffusers is neither vendored in the reference nor installed here), plus the
uint8 pre/post-processing kernels, which are exact integer -> fp16 / fp16 -> integer maps and must match bit for bit."""
import pytest
import torch

from oracle import taesd_oracle as T
from parity import referee

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make(h, w, n=1, seed=0):
    from live2diff_b200.taesd import B200TinyVAE, random_taesd_state_dict

    sd = random_taesd_state_dict(seed)
    return sd, B200TinyVAE(sd, h, w, max_batch=n, device=DEV)


@pytest.mark.parametrize("n,h,w", [(1, 64, 64), (2, 128, 64), (1, 512, 512)])
def test_taesd_encode_vs_oracle(n, h, w):
    sd, vae = make(h, w, n)
    x = (torch.rand(n, 3, h, w, generator=torch.Generator().manual_seed(1)) * 2 - 1).half()
    z = vae.encode(x.to(DEV)).latents
    assert z.shape == (n, 4, h // 8, w // 8)
    ref32 = T.encode(sd, x.float())
    sd16 = {k: v.to(DEV).half() for k, v in sd.items()}
    ref16 = T.encode(sd16, x.to(DEV))
    referee(z, ref32, ref16, f"taesd encode {n}x{h}x{w}", slack=2.5)


@pytest.mark.parametrize("n,h,w", [(1, 64, 64), (2, 64, 128), (1, 512, 512)])
def test_taesd_decode_vs_oracle(n, h, w):
    sd, vae = make(h, w, n, seed=3)
    z = (torch.randn(n, 4, h // 8, w // 8, generator=torch.Generator().manual_seed(2)) * 2).half()
    img = vae.decode(z.to(DEV), return_dict=False)[0]
    assert img.shape == (n, 3, h, w)
    ref32 = T.decode(sd, z.float())
    sd16 = {k: v.to(DEV).half() for k, v in sd.items()}
    ref16 = T.decode(sd16, z.to(DEV))
    referee(img, ref32, ref16, f"taesd decode {n}x{h}x{w}", slack=2.5)
    clipped = vae.decode(z.to(DEV), return_dict=False, clip=True)[0]
    assert torch.equal(clipped, img.clip(-1, 1))


def test_u8_pre_and_post_processing_are_exact():
    from live2diff_b200.taesd import B200TinyVAE  # noqa: F401

    _, vae = make(64, 64)
    u8 = torch.arange(256, dtype=torch.uint8).repeat(48)[: 64 * 64 * 3].reshape(1, 64, 64, 3)
    x = vae.preprocess_u8(u8.to(DEV))
    assert torch.equal(x.cpu(), T.preprocess_u8(u8).half())
    # every fp16 value in [-1.5, 1.5] on a fine grid: the uint8 map must equal the reference's fp16 -> float32 -> round path
    img = torch.linspace(-1.5, 1.5, 3 * 64 * 64).half().reshape(1, 3, 64, 64)
    out = vae.postprocess_u8(img.to(DEV))
    ref = ((img / 2 + 0.5).clamp(0, 1)).float().mul(255).round().to(torch.uint8).permute(0, 2, 3, 1)
    assert torch.equal(out.cpu(), ref)
    # round trip of a frame through pre -> post
    assert torch.equal(vae.postprocess_u8(x).cpu(), u8)
