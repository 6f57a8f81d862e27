"""a1: `B200ImageStream.__call__` = StreamAnimateDiffusionDepth.__call__ (pipeline:625-660) image in ->
image out, against the composition of the oracles (TAESD restatement -- unpinned, see oracle/taesd_oracle.py -- around the
pinned stream / UNet oracle), with injected noise.  The depth prior is supplied as a map (MiDaS itself is §8 f2)."""
import pytest
import torch

from live2diff_b200.weights import UNetDims, random_state_dict
from oracle import schedule_oracle as S
from oracle import taesd_oracle as T
from oracle import unet_oracle as O
from parity import referee

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_image_stream_matches_oracle_composition():
    from live2diff_b200.image_pipeline import B200ImageStream
    from live2diff_b200.taesd import B200TinyVAE, random_taesd_state_dict
    from live2diff_b200.unet_step import B200UNetStep

    d = UNetDims(block_out_channels=(64, 128, 128, 128), cross_attention_dim=96)
    n, H, W = 2, 128, 128
    h, w = H // 8, W // 8
    sd = random_state_dict(d, seed=11)
    vsd = random_taesd_state_dict(5)
    unet = B200UNetStep(sd, d, n, h, w, use_cuda_graph=True, device=DEV)
    vae = B200TinyVAE(vsd, H, W, max_batch=1, device=DEV)
    pipe = B200ImageStream(unet, vae, [30, 40])
    gen = torch.Generator().manual_seed(5)
    prompt = torch.randn(1, 77, 96, generator=gen)
    kv = unet.prepare_cache(n)
    for c in kv:
        c[:, :, :, :8] = torch.randn(c[:, :, :, :8].shape, generator=gen).half().to(DEV)
    kv32 = [c.float().cpu() for c in kv]
    pipe.prepare(prompt, kv)
    od = O.UNetDims(**d.__dict__)
    orc = S.StreamOracle(lambda s, t, **kw: O.unet_forward(sd, od, s, t, kw["encoder_hidden_states"],
                                                           kw["temporal_attention_mask"], kw["depth_sample"],
                                                           kw["kv_cache"], kw["pe_idx"], kw["update_idx"]),
                         kv32, prompt.half().float().repeat(n, 1, 1), [30, 40], (h, w))
    # the same composition evaluated in fp16 with torch ops on the GPU = the reference's own rounding order (referee)
    sd16 = {k: v.to(DEV).half() for k, v in sd.items()}
    vsd16 = {k: v.to(DEV).half() for k, v in vsd.items()}
    kv16 = [c.clone() for c in kv]
    orc16 = S.StreamOracle(lambda s, t, **kw: O.unet_forward(sd16, od, s, t, kw["encoder_hidden_states"],
                                                             kw["temporal_attention_mask"], kw["depth_sample"],
                                                             kw["kv_cache"], kw["pe_idx"], kw["update_idx"]),
                           kv16, prompt.half().to(DEV).repeat(n, 1, 1), [30, 40], (h, w), dtype=torch.float16, device=DEV)
    _, _, _, a, b = S.stream_constants([30, 40])
    worst, worst_u8, mean_u8 = 0.0, 0, 0.0
    for f in range(3):
        frame = torch.randint(0, 256, (H, W, 3), generator=gen, dtype=torch.uint8)
        depth = torch.rand(1, H, W, generator=gen)
        noise0 = torch.randn(1, 4, h, w, generator=gen).half()
        renoise = torch.randn(n - 1, 4, 1, h, w, generator=gen).half()
        out = pipe(frame.to(DEV), depth_map=depth.to(DEV), noise=noise0.to(DEV), renoise=renoise.to(DEV))
        out_u8 = vae.postprocess_u8(out)[0].cpu()
        # oracle composition (fp32)
        x = T.preprocess_u8(frame[None])
        x_t = T.encode_image(vsd, x, noise0.float(), float(a[0]), float(b[0]))
        dl = T.encode_depth_map(vsd, depth.half().float())
        x0 = orc.step(x_t[:, :, None], dl[:, :, None], renoise.float())
        ref = T.decode_image(vsd, x0[:, :, 0])
        ref_u8 = T.postprocess_u8(ref)[0]
        x16 = T.preprocess_u8(frame[None]).half().to(DEV)
        x_t16 = T.encode_image(vsd16, x16, noise0.to(DEV), float(a[0]), float(b[0]))
        dl16 = T.encode_depth_map(vsd16, depth.half().to(DEV))
        x016 = orc16.step(x_t16[:, :, None], dl16[:, :, None], renoise.to(DEV))
        ref16 = T.decode_image(vsd16, x016[:, :, 0])
        referee(out, ref, ref16, f"image stream frame {f}", slack=2.5)
        err = float((out.float().cpu() - ref).abs().max())
        worst = max(worst, err)
        du8 = (out_u8.int() - ref_u8.int()).abs()
        worst_u8, mean_u8 = max(worst_u8, int(du8.max())), max(mean_u8, float(du8.float().mean()))
    print(f"[image stream] worst |x_output - oracle| = {worst:.3e} (range [-1,1]); uint8: max diff {worst_u8}, mean diff {mean_u8:.3f}")
    assert worst < 0.25 and mean_u8 < 1.5     # loose absolute guard; the referee above is the criterion
    assert out.shape == (1, 3, H, W) and out_u8.shape == (H, W, 3)
    # uint8 in / uint8 out convenience path
    res = pipe.frame_u8(frame.to(DEV), depth_map=depth.to(DEV))
    assert res.shape == (H, W, 3) and res.dtype == torch.uint8
