"""smoke(): one small invocation of the hot path on the GPU, checked against the CPU oracle."""
import torch


def run_smoke(dev):
    from live2diff_b200 import _lib
    from live2diff_b200.stream_pipeline import B200StreamPipeline
    from live2diff_b200.unet_step import B200UNetStep
    from live2diff_b200.weights import UNetDims, random_state_dict
    from oracle import schedule_oracle as S
    from oracle import unet_oracle as O

    d = UNetDims(block_out_channels=(64, 128, 128, 128), cross_attention_dim=96)
    n, h, w = 2, 16, 16
    sd = random_state_dict(d, seed=11)
    unet = B200UNetStep(sd, d, n, h, w, use_cuda_graph=False, device=dev)
    pipe = B200StreamPipeline(unet, [30, 40])
    gen = torch.Generator().manual_seed(5)
    prompt = torch.randn(1, 77, 96, generator=gen)
    kv = unet.prepare_cache(n)
    for c in kv:
        c[:, :, :, :8] = torch.randn(c[:, :, :, :8].shape, generator=gen).half().to(dev)
    kv32 = [c.float().cpu() for c in kv]
    pipe.prepare(prompt, kv)
    od = O.UNetDims(**d.__dict__)
    orc = S.StreamOracle(lambda s, t, **kw: O.unet_forward(sd, od, s, t, kw["encoder_hidden_states"],
                                                           kw["temporal_attention_mask"], kw["depth_sample"],
                                                           kw["kv_cache"], kw["pe_idx"], kw["update_idx"]),
                         kv32, prompt.half().float().repeat(n, 1, 1), [30, 40], (h, w))
    l0 = _lib.lib().l2d_launch_count()
    worst = 0.0
    for f in range(3):
        x = torch.randn(1, 4, 1, h, w, generator=gen).half()
        dep = torch.randn(1, 4, 1, h, w, generator=gen).half()
        noise = torch.randn(n - 1, 4, 1, h, w, generator=gen).half()
        out = pipe(x.to(dev), dep.to(dev), noise=noise.to(dev))
        ref = orc.step(x.float(), dep.float(), noise.float())
        worst = max(worst, float((out.float().cpu() - ref).abs().max()) / max(float(ref.abs().max()), 1.0))
    torch.cuda.synchronize(dev)
    launches = _lib.lib().l2d_launch_count() - l0
    print(f"[smoke] 3 stream frames on {torch.cuda.get_device_name(dev)}: worst rel-to-scale error vs CPU oracle "
          f"{worst:.3e}; {launches} kernels launched by libl2d_b200.so")
    assert worst < 2e-2, worst
    assert launches > 0
    # the VAE legs of the frame: TAESD encode + decode of one 64x64 image against the oracle restatement
    from live2diff_b200.taesd import B200TinyVAE, random_taesd_state_dict
    from oracle import taesd_oracle as T

    vsd = random_taesd_state_dict(3)
    vae = B200TinyVAE(vsd, 64, 64, device=dev)
    img = (torch.rand(1, 3, 64, 64, generator=gen) * 2 - 1).half()
    z = vae.encode(img.to(dev)).latents
    rec = vae.decode(z, return_dict=False)[0]
    z_ref = T.encode(vsd, img.float())
    rec_ref = T.decode(vsd, z.float().cpu())
    ez = float((z.float().cpu() - z_ref).abs().max()) / max(float(z_ref.abs().max()), 1.0)
    er = float((rec.float().cpu() - rec_ref).abs().max()) / max(float(rec_ref.abs().max()), 1.0)
    print(f"[smoke] TAESD 64x64: encode rel-to-scale error {ez:.3e}, decode {er:.3e}")
    assert ez < 2e-2 and er < 2e-2, (ez, er)
