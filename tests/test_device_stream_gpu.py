"""device-resident stream state (`B200DeviceStream` over l2d_stream_*): whole-frame CUDA graph, ring
schedule advanced on the device, counter-based re-noise, pinned-host I/O, state save/load."""
import pytest
import torch

from live2diff_b200.weights import UNetDims, random_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TINY = UNetDims(block_out_channels=(64, 128, 128, 128), cross_attention_dim=96)
N, H, W = 2, 16, 16


def make(seed_w=7, graph=True, stream_seed=2, do_add_noise=True, unet=None):
    from live2diff_b200.device_stream import B200DeviceStream
    from live2diff_b200.unet_step import B200UNetStep

    if unet is None:
        unet = B200UNetStep(random_state_dict(TINY, seed=seed_w), TINY, N, H, W, use_cuda_graph=graph)
    return unet, B200DeviceStream(unet, [30, 40], seed=stream_seed, do_add_noise=do_add_noise, use_cuda_graph=graph)


def warm_caches(unet, seed):
    gen = torch.Generator().manual_seed(seed)
    kv = unet.prepare_cache(N)
    for c in kv:
        c[:, :, :, :8] = torch.randn(c[:, :, :, :8].shape, generator=gen).half().to(DEV)
    return kv


def close(a, b, what, tol=5e-3):
    err = float((a.float() - b.float()).abs().max())
    scale = max(float(b.float().abs().max()), 1.0)
    assert torch.isfinite(a.float()).all() and err <= tol * scale, f"{what}: max-abs diff {err:.3e} (scale {scale:.3e})"


@pytest.mark.parametrize("graph", [False, True])
def test_device_stream_matches_host_scheduled_pipeline(graph):
    """Same engine, same inputs, injected noise: the device-resident state machine must reproduce B200StreamPipeline
    (host-side schedule, parity-tested against the stream oracle) frame by frame, through fill phase and first wrap,
    and its on-device schedule must equal the host RingSchedule after every frame."""
    from live2diff_b200.stream_pipeline import B200StreamPipeline

    unet, ds = make(graph=graph)
    pipe = B200StreamPipeline(unet, [30, 40])
    gen = torch.Generator().manual_seed(3)
    prompt = torch.randn(1, 77, 96, generator=gen)
    kv_a, kv_b = warm_caches(unet, 11), warm_caches(unet, 11)
    pipe.prepare(prompt, kv_a)
    ds.prepare(prompt, kv_b)
    s0 = ds.schedule()
    assert (s0["valid"], s0["pe_idx"], s0["update_idx"], s0["frame"]) == (pipe.schedule.valid, pipe.schedule.pe_idx,
                                                                         pipe.schedule.update_idx, 0)
    for f in range(13):
        x = torch.randn(1, 4, 1, H, W, generator=gen).half().to(DEV)
        dep = torch.randn(1, 4, 1, H, W, generator=gen).half().to(DEV)
        noise = torch.randn(N - 1, 4, 1, H, W, generator=gen).half().to(DEV)
        ref = pipe(x, dep, noise=noise).clone()
        out = ds(x, dep, noise=noise).clone()
        close(out, ref, f"frame {f}", tol=0.0)     # same engine, same kernels, deterministic reductions: bit for bit
        s = ds.schedule()
        assert (s["valid"], s["pe_idx"], s["update_idx"], s["frame"]) == (pipe.schedule.valid, pipe.schedule.pe_idx,
                                                                         pipe.schedule.update_idx, f + 1), f"frame {f}"
    for i in (0, 17, 39):
        close(kv_b[i], kv_a[i], f"kv[{i}]", tol=0.0)
    assert ds.launches_per_frame > unet.launches_per_step > 0


def test_device_stream_pinned_host_io_and_prompt_update():
    unet, ds = make()
    _, ds2 = make(unet=unet)
    gen = torch.Generator().manual_seed(5)
    prompt = torch.randn(1, 77, 96, generator=gen)
    ds.prepare(prompt, warm_caches(unet, 1))
    ds2.prepare(prompt, warm_caches(unet, 1))
    host_out = torch.empty(1, 4, 1, H, W, dtype=torch.float16).pin_memory()
    for f in range(4):
        if f == 2:                                                    # update_prompt mid-stream (:368-376)
            prompt = torch.randn(1, 77, 96, generator=gen)
            ds.update_prompt(prompt)
            ds2.update_prompt(prompt.repeat(N, 1, 1))                 # N-row form
        x = torch.randn(1, 4, 1, H, W, generator=gen).half()
        dep = torch.randn(1, 4, 1, H, W, generator=gen).half()
        noise = torch.randn(N - 1, 4, 1, H, W, generator=gen).half()
        ds(x.pin_memory(), dep.pin_memory(), noise=noise.pin_memory(), out=host_out)
        torch.cuda.current_stream().synchronize()
        ref = ds2(x.to(DEV), dep.to(DEV), noise=noise.to(DEV))
        close(host_out.to(DEV), ref, f"pinned-host frame {f}")
    with pytest.raises(ValueError):
        ds(torch.zeros(1, 4, 1, H, W, dtype=torch.float16), dep.to(DEV))          # pageable host memory is refused
    with pytest.raises(ValueError):
        ds.update_prompt(torch.zeros(1, 10, 96))


def test_device_stream_internal_noise_is_seeded_and_state_round_trips():
    """Internal Philox re-noise: same seed -> same stream, other seed -> different; save_state / load_state moves a
    stream (buffers + schedule + frame counter + seed) onto another l2d_stream, which then continues identically."""
    unet, a = make(stream_seed=5)
    _, b = make(unet=unet, stream_seed=5)
    _, c = make(unet=unet, stream_seed=6)
    gen = torch.Generator().manual_seed(9)
    prompt = torch.randn(1, 77, 96, generator=gen)
    kvs = [warm_caches(unet, 2) for _ in range(3)]
    for s, kv in zip((a, b, c), kvs):
        s.prepare(prompt, kv)
    frames = [(torch.randn(1, 4, 1, H, W, generator=gen).half().to(DEV), torch.randn(1, 4, 1, H, W, generator=gen).half().to(DEV))
              for _ in range(9)]
    for f in range(5):
        oa, ob, oc = (s(*frames[f]).clone() for s in (a, b, c))
        close(ob, oa, f"same seed frame {f}")                        # (GroupNorm's shared-memory atomics reorder sums run to run)
        if f >= 1:                                                    # frame 0 has no re-noised row yet
            assert float((oc.float() - oa.float()).abs().max()) > 5e-2, "a different seed must change the stream"
    blob = a.save_state()
    assert len(blob) > 2 * (N - 1) * 4 * H * W * 2
    # migrate stream a onto a fresh stream object d (different construction seed: the blob carries the seed)
    _, d = make(unet=unet, stream_seed=123)
    kv_d = [t.clone() for t in kvs[0]]
    d.prepare(prompt, kv_d)
    d.load_state(blob)
    assert d.schedule() == a.schedule()
    for f in range(5, 9):
        oa, od = a(*frames[f]).clone(), d(*frames[f]).clone()
        close(od, oa, f"migrated stream frame {f}")
    assert d.schedule() == a.schedule() and a.schedule()["frame"] == 9
    # the re-noised buffer row is a[1] * x0 + b[1] * noise with noise ~ N(0,1): not degenerate
    buf = torch.frombuffer(bytearray(a.save_state()[-2 * (N - 1) * 4 * H * W * 2:-(N - 1) * 4 * H * W * 2]), dtype=torch.float16)
    assert torch.isfinite(buf).all() and 0.3 < float(buf.float().std()) < 3.0
    with pytest.raises(RuntimeError):
        d.load_state(b"\0" * len(blob))


def test_device_stream_without_noise():
    unet, ds = make(do_add_noise=False)
    from live2diff_b200.stream_pipeline import B200StreamPipeline

    pipe = B200StreamPipeline(unet, [30, 40], do_add_noise=False)
    gen = torch.Generator().manual_seed(4)
    prompt = torch.randn(1, 77, 96, generator=gen)
    pipe.prepare(prompt, warm_caches(unet, 3))
    ds.prepare(prompt, warm_caches(unet, 3))
    for f in range(4):
        x = torch.randn(1, 4, 1, H, W, generator=gen).half().to(DEV)
        dep = torch.randn(1, 4, 1, H, W, generator=gen).half().to(DEV)
        close(ds(x, dep).clone(), pipe(x, dep).clone(), f"no-noise frame {f}")
