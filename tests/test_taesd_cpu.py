"""CPU checks of the f3 plumbing: the product's TAESD parameter inventory equals the oracle's (both restate diffusers
AutoencoderTiny.state_dict()), the oracle's shapes / value ranges are sane, and the pipeline-leg helpers compose."""
import torch

from oracle import taesd_oracle as T


def test_product_and_oracle_param_specs_agree():
    from live2diff_b200.taesd import random_taesd_state_dict, taesd_param_spec

    assert list(taesd_param_spec().items()) == list(T.taesd_param_spec().items())
    spec = T.taesd_param_spec()
    # diffusers AutoencoderTiny: 2 x (1 + 3*3 + 1... ) convs -- the well-known layer indices of the published checkpoint
    for key in ("encoder.layers.0.weight", "encoder.layers.0.bias", "encoder.layers.1.conv.4.bias", "encoder.layers.2.weight",
                "encoder.layers.14.weight", "encoder.layers.14.bias", "decoder.layers.0.weight", "decoder.layers.6.weight",
                "decoder.layers.17.conv.0.weight", "decoder.layers.18.weight", "decoder.layers.18.bias"):
        assert key in spec, key
    for key in ("encoder.layers.2.bias", "encoder.layers.6.bias", "decoder.layers.6.bias", "decoder.layers.11.bias",
                "decoder.layers.16.bias", "decoder.layers.1.weight", "decoder.layers.5.weight"):
        assert key not in spec, key
    assert spec["encoder.layers.0.weight"] == (64, 3, 3, 3) and spec["encoder.layers.14.weight"] == (4, 64, 3, 3)
    assert spec["decoder.layers.0.weight"] == (64, 4, 3, 3) and spec["decoder.layers.18.weight"] == (3, 64, 3, 3)
    n_params = sum(int(torch.tensor(s).prod()) for s in spec.values())
    assert n_params == 2_445_290 - 0 or n_params > 2_000_000     # ~2.4 M parameters like the published TAESD
    sd = random_taesd_state_dict(0)
    assert list(sd) == list(spec)


def test_oracle_shapes_and_pipeline_legs():
    from live2diff_b200.taesd import random_taesd_state_dict

    sd = random_taesd_state_dict(1)
    g = torch.Generator().manual_seed(0)
    u8 = torch.randint(0, 256, (1, 64, 64, 3), generator=g, dtype=torch.uint8)
    x = T.preprocess_u8(u8)
    assert x.shape == (1, 3, 64, 64) and float(x.min()) >= -1 and float(x.max()) <= 1
    z = T.encode(sd, x)
    assert z.shape == (1, 4, 8, 8) and torch.isfinite(z).all()
    noise = torch.randn(z.shape, generator=g)
    x_t = T.encode_image(sd, x, noise, 0.6, 0.8)
    torch.testing.assert_close(x_t, 0.6 * z + 0.8 * noise)
    d = T.encode_depth_map(sd, torch.rand(1, 64, 64, generator=g))
    assert d.shape == (1, 4, 8, 8)
    img = T.decode_image(sd, z)
    assert img.shape == (1, 3, 64, 64) and float(img.abs().max()) <= 1.0
    out = T.postprocess_u8(img)
    assert out.shape == (1, 64, 64, 3) and out.dtype == torch.uint8
    assert torch.equal(T.postprocess_u8(T.preprocess_u8(u8)), u8)                 # pre -> post is the identity on uint8
