"""Pin the oracle's DAC-VAE encoder / decoder restatement (oracle/restate.py codec_encode / codec_decode) against an
independent implementation of the same Descript-DAC layer layout: ``transformers.models.dac`` (DacEncoder / DacDecoder,
installed in this image).  The reference's own codec, ``dacvae`` (pyproject.toml:19, un-pinned git dependency), is
absent from /root/reference, so this is the strongest pin available here: same layer list (Snake1d, dilated residual
units 1/3/9, strided convs with k = 2s and pad = ceil(s/2), transposed convs, tanh), same weights, bit-for-bit the same
torch ops up to fp32 rounding.  What stays un-pinned is only dacvae's bottleneck convention (in_proj -> first half =
mean, reference codec.py:68) and whatever else dacvae adds on top of the Descript layout.
"""
import math

import pytest
import torch

from _util import rel_l2

dac = pytest.importorskip("transformers.models.dac.modeling_dac")


def _tiny_codec_cfg():
    from sam_audio_b200.config import DACVAEConfig
    # same rates (hop 1920) and layer counts as the real codec, narrow channels so the CPU test stays sub-second
    return DACVAEConfig(encoder_dim=8, encoder_rates=[2, 8, 10, 12], latent_dim=32, decoder_dim=64,
                        decoder_rates=[12, 10, 8, 2], codebook_dim=16)


def _codec_sd(ccfg, seed=3):
    from sam_audio_b200.synthetic import codec_param_shapes
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for name, shape, kind in codec_param_shapes(ccfg):
        if kind == "alpha":
            sd[f"audio_codec.{name}"] = (1.0 + 0.3 * torch.randn(shape, generator=g)).abs() + 0.05
        elif kind == "bias":
            sd[f"audio_codec.{name}"] = 0.05 * torch.randn(shape, generator=g)
        else:
            fan = shape[1] * shape[2] if kind == "conv" else shape[0] * 2
            sd[f"audio_codec.{name}"] = torch.randn(shape, generator=g) / math.sqrt(fan)
    return sd


def _hf_modules(ccfg, sd):
    from transformers import DacConfig
    hc = DacConfig(encoder_hidden_size=ccfg.encoder_dim, downsampling_ratios=list(ccfg.encoder_rates),
                   decoder_hidden_size=ccfg.decoder_dim, upsampling_ratios=list(ccfg.decoder_rates),
                   hidden_size=ccfg.latent_dim, codebook_dim=ccfg.codebook_dim, n_codebooks=1, codebook_size=8)
    enc, dec = dac.DacEncoder(hc).eval(), dac.DacDecoder(hc).eval()

    def put(mod, name):
        with torch.no_grad():
            mod.weight.copy_(sd[f"audio_codec.{name}.weight"])
            mod.bias.copy_(sd[f"audio_codec.{name}.bias"])

    def put_alpha(mod, name):
        with torch.no_grad():
            mod.alpha.copy_(sd[f"audio_codec.{name}.alpha"])

    def put_unit(unit, name):
        put_alpha(unit.snake1, f"{name}.block.0"); put(unit.conv1, f"{name}.block.1")
        put_alpha(unit.snake2, f"{name}.block.2"); put(unit.conv2, f"{name}.block.3")

    n = len(ccfg.encoder_rates)
    put(enc.conv1, "encoder.block.0")
    for i, blk in enumerate(enc.block):
        p = f"encoder.block.{i + 1}"
        for j, unit in enumerate((blk.res_unit1, blk.res_unit2, blk.res_unit3)):
            put_unit(unit, f"{p}.block.{j}")
        put_alpha(blk.snake1, f"{p}.block.3"); put(blk.conv1, f"{p}.block.4")
    put_alpha(enc.snake1, f"encoder.block.{n + 1}"); put(enc.conv2, f"encoder.block.{n + 2}")

    put(dec.conv1, "decoder.model.0")
    for i, blk in enumerate(dec.block):
        p = f"decoder.model.{i + 1}"
        put_alpha(blk.snake1, f"{p}.block.0"); put(blk.conv_t1, f"{p}.block.1")
        for j, unit in enumerate((blk.res_unit1, blk.res_unit2, blk.res_unit3)):
            put_unit(unit, f"{p}.block.{j + 2}")
    put_alpha(dec.snake1, f"decoder.model.{n + 1}"); put(dec.conv2, f"decoder.model.{n + 2}")
    return enc, dec


@torch.inference_mode()
def test_codec_restatement_matches_transformers_dac():
    from oracle import restate
    ccfg = _tiny_codec_cfg()
    # float64 on both sides: the two implementations order a few fp32 roundings differently (HF multiplies by the
    # reciprocal in Snake), which 24 residual units amplify to ~3e-5 in fp32; in double the layer lists must agree
    sd = {k: v.double() for k, v in _codec_sd(ccfg).items()}
    enc, dec = (m.double() for m in _hf_modules(ccfg, sd))
    g = torch.Generator().manual_seed(9)
    wav = 0.5 * torch.randn(2, 1, 3 * ccfg.hop_length, generator=g, dtype=torch.float64)   # a multiple of the hop
    # encoder: HF trunk + the bottleneck convention of reference codec.py:65-70 (in_proj, first half = mean)
    z_hf = torch.nn.functional.conv1d(enc(wav), sd["audio_codec.quantizer.in_proj.weight"],
                                      sd["audio_codec.quantizer.in_proj.bias"])[:, : ccfg.codebook_dim]
    z = restate.codec_encode(sd, ccfg, wav)
    assert z.shape == z_hf.shape == (2, ccfg.codebook_dim, 3)
    assert rel_l2(z, z_hf) < 1e-12
    # decoder: out_proj (reference codec.py:86-89) + HF trunk
    lat = torch.randn(2, ccfg.codebook_dim, 3, generator=g, dtype=torch.float64)
    y_hf = dec(torch.nn.functional.conv1d(lat, sd["audio_codec.quantizer.out_proj.weight"],
                                          sd["audio_codec.quantizer.out_proj.bias"]))
    y = restate.codec_decode(sd, ccfg, lat)
    assert y.shape == y_hf.shape == (2, 1, 3 * ccfg.hop_length)
    assert rel_l2(y, y_hf) < 1e-12


@torch.inference_mode()
def test_snake_matches_transformers_dac():
    from oracle import restate
    g = torch.Generator().manual_seed(2)
    x = 3.0 * torch.randn(2, 6, 50, generator=g)
    s = dac.Snake1d(6)
    with torch.no_grad():
        s.alpha.copy_((1.0 + 0.5 * torch.randn(1, 6, 1, generator=g)).abs() + 0.01)
    assert torch.allclose(restate.snake(x, s.alpha), s(x), rtol=1e-6, atol=1e-6)
