"""CPU wiring test of the native VAE decoder (scope row f1): layouts, the folded post_quant_conv / scaling / conv_in
GEMM (incl. the constant-one channel that makes the folded bias exact at the border), the GEMM-softmax-GEMM mid-block
attention, up-block order and the padded conv_out, against the fp32 oracle with stand-in ops (no kernels involved)."""
import numpy as np
import torch

from test_wiring_cpu import patched  # noqa: F401  (fixture)


def _pair(cfg, seed=0):
    from imagharmony_b200.vae import AutoencoderKLDecoder
    from imagharmony_b200.weights import random_state_dict, shapes_of
    from oracle.vae_ref import VAEDecoderRef
    with torch.device("meta"):
        shapes = shapes_of(VAEDecoderRef(cfg))
    sd = {k: v.float() for k, v in random_state_dict(shapes, seed).items()}
    ref = VAEDecoderRef(cfg)
    ref.load_state_dict(sd)
    with torch.device("meta"):
        native = AutoencoderKLDecoder(cfg)
    native.load_state_dict(sd, assign=True)
    native.finalize()
    return native, ref.eval()


def test_vae_decoder_wiring_matches_oracle(patched):  # noqa: F811
    from imagharmony_b200.config import TINY_VAE
    native, ref = _pair(TINY_VAE, seed=3)
    assert list(native.state_dict().keys()) == list(ref.state_dict().keys())      # diffusers key contract
    z = torch.randn(2, 4, 4, 4, generator=torch.Generator().manual_seed(1)) * TINY_VAE.scaling_factor * 3
    with torch.no_grad():
        r = ref.decode(z)
        o = native.decode(z)
    assert o.shape == r.shape == (2, 3, 32, 32)
    assert torch.allclose(o, r, rtol=1e-4, atol=1e-4), (o - r).abs().max()


def test_vae_postprocess_matches_reference_semantics():
    from imagharmony_b200.vae import postprocess
    from oracle.vae_ref import postprocess_ref
    img = torch.randn(2, 3, 16, 16) * 1.5
    np.testing.assert_allclose(postprocess(img, "np"), postprocess_ref(img), rtol=0, atol=0)
    pil = postprocess(img, "pil")
    assert len(pil) == 2 and pil[0].size == (16, 16) and pil[0].mode == "RGB"
    u8 = np.asarray(pil[1])
    assert np.array_equal(u8, (postprocess_ref(img)[1] * 255).round().astype("uint8"))


def test_vae_tiled_decode_matches_oracle(patched):  # noqa: F811
    """pipe.enable_vae_tiling() (test.py:73): a latent larger than one tile is decoded tile-wise and cross-faded like
    [3P] AutoencoderKL.tiled_decode; a latent that fits one tile takes the plain path."""
    from imagharmony_b200.config import TINY_VAE            # sample_size 64 -> 8 x 8 latent tiles, stride 6, blend 16 px
    from oracle.vae_ref import tiled_decode_ref
    native, ref = _pair(TINY_VAE, seed=4)
    native.use_tiling = True
    z = torch.randn(1, 4, 14, 11, generator=torch.Generator().manual_seed(2)) * TINY_VAE.scaling_factor * 3
    with torch.no_grad():
        r = tiled_decode_ref(ref, z)
        o = native.decode(z)
    assert o.shape == r.shape
    assert torch.allclose(o, r, rtol=1e-4, atol=1e-4), (o - r).abs().max()
    small = z[:, :, :8, :8].contiguous()
    with torch.no_grad():
        assert torch.allclose(native.decode(small), ref.decode(small), rtol=1e-4, atol=1e-4)


def test_vae_scaled_residual_stream_is_exact_in_fp32(patched):  # noqa: F811
    """The force_upcast replacement (module docstring of vae.py): storing the residual stream as 2^-k * x with scale-
    invariant GroupNorms and alpha-scaled writers is the same function.  fp32 stand-in ops: identical to the oracle
    for k = 0, 7, 10 -- including a model whose stream reaches 1e5 (it would overflow fp16 without the scaling)."""
    from imagharmony_b200.config import TINY_VAE
    native, ref = _pair(TINY_VAE, seed=5)
    assert native.stream_scale == 2.0 ** -7          # TINY_VAE / SDXL_VAE ask for the upcast (force_upcast)
    with torch.no_grad():
        for m in (native, ref):                      # blow the residual stream up at its source
            m.decoder.conv_in.weight.mul_(4000.0)
            m.decoder.conv_in.bias.mul_(4000.0)
    z = torch.randn(1, 4, 6, 5, generator=torch.Generator().manual_seed(3)) * TINY_VAE.scaling_factor * 3
    with torch.no_grad():
        r = ref.decode(z)
    outs = []
    for k in (0, 7, 10):
        native.stream_scale = 2.0 ** -k
        native.finalize()
        with torch.no_grad():
            outs.append(native.decode(z))
        assert torch.allclose(outs[-1], r, rtol=2e-4, atol=2e-4), (k, (outs[-1] - r).abs().max())


def test_vae_blocked_attention_matches_oracle(patched, monkeypatch):  # noqa: F811
    """The mid-block attention in several query blocks (forced by a tiny score budget) and with padded keys (N % 8 != 0)."""
    from imagharmony_b200 import vae as V
    from imagharmony_b200.config import TINY_VAE
    native, ref = _pair(TINY_VAE, seed=6)
    monkeypatch.setattr(V.VAEAttention, "SCORE_BLOCK_BYTES", 128 * 2 * 304)    # 128 query rows per block
    for hw in ((19, 16), (15, 11)):                                            # N = 304 (3 blocks), N = 165 (padded to 168)
        z = torch.randn(1, 4, *hw, generator=torch.Generator().manual_seed(4)) * TINY_VAE.scaling_factor * 3
        with torch.no_grad():
            r, o = ref.decode(z), native.decode(z)
        assert torch.allclose(o, r, rtol=1e-4, atol=1e-4), (hw, (o - r).abs().max())
