"""Teacher-forced per-layer parity of the generator's forward AND backward kernels at the real layer shapes
(256 x 256 input, B = 1), through the C ABI on the GPU.  The checks live in tests/teacher_forced.py (read its
docstring): each kernel is fed the oracle's activations / gradients and must reproduce the oracle's fp32 result
to the north star's per-layer bf16 tolerance, rel-L2 <= 1e-2 (measured: 1.7e-3, the rounding of the stored bf16 output)."""
import pytest

import teacher_forced as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import irc_b200  # noqa: F401
    from irc_b200._native import CudaBackend
    T.Cfg.TOL, T.Cfg.B, T.Cfg.H, T.Cfg.W, T.Cfg.dev, T.Cfg.round_bf16 = 1e-2, 1, 256, 256, "cuda", True
    return T.make_ctx(CudaBackend())


def test_inc_7x7_reflect(ctx):
    T.check_inc_7x7_reflect(ctx)


@pytest.mark.parametrize("name", ["down1", "down2", "up1", "up2"])
def test_3x3_zero_pad_layers(ctx, name):
    T.check_3x3_zero_pad_layers(ctx, name)


@pytest.mark.parametrize("b", [0, 8])
def test_resnet_block(ctx, b):
    T.check_resnet_block(ctx, b)


def test_upsample_into_concat_and_transpose(ctx):
    T.check_upsample_into_concat_and_transpose(ctx)


def test_outc_7x7_tanh(ctx):
    T.check_outc_7x7_tanh(ctx)


@pytest.fixture(scope="module")
def dv_ctx(ctx):
    return T.make_dv_ctx(ctx.be)


def test_discriminator_layers(dv_ctx):
    T.check_discriminator_layers(dv_ctx)


def test_vgg_layers(dv_ctx):
    T.check_vgg_layers(dv_ctx)
