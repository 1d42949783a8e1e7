"""The teacher-forced per-layer checks of tests/teacher_forced.py driven on the CPU with the torch restatement of the
primitives (tests/ref_backend.py) in float32 frames at 32 x 32: proves that the checks themselves (which buffer holds
what, which oracle expression each launch must equal) are right, to 1e-4, before the GPU run applies them to the CUDA
kernels at 256 x 256 with the bf16 tolerance."""
import pytest
import torch

import teacher_forced as T


@pytest.fixture(scope="module")
def ctx():
    import irc_b200  # noqa: F401
    from irc_b200 import layout as L
    from ref_backend import RefBackend
    old = L.ACT_DTYPE
    L.ACT_DTYPE = torch.float32
    saved = (T.Cfg.TOL, T.Cfg.B, T.Cfg.H, T.Cfg.W, T.Cfg.dev, T.Cfg.round_bf16)
    T.Cfg.TOL, T.Cfg.B, T.Cfg.H, T.Cfg.W, T.Cfg.dev, T.Cfg.round_bf16 = 1e-4, 1, 32, 32, "cpu", False
    yield T.make_ctx(RefBackend())
    L.ACT_DTYPE = old
    T.Cfg.TOL, T.Cfg.B, T.Cfg.H, T.Cfg.W, T.Cfg.dev, T.Cfg.round_bf16 = saved


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
    from ref_backend import RefBackend
    return T.make_dv_ctx(RefBackend())


def test_discriminator_layers(dv_ctx):
    T.check_discriminator_layers(dv_ctx)


def test_vgg_layers(dv_ctx):
    T.check_vgg_layers(dv_ctx)
