"""irc_b200: B200-native implementation of the train / test hot path of
yavuzmurattas/Infrared-Colorization-with-ResNet-Generator-and-PatchGAN (Code/ir_colorization.py).

The names below mirror the reference module so that `import irc_b200 as ir_colorization`
is a drop-in for that path; everything executes in libirc_sm100.so (sm_100a)."""
from .modules import (Config, Downsample, Identity, IRColorizationModel, NLayerDiscriminator, ResnetBlock, ResnetUNetGenerator,  # noqa: F401
                      UpsampleAA, VGGPerceptual, get_filter, get_lr_lambda, get_norm_layer, init_net, init_weights,
                      ssim_loss_torch, tv_loss)
from .train import (batch_metrics, compute_metrics, main, run_test, tensor_to_rgb_image, train_kaist, validate_kaist)  # noqa: F401
from .train_step import TrainStep  # noqa: F401
