"""Input pipeline: the oracle's restatement of OpenCV's INTER_AREA + the reference loaders against the
fixtures produced by cv2 and the unmodified reference (oracle/make_golden_input.py), the product's host-side table builder
against the oracle's, and - on the GPU - the batched kernels against both, BIT-EXACT (bytes and float32 tensors)."""
import os

import numpy as np
import pytest
import torch

import input_pipeline as IP

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "input_pipeline.npz"))
CASES = ["general_kaist_ratio", "general_odd", "int_3x4", "two_by_two", "x_int_y_frac"]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("cn", [1, 3])
def test_oracle_resize_equals_opencv_fixture(name, cn):
    src, dst = GOLD[f"resize/{name}/c{cn}/src"], GOLD[f"resize/{name}/c{cn}/dst"]
    assert np.array_equal(IP.resize_area_u8(src, dst.shape[0], dst.shape[1]), dst)


def test_oracle_resize_full_kaist_geometry():
    src = np.random.default_rng(int(GOLD["resize/kaist_full/seed"][0])).integers(0, 256, (512, 640, 3), dtype=np.uint8)
    out = IP.resize_area_u8(src, 256, 256)
    assert int(out.astype(np.int64).sum()) == int(GOLD["resize/kaist_full/sum"][0])
    assert np.array_equal(out.reshape(-1)[::997], GOLD["resize/kaist_full/sample"])


@pytest.mark.parametrize("i", range(4))
@pytest.mark.parametrize("flip", [0, 1])
def test_oracle_pair_equals_reference_loader(i, flip):
    """irc:1132-1177 on decoded frames, incl. the frame whose maximum is <= 1 (no /255, irc:1142) and the paired flip"""
    ir, bgr = GOLD[f"pair/{i}/ir_u8"], GOLD[f"pair/{i}/bgr_u8"]
    assert np.array_equal(IP.ir_from_u8(ir, 32, bool(flip)), GOLD[f"pair/{i}/flip{flip}/ir"])
    assert np.array_equal(IP.rgb_from_bgr_u8(bgr, 32, bool(flip)), GOLD[f"pair/{i}/flip{flip}/rgb"])


@pytest.mark.parametrize("ssize,dsize", [(640, 256), (512, 256), (131, 57), (97, 31), (128, 32), (60, 32)])
def test_product_area_tables_equal_the_oracles(ssize, dsize):
    import irc_b200  # noqa: F401
    from irc_b200.data import area_table
    idx, w = area_table(ssize, dsize)
    oi, ow = IP.padded_tab(ssize, dsize)
    assert np.array_equal(idx, oi) and np.array_equal(w, ow) and w.dtype == np.float32
    assert np.allclose(w.sum(1), 1.0, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("cn", [1, 3])
def test_gpu_resize_is_bit_exact_with_opencv(name, cn):
    import irc_b200  # noqa: F401
    from irc_b200 import modules as M
    from irc_b200.data import GpuPairPreprocessor
    src, dst = GOLD[f"resize/{name}/c{cn}/src"], GOLD[f"resize/{name}/c{cn}/dst"]
    sh, sw = src.shape[:2]
    dh, dw = dst.shape[:2]
    be = M.backend()
    # a batch of three frames: the fixture, its vertical mirror, zeros
    batch = np.stack([src, src[::-1].copy(), np.zeros_like(src)]).reshape(3, sh, sw, cn)
    x = torch.from_numpy(batch).cuda()
    out = torch.zeros(3, dh, dw, cn, device="cuda", dtype=torch.uint8)
    vmax = torch.zeros(3, device="cuda", dtype=torch.int32)
    mode = {"general": 0, "int": 1, "2x2": 2}[IP.resize_mode(sh, sw, dh, dw)]
    tables = None
    if mode == 0:
        from irc_b200.data import area_table
        xi, xw = area_table(sw, dw); yi, yw = area_table(sh, dh)
        tables = tuple(torch.from_numpy(a).cuda() for a in (xi, xw, yi, yw))
    be.resize_area_u8(x, out, tables, mode, img_max=vmax)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    assert np.array_equal(got[0].reshape(dst.shape), dst)
    assert np.array_equal(got[1].reshape(dst.shape), IP.resize_area_u8(src[::-1].copy(), dh, dw))
    assert got[2].max() == 0
    assert vmax.tolist() == [int(got[0].max()), int(got[1].max()), 0]


@pytest.mark.gpu
def test_gpu_pair_preprocessor_equals_the_reference_loader_bit_for_bit():
    """the whole per-sample path of irc:1132-1177 for a batch: resize, BGR->RGB, /255 (or not), clip, paired flip, [-1, 1]"""
    import irc_b200  # noqa: F401
    from irc_b200.data import GpuPairPreprocessor
    for ids in ([0, 1], [3]):            # frames 0, 1 are 64 x 80 (general path, one of them with maximum <= 1); frame 3 is 64 x 64 (2 x 2)
        ir = torch.from_numpy(np.stack([GOLD[f"pair/{i}/ir_u8"] for i in ids]))
        bgr = torch.from_numpy(np.stack([GOLD[f"pair/{i}/bgr_u8"] for i in ids]))
        pre = GpuPairPreprocessor(32, tuple(ir.shape[1:3]), "cuda")
        for flips in ([False] * len(ids), [True] * len(ids), [True, False][:len(ids)]):
            out = pre(ir, bgr, flip=torch.tensor(flips))
            torch.cuda.synchronize()
            for j, i in enumerate(ids):
                f = int(flips[j])
                assert np.array_equal(out["ir"][j].cpu().numpy(), GOLD[f"pair/{i}/flip{f}/ir"]), (i, f)
                assert np.array_equal(out["rgb"][j].cpu().numpy(), GOLD[f"pair/{i}/flip{f}/rgb"]), (i, f)
    # 96 x 128 -> 32 x 32: integer factors 3 x 4
    pre = GpuPairPreprocessor(32, (96, 128), "cuda")
    out = pre(torch.from_numpy(GOLD["pair/2/ir_u8"][None]), torch.from_numpy(GOLD["pair/2/bgr_u8"][None]), flip=torch.tensor([False]))
    assert np.array_equal(out["ir"][0].cpu().numpy(), GOLD["pair/2/flip0/ir"]) and np.array_equal(out["rgb"][0].cpu().numpy(), GOLD["pair/2/flip0/rgb"])


@pytest.mark.gpu
def test_gpu_preprocessor_full_kaist_geometry_and_training_step():
    """640 x 512 frames -> 256 x 256 batch, bit-exact with the oracle, fed straight into a training iteration"""
    import irc_oracle as O
    import irc_b200  # noqa: F401
    from irc_b200.data import GpuPairPreprocessor
    rng = np.random.default_rng(5)
    B = 2
    ir = rng.integers(0, 256, (B, 512, 640), dtype=np.uint8); bgr = rng.integers(0, 256, (B, 512, 640, 3), dtype=np.uint8)
    pre = GpuPairPreprocessor(256, (512, 640), "cuda")
    out = pre(torch.from_numpy(ir), torch.from_numpy(bgr), flip=torch.tensor([True, False]))
    for j, f in enumerate((True, False)):
        assert np.array_equal(out["ir"][j].cpu().numpy(), IP.ir_from_u8(ir[j], 256, f))
        assert np.array_equal(out["rgb"][j].cpu().numpy(), IP.rgb_from_bgr_u8(bgr[j], 256, f))
    from irc_b200._native import CudaBackend
    from irc_b200.train_step import TrainStep
    ts = TrainStep(CudaBackend(), B, 256, 256, "cuda")
    ts.load(O.seeded_params(O.generator_shapes(), 1), O.seeded_params(O.discriminator_shapes(), 2), O.seeded_params(O.vgg_shapes(), 3, kaiming=True))
    ts.step(out["ir"], out["rgb"])
    assert all(np.isfinite(v) for v in ts.losses().values())
