"""Pin oracle/input_pipeline.py and write tests/golden/input_pipeline.npz.  Build container only (needs cv2 and
/root/reference): synthetic 8-bit frames are written as PNG files (lossless), read back and processed by the UNMODIFIED
reference's own functions (load_ir_image / load_rgb_image / ir_to_tensor, KAISTPairDataset._read_ir / _read_rgb /
__getitem__ with the flip forced both ways), and by cv2.resize directly; the numpy restatement must match all of them bit
for bit.  The fixture keeps the raw frames, OpenCV's resized bytes and the reference's final tensors."""
import os
import sys
import tempfile

import cv2
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/Code")
import input_pipeline as IP  # noqa: E402
import ir_colorization as R  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    rng = np.random.default_rng(2024)
    gold = {}
    # ---- cv2.resize itself, all three branches, grey and colour
    cases = [("general_kaist_ratio", 64, 80, 32, 32), ("general_odd", 97, 131, 31, 57), ("int_3x4", 96, 128, 32, 32), ("two_by_two", 66, 38, 33, 19),
             ("x_int_y_frac", 60, 128, 32, 32)]
    for name, sh, sw, dh, dw in cases:
        for cn in (1, 3):
            src = rng.integers(0, 256, (sh, sw) if cn == 1 else (sh, sw, 3), dtype=np.uint8)
            ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA)
            assert np.array_equal(ref, IP.resize_area_u8(src, dh, dw)), (name, cn)
            gold[f"resize/{name}/c{cn}/src"] = src; gold[f"resize/{name}/c{cn}/dst"] = ref
    # full KAIST geometry once (640 x 512 -> 256 x 256), checked here, only a checksum kept
    src = rng.integers(0, 256, (512, 640, 3), dtype=np.uint8)
    ref = cv2.resize(src, (256, 256), interpolation=cv2.INTER_AREA)
    assert np.array_equal(ref, IP.resize_area_u8(src, 256, 256))
    gold["resize/kaist_full/seed"] = np.array([77]); src = np.random.default_rng(77).integers(0, 256, (512, 640, 3), dtype=np.uint8)
    ref = cv2.resize(src, (256, 256), interpolation=cv2.INTER_AREA)
    assert np.array_equal(ref, IP.resize_area_u8(src, 256, 256))
    gold["resize/kaist_full/sum"] = np.array([int(ref.astype(np.int64).sum())]); gold["resize/kaist_full/sample"] = ref.reshape(-1)[::997].copy()

    # ---- the reference's loaders on real files
    size = 32
    with tempfile.TemporaryDirectory() as td:
        pairs = []
        for i, (sh, sw) in enumerate([(64, 80), (64, 80), (96, 128), (64, 64)]):
            ir = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
            if i == 1:
                ir = (ir > 250).astype(np.uint8)          # a frame whose resized maximum is <= 1: the reference skips the /255 (irc:1142)
            bgr = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
            pi, pr = os.path.join(td, f"ir{i}.png"), os.path.join(td, f"rgb{i}.png")
            cv2.imwrite(pi, ir); cv2.imwrite(pr, bgr)
            assert np.array_equal(cv2.imread(pi, cv2.IMREAD_GRAYSCALE), ir) and np.array_equal(cv2.imread(pr, cv2.IMREAD_COLOR), bgr)
            pairs.append((pi, pr, ir, bgr))
        ds = R.KAISTPairDataset.__new__(R.KAISTPairDataset)      # the constructor scans a dataset tree; only the readers are exercised
        ds.img_size, ds.augment = size, True
        ds.ir_paths = [p[0] for p in pairs]; ds.rgb_paths = [p[1] for p in pairs]
        for i, (pi, pr, ir, bgr) in enumerate(pairs):
            a = R.load_ir_image(pi, img_size=size); b = R.load_rgb_image(pr, img_size=size)
            assert np.array_equal(a, ds._read_ir(pi)) and np.array_equal(b, ds._read_rgb(pr))
            t = R.ir_to_tensor(a)[0].numpy()
            assert np.array_equal(t, IP.ir_from_u8(ir, size)), i
            for flip in (False, True):
                R.random.random = (lambda f=flip: 0.0 if f else 1.0)      # force the coin of irc:1166
                item = ds[i]
                want_ir, want_rgb = item["ir"].numpy(), item["rgb"].numpy()
                assert np.array_equal(want_ir, IP.ir_from_u8(ir, size, flip)), (i, flip)
                assert np.array_equal(want_rgb, IP.rgb_from_bgr_u8(bgr, size, flip)), (i, flip)
                gold[f"pair/{i}/flip{int(flip)}/ir"] = want_ir; gold[f"pair/{i}/flip{int(flip)}/rgb"] = want_rgb
            gold[f"pair/{i}/ir_u8"] = ir; gold[f"pair/{i}/bgr_u8"] = bgr
    np.savez_compressed(os.path.join(OUT, "input_pipeline.npz"), **gold)
    print("wrote input_pipeline.npz keys:", len(gold))


if __name__ == "__main__":
    main()
