"""CPU oracle of the reference's per-sample input pipeline (irc = Code/ir_colorization.py).  TEST INFRASTRUCTURE ONLY.

What it restates: KAISTPairDataset._read_ir / _read_rgb / __getitem__ (irc:1132-1177) and load_ir_image / load_rgb_image /
ir_to_tensor (irc:803-863) AFTER the file has been decoded: `cv2.resize(..., interpolation=cv2.INTER_AREA)` of the uint8
frame to img_size x img_size, BGR->RGB, the /255 scaling (IR: only when the resized frame's maximum exceeds 1, irc:1142-1146),
the clip, the paired horizontal flip (irc:1166-1168) and the [-1, 1] mapping (irc:1174-1175).

The resize lives in a third-party dependency that is absent from /root/reference: OpenCV (`opencv-python`, version not
pinned by the reference, README:112; 4.13.0 in this image).  Its INTER_AREA algorithm for 8-bit images is restated here from
the published implementation (modules/imgproc/src/resize.cpp) in exact float32 operation order:
  * both scale factors integers, 2 x 2:  (a + b + c + d + 2) >> 2                               (ResizeAreaFastVec_SIMD_8u)
  * both integers otherwise:             cvRound(float(sum) * float(1 / area))                   (resizeAreaFast_)
  * else (KAIST 640 x 512 -> 256 x 256): separable float tables from computeResizeAreaTab, per source row
    buf[dx] = sum_k S[sx_k] * alpha_k (k ascending), per destination row sum = beta_0 * buf_0, sum += beta_j * buf_j,
    cvRound (round half to even) and saturation                                                  (ResizeArea_Invoker)
PARITY PINNING: oracle/make_golden_input.py runs cv2.resize itself and the unmodified reference's loader functions on PNG
files written to a temporary directory and asserts bit-equality with this restatement (all three branches); the fixtures it
commits (tests/golden/input_pipeline.npz) are what the GPU kernel is tested against, together with this oracle."""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np


def area_tab(ssize: int, dsize: int) -> List[Tuple[int, int, np.float32]]:
    """computeResizeAreaTab for one axis: (destination index, source index, weight) in emission order"""
    scale = ssize / dsize
    out = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            out.append((dx, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            out.append((dx, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            out.append((dx, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return out


def padded_tab(ssize: int, dsize: int):
    """the same table as dense [dsize, K] index / weight arrays (weight 0 = unused slot; adding +0.0 changes nothing)"""
    t = area_tab(ssize, dsize)
    K = max(sum(1 for e in t if e[0] == d) for d in range(dsize))
    idx = np.zeros((dsize, K), np.int32); w = np.zeros((dsize, K), np.float32)
    fill = [0] * dsize
    for d, s, a in t:
        idx[d, fill[d]] = s; w[d, fill[d]] = a; fill[d] += 1
    return idx, w


def resize_mode(sh: int, sw: int, dh: int, dw: int) -> str:
    """which OpenCV branch INTER_AREA takes for an 8-bit image (shrinking only)"""
    if dh > sh or dw > sw:
        raise NotImplementedError("INTER_AREA enlargement falls back to a bilinear variant in OpenCV: not part of the KAIST path")
    if sh % dh == 0 and sw % dw == 0:
        return "2x2" if (sh == 2 * dh and sw == 2 * dw) else "int"
    return "general"


def resize_area_u8(src: np.ndarray, dh: int, dw: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA) for uint8 HxW or HxWxC"""
    assert src.dtype == np.uint8
    sh, sw = src.shape[:2]
    cn = 1 if src.ndim == 2 else src.shape[2]
    s = src.reshape(sh, sw, cn)
    mode = resize_mode(sh, sw, dh, dw)
    if mode != "general":
        fy, fx = sh // dh, sw // dw
        acc = s.reshape(dh, fy, dw, fx, cn).astype(np.int64).sum((1, 3))
        if mode == "2x2":
            out = ((acc + 2) >> 2).astype(np.uint8)
        else:
            scale = np.float32(1.0 / (fx * fy))
            out = np.clip(np.rint((acc.astype(np.float32) * scale).astype(np.float32)), 0, 255).astype(np.uint8)
    else:
        xi, xw = padded_tab(sw, dw)
        yi, yw = padded_tab(sh, dh)
        S = s.astype(np.float32)
        # horizontal: buf[sy, dx] = sum_k S[sy, xi[dx, k]] * xw[dx, k], k ascending, every product and sum rounded to float32
        buf = np.zeros((sh, dw, cn), np.float32)
        for k in range(xi.shape[1]):
            buf = (buf + (S[:, xi[:, k], :] * xw[None, :, k, None]).astype(np.float32)).astype(np.float32)
        acc = np.zeros((dh, dw, cn), np.float32)
        for k in range(yi.shape[1]):
            acc = (acc + (yw[:, k, None, None] * buf[yi[:, k]]).astype(np.float32)).astype(np.float32)
        out = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    return out.reshape(dh, dw) if src.ndim == 2 else out


def ir_from_u8(gray_u8: np.ndarray, size: int, flip: bool = False) -> np.ndarray:
    """decoded 8-bit grey frame -> 1 x size x size float32 in [-1, 1]  (irc:1132-1147, :1166-1174)"""
    q = resize_area_u8(gray_u8, size, size)
    img = q.astype(np.float32)
    if img.max() > 1.0:          # cv2.IMREAD_GRAYSCALE always yields uint8, so the /65535 branch (irc:1146) is unreachable
        img = img / np.float32(255.0)
    img = np.clip(img, 0.0, 1.0)
    if flip:
        img = np.fliplr(img).copy()
    return (img[None] * np.float32(2.0) - np.float32(1.0)).astype(np.float32)


def rgb_from_bgr_u8(bgr_u8: np.ndarray, size: int, flip: bool = False) -> np.ndarray:
    """decoded 8-bit BGR frame (cv2.imread order) -> 3 x size x size float32 RGB in [-1, 1]  (irc:1149-1158, :1166-1175)"""
    rgb = bgr_u8[:, :, ::-1]
    q = resize_area_u8(np.ascontiguousarray(rgb), size, size)
    img = q.astype(np.float32) / np.float32(255.0)
    img = np.clip(img, 0.0, 1.0)
    if flip:
        img = np.fliplr(img).copy()
    return (np.transpose(img, (2, 0, 1)) * np.float32(2.0) - np.float32(1.0)).astype(np.float32)
