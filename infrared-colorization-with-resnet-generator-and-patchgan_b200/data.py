"""GPU input pipeline: what KAISTPairDataset.__getitem__ (irc:1132-1177) and load_ir_image / load_rgb_image / ir_to_tensor
(irc:803-863) do to a DECODED frame - cv2.INTER_AREA resize to img_size x img_size, BGR->RGB, /255, clip, paired horizontal
flip, [-1, 1] - for a whole batch in two kernel launches per modality (libirc_sm100.so: irc_resize_area_u8, irc_u8_to_pm1),
bit-exact with OpenCV + the reference (tests/test_input_gpu.py).  Decoding JPEG/PNG files and scanning the KAIST directory
tree (irc:887-942, :1045-1120) stay on the host and are outside the accelerated path: this module starts from
uint8 tensors, e.g. the output of any decoder running in DataLoader workers, which then only have to ship raw bytes
(0.3 MB per IR frame, 1 MB per RGB frame) instead of resizing on the CPU."""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import modules as M


def area_table(ssize: int, dsize: int) -> Tuple[np.ndarray, np.ndarray]:
    """OpenCV's computeResizeAreaTab for one axis as dense [dsize, K] (source index, float32 weight) arrays; unused slots
    carry weight 0.  Pure index arithmetic on the host (double precision, like OpenCV), evaluated once per geometry."""
    scale = ssize / dsize
    rows = []
    for d in range(dsize):
        f1 = d * scale
        f2 = f1 + scale
        cell = min(scale, ssize - f1)
        s1, s2 = math.ceil(f1), math.floor(f2)
        s2 = min(s2, ssize - 1)
        s1 = min(s1, s2)
        ent = []
        if s1 - f1 > 1e-3:
            ent.append((s1 - 1, np.float32((s1 - f1) / cell)))
        ent += [(s, np.float32(1.0 / cell)) for s in range(s1, s2)]
        if f2 - s2 > 1e-3:
            ent.append((s2, np.float32(min(min(f2 - s2, 1.0), cell) / cell)))
        rows.append(ent)
    K = max(len(e) for e in rows)
    idx = np.zeros((dsize, K), np.int32); w = np.zeros((dsize, K), np.float32)
    for d, ent in enumerate(rows):
        for k, (s, a) in enumerate(ent):
            idx[d, k] = s; w[d, k] = a
    return idx, w


class GpuPairPreprocessor:
    """Batched, device-side equivalent of the reference's per-sample loader for frames of one source geometry.

        pre = GpuPairPreprocessor(img_size=256, src_hw=(512, 640), device="cuda")
        batch = pre(ir_u8, bgr_u8, flip)      # ir_u8 [B,Hs,Ws] / bgr_u8 [B,Hs,Ws,3] uint8 (host or device), flip [B] bool or None
        ts.step(batch["ir"], batch["rgb"])    # {'ir': Bx1xSxS, 'rgb': Bx3xSxS} float32 in [-1, 1], the DataLoader's format

    `flip=None` draws the paired horizontal flip with probability 0.5 per sample from `generator` (irc:1166)."""

    def __init__(self, img_size: int, src_hw: Tuple[int, int], device="cuda", rgb_is_bgr: bool = True, generator: Optional[torch.Generator] = None):
        self.S, (self.Hs, self.Ws), self.dev, self.bgr = img_size, src_hw, torch.device(device), rgb_is_bgr
        self.gen = generator
        Hs, Ws, S = self.Hs, self.Ws, img_size
        if S > Hs or S > Ws:
            raise NotImplementedError("INTER_AREA enlargement (OpenCV switches to a bilinear variant) is not part of the KAIST path")
        if Hs % S == 0 and Ws % S == 0:
            self.mode, self.tables = (2 if (Hs == 2 * S and Ws == 2 * S) else 1), None
        else:
            self.mode = 0
            xi, xw = area_table(Ws, S); yi, yw = area_table(Hs, S)
            f = lambda a: torch.from_numpy(a).to(self.dev).contiguous()
            self.tables = (f(xi), f(xw), f(yi), f(yw))
        self._buf: Dict[Tuple[int, int], Tuple[torch.Tensor, ...]] = {}

    def _buffers(self, B: int, C: int):
        key = (B, C)
        if key not in self._buf:
            self._buf[key] = (torch.empty(B, self.S, self.S, C, device=self.dev, dtype=torch.uint8), torch.zeros(B, device=self.dev, dtype=torch.int32))
        return self._buf[key]

    def _one(self, frames: torch.Tensor, C: int, flip_dev: Optional[torch.Tensor], is_ir: bool) -> torch.Tensor:
        be = M.backend()
        B = frames.shape[0]
        x = frames.to(self.dev, non_blocking=True).reshape(B, self.Hs, self.Ws, C).contiguous()
        small, vmax = self._buffers(B, C)
        be.resize_area_u8(x, small, self.tables, self.mode, img_max=vmax if is_ir else None)
        out = torch.empty(B, C, self.S, self.S, device=self.dev)
        be.u8_to_pm1(small, out, swap_rb=(self.bgr and C == 3), flip=flip_dev, img_max=vmax if is_ir else None)
        return out

    def __call__(self, ir_u8: torch.Tensor, rgb_u8: Optional[torch.Tensor] = None, flip=None) -> Dict[str, torch.Tensor]:
        B = ir_u8.shape[0]
        if flip is None:
            flip = torch.rand(B, generator=self.gen) < 0.5
        flip_dev = torch.as_tensor(flip).to(torch.uint8).to(self.dev).contiguous() if flip is not False else None
        out = {"ir": self._one(ir_u8, 1, flip_dev, True)}
        if rgb_u8 is not None:
            out["rgb"] = self._one(rgb_u8, 3, flip_dev, False)
        return out


class DevicePrefetcher:
    """Iterates a loader of dict batches one batch ahead of its consumer: the tensors of batch i + 1 are copied to the device on a
    side stream (pinned first when they are pageable) while batch i is being processed, and the consumer's stream only waits for
    the copy event.  Two persistent sets of device buffers alternate (no allocator traffic per batch); a yielded batch stays valid
    until the consumer asks for the batch after the next one.  On a CPU device it is a plain pass-through.  Non-tensor entries
    (names) are forwarded untouched.

        for batch in DevicePrefetcher(loader, device):      # batch['ir'] etc. already live on `device`
            fake = model(batch['ir'])"""

    def __init__(self, loader, device, pin: bool = True):
        self.loader, self.device, self.pin = loader, torch.device(device), pin
        self.stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None

    def __iter__(self):
        if self.stream is None:
            for batch in self.loader:
                yield batch
            return
        bufs = [dict(), dict()]          # slot -> key -> device tensor (allocated on the consumer's stream)
        free_ev = [None, None]           # consumer's work on the slot's previous batch has been queued up to here

        def stage(batch, k):
            out = {}
            for key, v in batch.items():
                if torch.is_tensor(v) and v.device.type == "cpu":
                    t = bufs[k].get(key)
                    if t is None or t.shape != v.shape or t.dtype != v.dtype:
                        t = bufs[k][key] = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                    out[key] = (t, v.pin_memory() if (self.pin and not v.is_pinned()) else v)
                else:
                    out[key] = v
            if free_ev[k] is not None:
                self.stream.wait_event(free_ev[k])
            with torch.cuda.stream(self.stream):
                for key, tv in out.items():
                    if isinstance(tv, tuple):
                        tv[0].copy_(tv[1], non_blocking=True)
                        out[key] = tv[0]
                ev = torch.cuda.Event()
                ev.record(self.stream)
            return out, ev

        it = iter(self.loader)
        k = 0
        nxt = None
        for batch in it:
            nxt = stage(batch, k)
            break
        while nxt is not None:
            cur, ev = nxt
            nxt = None
            for batch in it:                      # stage the next batch (other slot) before handing out the current one
                nxt = stage(batch, k ^ 1)
                break
            main = torch.cuda.current_stream(self.device)
            main.wait_event(ev)
            yield cur
            free_ev[k] = torch.cuda.Event()
            free_ev[k].record(torch.cuda.current_stream(self.device))
            k ^= 1
