"""Host-side mirror of `voroffset::DoubleCompressedImage`'s morphology operators
(src/vor2d/DoubleCompressedImage.h:96-111, .cpp:438-468,680-719), backed by the C ABI's vo_morph2d.

Like the reference's member functions, the operators mutate the image in place; `dilate(r)` sweeps
with R = r * rows and `erode(r)` with R = r (DoubleCompressedImage.cpp:685-686,698-699) - that
quirk lives inside the library so this boundary takes exactly the reference's argument.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .volume import DexelImage


class DoubleCompressedImage(DexelImage):
    def _bind(self, ctx=None, device: int = 0):
        self._ctx = ctx or _lib.default_context(device)
        return self

    @classmethod
    def from_image(cls, img: DexelImage, ctx=None, device: int = 0) -> "DoubleCompressedImage":
        return cls(img.rows, img.width, img.off.copy(), img.spans.copy())._bind(ctx, device)

    def _apply(self, op: str, r: float):
        ctx = getattr(self, "_ctx", None) or _lib.default_context(0)
        poff, pspans = _lib._u32p(), _lib._f64p()
        n = C.c_uint64()
        ms = C.c_double(0)
        spans = self.spans if self.spans.size else np.zeros((1, 2))
        ctx.check(ctx.lib.vo_morph2d(ctx.handle, _lib.OPS2D[op], self.rows, self.width, _lib.ptr(self.off),
                                     _lib.ptr(spans), float(r), C.byref(poff), C.byref(pspans), C.byref(n), C.byref(ms)))
        self.off, self.spans = ctx.take_host(poff, pspans, self.rows, int(n.value))
        self.last_ms = ms.value
        if op != "negate" and not self.isValid():      # vor_assert(isValid()) at .cpp:688,702
            raise RuntimeError("Assertion failed: isValid() == true")

    def dilate(self, r: float):
        self._apply("dilate", r)

    def erode(self, r: float):
        self._apply("erode", r)

    def close(self, r: float):
        self._apply("close", r)

    def open(self, r: float):
        self._apply("open", r)

    def negate(self):
        self._apply("negate", 0.0)
