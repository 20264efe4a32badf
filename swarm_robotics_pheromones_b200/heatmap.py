"""HeatMap — Phase-3/graph.py's offline loop on the GPU: per trajectory point,
`heatmap[|int(y*H)|, |int(x*W)|] += 20` then a whole-map Gaussian blur (scipy's gaussian_filter
semantics: separable, mode='reflect', truncate=4.0, fp64)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _ffi


def gaussian_weights(sigma: float, truncate: float = 4.0):
    """the 1-D kernel exactly as scipy.ndimage._filters._gaussian_kernel1d builds it (order 0);
    returns (radius, weights[centre .. right end])"""
    sd = float(sigma)
    radius = int(truncate * sd + 0.5)
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi_x = np.exp(-0.5 / sigma2 * x ** 2)
    phi_x = phi_x / phi_x.sum()
    return radius, np.ascontiguousarray(phi_x[radius:], dtype=np.float64)


class HeatMap:
    def __init__(self, height: int, width: int, sigma: float = 5.0, device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("pherofield needs a CUDA device (sm_100a); there is no CPU fallback")
        self._L = _ffi.lib()
        self.shape = (height, width)
        radius, w = gaussian_weights(sigma)
        h = C.c_void_p()
        rc = self._L.pf_heatmap_create(device, height, width, radius, w.ctypes.data, C.byref(h))
        if rc:
            raise _ffi.PFError(rc, (self._L.pf_heatmap_last_error(None) or b"").decode())
        self._h = h
        self.device = torch.device("cuda", device)

    def _ck(self, rc):
        if rc:
            raise _ffi.PFError(rc, (self._L.pf_heatmap_last_error(self._h) or b"").decode())

    def accumulate(self, points, amount: float = 20.0):
        """graph.py:17-24 for a list of [x, y, (z)] points"""
        p = np.asarray(points, np.float64)
        x = np.ascontiguousarray(p[:, 0])
        y = np.ascontiguousarray(p[:, 1])
        s = torch.cuda.current_stream(self.device).cuda_stream
        self._ck(self._L.pf_heatmap_accumulate(self._h, len(x), x.ctypes.data, y.ctypes.data, amount, s))

    def snapshot(self) -> np.ndarray:
        out = np.empty(self.shape, np.float64)
        s = torch.cuda.current_stream(self.device).cuda_stream
        self._ck(self._L.pf_heatmap_snapshot_host(self._h, out.ctypes.data, s))
        return out

    def close(self):
        if getattr(self, "_h", None):
            self._L.pf_heatmap_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
