"""CPU ORACLE helper (test infrastructure): a headless restatement of the Neural Photo Editor's event handlers
(reference NPE.py), parameterised by a `model` object with the API.IAN surface.  Running the same scripted session
once over the CUDA-backed IAN and once over `OracleModel` (the float64 oracle behind the same surface) checks the
drop-in claim end to end: same call sequence, same dtypes (NPE's Z silently becomes float64 after sample(),
NPE.py:319), same box arithmetic with integral floats (NPE.py:202), same blends.

Handlers restated: infer (NPE.py:239-274), move_mouse (143-156), paint (192-235), scroll (305-314),
sample (317-327), Reset (330-340), update_photo's display conversion (107-118).  Tk widgets are replaced by plain
state; `event` is (x, y) in canvas pixels.
"""
from __future__ import annotations

import numpy as np

from . import ian_numpy as on


class OracleModel:
    """API.IAN surface over the float64 oracle (IAN_simple graph), returning float32 like theano.function would."""

    def __init__(self, P):
        self.P = P

    def encode_images(self, x):
        return on.simple_encode(self.P, x).astype(np.float32)

    def sample_at(self, z):
        return on.simple_decode(self.P, z).astype(np.float32)

    def imgrad(self, c1, r1, c2, r2, z):
        return on.simple_imgrad(self.P, c1, r1, c2, r2, z).astype(np.float32)

    def imgradRGB(self, c1, r1, c2, r2, RGB, z):
        return on.simple_imgradRGB(self.P, c1, r1, c2, r2, RGB, z).astype(np.float32)


class NPESession:
    def __init__(self, model, brush_d=12):
        self.model = model
        self.d = brush_d                                        # d.set(12), NPE.py:101
        self.Z = np.zeros((10, 10), dtype=np.float32)           # NPE.py:70
        self.myRGB = np.zeros((1, 3, 64, 64), dtype=np.float32)  # NPE.py:87
        self.GIM = self.IM = self.RECON = None
        self.ERROR = None
        self.SAMPLE_FLAG = 0
        self.rect = [0.0, 0.0, 0.0, 0.0]                        # output.coords(pixel_rect): Tk returns floats
        self.display = None

    # ---- helpers ---------------------------------------------------------------------------------------
    def set_color(self, rgb255):
        """the colour chooser fills myRGB with one colour (NPE.py:~352-356)."""
        self.myRGB = np.broadcast_to(np.float32(rgb255).reshape(1, 3, 1, 1), (1, 3, 64, 64)).astype(np.float32).copy()

    def update_photo(self, data=None):                          # NPE.py:107-118
        if data is None:
            data = np.uint8(on.from_tanh(self.model.sample_at(np.float32([self.Z.flatten()]))[0]))
        self.display = on.npe_display(np.uint8(data))

    # ---- handlers --------------------------------------------------------------------------------------
    def infer(self, image_u8):                                  # NPE.py:239-274
        self.GIM = np.asarray(image_u8)
        self.IM = self.GIM
        s = self.model.encode_images(np.asarray([on.to_tanh(self.IM)], dtype=np.float32))
        self.Z = np.reshape(s[0], np.shape(self.Z))
        self.RECON = np.uint8(on.from_tanh(self.model.sample_at(np.float32([self.Z.flatten()]))[0]))
        self.ERROR = on.to_tanh(np.float32(self.IM)) - on.to_tanh(np.float32(self.RECON))
        self.SAMPLE_FLAG = 0
        self.update_photo(self.IM)

    def move_mouse(self, ex, ey):                               # NPE.py:143-156
        x, y = ex // 4, ey // 4
        bw = (self.d // 4) + 1
        xmin = max(min(x - bw // 2, 64 - bw), 0)
        ymin = max(min(y - bw // 2, 64 - bw), 0)
        self.rect = [float(4 * xmin), float(4 * ymin), float(4 * (xmin + bw)), float(4 * (ymin + bw))]

    def paint(self, ex, ey):                                    # NPE.py:192-235
        self.move_mouse(ex, ey)
        weight = 0.05
        x1, y1, x2, y2 = [c // 4 for c in self.rect]            # floats with integral values
        temp = np.asarray(self.model.imgradRGB(x1, y1, x2, y2, np.float32(on.to_tanh(self.myRGB)),
                                               np.float32([self.Z.flatten()]))[0])
        grad = temp.reshape((10, 10)) * (1 + (x2 - x1))
        self.Z = self.Z - weight * grad                         # `Z -= weight*grad` (float32 or float64 Z)
        if self.SAMPLE_FLAG:
            self.update_photo(None)
        else:
            xh = self.model.sample_at(np.float32([self.Z.flatten()]))[0]
            self.IM = on.npe_paint_blend(xh, self.RECON, self.ERROR)
            self.update_photo(self.IM)

    def scroll(self, delta):                                    # NPE.py:305-314
        weight = 0.1
        x1, y1, x2, y2 = [c // 4 for c in self.rect]
        grad = np.reshape(self.model.imgrad(x1, y1, x2, y2, np.float32([self.Z.flatten()]))[0], self.Z.shape) * (1 + (x2 - x1))
        self.Z = self.Z + np.sign(delta) * weight * grad
        self.update_photo(None)

    def sample(self, seed):                                     # NPE.py:317-327 (np.random.randn -> float64 Z)
        self.Z = np.random.RandomState(seed).randn(self.Z.shape[0], self.Z.shape[1])
        self.RECON = np.uint8(on.from_tanh(self.model.sample_at(np.float32([self.Z.flatten()]))[0]))
        self.ERROR = on.to_tanh(np.float32(self.IM)) - on.to_tanh(np.float32(self.RECON))
        self.SAMPLE_FLAG = 1
        self.update_photo(None)

    def reset(self):                                            # NPE.py:330-340
        self.IM = self.GIM
        self.Z = np.reshape(self.model.encode_images(np.asarray([on.to_tanh(self.IM)], dtype=np.float32))[0], np.shape(self.Z))
        self.RECON = np.uint8(on.from_tanh(self.model.sample_at(np.float32([self.Z.flatten()]))[0]))
        self.ERROR = on.to_tanh(np.float32(self.IM)) - on.to_tanh(np.float32(self.RECON))
        self.SAMPLE_FLAG = 0
        self.update_photo(self.IM)


def scripted_session(model, image_u8):
    """a fixed little editing session; returns the trace of states a test compares."""
    s = NPESession(model)
    trace = {}
    s.infer(image_u8)
    trace["z_infer"] = s.Z.copy(); trace["recon"] = s.RECON.copy()
    s.set_color((220, 40, 40))
    for k, (ex, ey) in enumerate([(100, 120), (104, 124), (108, 128)]):        # three <B1-Motion> events
        s.paint(ex, ey)
    trace["z_paint"] = s.Z.copy(); trace["im_paint"] = s.IM.copy(); trace["display_paint"] = s.display.copy()
    s.scroll(+120)
    trace["z_scroll"] = s.Z.copy()
    s.sample(7)
    s.paint(30, 200)                                                            # painting on a sample (SAMPLE_FLAG=1)
    trace["z_sample_paint"] = np.asarray(s.Z).copy(); trace["z_dtype_after_sample"] = str(np.asarray(s.Z).dtype)
    s.reset()
    trace["z_reset"] = s.Z.copy(); trace["display_reset"] = s.display.copy()
    return trace
