"""StateBuffer with the reference's call surface (/root/reference/src/state_buffer.py:3-27),
kept on the device so ``predict`` reads the 4-frame window without a 0.9 MB upload per step."""
import ctypes as C

import numpy as np

from . import _lib as L


class DeviceStates:
    """Handle returned by ``StateBuffer.getStateMinibatch()``: shaped like the reference's
    (batch, hist, h, w) uint8 buffer; ``DeepQNetwork.predict`` consumes it on the device, numpy
    consumers get a host copy through ``__array__``."""

    def __init__(self, buf):
        self._buf = buf
        self.shape = (buf.batch_size, buf.history_length) + buf.dims
        self.dtype = np.dtype(np.uint8)
        self.live_rows = 1                  # only row 0 is ever written (state_buffer.py:15-18)

    def device_ptr(self):
        return self._buf._device_ptr()

    def __array__(self, dtype=None, copy=None):
        a = self._buf._read(whole=True)
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, i):
        return np.asarray(self)[i]


class StateBuffer:
    def __init__(self, args, device=0, stream=None):
        self.history_length = args.history_length
        self.dims = (args.screen_height, args.screen_width)
        self.batch_size = args.batch_size
        self.device = device
        self._stream_obj = stream            # keep the stream alive as long as this object uses it
        self._stream = L.stream_ptr(stream)
        h = C.c_void_p()
        L.call("b200dqn_statebuf_create", device, self.dims[0], self.dims[1], self.history_length, self.batch_size,
               C.byref(h))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                L.load().b200dqn_statebuf_destroy(h)
            except Exception:
                pass

    def _device_ptr(self):
        p, b = C.c_void_p(), C.c_size_t()
        L.call("b200dqn_statebuf_device_ptr", self._h, C.byref(p), C.byref(b))
        return p.value

    def _read(self, whole):
        shape = ((self.batch_size,) if whole else ()) + (self.history_length,) + self.dims
        out = np.empty(shape, dtype=np.uint8)
        L.call("b200dqn_statebuf_read", self._h, L.np_ptr(out), int(whole), self._stream)
        return out

    @property
    def buffer(self):
        return self._read(whole=True)

    def add(self, observation):
        assert observation.shape == self.dims                          # :16
        obs = np.ascontiguousarray(observation, dtype=np.uint8)
        L.call("b200dqn_statebuf_add", self._h, L.np_ptr(obs), self._stream)

    def getState(self):
        return self._read(whole=False)

    def getStateMinibatch(self):
        return DeviceStates(self)

    def reset(self):
        L.call("b200dqn_statebuf_reset", self._h, self._stream)
