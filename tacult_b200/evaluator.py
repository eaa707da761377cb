"""Drop-in for `NNetWrapper.predict` (/root/reference/src/tacult/base/nn_wrapper.py:125-136).

`B200Evaluator(wrapper)` wraps a reference `NNetWrapper` (or anything with `.nnet.state_dict()`): `predict`
runs the batched forward of `_UtacNNet` (network.py:81-102) on the GPU through the C ABI; every other
attribute (train, save_checkpoint, load_checkpoint ...) is delegated to the wrapped object, and the GPU
copy of the weights is refreshed after calls that change them.
"""
import numpy as np

from . import _lib
from .engine import Engine


class B200Evaluator:
    _REFRESH_AFTER = ("train", "load_checkpoint")

    def __init__(self, wrapped, max_batch=4096, device=0, precision="f32"):
        self._wrapped = wrapped
        self._engine = Engine(max_batch, 2, device=device)
        self._mode = _lib.EVAL_NET_BF16 if str(precision).lower() == "bf16" else _lib.EVAL_NET_F32
        self.refresh_weights()

    @property
    def nnet(self):
        return self._wrapped.nnet

    def refresh_weights(self):
        module = getattr(self._wrapped, "nnet", self._wrapped)
        self._engine.load_weights(module.state_dict())

    def predict(self, obs):
        x = obs.detach().cpu().numpy() if hasattr(obs, "detach") else np.asarray(obs)
        pi, v = self._engine.forward(x.astype(np.float32, copy=False), self._mode)
        return pi, v.reshape(-1, 1)

    def __getattr__(self, name):
        attr = getattr(self._wrapped, name)
        if name in self._REFRESH_AFTER and callable(attr):
            def call_then_refresh(*a, **kw):
                out = attr(*a, **kw)
                self.refresh_weights()
                return out
            return call_then_refresh
        return attr
