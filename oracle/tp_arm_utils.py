"""Restatement of the arm_pytorch_utilities helpers the reference uses.

TEST INFRASTRUCTURE ONLY.  Third-party ("arm-pytorch-utilities>=0.4",
pyproject.toml:57), absent from this image: "parity unpinned".  Call sites:
src/pytorch_volumetric/sdf.py:122 (handle_batch_input), :166 (ensure_tensor),
:644-645 (rand.SavedRNG, rand.seed).  Behaviour evidenced by
tests/test_sdf.py:26-28: a (10,100,3) input yields (10,100) / (10,100,3).
"""
import functools
import random as _pyrandom
import types

import numpy as np
import torch


def handle_batch_input(n):
    """Flatten every leading batch dim of array arguments that have more than
    `n` dims, call, then restore the batch dims on every returned array."""

    def decorator(func):
        @functools.wraps(func)
        def wrapper(*args, **kwargs):
            batch_dims = None
            new_args = []
            first = next((a for a in args if torch.is_tensor(a) or isinstance(a, np.ndarray)), None)
            if first is not None and len(first.shape) < n:
                # fewer dims than expected: add leading singleton dims, squeeze them off the outputs
                k = n - len(first.shape)
                args2 = [a.reshape(*([1] * k), *a.shape) if (torch.is_tensor(a) or isinstance(a, np.ndarray)) else a
                         for a in args]
                ret = func(*args2, **kwargs)
                rets = ret if isinstance(ret, tuple) else (ret,)
                out = [r.reshape(r.shape[k:]) if (torch.is_tensor(r) or isinstance(r, np.ndarray)) else r
                       for r in rets]
                return tuple(out) if isinstance(ret, tuple) else out[0]
            for a in args:
                if (torch.is_tensor(a) or isinstance(a, np.ndarray)) and len(a.shape) > n:
                    if batch_dims is None:
                        batch_dims = tuple(a.shape[:-(n - 1)])
                    a = a.reshape(-1, *a.shape[-(n - 1):])
                new_args.append(a)
            ret = func(*new_args, **kwargs)
            if batch_dims is None:
                return ret
            single = not isinstance(ret, tuple)
            rets = (ret,) if single else ret
            out = []
            for r in rets:
                if r is None or not (torch.is_tensor(r) or isinstance(r, np.ndarray)):
                    out.append(r)
                else:
                    out.append(r.reshape(*batch_dims, *r.shape[1:]))
            return out[0] if single else tuple(out)

        return wrapper

    return decorator


def ensure_tensor(device, dtype, *args):
    out = tuple(a.to(device=device, dtype=dtype) if torch.is_tensor(a)
                else (None if a is None else torch.tensor(a, device=device, dtype=dtype)) for a in args)
    return out if len(out) > 1 else out[0]


def seed(s):
    _pyrandom.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)
    return s


class SavedRNG:
    def __enter__(self):
        self._py = _pyrandom.getstate()
        self._np = np.random.get_state()
        self._torch = torch.get_rng_state()
        return self

    def __exit__(self, *exc):
        _pyrandom.setstate(self._py)
        np.random.set_state(self._np)
        torch.set_rng_state(self._torch)
        return False


tensor_utils = types.SimpleNamespace(handle_batch_input=handle_batch_input, ensure_tensor=ensure_tensor)
rand = types.SimpleNamespace(seed=seed, SavedRNG=SavedRNG)
