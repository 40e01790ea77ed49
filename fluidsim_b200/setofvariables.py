"""Tensor-backed ``SetOfVariables`` (GPU counterpart of ``fluiddyn.calcul.setofvariables``).

The reference container is an ``np.ndarray`` subclass of shape ``(nvar, *shape_variable)`` with
``keys / nvar / info`` attributes and ``get_var / set_var / initialize``
(``/root/reference/fluidsim/base/setofvariables.py:12``; constructor uses at
``base/state.py:63-68,243-248`` and ``solvers/ns3d/solver.py:229-231``).  Here the storage is one
contiguous CUDA ``torch`` tensor (``.tensor``); ``get_var`` returns a view of it, so kernels
write in place exactly like the reference's in-place numpy updates.
"""

import numpy as np
import torch


class SetOfVariables:
    def __init__(
        self,
        input_array=None,
        keys=None,
        shape_variable=None,
        like=None,
        value=None,
        info=None,
        dtype=None,
        device=None,
    ):
        if input_array is not None:
            if keys is None:
                raise ValueError("keys should be provided with input_array")
            tensor = input_array
        elif like is not None:
            info = info if info is not None else like.info
            keys = like.keys
            dtype = dtype if dtype is not None else like.dtype
            tensor = torch.empty(like.shape, dtype=dtype, device=like.tensor.device)
        else:
            if keys is None or shape_variable is None:
                raise ValueError("keys and shape_variable are required")
            if dtype is None:
                dtype = torch.float64
            if isinstance(dtype, (np.dtype, type)) and not isinstance(dtype, torch.dtype):
                dtype = {np.dtype("float64"): torch.float64, np.dtype("complex128"): torch.complex128}[
                    np.dtype(dtype)
                ]
            device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
            tensor = torch.empty([len(keys)] + list(shape_variable), dtype=dtype, device=device)
        if value is not None and input_array is None:
            tensor.fill_(value)
        self.tensor = tensor
        self.keys = list(keys)
        self.nvar = len(self.keys)
        self.info = info
        # optional hook called whenever a writable view is handed out or the container is written
        # through its own methods (the state uses it to invalidate "state is dealiased", see state.py)
        self._on_touch = None

    def _touch(self):
        if self._on_touch is not None:
            self._on_touch()

    # ndarray-like surface used by the reference code paths we mirror
    @property
    def shape(self):
        return tuple(self.tensor.shape)

    @property
    def dtype(self):
        return self.tensor.dtype

    @property
    def ndim(self):
        return self.tensor.dim()

    def __getitem__(self, item):
        self._touch()
        return self.tensor[item]

    def __setitem__(self, item, value):
        if isinstance(value, SetOfVariables):
            value = value.tensor
        elif isinstance(value, np.ndarray):
            value = torch.from_numpy(value).to(self.tensor.device)
        self._touch()
        self.tensor[item] = value

    def __iadd__(self, other):
        self._touch()
        self.tensor += other.tensor if isinstance(other, SetOfVariables) else other
        return self

    def __array__(self, dtype=None, copy=None):
        a = self.tensor.detach().cpu().numpy()
        return a if dtype is None else a.astype(dtype)

    def get_var(self, arg):
        index = arg if isinstance(arg, int) else self.keys.index(arg)
        self._touch()
        return self.tensor[index]

    def set_var(self, arg, value):
        index = arg if isinstance(arg, int) else self.keys.index(arg)
        self[index] = value

    def initialize(self, value=0):
        self._touch()
        self.tensor.fill_(value)

    def fill(self, value):
        self._touch()
        self.tensor.fill_(value)

    def copy(self):
        return SetOfVariables(input_array=self.tensor.clone(), keys=self.keys, info=self.info)

    def numpy(self):
        return self.tensor.detach().cpu().numpy()
