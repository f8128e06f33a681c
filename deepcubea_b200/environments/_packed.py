"""Shared machinery of the GPU-backed environments: List[State] <-> packed uint8[N,S] and the device calls."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch

from .. import _lib, ops


class PackedEnvMixin:
    """Expects `self.env_id`, `self.state_dim`, `self._state_cls`, `self._attr` ("colors"/"tiles")."""

    def _device(self) -> torch.device:
        if not torch.cuda.is_available():
            raise _lib.DcbError("%s runs its batched operations on the GPU; no CUDA device is visible and there "
                                "is no CPU fallback" % type(self).__name__)
        return torch.device("cuda", torch.cuda.current_device())

    def pack(self, states: List) -> np.ndarray:
        """Stack states into uint8[N,S] (shipped pickles hold int64 payloads for cube3 / puzzle15)."""
        attr = self._attr
        if len(states) == 0:
            return np.zeros((0, self.state_dim), dtype=np.uint8)
        return np.ascontiguousarray(np.stack([getattr(s, attr) for s in states], axis=0).astype(np.uint8, copy=False))

    def unpack(self, arr: np.ndarray) -> List:
        cls, dt = self._state_cls, self.dtype
        arr = arr.astype(dt, copy=False)
        return [cls(x) for x in arr]

    def to_device(self, arr: np.ndarray) -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.uint8)).to(self._device(), non_blocking=False)

    # ---- tensor-native fast paths (used by the GPU search; no Python objects) -------------------------
    def expand_packed(self, parents: torch.Tensor):
        """u8[N,S] cuda -> (children u8[N,A,S], solved u8[N,A], hash i64[N,A])."""
        return ops.expand(self.env_id, parents)

    def next_state_packed(self, states: torch.Tensor, action: int) -> torch.Tensor:
        return ops.next_state(self.env_id, states, action)

    def is_solved_packed(self, states: torch.Tensor) -> torch.Tensor:
        return ops.is_solved(self.env_id, states)

    def nnet_input_packed(self, states: torch.Tensor) -> torch.Tensor:
        return ops.nnet_input(self.env_id, states)

    # ---- List[State] API ---------------------------------------------------------------------------------
    def _next_state_np(self, states_np: np.ndarray, action: int) -> Tuple[np.ndarray, List[float]]:
        out = ops.next_state(self.env_id, self.to_device(states_np), action).cpu().numpy()
        return out, [1.0 for _ in range(states_np.shape[0])]

    def _expand_np(self, states_np: np.ndarray) -> np.ndarray:
        children, _, _ = ops.expand(self.env_id, self.to_device(states_np), want_solved=False, want_hash=False)
        return children.cpu().numpy()

    def _is_solved_np(self, states_np: np.ndarray) -> np.ndarray:
        if states_np.shape[0] == 0:
            return np.zeros(0, dtype=bool)
        return ops.is_solved(self.env_id, self.to_device(states_np)).cpu().numpy().astype(bool)
