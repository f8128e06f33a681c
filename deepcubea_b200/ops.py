"""Tensor-level wrappers over the C ABI: PyTorch owns device memory and the stream, the library launches
the hand-written kernels on them.  Every function requires CUDA tensors and raises if the extension or
the GPU is missing (no CPU fallback)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import ENV_IDS, check, ptr


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise _lib.DcbError("%s must be a CUDA tensor (this path has no CPU implementation)" % name)
    if not t.is_contiguous():
        raise _lib.DcbError("%s must be contiguous" % name)


def env_shape(env_id: int) -> Tuple[int, int]:
    lib = _lib.load()
    return lib.dcb_env_state_bytes(env_id), lib.dcb_env_num_moves(env_id)


def expand(env_id: int, parents: torch.Tensor, want_solved: bool = True, want_hash: bool = True,
           out: Optional[torch.Tensor] = None):
    """parents u8[N,S] (cuda) -> children u8[N,A,S], solved u8[N,A] | None, hash i64[N,A] (u64 bits) | None."""
    lib = _lib.load()
    _need_cuda(parents, "parents")
    s, a = env_shape(env_id)
    n = parents.shape[0]
    assert parents.dtype == torch.uint8 and parents.shape[1] == s
    children = out if out is not None else torch.empty((n, a, s), dtype=torch.uint8, device=parents.device)
    solved = torch.empty((n, a), dtype=torch.uint8, device=parents.device) if want_solved else None
    hsh = torch.empty((n, a), dtype=torch.int64, device=parents.device) if want_hash else None
    check(lib.dcb_expand(env_id, ptr(parents), n, ptr(children), ptr(solved), ptr(hsh), _stream()), "dcb_expand")
    return children, solved, hsh


def next_state(env_id: int, states: torch.Tensor, action: int) -> torch.Tensor:
    lib = _lib.load()
    _need_cuda(states, "states")
    out = torch.empty_like(states)
    check(lib.dcb_next_state(env_id, ptr(states), states.shape[0], int(action), ptr(out), _stream()), "dcb_next_state")
    return out


def is_solved(env_id: int, states: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _need_cuda(states, "states")
    out = torch.empty((states.shape[0],), dtype=torch.uint8, device=states.device)
    check(lib.dcb_is_solved(env_id, ptr(states), states.shape[0], ptr(out), _stream()), "dcb_is_solved")
    return out


def hash_states(env_id: int, states: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _need_cuda(states, "states")
    out = torch.empty((states.shape[0],), dtype=torch.int64, device=states.device)
    check(lib.dcb_hash_states(env_id, ptr(states), states.shape[0], ptr(out), _stream()), "dcb_hash_states")
    return out


def nnet_input(env_id: int, states: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _need_cuda(states, "states")
    out = torch.empty_like(states)
    check(lib.dcb_nnet_input(env_id, ptr(states), states.shape[0], ptr(out), _stream()), "dcb_nnet_input")
    return out
