"""Inference half of the reference's utils/nnet_utils.py (device pick :122-130, load_nnet :134-152,
get_heuristic_fn :160-196, load_heuristic_fn :206-221).  The training loop and the per-GPU runner
processes (:53-118, :247-311) are outside the hot path: in this design the network runs in the search's
own process and stream, one process per GPU.
"""
from __future__ import annotations

import os
import re
from collections import OrderedDict
from typing import List, Optional, Tuple

import numpy as np
import torch
from torch import nn

from ..environments.environment_abstract import Environment


def get_available_gpu_nums() -> List[int]:
    devices: Optional[str] = os.environ.get("CUDA_VISIBLE_DEVICES")
    return [int(x) for x in devices.split(",")] if devices else []


def get_device() -> Tuple[torch.device, List[int], bool]:
    """The process's CURRENT CUDA device when there is one (cuda:0 unless the caller picked another with
    torch.cuda.set_device -- one process per GPU under torchrun; the reference always answers cuda:0 because it isolates its
    per-GPU runner processes with CUDA_VISIBLE_DEVICES, nnet_utils.py:122-130).  An unset CUDA_VISIBLE_DEVICES means "all GPUs"
    (the reference requires it to be set)."""
    devices = get_available_gpu_nums()
    if torch.cuda.is_available():
        if not devices:
            devices = list(range(torch.cuda.device_count()))
        return torch.device("cuda", torch.cuda.current_device()), devices, True
    return torch.device("cpu"), devices, False


def states_nnet_to_pytorch_input(states_nnet: List[np.ndarray], device) -> List[torch.Tensor]:
    return [torch.tensor(x, device=device) for x in states_nnet]


def load_nnet(model_file: str, nnet: nn.Module, device: torch.device = None) -> nn.Module:
    """Load `model_state_dict.pt`, stripping the DataParallel `module.` prefix (nnet_utils.py:134-152)."""
    state_dict = torch.load(model_file, map_location=device) if device is not None else torch.load(model_file)
    clean = OrderedDict((re.sub(r"^module\.", "", k), v) for k, v in state_dict.items())
    nnet.load_state_dict(clean)
    nnet.eval()
    return nnet


def get_heuristic_fn(nnet: nn.Module, device: torch.device, env: Environment, clip_zero: bool = False,
                     batch_size: Optional[int] = None):
    """Reference-shaped heuristic: List[State] (or nnet-format arrays) -> float64 ndarray holding fp32 values.
    The returned function also carries `.device_fn`, the zero-copy form the CUDA search calls
    (nnet-input u8 tensor on the device -> f32 tensor on the device)."""
    nnet.eval()

    def device_fn(x: torch.Tensor) -> torch.Tensor:
        with torch.no_grad():
            n = x.shape[0]
            step = batch_size if batch_size is not None else max(n, 1)
            outs = []
            for i in range(0, n, step):
                o = nnet(x[i:i + step])
                outs.append(o[:, 0] if o.dim() == 2 else o)
            out = torch.cat(outs) if len(outs) != 1 else outs[0]
            return out.float().contiguous()

    def heuristic_fn(states: List, is_nnet_format: bool = False) -> np.ndarray:
        if is_nnet_format:
            arrays = states
        else:
            arrays = env.state_to_nnet_input(states)
        n = arrays[0].shape[0]
        if n == 0:
            return np.zeros(0)
        x = torch.tensor(arrays[0], device=device)
        cost_to_go = device_fn(x).cpu().numpy().astype(np.float64)
        assert cost_to_go.shape[0] == n
        return np.maximum(cost_to_go, 0.0) if clip_zero else cost_to_go

    heuristic_fn.device_fn = device_fn
    heuristic_fn.clip_zero = clip_zero
    return heuristic_fn


def load_heuristic_fn(nnet_dir: str, device: torch.device, on_gpu: bool, nnet: nn.Module, env: Environment,
                      clip_zero: bool = False, gpu_num: int = -1, batch_size: Optional[int] = None,
                      precision: Optional[str] = None):
    """nnet_utils.py:206-221.  `precision` (or $DCB_NNET_PRECISION) selects the inference arithmetic of the
    network: fp16x3 (hand-written tcgen05 layers, fp32-parity: max |err| 3e-5 vs the fp64 network; the default on a GPU) | fp32
    (cuBLAS SGEMM, the reference's arithmetic) | tf32 | bf16 | fp16 (tcgen05, reduced precision).  Networks the tcgen05 path does
    not encode (no one-hot input) fall back to fp32 unless a precision was asked for explicitly."""
    if gpu_num >= 0 and on_gpu:
        os.environ["CUDA_VISIBLE_DEVICES"] = str(gpu_num)
    nnet = load_nnet("%s/model_state_dict.pt" % nnet_dir, nnet, device=device)
    nnet.eval()
    nnet.to(device)
    if on_gpu:
        mode = precision or os.environ.get("DCB_NNET_PRECISION")
        if mode is None:
            mode = "fp16x3" if getattr(nnet, "one_hot_depth", 0) > 0 else "fp32"
        if mode in ("fp16x3", "fp16"):            # hand-written tcgen05 dense layers
            from ..nnet.tc_resnet import TcResnet
            tc = TcResnet(nnet, device, mode=mode, chunk=max(batch_size or 0, 1 << 18))
            fn = get_heuristic_fn(nnet, device, env, clip_zero=clip_zero, batch_size=batch_size)
            fn.device_fn = tc
            return fn
        from ..nnet.folded import FoldedResnet
        nnet = FoldedResnet(nnet, mode=mode).to(device)
    return get_heuristic_fn(nnet, device, env, clip_zero=clip_zero, batch_size=batch_size)
