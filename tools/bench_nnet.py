"""Throughput of the cost-to-go network paths on one GPU (rows/s, TFLOP/s)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepcubea_b200.nnet.folded import FoldedResnet
from deepcubea_b200.nnet.tc_resnet import TcResnet
from deepcubea_b200.utils.pytorch_models import ResnetModel
torch.manual_seed(0)
dev = torch.device("cuda")
model = ResnetModel(54, 6, 5000, 1000, 4, 1, True).eval()
x = torch.randint(0, 6, (131072, 54), device=dev, dtype=torch.uint8)
only = sys.argv[1] if len(sys.argv) > 1 else ""           # e.g. "tc" -> only the hand-written paths
makers = {"torch fp32": lambda: FoldedResnet(model, "fp32").to(dev), "torch tf32": lambda: FoldedResnet(model, "tf32").to(dev),
          "torch bf16": lambda: FoldedResnet(model, "bf16").to(dev), "tc fp16x3": lambda: TcResnet(model, dev, "fp16x3"),
          "tc fp16": lambda: TcResnet(model, dev, "fp16")}
paths = {k: mk() for k, mk in makers.items() if only in k}
for name, f in paths.items():
    for _ in range(2): f(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): f(x)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print("%-12s %8.2f ms / %d rows  -> %6.2f M rows/s, %7.1f dense-equivalent TFLOP/s (29.24 MFLOP/row)" % (name, ms, x.shape[0], x.shape[0] / ms / 1e3, 29.24e6 * x.shape[0] / ms / 1e9))
    if hasattr(f, "gemm_events"):                          # per-layer CUDA-event times of one more pass
        f.gemm_events = []
        f(x); torch.cuda.synchronize()
        print("    per launch (us): " + " ".join("%.0f" % (a.elapsed_time(b) * 1e3) for a, b, _ in f.gemm_events))
        f.gemm_events = None
