"""Workload for ncu -k regex:resnet_gemm: the cost-to-go network's tcgen05 layers on 131072 rows (fp16x3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepcubea_b200.nnet.tc_resnet import TcResnet
from deepcubea_b200.utils.pytorch_models import ResnetModel
torch.manual_seed(0)
dev = torch.device("cuda")
tc = TcResnet(ResnetModel(54, 6, 5000, 1000, 4, 1, True).eval(), dev, sys.argv[1] if len(sys.argv) > 1 else "fp16x3", chunk=1 << 17)
x = torch.randint(0, 6, (131072, 54), device=dev, dtype=torch.uint8)
for _ in range(2):
    tc(x)
torch.cuda.synchronize()
print("done")
