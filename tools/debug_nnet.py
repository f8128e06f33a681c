"""Experiment: does K-chunked accumulation (fresh TMEM accumulator per chunk, fp32 RN adds between chunks) bring the
fp16x3 network under 1e-4?  Emulated with per-chunk GEMM calls + torch epilogue."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from deepcubea_b200 import _lib
from deepcubea_b200.nnet.tc_resnet import TcResnet
from deepcubea_b200.utils.pytorch_models import ResnetModel
from oracle.oracle_env import OracleCube3
dev = torch.device("cuda"); lib = _lib.load(); p = _lib.ptr
st = torch.cuda.current_stream().cuda_stream
m = ResnetModel(54, 6, 5000, 1000, 4, 1, True)
sd = torch.load("assets/saved_models/cube3/current/model_state_dict.pt", map_location="cpu")
m.load_state_dict({k.replace("module.", "", 1): v for k, v in sd.items()}); m.eval()
g = np.load("tests/golden/nnet_cube3.npz")
x = torch.from_numpy(OracleCube3().nnet_input(g["states"])).cuda()
m64 = copy.deepcopy(m).double().to(dev); m64._encode = lambda t: torch.nn.functional.one_hot(t.long(), 6).double().flatten(1)
with torch.no_grad(): ref64 = m64(x)[:, 0]
tc = TcResnet(m, dev, "fp16x3")
print("current kernel: max err vs fp64 %.3g, vs golden %.3g" % ((tc(x).double() - ref64).abs().max().item(), np.abs(tc(x).cpu().numpy() - g["ctg"]).max()))
M = x.shape[0]
def split(v):
    hi = v.half(); return hi, (v - hi.float()).half()
def gemm_chunked(layer, a_hi, a_lo, chunk, sep_lo):
    acc = torch.zeros((M, layer.np_), dtype=torch.float32, device=dev)
    zb = torch.zeros(layer.np_, device=dev)
    for k0 in range(0, layer.kp, chunk):
        k1 = min(layer.kp, k0 + chunk)
        ah = a_hi[:, k0:k1].contiguous(); al = a_lo[:, k0:k1].contiguous() if a_lo is not None else None
        wh = layer.w_hi[:, k0:k1].contiguous(); wl = layer.w_lo[:, k0:k1].contiguous()
        o_hi = torch.empty((M, layer.np_), dtype=torch.float16, device=dev); o_f = torch.empty((M, layer.np_), dtype=torch.float32, device=dev)
        if sep_lo:
            # main product alone, then the two small products in their own accumulation
            _lib.check(lib.dcb_resnet_gemm(p(ah), None, p(wh), None, p(zb), 1.0, None, None, 0, p(o_hi), None, p(o_f), M, layer.np_, k1 - k0, st)); acc += o_f
            _lib.check(lib.dcb_resnet_gemm(p(ah), None, p(wl), None, p(zb), 1.0, None, None, 0, p(o_hi), None, p(o_f), M, layer.np_, k1 - k0, st)); acc += o_f
            if al is not None:
                _lib.check(lib.dcb_resnet_gemm(p(al), None, p(wh), None, p(zb), 1.0, None, None, 0, p(o_hi), None, p(o_f), M, layer.np_, k1 - k0, st)); acc += o_f
        else:
            _lib.check(lib.dcb_resnet_gemm(p(ah), p(al), p(wh), p(wl), p(zb), 1.0, None, None, 0, p(o_hi), None, p(o_f), M, layer.np_, k1 - k0, st)); acc += o_f
    return acc * layer.scale + layer.bias
def forward(chunk_first, chunk_rest, sep_lo):
    a0 = torch.empty((M, 384), dtype=torch.float16, device=dev)
    _lib.check(lib.dcb_onehot_fp16(p(x), M, 54, 6, 384, p(a0), st))
    L = tc.layers
    h = torch.relu(gemm_chunked(L[0], a0, None, 384, sep_lo)); hi, lo = split(h)
    v = torch.relu(gemm_chunked(L[1], hi, lo, chunk_first, sep_lo)); xh, xl = split(v)
    for k in range(4):
        t = torch.relu(gemm_chunked(L[2 + 2 * k], xh, xl, chunk_rest, sep_lo)); th, tl = split(t)
        v = torch.relu(gemm_chunked(L[3 + 2 * k], th, tl, chunk_rest, sep_lo) + xh.float() + xl.float()); xh, xl = split(v)
    out = ((xh.float() + xl.float())[:, :1000] @ tc.w_out) + tc.b_out
    return out
for cf, cr, sep in ((5120, 1024, False), (1024, 1024, False), (512, 512, False), (256, 256, False), (1024, 1024, True), (256, 256, True), (64, 64, False)):
    o = forward(cf, cr, sep)
    print("chunk fc2=%4d rest=%4d sep_lo=%d : max err vs fp64 %.3g  vs golden %.3g" % (cf, cr, sep, (o.double() - ref64).abs().max().item(), np.abs(o.cpu().numpy() - g["ctg"]).max()))
