"""GPU: the tcgen05 dense layers against PyTorch references (fp64 for single GEMMs, the fp32 network end to end)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _split(x):
    hi = x.to(torch.float16)
    return hi, (x - hi.float()).to(torch.float16)


@pytest.mark.parametrize("m", [1, 127, 128, 129, 1000, 33333])
@pytest.mark.parametrize("np_,kp", [(256, 64), (1024, 384), (512, 1024)])
@pytest.mark.parametrize("nprod", [1, 2, 3])
def test_gemm_layer_vs_fp64(m, np_, kp, nprod):
    from deepcubea_b200 import _lib
    lib = _lib.load()
    g = torch.Generator(device="cuda"); g.manual_seed(m + np_ + kp + nprod)
    a = torch.rand((m, kp), generator=g, device="cuda") * 4.0
    w = (torch.rand((np_, kp), generator=g, device="cuda") - 0.5) * 512.0
    bias = torch.randn(np_, generator=g, device="cuda")
    skip = torch.rand((m, np_), generator=g, device="cuda") * 2.0
    a_hi, a_lo = _split(a); w_hi, w_lo = _split(w); s_hi, s_lo = _split(skip)
    if nprod == 2:
        a = a_hi.float(); a_lo = None            # A exactly representable
    if nprod == 1:
        a = a_hi.float(); w = w_hi.float(); a_lo = None; w_lo = None
    scale = 2.0 ** -9
    out_hi = torch.empty((m, np_), dtype=torch.float16, device="cuda"); out_lo = torch.empty_like(out_hi)
    out_f = torch.empty((m, np_), dtype=torch.float32, device="cuda")
    p = _lib.ptr
    _lib.check(lib.dcb_resnet_gemm(p(a_hi), p(a_lo), kp, p(w_hi), p(w_lo), kp, p(bias), scale, p(s_hi), p(s_lo), 1, p(out_hi), p(out_lo), p(out_f),
                                   None, None, None, None, m, np_, kp, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = torch.relu((a.double() @ w.double().t()) * scale + bias.double() + s_hi.double() + s_lo.double())
    err = (out_f.double() - ref).abs().max().item()
    mag = ref.abs().max().item()
    # fp32 accumulation over K <= 1024 terms of magnitude ~ 4*256*2^-9 = 2: a few ulp of the partial sums
    tol = 2e-4 * max(mag, 1.0) if nprod == 3 else 6e-4 * max(mag, 1.0)
    assert err < tol, (err, mag)
    # hi + lo reproduces the fp32 output to 2^-21 relative
    rec = out_hi.float() + out_lo.float()
    assert ((rec - out_f).abs() <= 1e-6 * out_f.abs() + 1e-7).all()


def test_gemm_k_chunks_chain_through_partial_sums():
    """A K=2048 layer as two K=1024 launches chained through the fp32 partial sum == one launch (to fp32 rounding)."""
    from deepcubea_b200 import _lib
    lib = _lib.load(); p = _lib.ptr
    m, np_, kp = 777, 512, 2048
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    a = torch.rand((m, kp), generator=g, device="cuda"); w = (torch.rand((np_, kp), generator=g, device="cuda") - 0.5) * 4
    a_hi, a_lo = _split(a); w_hi, w_lo = _split(w)
    bias = torch.randn(np_, generator=g, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    part = torch.empty((m, np_), dtype=torch.float32, device="cuda")
    o_hi = torch.empty((m, np_), dtype=torch.float16, device="cuda"); o_f = torch.empty((m, np_), dtype=torch.float32, device="cuda")
    _lib.check(lib.dcb_resnet_gemm(p(a_hi), p(a_lo), kp, p(w_hi), p(w_lo), kp, None, 1.0, None, None, 0, None, None, None, None, p(part), None, None, m, np_, 1024, st))
    _lib.check(lib.dcb_resnet_gemm(a_hi.data_ptr() + 2048, a_lo.data_ptr() + 2048, kp, w_hi.data_ptr() + 2048, w_lo.data_ptr() + 2048, kp, p(bias), 0.5,
                                   None, None, 0, p(o_hi), None, p(o_f), p(part), None, None, None, m, np_, 1024, st))
    torch.cuda.synchronize()
    ref = (a.double() @ w.double().t()) * 0.5 + bias.double()
    assert (o_f.double() - ref).abs().max().item() < 2e-4


def _model(seed=0):
    from deepcubea_b200.utils.pytorch_models import ResnetModel
    torch.manual_seed(seed)
    m = ResnetModel(54, 6, 5000, 1000, 4, 1, True)
    w = os.path.join(ROOT, "assets", "saved_models", "cube3", "current", "model_state_dict.pt")
    trained = os.path.exists(w)
    if trained:
        sd = torch.load(w, map_location="cpu")
        m.load_state_dict({k.replace("module.", "", 1): v for k, v in sd.items()})
    else:
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm1d):
                mod.running_mean.normal_(0, 0.1); mod.running_var.uniform_(0.5, 1.5)
    return m.eval(), trained


def test_tc_network_matches_fp32_network():
    """fp16x3 mode: cost-to-go within 1e-4 of the fp32 PyTorch network (north-star tolerance)."""
    from deepcubea_b200.nnet.folded import FoldedResnet
    from deepcubea_b200.nnet.tc_resnet import TcResnet
    model, trained = _model()
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    x = torch.randint(0, 6, (5000, 54), generator=g, device="cuda", dtype=torch.uint8)
    x[0] = (torch.arange(54, device="cuda") // 9).to(torch.uint8)                  # the goal state's input
    import copy
    m64 = copy.deepcopy(model).double().to(dev)
    m64._encode = lambda t: torch.nn.functional.one_hot(t.long(), 6).double().flatten(1)
    with torch.no_grad():
        ref64 = m64(x)[:, 0]
    ref32 = FoldedResnet(model, "fp32").to(dev)(x).double()
    got3 = TcResnet(model, dev, "fp16x3")(x).double()
    got1 = TcResnet(model, dev, "fp16")(x).double()
    e32 = (ref32 - ref64).abs().max().item()
    e3 = (got3 - ref64).abs().max().item()
    e1 = (got1 - ref64).abs().max().item()
    print("max |err| vs fp64: torch fp32 %.3g, tc fp16x3 %.3g, tc fp16 %.3g (trained=%s, |out| max %.3g)" % (e32, e3, e1, trained, ref64.abs().max().item()))
    assert (got3 - ref32).abs().max().item() < 1e-4       # north-star tolerance
    assert e3 < 1e-4
    assert e1 < 0.05


def test_tc_network_golden_reference_values(golden_dir):
    """Cost-to-go of the REFERENCE network (trained weights, CPU fp32, tests/golden/nnet_cube3.npz)."""
    w = os.path.join(ROOT, "assets", "saved_models", "cube3", "current", "model_state_dict.pt")
    if not os.path.exists(w):
        pytest.skip("trained weights not present (assets/)")
    from deepcubea_b200.nnet.tc_resnet import TcResnet
    from oracle.oracle_env import OracleCube3
    g = np.load(golden_dir + "/nnet_cube3.npz")
    model, _ = _model()
    x = torch.from_numpy(OracleCube3().nnet_input(g["states"])).cuda()
    got = TcResnet(model, torch.device("cuda"), "fp16x3")(x).cpu().numpy()
    assert np.abs(got - g["ctg"]).max() < 1e-4


def test_eval_nodes_equals_gathered_input_path():
    """dcb_onehot_fp16_nodes (one-hot straight from the node arena) == dcb_gather_nnet_input + dcb_onehot_fp16."""
    import random
    from deepcubea_b200 import _lib
    from deepcubea_b200.nnet.tc_resnet import TcResnet
    from oracle.oracle_env import OracleCube3
    lib = _lib.load(); p = _lib.ptr
    env = OracleCube3()
    np.random.seed(5); random.seed(5)
    states, _ = env.generate_states(3000, (0, 20))
    arena = torch.from_numpy(np.concatenate([states.reshape(-1), np.zeros(64, np.uint8)])).cuda()
    ids = torch.from_numpy(np.random.permutation(3000)[:1777].astype(np.int32)).cuda()
    st = torch.cuda.current_stream().cuda_stream
    a = torch.empty((1777, 384), dtype=torch.float16, device="cuda"); b = torch.empty_like(a)
    _lib.check(lib.dcb_onehot_fp16_nodes(0, p(arena), p(ids), 1777, 6, 384, p(a), st))
    x = torch.from_numpy(env.nnet_input(states[ids.cpu().numpy()])).cuda()
    _lib.check(lib.dcb_onehot_fp16(p(x), 1777, 54, 6, 384, p(b), st))
    assert torch.equal(a, b)
    ref = torch.nn.functional.one_hot(x.long(), 6).flatten(1).half()
    assert torch.equal(a[:, :324], ref) and float(a[:, 324:].abs().max()) == 0.0
    model, _ = _model()
    tc = TcResnet(model, torch.device("cuda"), "fp16x3")
    assert torch.equal(tc.eval_nodes(0, arena, ids, 1777), tc(x))


def test_fused_fc_out_dot_partials():
    """dcb_resnet_gemm with d_dot_w: per-tile partial dot products of the layer output == (relu(A W^T + b + skip)) . w."""
    from deepcubea_b200 import _lib
    lib = _lib.load(); p = _lib.ptr
    m, np_, kp = 1000, 1024, 1024
    g = torch.Generator(device="cuda"); g.manual_seed(11)
    a = torch.rand((m, kp), generator=g, device="cuda"); w = (torch.rand((np_, kp), generator=g, device="cuda") - 0.5) * 2
    bias = torch.randn(np_, generator=g, device="cuda"); skip = torch.rand((m, np_), generator=g, device="cuda")
    wd = torch.randn(np_, generator=g, device="cuda"); wd[1000:] = 0
    a_hi, a_lo = _split(a); w_hi, w_lo = _split(w); s_hi, s_lo = _split(skip)
    part = torch.zeros((m, np_ // 256), dtype=torch.float32, device="cuda")
    _lib.check(lib.dcb_resnet_gemm(p(a_hi), p(a_lo), kp, p(w_hi), p(w_lo), kp, p(bias), 1.0, p(s_hi), p(s_lo), 1, None, None, None, None, None,
                                   p(wd), p(part), m, np_, kp, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    out = torch.relu(a.double() @ w.double().t() + bias.double() + s_hi.double() + s_lo.double())
    ref = (out.view(m, 4, 256) * wd.double().view(1, 4, 256)).sum(dim=2)
    assert (part.double() - ref).abs().max().item() < 5e-4 * max(1.0, ref.abs().max().item())
