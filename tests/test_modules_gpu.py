"""-m gpu: module-level coverage on the CUDA path that the network tests do not reach — QuantConv2d (with and without bias,
reference quant_modules.py:605-736) and the HAWQ checkpoint formats loaded into the CUDA engine (quant_train.py:304-318,665-670)."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import hawq_b200 as hb
from hawq_b200 import checkpoint as ck
from hawq_b200.synthetic import synthetic_batch, synthetic_float_resnet
from tests import abi_model

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _ConvNet(nn.Module):
    """quant_input -> QuantConv2d 3x3 (bias) -> ReLU -> QuantAct -> QuantConv2d 1x1 (bias=None) -> ReLU -> QuantAct"""

    def __init__(self, a_bits):
        super().__init__()
        g = torch.Generator().manual_seed(5)
        c1, c2 = nn.Conv2d(64, 128, 3, 1, 1, bias=True), nn.Conv2d(128, 64, 1, 1, 0, bias=False)
        with torch.no_grad():
            for c in (c1, c2):
                c.weight.copy_(torch.randn(c.weight.shape, generator=g) * 0.05)
            c1.bias.copy_(torch.randn(128, generator=g) * 0.1)
        mode = "asymmetric" if a_bits == 4 else "symmetric"
        self.quant_input = hb.QuantAct(activation_bit=8, act_range_momentum=0.99)
        self.conv1 = hb.QuantConv2d(weight_bit=a_bits, bias_bit=32, per_channel=True)
        self.conv1.set_param(c1)
        self.act1 = hb.QuantAct(activation_bit=a_bits, quant_mode=mode, act_range_momentum=0.99)
        self.conv2 = hb.QuantConv2d(weight_bit=a_bits, bias_bit=32, per_channel=True)
        self.conv2.set_param(c2)
        self.act2 = hb.QuantAct(activation_bit=8, act_range_momentum=0.99)

    def forward(self, x):
        x, sf = self.quant_input(x)
        x, w_sf = self.conv1(x, sf)
        x, sf = self.act1(F.relu(x), sf, w_sf)
        x, w_sf = self.conv2(x, sf)
        x, sf = self.act2(F.relu(x), sf, w_sf)
        return x, sf


@pytest.mark.parametrize("a_bits", [8, 4])
def test_quantconv2d_frozen_cuda_equals_abi_model(a_bits, monkeypatch):
    """A graph of QuantConv2d modules (bias and bias=None) frozen on the CUDA kernels gives the integers of the numpy ABI model
    (tests/abi_model.py -> oracle/int_ref.py) run through the same host logic, and agrees with its own un-frozen float emulation
    (the reference arithmetic) on the calibration batch."""
    torch.manual_seed(0)
    x = torch.randn(3, 64, 14, 14)
    net = _ConvNet(a_bits).eval()
    with torch.no_grad():
        f_out, f_sf = net(x)                             # calibration / float emulation (un-frozen)
    hb.freeze_model(net)
    with torch.no_grad():
        g_out, g_sf = net(x.to(DEV))
        got = g_out.int_tensor().cpu().numpy()
    torch.cuda.synchronize()
    # same frozen graph on the numpy ABI model
    net_c = _ConvNet(a_bits).eval()
    net_c.load_state_dict(net.state_dict())
    hb.freeze_model(net_c)
    abi_model.install_cpu_backend(monkeypatch)
    with torch.no_grad():
        c_out, c_sf = net_c(x)
        want = c_out.int_tensor().numpy()
    assert np.array_equal(got, want)
    assert torch.equal(g_sf.cpu().reshape(-1), c_sf.reshape(-1))
    # un-frozen float emulation of the same modules (frozen ranges == calibration ranges of this batch after one update)
    assert np.array_equal(np.round((f_out / f_sf.view(-1)).numpy()).astype(np.int64), want.astype(np.int64))


def test_quantized_checkpoint_into_the_cuda_engine(tmp_path):
    """quantized_checkpoint.pth.tar written from a frozen CUDA forward becomes the whole plan of a skeleton with different float
    weights; the CUDA-graph engine reproduces the logits bit for bit.  The --resume-quantize path (float weights + ranges through
    load_state_dict) rebuilds the plan on the GPU."""
    x = synthetic_batch(4, 3)
    qa = hb.build_synthetic_qresnet("resnet18", "uniform4", calib_batch=2)
    with torch.no_grad():
        la = qa(x.to(DEV))
    path = os.path.join(tmp_path, "quantized_checkpoint.pth.tar")
    ck.save_quantized_checkpoint(qa, path)

    qb = hb.q_resnet18(synthetic_float_resnet("resnet18", 123))
    hb.stamp_bit_config(qb, hb.get_bit_config("resnet18", "uniform4"))
    qb.eval()
    assert ck.apply_integer_checkpoint(qb, path) == 21 + 27
    eng = hb.compile_model(qb, x.to(DEV))
    assert torch.equal(eng(x.to(DEV)), la)

    qc = hb.q_resnet18(synthetic_float_resnet("resnet18", 123))
    hb.stamp_bit_config(qc, hb.get_bit_config("resnet18", "uniform4"))
    qc.eval()
    with torch.no_grad():
        qc(x)                                             # calibration forward (float path, CPU)
    hb.freeze_model(qc)
    with torch.no_grad():
        assert not torch.equal(qc(x.to(DEV)), la)         # plan built from the seed-123 weights
    res = ck.load_quantized_checkpoint(qc, {"state_dict": {"module." + k: v for k, v in qa.state_dict().items()}})
    assert not res.unexpected_keys
    with torch.no_grad():
        assert torch.equal(qc(x.to(DEV)), la)             # load_state_dict wrote in place: the cached GPU plan is not reused


def test_unfrozen_quantact_runs_on_cuda():
    """Calibration after model.cuda() (the reference's normal workflow, quant_train.py): the un-frozen float emulation of
    QuantAct / QuantBnConv2d runs on CUDA tensors and matches the CPU result."""
    net_c = _ConvNet(8).eval()
    net_g = _ConvNet(8).eval().to(DEV)
    torch.manual_seed(1)
    x = torch.randn(2, 64, 8, 8)
    with torch.no_grad():
        oc, sc = net_c(x)
        og, sg = net_g(x.to(DEV))
    assert torch.allclose(sg.cpu(), sc, rtol=1e-6) and torch.equal(torch.round(og.cpu() / sg.cpu().view(-1)), torch.round(oc / sc.view(-1)))
