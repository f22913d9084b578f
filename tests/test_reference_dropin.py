"""Drop-in claim, checked in the build container (skipped where /root/reference is absent): the reference's UNMODIFIED graph
code (utils/models/q_resnet.py: Q_ResNet50 / Q_ResUnitBn forward with nn.ReLU, nn.MaxPool2d, `x + identity`, x.view) runs on
this package's quant modules, both un-frozen (calibration, equal ranges) and frozen (integer path through IntActivation
payloads; C-ABI launchers replaced by the numpy ABI model on CPU) and reproduces the reference's logits bit for bit."""
import numpy as np
import pytest
import torch

import hawq_b200 as hb
from hawq_b200.build import build_library
from hawq_b200.synthetic import synthetic_batch, synthetic_float_resnet
from oracle import ref_harness as rh
from tests import abi_model
from tests.util import load_net_golden

pytestmark = pytest.mark.skipif(not rh.available(), reason="reference checkout not present")


@pytest.mark.parametrize("arch,scheme", [("resnet18", "uniform4"), ("resnet50", "bops_0.5")])
def test_reference_graph_code_on_our_modules(monkeypatch, arch, scheme):
    build_library()
    ns = rh.load()
    qr = ns.q_resnet
    # the reference graph file did `from ..quant_modules import *`: rebind those names to this package's classes
    for name in ("QuantAct", "QuantBnConv2d", "QuantLinear", "QuantAveragePool2d", "QuantConv2d"):
        monkeypatch.setattr(qr, name, getattr(hb, name))
    logits_g, meta = load_net_golden(arch, scheme)
    net = synthetic_float_resnet(arch, 0)
    q = {"resnet18": qr.q_resnet18, "resnet50": qr.q_resnet50}[arch](net)        # reference Q_ResNet*, our modules inside
    assert type(q).__module__.startswith("utils.models")
    assert isinstance(q.quant_input, hb.QuantAct)
    rh.stamp_like_quant_train(q, ns.bit_config_dict["bit_config_%s_%s" % (arch, scheme)])
    q.eval()
    with torch.no_grad():
        q(synthetic_batch(*meta["calib"]))                                     # un-frozen: float calibration pass
    for name, m in q.named_modules():
        if isinstance(m, hb.QuantAct):
            assert float(m.x_min) == meta["acts"][name]["x_min"] and float(m.x_max) == meta["acts"][name]["x_max"], name
    hb.freeze_model(q)                                                        # reference-style traversal, our implementation
    abi_model.install_cpu_backend(monkeypatch)
    with torch.no_grad():
        out = q(synthetic_batch(*meta["input"]))
    assert np.array_equal(out.numpy(), logits_g)


def test_reference_mobilenetv2_graph_code_on_our_modules_unfrozen(monkeypatch):
    """utils/models/q_mobilenetv2.py (Q_LinearBottleneck / Q_MobileNetV2: nn.ReLU6, depthwise QuantBnConv2d, `x + identity` without
    ReLU, QuantConv2d classifier) built on this package's modules: the calibration forward gives the reference's ranges and the
    same fp32 outputs as the reference's own modules.  (The frozen integer path of this family is not built: DESIGN.md row f3.)"""
    import json
    import os
    from hawq_b200.synthetic import synthetic_float_mobilenetv2
    ns = rh.load()
    qmb = ns.q_mobilenetv2
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "net_mobilenetv2_w1_uniform8.npz"))
    meta = json.loads(str(z["meta"]))
    x = synthetic_batch(*meta["calib"])
    ref = rh.build_reference_qmobilenetv2("uniform8", synthetic_float_mobilenetv2(0), x, freeze=False)     # reference modules
    with torch.no_grad():
        want = ref(synthetic_batch(*meta["input"]))          # second un-frozen forward (ranges keep moving, like ours below)
    for name in ("QuantAct", "QuantBnConv2d", "QuantLinear", "QuantAveragePool2d", "QuantConv2d"):
        monkeypatch.setattr(qmb, name, getattr(hb, name))
    q = qmb.q_mobilenetv2_w1(synthetic_float_mobilenetv2(0))
    assert type(q).__module__.startswith("utils.models") and isinstance(q.features.stage2.unit1.conv2, hb.QuantBnConv2d)
    rh.stamp_like_quant_train(q, ns.bit_config_dict["bit_config_mobilenetv2_w1_uniform8"])
    q.eval()
    with torch.no_grad():
        q(x)
    for name, m in q.named_modules():
        if isinstance(m, hb.QuantAct):
            assert float(m.x_min) == meta["acts"][name]["x_min"] and float(m.x_max) == meta["acts"][name]["x_max"], name
    with torch.no_grad():
        got = q(synthetic_batch(*meta["input"]))
    assert torch.equal(got, want)
