"""TEST INFRASTRUCTURE: drive the UNMODIFIED reference (/root/reference) on CPU.

Only usable in the build container (the GPU box has no /root/reference).  Used by
``tests/golden/make_golden.py`` to produce the committed golden vectors and by a few
``-m "not gpu"`` tests (skipped when the reference is absent) that pin the restatements
in ``oracle/`` and the host logic in ``hawq_b200`` directly against the reference.

Shims (the reference itself is never edited):
  * dummy ``pytorchcv`` modules, because ``utils/__init__.py:1-4`` imports every model file and
    ``utils/models/q_resnet.py:10-11`` imports names from pytorchcv that ResNets never use;
  * ``torch.Tensor.cuda`` -> identity on GPU-less hosts: the reference hard-codes ``.cuda()`` at
    ``utils/quantization_utils/quant_utils.py:212-213,251,299``.
"""
import importlib.util
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("HAWQ_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "utils", "quantization_utils", "quant_modules.py"))


_loaded = {}


def load():
    """Import the reference; returns a namespace with quant_modules, quant_utils, q_resnet, bit_config_dict."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    for n in ("pytorchcv", "pytorchcv.models", "pytorchcv.models.common", "pytorchcv.models.shufflenetv2"):
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules["pytorchcv.models.common"].ConvBlock = object
    sys.modules["pytorchcv.models.shufflenetv2"].ShuffleUnit = object
    sys.modules["pytorchcv.models.shufflenetv2"].ShuffleInitBlock = object
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    # the reference's top-level package is called "utils"; make sure nothing else shadows it
    for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
        del sys.modules[k]
    import utils.quantization_utils.quant_modules as qm  # noqa
    import utils.quantization_utils.quant_utils as qu  # noqa
    import utils.models.q_resnet as qr  # noqa
    import utils.models.q_mobilenetv2 as qmb  # noqa
    spec = importlib.util.spec_from_file_location("_ref_bit_config", os.path.join(REF_ROOT, "bit_config.py"))
    bc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bc)
    ns = types.SimpleNamespace(quant_modules=qm, quant_utils=qu, q_resnet=qr, q_mobilenetv2=qmb,
                               bit_config_dict=bc.bit_config_dict)
    _loaded["ns"] = ns
    return ns


def stamp_like_quant_train(model, bit_config):
    """Attribute stamping exactly as reference quant_train.py:267-299 with the CLI defaults
    (--bias-bit 32, channel-wise, act-range-momentum 0.99, percentiles 0, --fix-BN)."""
    n = 0
    for name, m in model.named_modules():
        if name in bit_config.keys():
            n += 1
            setattr(m, 'quant_mode', 'symmetric')
            setattr(m, 'bias_bit', 32)
            setattr(m, 'quantize_bias', True)
            setattr(m, 'per_channel', True)
            setattr(m, 'act_percentile', 0)
            setattr(m, 'act_range_momentum', 0.99)
            setattr(m, 'weight_percentile', 0)
            setattr(m, 'fix_flag', False)
            setattr(m, 'fix_BN', True)
            setattr(m, 'fix_BN_threshold', None)
            setattr(m, 'training_BN_mode', True)
            setattr(m, 'fixed_point_quantization', False)
            v = bit_config[name]
            bitwidth = v[0] if type(v) is tuple else v
            if hasattr(m, 'activation_bit'):
                setattr(m, 'activation_bit', bitwidth)
                if bitwidth == 4:
                    setattr(m, 'quant_mode', 'asymmetric')
            else:
                setattr(m, 'weight_bit', bitwidth)
    assert n == len(bit_config), (n, len(bit_config))


def build_reference_qresnet(arch, scheme, float_model, calib):
    """q_resnetXX(float skeleton) -> stamp -> one calibration forward (running_stat) -> freeze."""
    ns = load()
    ctor = {"resnet18": ns.q_resnet.q_resnet18, "resnet50": ns.q_resnet.q_resnet50,
            "resnet101": ns.q_resnet.q_resnet101}[arch]
    q = ctor(float_model)
    stamp_like_quant_train(q, ns.bit_config_dict["bit_config_%s_%s" % (arch, scheme)])
    q.eval()
    with torch.no_grad():
        q(calib)
    ns.quant_modules.freeze_model(q)
    return q


def build_reference_qmobilenetv2(scheme, float_model, calib, freeze=True):
    """q_mobilenetv2_w1(float skeleton) -> stamp -> one calibration forward (running_stat) -> freeze (reference quant_train.py flow)."""
    ns = load()
    q = ns.q_mobilenetv2.q_mobilenetv2_w1(float_model)
    stamp_like_quant_train(q, ns.bit_config_dict["bit_config_mobilenetv2_w1_%s" % scheme])
    q.eval()
    with torch.no_grad():
        q(calib)
    if freeze:
        ns.quant_modules.freeze_model(q)
    return q


def run_with_act_hooks(q, x):
    """Frozen forward; returns (logits, {QuantAct name: int64 activation integers round(out/scale)})."""
    ns = load()
    acts = {}
    hooks = []
    for name, m in q.named_modules():
        if type(m) is ns.quant_modules.QuantAct:
            def hook(mod, inp, out, name=name):
                acts[name] = torch.round(out[0] / out[1].view(-1)).to(torch.int64)
            hooks.append(m.register_forward_hook(hook))
    with torch.no_grad():
        logits = q(x)
    for h in hooks:
        h.remove()
    return logits, acts
