"""Bit-allocation tables and the attribute-stamping contract.

The tables are the reference's ``bit_config_dict`` entries for the ResNet
family (reference ``bit_config.py:3-3054``) and for MobileNetV2 (``:3602-4202``), carried as data in
``data/bit_configs_resnet.json`` / ``data/bit_configs_mobilenetv2.json`` (module name -> 4 / 8 / 16, optional 'hook').

``stamp_bit_config`` reproduces how the reference trainer writes those numbers
onto the quant modules by plain ``setattr`` (reference ``quant_train.py:264-299``):
everything is symmetric/per-channel/bias_bit=32 except that 4-bit *activations*
are switched to 'asymmetric' (unsigned 0..15, no zero point).
"""
import json
import os

_DATA = [os.path.join(os.path.dirname(__file__), "data", n) for n in ("bit_configs_resnet.json", "bit_configs_mobilenetv2.json")]
_cache = None


def bit_config_dict():
    """{"bit_config_<arch>_<scheme>": {module_name: bits | (bits, 'hook')}} (insertion ordered)."""
    global _cache
    if _cache is None:
        raw = {}
        for path in _DATA:
            with open(path) as f:
                raw.update(json.load(f))
        out = {}
        for key, entries in raw.items():
            d = {}
            for e in entries:
                d[e[0]] = e[1] if len(e) == 2 else (e[1], e[2])
            out[key] = d
        _cache = out
    return _cache


def get_bit_config(arch, scheme):
    key = "bit_config_%s_%s" % (arch, scheme)
    table = bit_config_dict()
    if key not in table:
        raise KeyError("unknown bit config %r (have: %s)" % (key, ", ".join(sorted(table))))
    return table[key]


def stamp_bit_config(model, bit_config, bias_bit=32, channel_wise=True, act_percentile=0,
                     act_range_momentum=0.99, weight_percentile=0, fix_BN=True,
                     fix_BN_threshold=None, fixed_point_quantization=False):
    """Write the quantisation attributes onto every module named in ``bit_config``.

    Defaults are the reference CLI defaults that matter on the forward path
    (``quant_train.py:104-151``: --bias-bit 32, channel-wise on, momentum 0.99) with
    ``fix_BN`` on, which is what every published HAWQ-V3 ResNet uses.
    Returns the number of matched modules (the reference logs whether it equals len(bit_config)).
    """
    matched = 0
    for name, m in model.named_modules():
        if name not in bit_config:
            continue
        matched += 1
        m.quant_mode = "symmetric"
        m.bias_bit = bias_bit
        m.quantize_bias = (bias_bit != 0)
        m.per_channel = channel_wise
        m.act_percentile = act_percentile
        m.act_range_momentum = act_range_momentum
        m.weight_percentile = weight_percentile
        m.fix_flag = False
        m.fix_BN = fix_BN
        m.fix_BN_threshold = fix_BN_threshold
        m.training_BN_mode = fix_BN
        m.fixed_point_quantization = fixed_point_quantization
        v = bit_config[name]
        bits = v[0] if isinstance(v, tuple) else v
        if hasattr(m, "activation_bit"):
            m.activation_bit = bits
            if bits == 4:
                m.quant_mode = "asymmetric"
        else:
            m.weight_bit = bits
    return matched
