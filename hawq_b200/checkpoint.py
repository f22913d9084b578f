"""HAWQ checkpoint formats in and out of the engine.

Three formats of the reference, each handled the way the reference itself handles it:

* ``checkpoint.pth.tar`` written during quantization-aware training and read back by ``--resume --resume-quantize``
  (``quant_train.py:304-318``): a ``state_dict`` with DataParallel ``module.`` prefixes.  :func:`load_quantized_checkpoint`
  applies the reference's key filter (drops ``num_batches_tracked`` / ``weight_integer`` / ``bias_integer``, strips the
  prefix, ``strict=False``); the engine then derives its integers from the float weights + BN statistics + activation ranges
  exactly like the reference's frozen forward does.
* ``quantized_checkpoint.pth.tar`` written after validation (``quant_train.py:665-670``): five dictionaries holding only the
  integers and scales (``weight_integer``, ``bias_integer``, ``convbn_scaling_factor``, ``fc_scaling_factor``,
  ``act_scaling_factor``).  :func:`save_quantized_checkpoint` writes it from a model that has run a frozen forward;
  :func:`apply_integer_checkpoint` turns it back into an engine plan on a model skeleton whose float weights are irrelevant
  (no BN fold, no re-quantisation: the stored integers ARE the plan).
* the TVM deployment parameters ``weights.npy`` / ``bias.npy`` (``tvm_benchmark/hawq_utils_resnet50.py:111-153,334-368``):
  HWOI kernels, int8 or eight 4-bit values per int32 (first value in the top nibble, ``:21-30``), renamed
  ``stage%d_unit%d_qconv%d_weight`` / ``..._qsc_weight`` / ``conv0_weight`` / ``fc_weight``.  :func:`export_tvm_params`.
"""
import os

import numpy as np
import torch

from .modules import QuantAct, QuantBnConv2d, QuantConv2d, QuantLinear

_GROUPS = ("convbn_scaling_factor", "fc_scaling_factor", "weight_integer", "bias_integer", "act_scaling_factor")


# ------------------------------------------------------------------------------------------------ checkpoint.pth.tar
def filter_resume_state_dict(state_dict):
    """The key filter of ``quant_train.py:307-314``."""
    out = {}
    for key, value in state_dict.items():
        if "num_batches_tracked" in key or "weight_integer" in key or "bias_integer" in key:
            continue
        out[key.replace("module.", "")] = value
    return out


def load_quantized_checkpoint(model, checkpoint, map_location="cpu"):
    """``--resume-quantize``: load a QAT checkpoint (path or already-loaded dict) into a quantized graph.
    Returns the ``load_state_dict`` result (missing / unexpected keys), ``strict=False`` like the reference."""
    if isinstance(checkpoint, (str, os.PathLike)):
        checkpoint = torch.load(checkpoint, map_location=map_location, weights_only=False)
    state = checkpoint["state_dict"] if "state_dict" in checkpoint else checkpoint
    return model.load_state_dict(filter_resume_state_dict(state), strict=False)


# ------------------------------------------------------------------------------------------------ quantized_checkpoint.pth.tar
def quantized_checkpoint_dict(model, prefix="module."):
    """The dictionary ``quant_train.py:665-670`` saves (keys carry the DataParallel prefix the reference's model has)."""
    sd = model.state_dict()
    return {g: {prefix + k: v.detach().clone() for k, v in sd.items() if g in k} for g in _GROUPS}


def save_quantized_checkpoint(model, path, prefix="module."):
    """Write ``quantized_checkpoint.pth.tar``.  Like in the reference the integer buffers are a by-product of a frozen
    forward: call this after one (``validate()`` in the reference, any frozen forward here)."""
    ck = quantized_checkpoint_dict(model, prefix)
    for name, m in model.named_modules():
        if isinstance(m, (QuantBnConv2d, QuantConv2d, QuantLinear)) and not bool(torch.count_nonzero(m.weight_integer)):
            raise RuntimeError("save_quantized_checkpoint: %s.weight_integer is empty: run a frozen forward first" % name)
    torch.save(ck, path)
    return ck


def _strip(d, prefix="module."):
    return {(k[len(prefix):] if k.startswith(prefix) else k): v for k, v in d.items()}


def apply_integer_checkpoint(model, checkpoint, map_location="cpu"):
    """Install the integers and scales of a ``quantized_checkpoint.pth.tar`` as the frozen plan of ``model`` (a quantized graph
    of the same architecture and bit configuration; its float weights are not used) and freeze it.
    Returns the number of quant modules configured."""
    if isinstance(checkpoint, (str, os.PathLike)):
        checkpoint = torch.load(checkpoint, map_location=map_location, weights_only=False)
    missing = [g for g in _GROUPS if g not in checkpoint]
    if missing:
        raise KeyError("not a quantized_checkpoint: missing %s" % missing)
    w_int = _strip(checkpoint["weight_integer"])
    b_int = _strip(checkpoint["bias_integer"])
    w_sf = {**_strip(checkpoint["convbn_scaling_factor"]), **_strip(checkpoint["fc_scaling_factor"])}
    a_sf = _strip(checkpoint["act_scaling_factor"])
    done = 0
    for name, m in model.named_modules():
        if isinstance(m, QuantAct):
            key = name + ".act_scaling_factor"
            if key not in a_sf:
                raise KeyError("quantized checkpoint has no %s" % key)
            m.load_frozen_scale(a_sf[key])
            done += 1
        elif isinstance(m, (QuantBnConv2d, QuantConv2d, QuantLinear)):
            sf_name = {QuantBnConv2d: "convbn_scaling_factor", QuantConv2d: "conv_scaling_factor", QuantLinear: "fc_scaling_factor"}[type(m)]
            kw, kb, ks = name + ".weight_integer", name + ".bias_integer", name + "." + sf_name
            if kw not in w_int or ks not in w_sf:
                raise KeyError("quantized checkpoint has no %s / %s" % (kw, ks))
            m.load_frozen_integers(w_sf[ks], w_int[kw], b_int.get(kb))
            done += 1
    from .modules import freeze_model
    freeze_model(model)
    return done


# ------------------------------------------------------------------------------------------------ TVM parameter files
def pack_int4_tvm(a):
    """``pack_int32_to_int4`` (hawq_utils_resnet50.py:21-30), vectorised: [I,J,K,L] -> int32 [I,J,K,L//8], value m of each
    group of eight in bits (7-m)*4 .. (7-m)*4+3 (two's-complement nibbles)."""
    a = np.asarray(a).astype(np.int64)
    i, j, k, ln = a.shape
    groups = ln // 8
    v = (a[..., :groups * 8].reshape(i, j, k, groups, 8) & 0xF).astype(np.uint64)
    shifts = ((7 - np.arange(8)) * 4).astype(np.uint64)
    packed = np.bitwise_or.reduce(v << shifts, axis=-1) if groups else np.zeros((i, j, k, 0), dtype=np.uint64)
    return packed.astype(np.uint32).view(np.int32).reshape(i, j, k, groups)


def unpack_int4_tvm(p):
    """``unpack_int4_to_int32`` (hawq_utils_resnet50.py:32-42): inverse of :func:`pack_int4_tvm`, unsigned nibbles 0..15."""
    p = np.asarray(p).astype(np.int32).view(np.uint32).astype(np.uint64)
    shifts = ((7 - np.arange(8)) * 4).astype(np.uint64)
    out = (p[..., None] >> shifts) & np.uint64(0xF)
    return out.reshape(*p.shape[:-1], p.shape[-1] * 8).astype(np.int32)


def _np(t):
    return t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)


def export_tvm_params(checkpoint, kernel_dtype="int8", num_stages=4, units=(3, 4, 6, 3), convs_per_unit=3):
    """(weights, bias) dictionaries of ``save_weights`` / ``save_bias`` (hawq_utils_resnet50.py:111-153, 334-368) from a
    ``quantized_checkpoint`` dictionary.  ``kernel_dtype``: 'int8' | 'int4' (uniform, as the reference), or a mapping from the
    reference's parameter name (e.g. ``stage2_unit1_qconv3_weight``) to one of them for mixed-precision models; the first
    convolution and the classifier are always int8."""
    w_in = _strip(checkpoint["weight_integer"])
    b_in = _strip(checkpoint["bias_integer"])

    def dtype_of(new_name):
        if new_name in ("conv0_weight", "fc_weight"):
            return "int8"
        return kernel_dtype if isinstance(kernel_dtype, str) else kernel_dtype[new_name]

    def weight(old, new):
        if old not in w_in:
            raise KeyError("%s is not in the params" % old)
        t = _np(w_in[old]).astype(np.int32)
        if t.ndim == 4:
            t = np.transpose(t, (2, 3, 0, 1))                       # OIHW -> HWOI
        dt = dtype_of(new)
        if dt == "int4":
            return pack_int4_tvm(t)
        if dt != "int8":
            raise ValueError("kernel dtype %r not supported" % (dt,))
        return t.astype(np.int8)

    def bias(old):
        if old not in b_in:
            raise KeyError("%s is not in the params" % old)
        return _np(b_in[old]).reshape(1, 1, 1, -1).astype(np.int32)

    weights = {"conv0_weight": weight("quant_init_convbn.weight_integer", "conv0_weight")}
    biases = {"conv0_bias": bias("quant_init_convbn.bias_integer")}
    for i in range(num_stages):
        for j in range(units[i]):
            for k in range(convs_per_unit):
                base = "stage%d.unit%d.quant_convbn%d" % (i + 1, j + 1, k + 1)
                new = "stage%d_unit%d_qconv%d" % (i + 1, j + 1, k + 1)
                weights[new + "_weight"] = weight(base + ".weight_integer", new + "_weight")
                biases[new + "_bias"] = bias(base + ".bias_integer")
            base = "stage%d.unit%d.quant_identity_convbn" % (i + 1, j + 1)
            if j == 0 and (convs_per_unit == 3 or base + ".weight_integer" in w_in):
                new = "stage%d_unit%d_qsc" % (i + 1, j + 1)
                weights[new + "_weight"] = weight(base + ".weight_integer", new + "_weight")
                biases[new + "_bias"] = bias(base + ".bias_integer")
    weights["fc_weight"] = weight("quant_output.weight_integer", "fc_weight")
    biases["fc_bias"] = bias("quant_output.bias_integer")[0, 0, 0, :]
    return weights, biases


def save_tvm_params(checkpoint, save_path, **kw):
    """Write ``weights.npy`` and ``bias.npy`` (pickled dictionaries, as the reference's ``np.save`` of a dict)."""
    weights, biases = export_tvm_params(checkpoint, **kw)
    np.save(os.path.join(save_path, "weights.npy"), weights)
    np.save(os.path.join(save_path, "bias.npy"), biases)
    return weights, biases
