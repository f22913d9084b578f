"""Host-side quantisation math: scale computation, weight/bias integerisation, BN folding, and the
float ("fake-quant") arithmetic used ONLY while a module is un-frozen (range calibration / QAT-style forward).

These are the freeze-time and calibration-time counterparts of reference
``utils/quantization_utils/quant_utils.py`` (file:line cited per function).  They run in torch fp32/fp64 exactly
like the reference so that the integers handed to the CUDA engine (weight_integer, bias_integer, scales, dyadic
pairs) are bit-identical to what the reference would compute.  The frozen inference path never calls the float
arithmetic here: it runs on the CUDA kernels behind the C ABI.
"""
import math

import torch


def symmetric_scale(bits, lo, hi, per_channel=False):
    """scale = clamp(max(|lo|, |hi|), 1e-8) / (2^(bits-1) - 1)      (quant_utils.py:128-152)."""
    with torch.no_grad():
        n = 2 ** (bits - 1) - 1
        if per_channel:
            mag = torch.stack([lo.abs(), hi.abs()], dim=1).max(dim=1).values
        else:
            mag = max(lo.abs(), hi.abs())
        return torch.clamp(mag, min=1e-8) / n


def asymmetric_scale(bits, lo, hi):
    """scale = clamp(hi - lo, 1e-8) / (2^bits - 1); HAWQ never applies the zero point (quant_utils.py:155-185,
    quant_modules.py:265-270), values live in [0, 2^bits - 1]."""
    with torch.no_grad():
        return torch.clamp(hi - lo, min=1e-8) / float(2 ** bits - 1)


def _like(scale, x):
    if x.dim() == 4:
        return scale.view(-1, 1, 1, 1)
    if x.dim() == 2:
        return scale.view(-1, 1)
    return scale.view(-1)


def quantize(x, bits, scale, signed=True):
    """clamp(round((1/scale) * x))  — SymmetricQuantFunction / AsymmetricQuantFunction forward
    (quant_utils.py:231-258, 275-308 with linear_quantize :73-97).  Signed range includes -2^(bits-1)."""
    with torch.no_grad():
        q = torch.round(1. / _like(scale, x) * x)
        if signed:
            n = 2 ** (bits - 1) - 1
            return torch.clamp(q, -n - 1, n)
        return torch.clamp(q, 0, 2 ** bits - 1)


def clamp_range(bits, mode):
    if mode == "symmetric":
        return -(2 ** (bits - 1)), 2 ** (bits - 1) - 1
    return 0, 2 ** bits - 1


def fold_bn(conv, bn):
    """Folded weight / bias of conv+BN with running statistics (quant_modules.py:441-449)."""
    std = torch.sqrt(bn.running_var.detach() + bn.eps)
    factor = bn.weight / std
    w = conv.weight * factor.reshape([conv.out_channels, 1, 1, 1])
    b = conv.bias if conv.bias is not None else torch.zeros_like(bn.running_mean)
    b = (b - bn.running_mean.detach()) * factor + bn.bias
    return w, b


def per_channel_minmax(w2d, percentile=0):
    """Row-wise weight range; percentile branch as quant_modules.py:455-467."""
    if percentile == 0:
        return w2d.min(dim=1).values, w2d.max(dim=1).values
    n = w2d.shape[1]
    lo_idx = math.ceil(n * (100 - percentile) * 0.01)
    hi_idx = math.ceil(n * percentile * 0.01)
    return torch.kthvalue(w2d, k=lo_idx, dim=1).values, torch.kthvalue(w2d, k=hi_idx, dim=1).values


def percentile_minmax(flat, lower_percentile, upper_percentile):
    """get_percentile_min_max (quant_utils.py:40-70) on a 1-D tensor, tensors out."""
    n = flat.shape[0]
    lower_index = round(n * (1 - lower_percentile * 0.01))
    upper_index = round(n * upper_percentile * 0.01)
    hi = torch.kthvalue(flat, k=upper_index).values
    lo = hi * 0 if lower_percentile == 0 else -torch.kthvalue(-flat, k=lower_index).values
    return lo, hi


def requant_ratio(a_sf, w_sf, out_sf):
    """fp64(fp32(fp64(a)*fp64(w))) / fp64(fp32(out))   (quant_utils.py:394-397)."""
    prod = a_sf.type(torch.double) * w_sf.type(torch.double)
    return prod.type(torch.float).type(torch.double) / out_sf.type(torch.float).type(torch.double)


def dyadic_pairs(ratios):
    """batch_frexp (quant_utils.py:188-213) for a 1-D double tensor -> (list m, list e) via the C library's host
    helper (frexp, mantissa * 2^31 rounded half-up, e = 31 - exp)."""
    from . import _lib
    out = [_lib.dyadic(float(r)) for r in ratios.reshape(-1).tolist()]
    return [m for m, _ in out], [e for _, e in out]


# ---------------------------------------------------------------------------------------------------------
# float emulation, used only while un-frozen (calibration): same arithmetic as fixedpoint_fn
# ---------------------------------------------------------------------------------------------------------
def _rs(t, z):
    return t.view(1, -1, 1, 1) if z.dim() == 4 else t.view(1, -1)


def _float_requant(z_int, a_sf, w_sf, out_sf, z):
    import numpy as np
    ratio = _rs(requant_ratio(a_sf, w_sf, out_sf), z)
    # batch_frexp goes through the host (quant_utils.py:202: `.cpu().numpy()`) and returns tensors on the data's device
    mant, ex = np.frexp(ratio.detach().reshape(-1).cpu().double().numpy())
    m = torch.tensor([math.floor(v * 2.0 ** 31 + 0.5) for v in mant.tolist()], dtype=torch.double).view(ratio.shape).to(z_int.device)
    e = torch.from_numpy(31.0 - ex).view(ratio.shape).to(z_int.device)
    return torch.round(z_int.type(torch.double) * m / (2.0 ** e))


def float_case0(z, bits, mode, out_sf, a_sf, w_sf):
    """fixedpoint_fn case 0 in float (quant_utils.py:390-413)."""
    with torch.no_grad():
        out_sf, a_sf, w_sf = _rs(out_sf, z), _rs(a_sf, z), _rs(w_sf, z)
        z_int = torch.round(z / a_sf / w_sf)
        lo, hi = clamp_range(bits, mode)
        return torch.clamp(_float_requant(z_int, a_sf, w_sf, out_sf, z).type(torch.float), lo, hi)


def float_case1(z, out_sf, a_sf, w_sf, identity, id_sf, id_w_sf):
    """fixedpoint_fn case 1 in float (quant_utils.py:416-456)."""
    with torch.no_grad():
        out_sf, a_sf, w_sf = _rs(out_sf, z), _rs(a_sf, z), _rs(w_sf, z)
        id_sf, id_w_sf = _rs(id_sf, z), _rs(id_w_sf, z)
        x_int = torch.round(identity / id_sf / id_w_sf)
        y_int = torch.round((z - identity) / a_sf / w_sf)
        o = _float_requant(x_int, id_sf, id_w_sf, out_sf, z) + _float_requant(y_int, a_sf, w_sf, out_sf, z)
        return o.type(torch.float)
