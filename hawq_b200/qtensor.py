"""Integer payload tensors and the frozen forward of the quant modules.

``IntActivation`` is what frozen engine modules hand to each other instead of the reference's fp32
``integer * scale`` tensors: a ``torch.Tensor`` wrapper subclass (logical shape NCHW, dtype fp32 as far as graph
code can tell) whose *node* carries either concrete integers in HBM (NHWC int8 / packed uint4 / uint16 or int32
residual stream) or a *pending* producer (a convolution, a residual sum, the stem, an average pool) that has not
been launched yet.  Pending producers are how the reference's module-by-module graph
(conv -> ReLU -> QuantAct, conv -> add -> QuantAct -> ReLU -> next QuantAct, reference
``utils/models/q_resnet.py:231-260``) turns into one fused kernel per convolution: the producer is launched when the
consuming ``QuantAct`` is reached, with that activation's dyadic requantisation (and, for residual sums, the NEXT
unit's low-bit activation) folded into the kernel epilogue.

``nn.ReLU``, ``nn.MaxPool2d``, ``+``, ``.view`` on an ``IntActivation`` are intercepted through
``__torch_function__`` and recorded on the node; anything else is not an integer-path operation and raises.
"""
import contextlib
import os
import threading

import numpy as np

import torch

from . import ops, quant_math as qmath
from ._lib import EPI_RAW_I32, EPI_REQUANT, EPI_RESIDUAL, EP_RATIOS_LE_2P20, ERR_UNSUPPORTED, HawqError


class EngineConfig(threading.local):
    """Per-thread execution mode of the frozen forward (no process-global state: several engines / threads may run at once).

    residual_bits: storage of the post-ReLU residual stream: 32 (always exact) or 16 (uint16 + sticky overflow flag
        HAWQ_FLAG_RESIDUAL_OVERFLOW; ``CompiledModel`` re-runs in 32-bit mode when the flag is raised).
    fast_kernels: False withholds every HAWQ_EP_RATIOS_* promise (always-saturating generic kernels).
    checked: True when the caller reads the device status word after the forward (``CompiledModel`` does).  The plain eager
        frozen forward does not, so it never promises HAWQ_EP_RATIOS_LE_2P20: the WIDE kernels only *flag* an int32 overflow and
        rely on the host to re-run, whereas ratios <= 1 cannot overflow and the generic kernels saturate like the reference."""
    residual_bits = 32
    fast_kernels = True
    checked = False
    # HBM container of 4-bit activations: 8 = one value per byte (the int8 kernels consume them directly: Blackwell has no int4 MMA,
    # so packed nibbles must be expanded on chip before every use), 4 = packed nibbles (half the bytes of those tensors, expansion by
    # converter warps).  Same integers either way; today the byte container is faster on every ResNet layer (profiles/r02), the
    # packed one is kept for bandwidth-bound deployments and is what the kernel tests exercise with a_bits = 4.
    a4_container = 4 if os.environ.get("HAWQ_B200_A4_STORAGE", "byte") == "packed" else 8
    dual = os.environ.get("HAWQ_B200_DUAL", "1") != "0"   # resize units: identity conv + last conv in one kernel (uint16 stream only)


config = EngineConfig()


@contextlib.contextmanager
def engine_mode(residual_bits=32, fast_kernels=True, checked=False):
    """Scoped execution mode (restored on exit, also when the forward raises)."""
    saved = (config.residual_bits, config.fast_kernels, config.checked)
    config.residual_bits, config.fast_kernels, config.checked = residual_bits, fast_kernels, checked
    try:
        yield config
    finally:
        config.residual_bits, config.fast_kernels, config.checked = saved


def _ratio_flags(*pairs):
    """HAWQ_EP_RATIOS_* promise for this thread's execution mode (see EngineConfig)."""
    if not config.fast_kernels:
        return 0
    f = ops.ratio_flags(*pairs)
    if f == EP_RATIOS_LE_2P20 and not config.checked:
        return 0
    return f


class Node:
    """Payload of an IntActivation.  kind:
         'int'      concrete integers: data (NHWC), bits, signed
         'conv'     pending convolution: mod, src(Node 'int'), a_sf, relu, pool
         'sum'      pending conv + identity: a (Node conv), b (Node conv | int)
         'residual' pending case-1 requant of a 'sum' by QuantAct `act` (+relu)
         'stem'     pending stem conv + pool + 16-bit requant by QuantAct `act` (+relu)
         'avgpool'  pending integer average pool of src
    """
    __slots__ = ("kind", "shape", "data", "bits", "signed", "mod", "src", "a_sf", "relu", "pool", "a", "b", "act",
                 "args", "k")

    def __init__(self, kind, shape, **kw):
        self.kind, self.shape = kind, tuple(shape)
        self.data = self.mod = self.src = self.a_sf = self.a = self.b = self.act = self.args = self.pool = None
        self.bits, self.signed, self.relu, self.k = 0, True, False, 0
        for k, v in kw.items():
            setattr(self, k, v)

    def become_int(self, data, bits, signed):
        self.kind, self.data, self.bits, self.signed = "int", data, bits, signed
        self.mod = self.src = self.a = self.b = self.act = self.args = None


class IntActivation(torch.Tensor):
    @staticmethod
    def __new__(cls, node, device, shape=None):
        t = torch.Tensor._make_wrapper_subclass(cls, tuple(shape if shape is not None else node.shape),
                                                dtype=torch.float32, device=device, requires_grad=False)
        t.node = node
        return t

    def __repr__(self):
        n = self.node
        return "IntActivation(kind=%s, shape=%s, bits=%s, device=%s)" % (n.kind, tuple(self.shape), n.bits, self.device)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", str(func))
        if name in ("relu", "relu_"):
            return _relu(args[0])
        if name in ("max_pool2d", "_max_pool2d", "max_pool2d_with_indices"):
            return _max_pool(*args, **kwargs)
        if name in ("add", "__add__", "__radd__", "__iadd__", "add_"):
            return _add(args[0], args[1])
        if name in ("view", "reshape", "flatten"):
            return _reshape(name, *args, **kwargs)
        if name in ("size", "dim", "__get__", "shape", "__repr__", "__str__", "is_cuda", "device"):
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        raise NotImplementedError(
            "torch op %r on an IntActivation is not part of the HAWQ integer path (supported between frozen "
            "modules: ReLU, MaxPool2d(3,2,1) after the stem, +, view/flatten). Use .dequantize() for a float tensor." % name)

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        raise NotImplementedError("aten op %s reached an IntActivation: only the HAWQ integer-path ops are defined on it" % func)

    def dequantize(self, scale):
        """fp32 NCHW ``integer * scale`` tensor (what the reference would have produced at this edge)."""
        n = materialize(self.node, self.device)
        s = float(scale.reshape(-1)[0]) if torch.is_tensor(scale) else float(scale)
        shp = self.node.shape
        nb, c = shp[0], shp[1]
        hh, ww = (shp[2], shp[3]) if len(shp) == 4 else (1, 1)
        out = torch.empty((nb, c, hh, ww), dtype=torch.float32, device=self.device)
        ops.dequant(n.data, nb, hh, ww, c, n.bits, n.signed, s, out)
        return out.view(*self.shape)

    def int_tensor(self):
        """Concrete integers as an int32 tensor in logical (NCHW / NC) order — for tests and debugging."""
        n = materialize(self.node, self.device)
        return unpack_to_int32(n).view(self.node.shape[0], *([self.node.shape[2], self.node.shape[3]] if len(self.node.shape) == 4 else []),
                                       self.node.shape[1]).permute(*((0, 3, 1, 2) if len(self.node.shape) == 4 else (0, 1))).contiguous()


def unpack_to_int32(n):
    """Node 'int' -> flat int32 tensor in storage (NHWC) order."""
    d = n.data
    if n.bits == 32:
        return d.view(torch.int32).reshape(-1).clone()
    if n.bits == 16:
        v = d.view(torch.int16).reshape(-1).to(torch.int32)
        return v if n.signed else (v & 0xFFFF)
    if n.bits == 8:
        return d.view(torch.int8 if n.signed else torch.uint8).reshape(-1).to(torch.int32)
    b = d.view(torch.uint8).reshape(-1, 4).to(torch.int32)            # hawq nibble order
    return torch.cat([b & 0xF, b >> 4], dim=1).reshape(-1)


# ------------------------------------------------------------------------------------------------ helpers
def _cpu_f32(t):
    return t.detach().to("cpu", torch.float32).reshape(-1)


def _key(t):
    return None if t is None else _cpu_f32(t).numpy().tobytes()


def _buf_key(t):
    return None if t is None else (id(t), t._version, t.data_ptr())


def _dev_of(x):
    return x.device


def _alloc(device, numel, bits):
    if bits == 32:
        return torch.empty(numel, dtype=torch.int32, device=device)
    if bits == 16:
        return torch.empty(numel, dtype=torch.int16, device=device)
    if bits == 8:
        return torch.empty(numel, dtype=torch.int8, device=device)
    return torch.empty(numel // 2, dtype=torch.uint8, device=device)


def _act_clamp(act):
    return qmath.clamp_range(act.activation_bit, act.quant_mode)


def _store_bits(act):
    """Bits per value of `act`'s output in HBM (see EngineConfig.a4_container)."""
    b = act.activation_bit
    return config.a4_container if b == 4 else b


def _store_signed(act):
    """Whether the stored integers may be read as signed bytes: symmetric activations, and 4-bit values in byte containers (0..15)."""
    return act.quant_mode == "symmetric" or (act.activation_bit == 4 and _store_bits(act) == 8)


def _frozen_scale(act):
    """Scale of a frozen QuantAct, CPU fp32 [1]; refreshes the act_scaling_factor buffer like the reference forward."""
    c = act.__dict__.setdefault("_hawq_cache", {})
    # keyed on buffer identity + in-place version (no device->host copy on a hit: the buffers may live on the GPU and
    # this runs inside CUDA-graph capture); `x_min += lo` bumps _version, `x_min = ...` rebinds the buffer
    key = ("scale", _buf_key(act.x_min), _buf_key(act.x_max), act.activation_bit, act.quant_mode, act.__dict__.get("_override_gen", 0))
    if c.get("scale_key") != key:
        sf = act.current_scale().detach().to("cpu", torch.float32).reshape(1)
        c["scale_key"], c["scale"] = key, sf
        act.act_scaling_factor = sf.to(act.x_min.device)
        c["me"] = {}
        c["gen"] = c.get("gen", 0) + 1
    return c["scale"]


def _act_tag(kind, act):
    """Cache tag of per-channel epilogue parameters that depend on QuantAct `act` (invalidated when its scale changes)."""
    return (kind, id(act), act._hawq_cache["gen"])


def _dyadic(act, a_sf, w_sf, tag):
    """Cached (m list, e list) of the ratio (a_sf * w_sf) / act_scale (quant_utils.py:394-400)."""
    c = act._hawq_cache
    key = (tag, _key(a_sf), _key(w_sf))
    if key not in c["me"]:
        ratio = qmath.requant_ratio(_cpu_f32(a_sf), _cpu_f32(w_sf), c["scale"])
        c["me"][key] = qmath.dyadic_pairs(ratio)
    return c["me"][key]


def _ones():
    return torch.ones(1)


# ------------------------------------------------------------------------------------------------ conv params
_OWN_BUFFERS = ("weight_integer", "bias_integer", "convbn_scaling_factor", "conv_scaling_factor", "fc_scaling_factor")


def _param_versions(mod):
    """In-place modification counters of every float parameter / statistic the integer plan of `mod` derives from:
    ``load_state_dict`` and optimizer steps write in place, so a changed tuple invalidates the cached plan (the reference
    recomputes its integers on every forward, quant_modules.py:440-484; SURVEY.md 8(b) lifecycle row)."""
    vs = []
    for name, t in list(mod.named_parameters()) + list(mod.named_buffers()):
        if name.rsplit(".", 1)[-1] in _OWN_BUFFERS or "num_batches_tracked" in name:
            continue
        vs.append((id(t), t._version))
    vs.append(mod.__dict__.get("_override_gen", 0))          # integers installed from a quantized checkpoint
    return tuple(vs)



def _conv_cache(mod, a_sf, a_bits, device):
    """Device-resident integer parameters of a frozen conv module for input scale a_sf / input width a_bits."""
    c = mod.__dict__.setdefault("_hawq_cache", {})
    key = (_key(a_sf), a_bits, str(device), mod.weight_bit, mod.per_channel, mod.bias_bit, mod.quantize_bias, _param_versions(mod))
    ent = c.get(key)
    if ent is not None:
        return ent
    if c:              # a different scale / bit width / parameter version: start a new plan.  The old dict is not cleared in
        c = mod.__dict__["_hawq_cache"] = {}     # place: a CompiledModel may still replay graphs that read its buffers
    with torch.no_grad():
        conv = mod.conv
        src_dev = conv.weight.device
        w_sf, w_int, b_int, _ = mod.integer_params(a_sf.to(src_dev))
        w_sf = w_sf.detach().to("cpu", torch.float32)
        cout, cin, kh, kw = w_int.shape
        if conv.groups != 1 or conv.dilation[0] != 1 or conv.dilation[1] != 1 or kh != kw or conv.stride[0] != conv.stride[1]:
            raise NotImplementedError("hawq_b200 convolutions: groups=1, dilation=1, square kernels/strides only")
        w = w_int.detach().to("cpu").permute(0, 2, 3, 1).contiguous().to(torch.int8)      # OHWI
        bias = (b_int.detach().to("cpu").to(torch.int64).numpy() if b_int is not None else np.zeros(cout, dtype=np.int64))
        stem = (cin == 3 and kh == 7 and conv.stride[0] == 2 and conv.padding[0] == 3 and cout == 64)
        w256 = None
        if stem:
            wp = torch.zeros((cout, 8, 8, 4), dtype=torch.int8)       # kernel rows 7 -> 8, taps 7 -> 8, channels 3 -> 4 (zeros)
            wp[:, :7, :7, :3] = w
            w256 = wp.to(device)                                      # fused tcgen05 stem: K = 256
            w = wp[:, :7].contiguous()                                # two-kernel path: K = 224
        else:
            if cin % 64 != 0 or cout % 64 != 0:
                raise NotImplementedError("hawq_b200 convolutions need Cin and Cout multiples of 64 (got %d, %d)" % (cin, cout))
            if a_bits == 4:
                ops.permute_weights_for_i4(w)
        tiled = (not stem) and torch.device(device).type == "cuda"
        ent = dict(w=ops.upload_weights(w, device) if tiled else w.to(device), w_layout=1 if tiled else 0, w_sf=w_sf, bias=bias, cout=cout, cin=cin, k=kh, stride=conv.stride[0],
                   pad=conv.padding[0], stem=stem, chan={}, w256=w256)
    c[key] = ent
    return ent


def _chan_tensor(ent, tag, m, e, device):
    t = ent["chan"].get(tag)
    if t is None:
        t = ent["chan"][tag] = ops.make_chan(ent["bias"], m, e).to(device)
    return t


# ------------------------------------------------------------------------------------------------ lazy ops
def _relu(x):
    n = x.node
    if n.kind in ("conv", "residual", "stem"):
        if n.kind == "conv" and n.pool is not None:
            raise NotImplementedError("ReLU after a pooled convolution must follow its QuantAct")
        m = Node(n.kind, n.shape, mod=n.mod, src=n.src, a_sf=n.a_sf, relu=True, pool=n.pool, a=n.a, b=n.b, act=n.act,
                 args=n.args)
        return IntActivation(m, x.device, x.shape)
    if n.kind == "int":
        if not n.signed:
            return x
        dt = {8: torch.int8, 16: torch.int16, 32: torch.int32}[n.bits]
        m = Node("int", n.shape, data=torch.clamp_min(n.data.view(dt), 0), bits=n.bits, signed=n.signed)
        return IntActivation(m, x.device, x.shape)
    raise NotImplementedError("ReLU on a pending %s" % n.kind)


def _max_pool(x, kernel_size, stride=None, padding=0, dilation=1, ceil_mode=False, return_indices=False):
    n = x.node
    one = lambda v: v[0] if isinstance(v, (tuple, list)) else v
    k, s, p = one(kernel_size), one(stride if stride is not None else kernel_size), one(padding)
    if n.kind != "conv" or (k, s, p) != (3, 2, 1) or one(dilation) != 1 or ceil_mode or return_indices or n.relu:
        raise NotImplementedError("integer max-pool is fused only as MaxPool2d(3,2,1) directly after a convolution")
    nb, c, hh, ww = n.shape
    shape = (nb, c, (hh + 2 - 3) // 2 + 1, (ww + 2 - 3) // 2 + 1)
    m = Node("conv", shape, mod=n.mod, src=n.src, a_sf=n.a_sf, relu=False, pool=(3, 2, 1))
    return IntActivation(m, x.device)


def _add(a, b):
    if not (isinstance(a, IntActivation) and isinstance(b, IntActivation)):
        raise NotImplementedError("IntActivation + non-IntActivation")
    na, nb = a.node, b.node
    if na.kind != "conv" or na.relu or na.pool is not None:
        if nb.kind == "conv" and not nb.relu and nb.pool is None:
            na, nb = nb, na
        else:
            raise NotImplementedError("integer residual add expects conv_output + identity")
    if tuple(na.shape) != tuple(nb.shape):
        raise RuntimeError("shape mismatch in residual add: %s vs %s" % (na.shape, nb.shape))
    return IntActivation(Node("sum", na.shape, a=na, b=nb), a.device)


def _reshape(name, x, *shape, **kw):
    with torch._C.DisableTorchFunctionSubclass():
        meta = torch.empty(x.shape, device="meta")
        new = getattr(meta, name)(*shape, **kw).shape
    n = x.node
    if n.kind != "int" or len(n.shape) == 4 and (n.shape[2] != 1 or n.shape[3] != 1):
        if tuple(new) == tuple(x.shape):
            return x
        raise NotImplementedError("view() of an IntActivation is only supported on [N,C,1,1] -> [N,C]")
    return IntActivation(n, x.device, tuple(new))


# ------------------------------------------------------------------------------------------------ launches
def _conv_out_hw(n, ent):
    nb, _, hh, ww = n.src.shape
    ho = (hh + 2 * ent["pad"] - ent["k"]) // ent["stride"] + 1
    wo = (ww + 2 * ent["pad"] - ent["k"]) // ent["stride"] + 1
    return nb, hh, ww, ho, wo


def _launch_conv(n, ent, ep, chan, device, out=None, out_low=None, res=None, res_chan=None):
    nb, hh, ww, _, _ = _conv_out_hw(n, ent)
    d = ops.conv_desc(nb, hh, ww, ent["cin"], ent["cout"], ent["k"], ent["k"], ent["stride"], ent["pad"], n.src.bits, ent["w_layout"])
    ops.conv2d(n.src.data, d, ep, ent["w"], chan, res=res, res_chan=res_chan, out=out, out_low=out_low)


def _check_src(n):
    s = n.src
    if s.kind != "int" or s.bits not in (4, 8):
        raise NotImplementedError("convolution input must be a concrete 4/8-bit IntActivation")
    if s.bits == 8 and not s.signed:
        raise NotImplementedError("8-bit unsigned (asymmetric) activations are not supported by the int8 kernels")


def _conv_case0(n, act, device):
    """conv [+ReLU] -> QuantAct case 0, one kernel."""
    _check_src(n)
    ent = _conv_cache(n.mod, n.a_sf, n.src.bits, device)
    if ent["stem"]:
        raise NotImplementedError("the stem convolution is only supported as conv -> MaxPool2d(3,2,1) -> 16-bit QuantAct")
    m, e = _dyadic(act, n.a_sf, ent["w_sf"], "case0")
    chan = _chan_tensor(ent, _act_tag("c0", act), m, e, device)
    nb, _, _, ho, wo = _conv_out_hw(n, ent)
    bits = _store_bits(act)
    lo, hi = _act_clamp(act)
    out = _alloc(device, nb * ho * wo * ent["cout"], bits)
    ep = ops.epilogue(EPI_REQUANT, relu=n.relu, out_bits=bits, clamp=(lo, hi),
                      flags=_ratio_flags((m, e)))
    _launch_conv(n, ent, ep, chan, device, out=out)
    return Node("int", (nb, ent["cout"], ho, wo), data=out, bits=bits, signed=_store_signed(act))


def _conv_raw(n, device):
    """identity-branch conv: int32 accumulator + bias."""
    _check_src(n)
    ent = _conv_cache(n.mod, n.a_sf, n.src.bits, device)
    chan = _chan_tensor(ent, "raw", [0] * ent["cout"], [1] * ent["cout"], device)
    nb, _, _, ho, wo = _conv_out_hw(n, ent)
    out = _alloc(device, nb * ho * wo * ent["cout"], 32)
    _launch_conv(n, ent, ops.epilogue(EPI_RAW_I32, flags=_ratio_flags()), chan, device, out=out)
    return out, ent


def _launch_residual(r, low_act, device):
    """'residual' node -> concrete residual stream (and optionally the next QuantAct's low-bit output)."""
    act, s = r.act, r.a
    conv, ident = s.a, s.b
    _check_src(conv)
    ent = _conv_cache(conv.mod, conv.a_sf, conv.src.bits, device)
    a_sf, w_sf, id_sf, id_w_sf = r.args
    m2, e2 = _dyadic(act, a_sf, w_sf, "case1-main")
    chan = _chan_tensor(ent, _act_tag("c1", act), m2, e2, device)
    res_chan = None
    pairs = [(m2, e2)]
    dual = None
    if ident.kind == "conv":
        _check_src(ident)
        ient = _conv_cache(ident.mod, ident.a_sf, ident.src.bits, device)
        m1, e1 = _dyadic(act, id_sf, id_w_sf, "case1-idconv")
        res_chan = _chan_tensor(ient, _act_tag("c1res", act), m1, e1, device)
        res_kind, res_bits, res_me = 1, 32, (0, 1)
        pairs.append((m1, e1))
        # resize units: both 1x1 convolutions in one kernel (two TMEM accumulators) when the fast uint16 stream is in use
        if (config.dual and r.relu and config.residual_bits == 16 and ent["k"] == 1 and ent["stride"] == 1 and ent["pad"] == 0
                and ient["k"] == 1 and ient["pad"] == 0 and ent["w_layout"] == 1 and ient["w_layout"] == 1
                and conv.src.bits == ident.src.bits):
            dual = ient
        else:
            res, _ = _conv_raw(ident, device)
    else:
        ident = materialize(ident, device)
        if ident.bits not in (16, 32):
            raise NotImplementedError("identity operand must be the 16/32-bit residual stream")
        res = ident.data
        m1, e1 = _dyadic(act, id_sf, id_w_sf, "case1-id")
        res_kind, res_bits, res_me = 0, ident.bits, (m1[0], e1[0])
        pairs.append((m1[0], e1[0]))
    nb, _, _, ho, wo = _conv_out_hw(conv, ent)
    numel = nb * ho * wo * ent["cout"]
    y_bits = config.residual_bits if r.relu else 32
    y = _alloc(device, numel, y_bits)
    low = low_node = None
    kw = {}
    if low_act is not None:
        lscale = _frozen_scale(low_act)
        lm, le = _dyadic(low_act, _frozen_scale(act), _ones(), "case0")
        lo, hi = _act_clamp(low_act)
        low = _alloc(device, numel, _store_bits(low_act))
        kw = dict(low_bits=_store_bits(low_act), low_me=(lm[0], le[0]), low_clamp=(lo, hi))
        pairs.append((lm[0], le[0]))
        low_node = Node("int", (nb, ent["cout"], ho, wo), data=low, bits=_store_bits(low_act), signed=_store_signed(low_act))
    ep = ops.epilogue(EPI_RESIDUAL, relu=r.relu, res_kind=res_kind, res_bits=res_bits, res_me=res_me, y_bits=y_bits,
                      flags=_ratio_flags(*pairs), **kw)
    if dual is not None and ep.flags == 0:       # no ratio promise (saturating generic kernels): two launches
        res, _ = _conv_raw(ident, device)
        dual = None
    if dual is not None:
        nb2, hh2, ww2, _, _ = _conv_out_hw(ident, dual)
        _, hh, ww, _, _ = _conv_out_hw(conv, ent)
        d1 = ops.conv_desc(nb, hh, ww, ent["cin"], ent["cout"], 1, 1, 1, 0, conv.src.bits, 1)
        d2 = ops.conv_desc(nb2, hh2, ww2, dual["cin"], dual["cout"], 1, 1, dual["stride"], 0, ident.src.bits, 1)
        try:
            ops.conv2d_dual(conv.src.data, d1, ep, ent["w"], chan, d2, ident.src.data, dual["w"], res_chan, out=y, out_low=low)
        except HawqError as err:
            if err.code != ERR_UNSUPPORTED:
                raise
            res, _ = _conv_raw(ident, device)          # the library declined this combination: two launches, same results
            _launch_conv(conv, ent, ep, chan, device, out=y, out_low=low, res=res, res_chan=res_chan)
    else:
        _launch_conv(conv, ent, ep, chan, device, out=y, out_low=low, res=res, res_chan=res_chan)
    r.shape = (nb, ent["cout"], ho, wo)
    r.become_int(y, y_bits, signed=(y_bits == 32))
    return low_node


def _launch_stem(st, low_act, device):
    """'stem' node: 7x7 conv (+bias, 16-bit requant, ReLU) -> int16, then max-pool -> residual stream (+ low-bit copy)."""
    conv, act = st.a, st.act
    src = conv.src
    if src.kind != "int" or src.bits != 8 or not src.signed:
        raise NotImplementedError("stem input must be signed int8")
    if not st.relu:
        raise NotImplementedError("the fused stem expects the reference order conv -> pool -> QuantAct(16) -> ReLU")
    ent = _conv_cache(conv.mod, conv.a_sf, 8, device)
    m, e = _dyadic(act, conv.a_sf, ent["w_sf"], "case0")
    chan = _chan_tensor(ent, _act_tag("c0", act), m, e, device)
    nb, _, hh, ww = src.shape
    ho, wo = (hh + 6 - 7) // 2 + 1, (ww + 6 - 7) // 2 + 1
    lo, hi = _act_clamp(act)
    po, qo = (ho + 2 - 3) // 2 + 1, (wo + 2 - 3) // 2 + 1
    numel = nb * po * qo * 64
    y_bits = config.residual_bits
    y = _alloc(device, numel, y_bits)
    low = low_node = None
    low_bits, low_me, low_clamp = 0, (0, 1), (0, 0)
    if low_act is not None:
        _frozen_scale(low_act)
        lm, le = _dyadic(low_act, _frozen_scale(act), _ones(), "case0")
        low_bits, low_me, low_clamp = _store_bits(low_act), (lm[0], le[0]), _act_clamp(low_act)
        low = _alloc(device, numel, low_bits)
        low_node = Node("int", (nb, 64, po, qo), data=low, bits=low_bits, signed=_store_signed(low_act))
    fused = False
    if config.fast_kernels and min(e) >= 31 and (low_bits == 0 or low_me[0] == 0 or 31 <= low_me[1] <= 51):
        try:                    # one kernel: convolution, pool, requantisation, low-bit copy (the int16 tensor never exists)
            ops.stem_pool(src.data, ent["w256"], chan, (lo, hi), nb, hh, ww, y_bits, y, low_bits, low_me, low_clamp, low)
            fused = True
        except HawqError as err:
            if err.code != ERR_UNSUPPORTED:
                raise
    if not fused:
        t16 = torch.empty(nb * ho * wo * 64, dtype=torch.int16, device=device)
        ops.stem_conv(src.data, ent["w"], chan, (lo, hi), t16, nb, hh, ww)
        ops.maxpool_requant(t16, nb, ho, wo, 64, y_bits, y, low_bits, low_me, low_clamp, low)
    st.shape = (nb, 64, po, qo)
    st.become_int(y, y_bits, signed=(y_bits == 32))
    return low_node


def materialize(n, device):
    """Force a node to concrete integers (launching its producer without further fusion)."""
    if n.kind == "int":
        return n
    if n.kind == "residual":
        _launch_residual(n, None, device)
        return n
    if n.kind == "stem":
        _launch_stem(n, None, device)
        return n
    raise NotImplementedError("a pending %s has no integer value before its QuantAct" % n.kind)


# ------------------------------------------------------------------------------------------------ module forwards
def _require_cuda(x, what):
    if not x.is_cuda:
        raise RuntimeError("%s is frozen: its forward runs on the hawq_b200 CUDA kernels and needs a CUDA input "
                           "(got %s). Un-freeze the model for CPU calibration; there is no CPU fallback." % (what, x.device))


def act_forward(act, x, a_sf, w_sf, identity, id_sf, id_w_sf):
    """Frozen QuantAct.forward (reference quant_modules.py:205-303) on the integer path."""
    scale = _frozen_scale(act)
    dev = _dev_of(x)
    if not isinstance(x, IntActivation):
        _require_cuda(x, "QuantAct")
        if a_sf is not None:
            raise NotImplementedError("a frozen QuantAct in the middle of a graph expects the IntActivation produced by "
                                      "the previous frozen module, not a float tensor")
        if x.dim() != 4:
            raise NotImplementedError("input quantisation expects NCHW")
        nb, c, hh, ww = x.shape
        lo, hi = _act_clamp(act)
        if act.activation_bit != 8 or act.quant_mode != "symmetric":
            raise NotImplementedError("network input quantisation is 8-bit symmetric in HAWQ ResNets")
        out = torch.empty(nb * hh * ww * c, dtype=torch.int8, device=dev)
        ops.quantize_input(x.contiguous().float(), float(scale), (lo, hi), out)
        return (IntActivation(Node("int", (nb, c, hh, ww), data=out, bits=8, signed=True), dev), scale)
    n = x.node
    if identity is not None:                                   # case 1: becomes a pending residual
        if n.kind != "sum":
            raise NotImplementedError("QuantAct with identity expects x = conv_output + identity")
        if w_sf is None:
            raise RuntimeError("case 1 needs the weight scaling factor of the last convolution")
        if id_w_sf is None:
            id_w_sf = _ones()
        r = Node("residual", n.shape, a=n, act=act, args=(a_sf, w_sf, id_sf, id_w_sf))
        return (IntActivation(r, dev), scale)
    if a_sf is None:                                           # already-quantised input handed in by the caller
        if n.kind != "int":
            raise NotImplementedError("QuantAct without a previous scale expects concrete integers")
        return (x, scale)
    if n.kind == "conv":
        if n.pool is not None:                                 # stem: conv -> pool -> QuantAct(16) [-> ReLU]
            st = Node("stem", n.shape, a=n, act=act)
            if act.activation_bit != 16:
                raise NotImplementedError("pooled convolution must be followed by the 16-bit quant_act_int32")
            return (IntActivation(st, dev), scale)
        return (IntActivation(_conv_case0(n, act, dev), dev), scale)
    if n.kind in ("residual", "stem"):                         # fuse this activation into the producer's epilogue
        low = (_launch_residual if n.kind == "residual" else _launch_stem)(n, act, dev)
        return (IntActivation(low, dev), scale)
    if n.kind == "avgpool":
        src = materialize(n.src, dev)
        m, e = _dyadic(act, a_sf, _ones(), "case0")
        lo, hi = _act_clamp(act)
        nb, c, hh, ww = src.shape
        if act.activation_bit != 8 or act.quant_mode != "symmetric":
            raise NotImplementedError("the pooled tail is 8-bit symmetric in HAWQ ResNets")
        out = torch.empty(nb * c, dtype=torch.int8, device=dev)
        ops.avgpool_requant(src.data, nb, hh * ww, c, src.bits, (m[0], e[0]), (lo, hi), out)
        return (IntActivation(Node("int", (nb, c, 1, 1), data=out, bits=8, signed=True), dev), scale)
    if n.kind == "int":                                        # stand-alone requant of the residual stream
        if n.bits not in (16, 32):
            raise NotImplementedError("stand-alone requantisation expects the 16/32-bit residual stream")
        ws = w_sf if w_sf is not None else _ones()
        m, e = _dyadic(act, a_sf, ws, "case0")
        c = n.shape[1]
        per_ch = len(m) > 1
        chan = ops.make_chan([0] * len(m), m, e).to(dev)
        rows = int(np.prod(n.shape)) // c
        lo, hi = _act_clamp(act)
        out = _alloc(dev, rows * c, _store_bits(act))
        ops.requant(n.data, rows, c, n.bits, chan, 1 if per_ch else 0, False, _store_bits(act), (lo, hi), out)
        return (IntActivation(Node("int", n.shape, data=out, bits=_store_bits(act), signed=_store_signed(act)), dev), scale)
    raise NotImplementedError("QuantAct on a pending %s" % n.kind)


def conv_forward(mod, x, a_sf):
    """Frozen QuantBnConv2d / QuantConv2d forward: records a pending convolution (launched by its consumer)."""
    if not isinstance(x, IntActivation):
        _require_cuda(x, type(mod).__name__)
        raise NotImplementedError("a frozen %s expects the IntActivation produced by a frozen QuantAct" % type(mod).__name__)
    if a_sf is None:
        raise ValueError("pre_act_scaling_factor is required")
    src = materialize(x.node, x.device)
    ent = _conv_cache(mod, a_sf, 8 if src.bits not in (4, 8) else src.bits, x.device)
    nb, _, hh, ww = src.shape
    ho = (hh + 2 * ent["pad"] - ent["k"]) // ent["stride"] + 1
    wo = (ww + 2 * ent["pad"] - ent["k"]) // ent["stride"] + 1
    n = Node("conv", (nb, ent["cout"], ho, wo), mod=mod, src=src, a_sf=a_sf)
    return (IntActivation(n, x.device), ent["w_sf"])


def linear_forward(mod, x, a_sf):
    """Frozen QuantLinear.forward (quant_modules.py:79-130) -> fp32 logits (a real torch tensor)."""
    if not isinstance(x, IntActivation):
        _require_cuda(x, "QuantLinear")
        raise NotImplementedError("a frozen QuantLinear expects the IntActivation produced by a frozen QuantAct")
    n = materialize(x.node, x.device)
    if n.bits != 8 or not n.signed:
        raise NotImplementedError("QuantLinear input must be signed int8")
    dev = x.device
    c = mod.__dict__.setdefault("_hawq_cache", {})
    key = (_key(a_sf), str(dev), mod.weight_bit, mod.per_channel, _param_versions(mod))
    ent = c.get(key)
    if ent is None:
        if c:
            c = mod.__dict__["_hawq_cache"] = {}
        with torch.no_grad():
            w_sf, w_int, b_int, bias_sf = mod.integer_params(a_sf.to(mod.weight.device))
            cout, k = w_int.shape
            if k % 64 != 0:
                raise NotImplementedError("QuantLinear in_features must be a multiple of 64")
            cpad = (cout + 63) // 64 * 64
            w = torch.zeros((cpad, k), dtype=torch.int8)
            w[:cout] = w_int.detach().to("cpu").to(torch.int8)
            bias = np.zeros(cpad, dtype=np.int64)
            if b_int is not None:
                bias[:cout] = b_int.detach().to("cpu").to(torch.int64).numpy()
            fs = torch.zeros(cpad, dtype=torch.float32)
            fs[:cout] = bias_sf.detach().to("cpu", torch.float32).reshape(-1)   # fc_scaling_factor * act scale, fp32
            ent = c[key] = dict(w=w.to(dev), chan=ops.make_chan(bias, [0] * cpad, [1] * cpad).to(dev), fscale=fs.to(dev),
                                cout=cout, cpad=cpad, k=k)
    nb = n.shape[0]
    out = torch.empty((nb, ent["cout"]), dtype=torch.float32, device=dev)
    ops.linear(n.data, ent["w"], ent["chan"], ent["fscale"], out, nb, ent["k"], ent["cout"], ent["cpad"])
    return out


def avgpool_forward(mod, x, sf):
    n = x.node
    shape = n.shape
    if len(shape) != 4 or mod.kernel_size != shape[2] or shape[2] != shape[3] or mod.padding != 0:
        raise NotImplementedError("integer average pooling is global (kernel == feature map)")
    p = Node("avgpool", (shape[0], shape[1], 1, 1), src=n)
    return (IntActivation(p, x.device), sf.view(-1) if sf is not None else sf)
