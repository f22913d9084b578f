"""TEST INFRASTRUCTURE (oracle, never imported by the product): torch-CPU restatement of the
reference's quantized ("fake-quant") forward, i.e. what HAWQ actually executes.

Activations travel as fp32 ``integer * scale`` tensors, convolutions are fp32 ``F.conv2d`` on
integer-valued tensors, requantisation is fp64 arithmetic with ``torch.round`` (half-to-even),
dyadic multipliers come from ``np.frexp`` + Decimal ROUND_HALF_UP.  Every function cites the
reference lines it follows (paths relative to /root/reference).  This file is pinned against the
unmodified reference by ``tests/test_oracle_vs_reference.py`` (build container) and against the
committed golden vectors in ``tests/golden`` (everywhere).  It doubles as the CPU baseline
(``cpu_baseline.kind == "port"``) in bench.py because the Python reference cannot travel to the GPU box.
"""
import decimal
from decimal import Decimal

import numpy as np
import torch
import torch.nn.functional as F

QM = "utils/quantization_utils/quant_modules.py"
QU = "utils/quantization_utils/quant_utils.py"


# ----------------------------------------------------------------------------- quant math (L1)
def sym_scale(bits, mn, mx, per_channel):
    """quant_utils.py:128-152 symmetric_linear_quantization_params."""
    n = 2 ** (bits - 1) - 1
    if per_channel:
        s, _ = torch.max(torch.stack([mn.abs(), mx.abs()], dim=1), dim=1)
        return torch.clamp(s, min=1e-8) / n
    s = max(mn.abs(), mx.abs())
    return torch.clamp(s, min=1e-8) / n


def asym_scale(bits, mn, mx):
    """quant_utils.py:155-185 (zero point is computed there but never applied; we drop it)."""
    n = 2 ** bits - 1
    return torch.clamp(mx - mn, min=1e-8) / float(n)


def _bcast(scale, x):
    """quant_utils.py:84-93 reshape rule of linear_quantize."""
    if x.dim() == 4:
        return scale.view(-1, 1, 1, 1)
    if x.dim() == 2:
        return scale.view(-1, 1)
    return scale.view(-1)


def quant_sym(x, bits, scale):
    """quant_utils.py:231-258 SymmetricQuantFunction.forward (+ linear_quantize :73-97)."""
    n = 2 ** (bits - 1) - 1
    q = torch.round(1. / _bcast(scale, x) * x + torch.tensor(0.))
    return torch.clamp(q, -n - 1, n)


def quant_asym(x, bits, scale):
    """quant_utils.py:275-308 AsymmetricQuantFunction.forward with zero_point 0."""
    n = 2 ** bits - 1
    q = torch.round(1. / _bcast(scale, x) * x + torch.tensor(0))
    return torch.clamp(q, 0, n)


def batch_frexp(r):
    """quant_utils.py:188-213: mantissa*2^31 rounded HALF_UP (may reach 2^31, not renormalised), e = 31 - exp."""
    shape = r.size()
    mant, ex = np.frexp(r.view(-1).cpu().numpy())
    ms = [int(Decimal(m * (2 ** 31)).quantize(Decimal('1'), rounding=decimal.ROUND_HALF_UP)) for m in mant]
    m = torch.from_numpy(np.array(ms)).view(shape)
    e = torch.from_numpy(31. - ex).view(shape)
    return m, e


def _new_scale(a_sf, w_sf, z_sf):
    """quant_utils.py:394-397 ("follow TVM's computation")."""
    A = a_sf.type(torch.double) * w_sf.type(torch.double)
    B = A.type(torch.float).type(torch.double)
    C = z_sf.type(torch.float).type(torch.double)
    return B / C


def _shape4(t, z):
    return t.view(1, -1, 1, 1) if z.dim() == 4 else t.view(1, -1)


def fixedpoint_case0(z, bits, mode, z_sf, a_sf, w_sf):
    """quant_utils.py:390-413."""
    n = 2 ** (bits - 1) - 1 if mode == 'symmetric' else 2 ** bits - 1
    z_sf, a_sf, w_sf = _shape4(z_sf, z), _shape4(a_sf, z), _shape4(w_sf, z)
    z_int = torch.round(z / a_sf / w_sf)
    m, e = batch_frexp(_shape4(_new_scale(a_sf, w_sf, z_sf), z))
    out = torch.round(z_int.type(torch.double) * m.type(torch.double) / (2.0 ** e))
    if mode == 'symmetric':
        return torch.clamp(out.type(torch.float), -n - 1, n)
    return torch.clamp(out.type(torch.float), 0, n)


def fixedpoint_case1(z, z_sf, a_sf, w_sf, identity, id_sf, id_w_sf):
    """quant_utils.py:416-456 (no clamp)."""
    z_sf, a_sf, w_sf = _shape4(z_sf, z), _shape4(a_sf, z), _shape4(w_sf, z)
    id_sf, id_w_sf = _shape4(id_sf, z), _shape4(id_w_sf, z)
    wx_int = torch.round(identity / id_sf / id_w_sf)
    m1, e1 = batch_frexp(_shape4(_new_scale(id_sf, id_w_sf, z_sf), z))
    o1 = wx_int.type(torch.double) * m1.type(torch.double)
    o1 = torch.round(o1 / (2.0 ** e1))
    wy_int = torch.round((z - identity) / a_sf / w_sf)
    m2, e2 = batch_frexp(_shape4(_new_scale(a_sf, w_sf, z_sf), z))
    o2 = wy_int.type(torch.double) * m2.type(torch.double)
    o2 = torch.round(o2 / (2.0 ** e2))
    return (o1 + o2).type(torch.float)


# ----------------------------------------------------------------------------- modules (L2), functional
class ActState:
    """State of one QuantAct (quant_modules.py:157-182): range buffers + stamped attributes."""

    def __init__(self, bits=4, mode='symmetric', momentum=0.95):
        self.bits, self.mode, self.momentum = bits, mode, momentum
        self.x_min = torch.zeros(1)
        self.x_max = torch.zeros(1)
        self.running_stat = True
        self.scale = torch.zeros(1)

    def _update_range(self, x):
        """quant_modules.py:233-258 with act_percentile == 0."""
        x_min, x_max = x.data.min(), x.data.max()
        if self.x_min == self.x_max:
            self.x_min = self.x_min + x_min
            self.x_max = self.x_max + x_max
        elif self.momentum == -1:
            self.x_min = min(self.x_min, x_min)
            self.x_max = max(self.x_max, x_max)
        else:
            self.x_min = self.x_min * self.momentum + x_min * (1 - self.momentum)
            self.x_max = self.x_max * self.momentum + x_max * (1 - self.momentum)

    def __call__(self, x, a_sf=None, w_sf=None, identity=None, id_sf=None, id_w_sf=None):
        """quant_modules.py:205-303 (no multi-branch list case; ResNets never use it)."""
        if self.running_stat:
            self._update_range(x)
        if self.mode == 'symmetric':
            self.scale = sym_scale(self.bits, self.x_min, self.x_max, False)
        elif self.mode == 'asymmetric':
            self.scale = asym_scale(self.bits, self.x_min, self.x_max)
        else:
            raise ValueError("unknown quant mode: {}".format(self.mode))
        if a_sf is None:
            q = quant_sym(x, self.bits, self.scale) if self.mode == 'symmetric' else quant_asym(x, self.bits, self.scale)
        elif identity is None:
            if w_sf is None:
                w_sf = torch.ones(1)
            q = fixedpoint_case0(x, self.bits, self.mode, self.scale, a_sf, w_sf)
        else:
            if id_w_sf is None:
                id_w_sf = torch.ones(1)
            q = fixedpoint_case1(x, self.scale, a_sf, w_sf, identity, id_sf, id_w_sf)
        return q * self.scale.view(-1), self.scale


class ConvBnState:
    """QuantBnConv2d in its folded-BN branch (quant_modules.py:440-494), per_channel, 32-bit bias."""

    def __init__(self, conv, bn, wbits=4, bias_bits=32):
        self.conv, self.bn, self.wbits, self.bias_bits = conv, bn, wbits, bias_bits
        self.weight_integer = None
        self.bias_integer = None
        self.w_sf = None

    def __call__(self, x, a_sf):
        conv, bn = self.conv, self.bn
        running_std = torch.sqrt(bn.running_var.detach() + bn.eps)
        scale_factor = bn.weight / running_std
        scaled_weight = conv.weight * scale_factor.reshape([conv.out_channels, 1, 1, 1])
        if conv.bias is not None:
            scaled_bias = conv.bias
        else:
            scaled_bias = torch.zeros_like(bn.running_mean)
        scaled_bias = (scaled_bias - bn.running_mean.detach()) * scale_factor + bn.bias
        w2 = scaled_weight.data.contiguous().view(conv.out_channels, -1)
        w_min, w_max = w2.min(dim=1).values, w2.max(dim=1).values
        self.w_sf = sym_scale(self.wbits, w_min, w_max, True)
        self.weight_integer = quant_sym(scaled_weight, self.wbits, self.w_sf)
        bias_sf = self.w_sf.view(1, -1) * a_sf.view(1, -1)
        self.bias_integer = quant_sym(scaled_bias, self.bias_bits, bias_sf)
        x_int = x / a_sf.view(1, -1, 1, 1)
        out = F.conv2d(x_int, self.weight_integer, self.bias_integer, conv.stride, conv.padding,
                       conv.dilation, conv.groups) * bias_sf.view(1, -1, 1, 1)
        return out, self.w_sf


class ConvState:
    """QuantConv2d (quant_modules.py:675-736): no BN, optional bias, per_channel."""

    def __init__(self, conv, wbits=4, bias_bits=32):
        self.conv, self.wbits, self.bias_bits = conv, wbits, bias_bits
        self.weight_integer = self.bias_integer = self.w_sf = None

    def __call__(self, x, a_sf):
        conv = self.conv
        w2 = conv.weight.data.contiguous().view(conv.out_channels, -1)
        self.w_sf = sym_scale(self.wbits, w2.min(dim=1).values, w2.max(dim=1).values, True)
        self.weight_integer = quant_sym(conv.weight, self.wbits, self.w_sf)
        bias_sf = self.w_sf.view(1, -1) * a_sf.view(1, -1)
        if conv.bias is not None:
            self.bias_integer = quant_sym(conv.bias, self.bias_bits, bias_sf)
            b = self.bias_integer
        else:
            self.bias_integer = None
            b = torch.zeros_like(bias_sf.view(-1))
        x_int = x / a_sf.view(1, -1, 1, 1)
        out = F.conv2d(x_int, self.weight_integer, b, conv.stride, conv.padding, conv.dilation,
                       conv.groups) * bias_sf.view(1, -1, 1, 1)
        return out, self.w_sf


class LinearState:
    """QuantLinear (quant_modules.py:79-130), per_channel, 32-bit bias."""

    def __init__(self, linear, wbits=4, bias_bits=32):
        self.linear, self.wbits, self.bias_bits = linear, wbits, bias_bits
        self.weight_integer = self.bias_integer = self.w_sf = None
        self.acc_integer = None

    def __call__(self, x, a_sf):
        w = self.linear.weight
        wt = w.data.detach()
        w_min, _ = torch.min(wt, dim=1)
        w_max, _ = torch.max(wt, dim=1)
        self.w_sf = sym_scale(self.wbits, w_min, w_max, True)
        self.weight_integer = quant_sym(w, self.wbits, self.w_sf)
        bias_sf = self.w_sf.view(1, -1) * a_sf.view(1, -1)
        self.bias_integer = quant_sym(self.linear.bias, self.bias_bits, bias_sf)
        x_int = x / a_sf.view(1, -1)
        self.acc_integer = torch.round(F.linear(x_int, weight=self.weight_integer, bias=self.bias_integer))
        return self.acc_integer * bias_sf[0].view(1, -1)


def int_avgpool(x, sf, pool):
    """QuantAveragePool2d.forward quant_modules.py:585-602 + quant_utils.py:324-341 (trunc(x + 0.01))."""
    sf = sf.view(-1)
    x_int = torch.round(x / sf)
    x_int = pool(x_int)
    x_int = torch.trunc(x_int + 0.01)
    return x_int * sf, sf


# ----------------------------------------------------------------------------- graphs (L3)
class FakeQuantResNet:
    """Restatement of Q_ResNet18/50/101 + Q_ResBlockBn/Q_ResUnitBn (utils/models/q_resnet.py:16-316).

    ``float_model`` only has to expose the pytorchcv attribute layout (features.init_block.conv.{conv,bn},
    features.stageN.unitM.{body.convK.{conv,bn}, identity_conv, resize_identity}, output).
    Module names match the reference's ``named_modules()`` so bit configs and goldens key identically.
    """

    def __init__(self, arch, float_model, bit_config, momentum=0.99):
        self.arch = arch
        self.units_per_stage = {"resnet18": [2, 2, 2, 2], "resnet50": [3, 4, 6, 3], "resnet101": [3, 4, 23, 3]}[arch]
        self.bottleneck = arch != "resnet18"
        self.init_name = "quant_init_block_convbn" if arch == "resnet18" else "quant_init_convbn"
        self.acts, self.convs = {}, {}
        f = float_model.features
        self._add_act("quant_input")
        self._add_conv(self.init_name, f.init_block.conv)
        self._add_act("quant_act_int32")
        self.resize = {}
        for s, n in enumerate(self.units_per_stage):
            for u in range(n):
                p = "stage%d.unit%d" % (s + 1, u + 1)
                unit = getattr(getattr(f, "stage%d" % (s + 1)), "unit%d" % (u + 1))
                self.resize[p] = unit.resize_identity
                self._add_act(p + ".quant_act")
                self._add_conv(p + ".quant_convbn1", unit.body.conv1)
                self._add_act(p + ".quant_act1")
                self._add_conv(p + ".quant_convbn2", unit.body.conv2)
                if self.bottleneck:
                    self._add_act(p + ".quant_act2")
                    self._add_conv(p + ".quant_convbn3", unit.body.conv3)
                if unit.resize_identity:
                    self._add_conv(p + ".quant_identity_convbn", unit.identity_conv)
                self._add_act(p + ".quant_act_int32")
        self.pool7 = torch.nn.AvgPool2d(kernel_size=7, stride=1, padding=0)
        self._add_act("quant_act_output")
        self.fc = LinearState(float_model.output)
        # stamp (quant_train.py:264-299)
        seen = 0
        for name, v in bit_config.items():
            bits = v[0] if isinstance(v, tuple) else v
            if name in self.acts:
                a = self.acts[name]
                a.bits, a.momentum = bits, momentum
                a.mode = 'asymmetric' if bits == 4 else 'symmetric'
                seen += 1
            elif name in self.convs:
                self.convs[name].wbits = bits
                seen += 1
            elif name == "quant_output":
                self.fc.wbits = bits
                seen += 1
        assert seen == len(bit_config), (seen, len(bit_config))
        self.trace = None

    def _add_act(self, name):
        self.acts[name] = ActState()

    def _add_conv(self, name, cb):
        self.convs[name] = ConvBnState(cb.conv, cb.bn)

    def freeze(self):
        """freeze_model (quant_modules.py:739-758): stop range updates."""
        for a in self.acts.values():
            a.running_stat = False

    def load_act_ranges(self, ranges):
        for k, (mn, mx) in ranges.items():
            self.acts[k].x_min = torch.tensor([mn], dtype=torch.float32)
            self.acts[k].x_max = torch.tensor([mx], dtype=torch.float32)

    def act_ranges(self):
        return {k: (float(a.x_min), float(a.x_max)) for k, a in self.acts.items()}

    def _act(self, name, *args, **kw):
        out = self.acts[name](*args, **kw)
        if self.trace is not None and (self.trace_names is None or name in self.trace_names):
            self.trace[name] = torch.round(out[0] / out[1].view(-1)).to(torch.int64)
        return out

    def _unit(self, p, x, sf32):
        """Q_ResUnitBn.forward q_resnet.py:231-260 / Q_ResBlockBn.forward :291-316."""
        relu = F.relu
        if self.resize[p]:
            x, a_sf = self._act(p + ".quant_act", x, sf32)
            id_a_sf = a_sf.clone()
            identity, id_w_sf = self.convs[p + ".quant_identity_convbn"](x, a_sf)
        else:
            identity = x
            x, a_sf = self._act(p + ".quant_act", x, sf32)
        x, w_sf = self.convs[p + ".quant_convbn1"](x, a_sf)
        x = relu(x)
        x, a_sf = self._act(p + ".quant_act1", x, a_sf, w_sf)
        x, w_sf = self.convs[p + ".quant_convbn2"](x, a_sf)
        if self.bottleneck:
            x = relu(x)
            x, a_sf = self._act(p + ".quant_act2", x, a_sf, w_sf)
            x, w_sf = self.convs[p + ".quant_convbn3"](x, a_sf)
        x = x + identity
        if self.resize[p]:
            x, a_sf = self._act(p + ".quant_act_int32", x, a_sf, w_sf, identity, id_a_sf, id_w_sf)
        else:
            x, a_sf = self._act(p + ".quant_act_int32", x, a_sf, w_sf, identity, sf32, None)
        return relu(x), a_sf

    @torch.no_grad()
    def forward(self, x, trace=False):
        """Q_ResNet50.forward q_resnet.py:114-135 (== Q_ResNet18.forward :53-74).  trace: False, True (every QuantAct output as
        integers) or a collection of QuantAct names (only those: the full trace of a batch of 128 is several GB)."""
        self.trace = {} if trace else None
        self.trace_names = None if trace is True or not trace else set(trace)
        x, a_sf = self._act("quant_input", x)
        x, w_sf = self.convs[self.init_name](x, a_sf)
        x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
        x, a_sf = self._act("quant_act_int32", x, a_sf, w_sf)
        x = F.relu(x)
        for s, n in enumerate(self.units_per_stage):
            for u in range(n):
                x, a_sf = self._unit("stage%d.unit%d" % (s + 1, u + 1), x, a_sf)
        x, a_sf = int_avgpool(x, a_sf, self.pool7)
        x, a_sf = self._act("quant_act_output", x, a_sf)
        x = x.view(x.size(0), -1)
        return self.fc(x, a_sf)

    __call__ = forward

    def harvest(self):
        """Frozen integer parameters after a forward: what quant_train.py:665-670 dumps, plus act metadata."""
        convs = {}
        for k, c in self.convs.items():
            convs[k] = dict(weight_integer=c.weight_integer.clone(), bias_integer=c.bias_integer.clone(),
                            w_sf=c.w_sf.clone(), stride=c.conv.stride[0], pad=c.conv.padding[0])
        acts = {k: dict(scale=a.scale.clone(), bits=a.bits, mode=a.mode) for k, a in self.acts.items()}
        fc = dict(weight_integer=self.fc.weight_integer.clone(), bias_integer=self.fc.bias_integer.clone(),
                  w_sf=self.fc.w_sf.clone())
        return dict(arch=self.arch, convs=convs, acts=acts, fc=fc, resize=dict(self.resize),
                    units_per_stage=list(self.units_per_stage), bottleneck=self.bottleneck,
                    init_name=self.init_name)


class FakeQuantMobileNetV2:
    """Restatement of Q_MobileNetV2 + Q_LinearBottleneck (utils/models/q_mobilenetv2.py:12-212) on a float model with the
    pytorchcv attribute layout.  Module names match the reference's ``named_modules()`` (bit_config.py:3602-4202).
    What differs from the ResNets: ReLU6 in the float domain before the QuantAct that follows a convolution (a clamp at 6.0 of
    acc * scale, i.e. a per-channel clamp of the accumulator), depthwise convolutions (groups = channels), no activation after
    conv3 and none after the residual sum (the 16-bit stream is signed), the unit input itself as identity operand, and a 1x1
    QuantConv2d as classifier whose fp32 output are the logits."""

    def __init__(self, float_model, bit_config, momentum=0.99):
        self.acts, self.convs = {}, {}
        f = float_model.features
        self.acts["quant_input"] = ActState()
        self.convs["init_block"] = ConvBnState(f.init_block.conv, f.init_block.bn)
        self.acts["quant_act_int32"] = ActState()
        self.units = []
        s = 1
        while hasattr(f, "stage%d" % s):
            stage, u = getattr(f, "stage%d" % s), 1
            while hasattr(stage, "unit%d" % u):
                unit = getattr(stage, "unit%d" % u)
                p = "features.stage%d.unit%d" % (s, u)
                c1, c3 = unit.conv1.conv, unit.conv3.conv
                residual = (c1.in_channels == c3.out_channels) and unit.conv2.conv.stride[0] == 1
                self.units.append((p, residual))
                self.acts[p + ".quant_act"] = ActState()
                self.convs[p + ".conv1"] = ConvBnState(unit.conv1.conv, unit.conv1.bn)
                self.acts[p + ".quant_act1"] = ActState()
                self.convs[p + ".conv2"] = ConvBnState(unit.conv2.conv, unit.conv2.bn)
                self.acts[p + ".quant_act2"] = ActState()
                self.convs[p + ".conv3"] = ConvBnState(unit.conv3.conv, unit.conv3.bn)
                self.acts[p + ".quant_act_int32"] = ActState()
                u += 1
            s += 1
        self.acts["quant_act_before_final_block"] = ActState()
        self.convs["features.final_block"] = ConvBnState(f.final_block.conv, f.final_block.bn)
        self.acts["quant_act_int32_final"] = ActState()
        self.pool = f.final_pool
        self.acts["quant_act_output"] = ActState()
        self.out = ConvState(float_model.output)
        seen = 0
        for name, v in bit_config.items():       # stamp (quant_train.py:264-299)
            bits = v[0] if isinstance(v, tuple) else v
            if name in self.acts:
                a = self.acts[name]
                a.bits, a.momentum = bits, momentum
                a.mode = 'asymmetric' if bits == 4 else 'symmetric'
                seen += 1
            elif name in self.convs:
                self.convs[name].wbits = bits
                seen += 1
            elif name == "output":
                self.out.wbits = bits
                seen += 1
            elif name.rsplit(".", 1)[0] in self.convs and name.rsplit(".", 1)[1] in ("conv", "bn"):
                seen += 1      # the uniform tables also list `features.stage4.unit5.conv1.{conv,bn}` (bit_config.py:3690-3691): the
                               # trainer stamps attributes onto the wrapped nn.Conv2d / BatchNorm2d, which nothing reads
        assert seen == len(bit_config), (seen, len(bit_config))
        self.trace = None

    def freeze(self):
        for a in self.acts.values():
            a.running_stat = False

    def load_act_ranges(self, ranges):
        for k, (mn, mx) in ranges.items():
            self.acts[k].x_min = torch.tensor([mn], dtype=torch.float32)
            self.acts[k].x_max = torch.tensor([mx], dtype=torch.float32)

    def _act(self, name, *args, **kw):
        out = self.acts[name](*args, **kw)
        if self.trace is not None:
            self.trace[name] = torch.round(out[0] / out[1].view(-1)).to(torch.int64)
        return out

    def _unit(self, p, residual, x, sf32):
        """Q_LinearBottleneck.forward q_mobilenetv2.py:58-93."""
        identity = x
        x, a_sf = self._act(p + ".quant_act", x, sf32)
        x, w_sf = self.convs[p + ".conv1"](x, a_sf)
        x = F.relu6(x)
        x, a_sf = self._act(p + ".quant_act1", x, a_sf, w_sf)
        x, w_sf = self.convs[p + ".conv2"](x, a_sf)
        x = F.relu6(x)
        x, a_sf = self._act(p + ".quant_act2", x, a_sf, w_sf)
        x, w_sf = self.convs[p + ".conv3"](x, a_sf)
        if residual:
            x = x + identity
            return self._act(p + ".quant_act_int32", x, a_sf, w_sf, identity, sf32, None)
        return self._act(p + ".quant_act_int32", x, a_sf, w_sf)

    @torch.no_grad()
    def forward(self, x, trace=False):
        """Q_MobileNetV2.forward q_mobilenetv2.py:182-212."""
        self.trace = {} if trace else None
        x, a_sf = self._act("quant_input", x)
        x, w_sf = self.convs["init_block"](x, a_sf)
        x = F.relu6(x)
        x, a_sf = self._act("quant_act_int32", x, a_sf, w_sf)
        for p, residual in self.units:
            x, a_sf = self._unit(p, residual, x, a_sf)
        x, a_sf = self._act("quant_act_before_final_block", x, a_sf)
        x, w_sf = self.convs["features.final_block"](x, a_sf)
        x = F.relu6(x)
        x, a_sf = self._act("quant_act_int32_final", x, a_sf, w_sf)
        x, a_sf = int_avgpool(x, a_sf, self.pool)
        x, a_sf = self._act("quant_act_output", x, a_sf)
        x, _ = self.out(x, a_sf)
        return x.view(x.size(0), -1)

    __call__ = forward

    def harvest(self):
        """Frozen integer parameters after a forward (layout of FakeQuantResNet.harvest, plus groups and the unit list)."""
        convs = {}
        for k, c in self.convs.items():
            convs[k] = dict(weight_integer=c.weight_integer.clone(), bias_integer=c.bias_integer.clone(), w_sf=c.w_sf.clone(),
                            stride=c.conv.stride[0], pad=c.conv.padding[0], groups=c.conv.groups)
        acts = {k: dict(scale=a.scale.clone(), bits=a.bits, mode=a.mode) for k, a in self.acts.items()}
        out = dict(weight_integer=self.out.weight_integer.clone(), w_sf=self.out.w_sf.clone(),
                   bias_integer=None if self.out.bias_integer is None else self.out.bias_integer.clone())
        return dict(arch="mobilenetv2_w1", convs=convs, acts=acts, output=out, units=list(self.units),
                    pool=int(self.pool.kernel_size if isinstance(self.pool.kernel_size, int) else self.pool.kernel_size[0]))
