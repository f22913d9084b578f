"""Quantized ResNet-18/50/101 graphs with the reference's module names.

Counterpart of reference ``utils/models/q_resnet.py`` (Q_ResNet18 :16-74, Q_ResNet50 :77-135, Q_ResNet101 :138-196,
Q_ResUnitBn :199-260, Q_ResBlockBn :263-316, factories :319-331).  ``named_modules()`` / ``state_dict()`` keys are
identical (``quant_input``, ``quant_init[_block]_convbn``, ``quant_act_int32``, ``stageS.unitU.quant_*``,
``final_pool``, ``quant_act_output``, ``quant_output``), so the reference's bit configs and checkpoints apply as is.
One generic network class covers the three depths; a unit is either *basic* (two 3x3 convs) or *bottleneck*.

The forward is written against the (tensor, scale) convention only, so the same code serves the un-frozen float
calibration pass and the frozen integer pass (where the tensors are ``IntActivation`` payloads and ReLU / max-pool /
add are recorded lazily and fused into the convolution epilogues).
"""
import torch.nn as nn
import torch.nn.functional as F

from .modules import QuantAct, QuantAveragePool2d, QuantBnConv2d, QuantLinear


class QResidualUnit(nn.Module):
    """quant_act -> [identity conv] -> conv1 -> ReLU -> act1 -> conv2 [-> ReLU -> act2 -> conv3] -> +identity ->
    quant_act_int32 (case 1) -> ReLU."""

    def __init__(self, bottleneck):
        super().__init__()
        self.bottleneck = bottleneck

    def set_param(self, unit):
        self.resize_identity = unit.resize_identity
        self.quant_act = QuantAct()
        body = unit.body
        self.quant_convbn1 = QuantBnConv2d()
        self.quant_convbn1.set_param(body.conv1.conv, body.conv1.bn)
        self.quant_act1 = QuantAct()
        self.quant_convbn2 = QuantBnConv2d()
        self.quant_convbn2.set_param(body.conv2.conv, body.conv2.bn)
        if self.bottleneck:
            self.quant_act2 = QuantAct()
            self.quant_convbn3 = QuantBnConv2d()
            self.quant_convbn3.set_param(body.conv3.conv, body.conv3.bn)
        if self.resize_identity:
            self.quant_identity_convbn = QuantBnConv2d()
            self.quant_identity_convbn.set_param(unit.identity_conv.conv, unit.identity_conv.bn)
        self.quant_act_int32 = QuantAct()

    def forward(self, x, scaling_factor_int32=None):
        residual_in = x                                   # 16-bit stream, post-ReLU
        x, a_sf = self.quant_act(x, scaling_factor_int32)
        if self.resize_identity:
            id_a_sf = a_sf.clone()
            identity, id_w_sf = self.quant_identity_convbn(x, a_sf)
        else:
            identity, id_a_sf, id_w_sf = residual_in, scaling_factor_int32, None
        x, w_sf = self.quant_convbn1(x, a_sf)
        x, a_sf = self.quant_act1(F.relu(x), a_sf, w_sf)
        x, w_sf = self.quant_convbn2(x, a_sf)
        if self.bottleneck:
            x, a_sf = self.quant_act2(F.relu(x), a_sf, w_sf)
            x, w_sf = self.quant_convbn3(x, a_sf)
        x = x + identity
        x, a_sf = self.quant_act_int32(x, a_sf, w_sf, identity, id_a_sf, id_w_sf)
        return F.relu(x), a_sf


class Q_ResBlockBn(QResidualUnit):
    def __init__(self):
        super().__init__(bottleneck=False)


class Q_ResUnitBn(QResidualUnit):
    def __init__(self):
        super().__init__(bottleneck=True)


class QResNet(nn.Module):
    def __init__(self, model, units_per_stage, bottleneck, init_name):
        super().__init__()
        features = model.features
        self.init_name = init_name
        self.channel = list(units_per_stage)
        self.quant_input = QuantAct()
        init = QuantBnConv2d()
        init.set_param(features.init_block.conv.conv, features.init_block.conv.bn)
        setattr(self, init_name, init)
        self.quant_act_int32 = QuantAct()
        self.pool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.act = nn.ReLU()
        for s, n in enumerate(self.channel):
            stage = nn.Module()
            src_stage = getattr(features, "stage%d" % (s + 1))
            for u in range(n):
                q = Q_ResUnitBn() if bottleneck else Q_ResBlockBn()
                q.set_param(getattr(src_stage, "unit%d" % (u + 1)))
                setattr(stage, "unit%d" % (u + 1), q)
            setattr(self, "stage%d" % (s + 1), stage)
        self.final_pool = QuantAveragePool2d(kernel_size=7, stride=1)
        self.quant_act_output = QuantAct()
        self.quant_output = QuantLinear()
        self.quant_output.set_param(model.output)

    def units(self):
        for s, n in enumerate(self.channel):
            stage = getattr(self, "stage%d" % (s + 1))
            for u in range(n):
                yield "stage%d.unit%d" % (s + 1, u + 1), getattr(stage, "unit%d" % (u + 1))

    def forward(self, x):
        x, a_sf = self.quant_input(x)
        x, w_sf = getattr(self, self.init_name)(x, a_sf)
        x = self.pool(x)
        x, a_sf = self.quant_act_int32(x, a_sf, w_sf)
        x = self.act(x)
        for _, unit in self.units():
            x, a_sf = unit(x, a_sf)
        x = self.final_pool(x, a_sf)
        x, a_sf = self.quant_act_output(x, a_sf)
        x = x.view(x.size(0), -1)
        return self.quant_output(x, a_sf)


class Q_ResNet18(QResNet):
    def __init__(self, model):
        super().__init__(model, [2, 2, 2, 2], False, "quant_init_block_convbn")


class Q_ResNet50(QResNet):
    def __init__(self, model):
        super().__init__(model, [3, 4, 6, 3], True, "quant_init_convbn")


class Q_ResNet101(QResNet):
    def __init__(self, model):
        super().__init__(model, [3, 4, 23, 3], True, "quant_init_convbn")


def q_resnet18(model):
    return Q_ResNet18(model)


def q_resnet50(model):
    return Q_ResNet50(model)


def q_resnet101(model):
    return Q_ResNet101(model)


quantize_arch_dict = {"resnet18": q_resnet18, "resnet50": q_resnet50, "resnet101": q_resnet101}
