"""Quantized MobileNetV2 graph on the hawq_b200 modules (module API of reference ``utils/models/q_mobilenetv2.py:12-262``).

Same module names, constructor reads and forward wiring as the reference's ``Q_LinearBottleneck`` / ``Q_MobileNetV2`` (so its
bit configs, ``bit_config.py:3602-4202``, and checkpoints key identically):

  unit:  quant_act (16 -> low bit, from the previous unit's scale) -> conv1 1x1 + ReLU6 -> quant_act1 -> conv2 depthwise 3x3 +
         ReLU6 -> quant_act2 -> conv3 1x1 (linear) -> quant_act_int32 (case 0, or case 1 with the unit input as identity: no ReLU)
  net:   quant_input -> init_block 3x3/2 + ReLU6 -> quant_act_int32 -> units -> quant_act_before_final_block -> final_block 1x1 +
         ReLU6 -> quant_act_int32_final -> final_pool -> quant_act_output -> output (QuantConv2d 1x1) -> logits

STATUS: the graph runs un-frozen (calibration / evaluation of the fake-quant arithmetic, same numbers
as the reference, tests/test_mobilenetv2_cpu.py) and the oracle for the frozen integers is pinned to the reference
(oracle/fakequant.py FakeQuantMobileNetV2, tests/golden/net_mobilenetv2_w1_*.npz).  The frozen integer path is NOT built: it
needs a depthwise kernel, ReLU6 as a per-channel output clamp and a signed residual stream; a frozen forward raises
NotImplementedError from the convolution planner instead of computing something else.
"""
import torch.nn as nn

from .modules import QuantAct, QuantAveragePool2d, QuantBnConv2d, QuantConv2d
from .synthetic import MOBILENETV2_CHANNELS


class Q_LinearBottleneck(nn.Module):
    def __init__(self, model, in_channels, out_channels, stride, expansion, remove_exp_conv=False):
        super().__init__()
        self.residual = (in_channels == out_channels) and (stride == 1)
        self.use_exp_conv = expansion or (not remove_exp_conv)
        self.activatition_func = nn.ReLU6()          # (attribute name as in the reference)
        self.quant_act = QuantAct()
        if self.use_exp_conv:
            self.conv1 = QuantBnConv2d()
            self.conv1.set_param(model.conv1.conv, model.conv1.bn)
            self.quant_act1 = QuantAct()
        self.conv2 = QuantBnConv2d()
        self.conv2.set_param(model.conv2.conv, model.conv2.bn)
        self.quant_act2 = QuantAct()
        self.conv3 = QuantBnConv2d()
        self.conv3.set_param(model.conv3.conv, model.conv3.bn)
        self.quant_act_int32 = QuantAct()

    def forward(self, x, scaling_factor_int32=None):
        identity = x if self.residual else None
        x, a_sf = self.quant_act(x, scaling_factor_int32, None, None, None, None)
        if self.use_exp_conv:
            x, w_sf = self.conv1(x, a_sf)
            x = self.activatition_func(x)
            x, a_sf = self.quant_act1(x, a_sf, w_sf, None, None)
        x, w_sf = self.conv2(x, a_sf)
        x = self.activatition_func(x)
        x, a_sf = self.quant_act2(x, a_sf, w_sf, None, None)
        x, w_sf = self.conv3(x, a_sf)               # linear: no activation after the projection
        if self.residual:
            x = x + identity
            return self.quant_act_int32(x, a_sf, w_sf, identity, scaling_factor_int32, None)
        return self.quant_act_int32(x, a_sf, w_sf, None, None, None)


class Q_MobileNetV2(nn.Module):
    def __init__(self, model, channels=None, remove_exp_conv=False):
        super().__init__()
        self.channels = [list(c) for c in (channels or MOBILENETV2_CHANNELS)]
        self.activatition_func = nn.ReLU6()
        self.quant_input = QuantAct()
        self.init_block = QuantBnConv2d()
        self.init_block.set_param(model.features.init_block.conv, model.features.init_block.bn)
        self.quant_act_int32 = QuantAct()
        self.features = nn.Sequential()
        cin = model.features.init_block.conv.out_channels
        for i, stage_channels in enumerate(self.channels):
            stage = nn.Sequential()
            src = getattr(model.features, "stage%d" % (i + 1))
            for j, cout in enumerate(stage_channels):
                stride = 2 if (j == 0 and i != 0) else 1
                stage.add_module("unit%d" % (j + 1), Q_LinearBottleneck(getattr(src, "unit%d" % (j + 1)), cin, cout, stride,
                                                                         expansion=(i != 0 or j != 0), remove_exp_conv=remove_exp_conv))
                cin = cout
            self.features.add_module("stage%d" % (i + 1), stage)
        self.quant_act_before_final_block = QuantAct()
        self.features.add_module("final_block", QuantBnConv2d())
        self.features.final_block.set_param(model.features.final_block.conv, model.features.final_block.bn)
        self.quant_act_int32_final = QuantAct()
        self.features.add_module("final_pool", QuantAveragePool2d())
        self.features.final_pool.set_param(model.features.final_pool)
        self.quant_act_output = QuantAct()
        self.output = QuantConv2d()
        self.output.set_param(model.output)

    def forward(self, x):
        if self.init_block.fix_flag:
            raise NotImplementedError("hawq_b200 has no frozen (integer) path for MobileNetV2: the engine lacks a depthwise kernel, the ReLU6 "
                                      "clamp and a signed residual stream (DESIGN.md section 2, row f3). Un-frozen forwards work.")
        x, a_sf = self.quant_input(x)
        x, w_sf = self.init_block(x, a_sf)
        x = self.activatition_func(x)
        x, a_sf = self.quant_act_int32(x, a_sf, w_sf, None, None)
        for i, stage_channels in enumerate(self.channels):
            stage = getattr(self.features, "stage%d" % (i + 1))
            for j in range(len(stage_channels)):
                x, a_sf = getattr(stage, "unit%d" % (j + 1))(x, a_sf)
        x, a_sf = self.quant_act_before_final_block(x, a_sf, None, None, None, None)
        x, w_sf = self.features.final_block(x, a_sf)
        x = self.activatition_func(x)
        x, a_sf = self.quant_act_int32_final(x, a_sf, w_sf, None, None, None)
        x = self.features.final_pool(x, a_sf)
        x, a_sf = self.quant_act_output(x, a_sf, None, None, None, None)
        x, _ = self.output(x, a_sf)
        return x.view(x.size(0), -1)


def q_mobilenetv2_w1(model):
    """Quantized MobileNetV2-1.0 from a float model with the pytorchcv attribute layout (reference ``q_mobilenetv2_w1``)."""
    return Q_MobileNetV2(model)
