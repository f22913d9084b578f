"""Un-frozen QuantAct on concatenated branches (the Inception-style input `(x, [scale_i], [channels_i])`,
quant_modules.py:219-223,275-286) against vectors produced by the reference's own QuantAct."""
import json

import pytest
import torch

import hawq_b200 as hb
from tests.util import load_golden


def test_multibranch_quantact_matches_reference_vectors():
    g = load_golden("kat_multibranch.npz")
    specs = json.loads(str(g["specs"]))
    scales = [torch.tensor([0.021]), torch.tensor([0.0173]), torch.tensor([0.05])]
    chans = [3, 5, 2]
    for i, spec in enumerate(specs):
        act = hb.QuantAct(activation_bit=spec["bits"], quant_mode=spec["mode"])
        x = torch.from_numpy(g["mb_%d_x" % i])
        y, sf = act((x.clone(), [s.clone() for s in scales], chans))
        assert torch.equal(sf.view(-1), torch.from_numpy(g["mb_%d_sf" % i]))
        assert torch.equal(y, torch.from_numpy(g["mb_%d_y" % i])), spec
    with pytest.raises(ValueError):
        hb.QuantAct()((x, [s.clone() for s in scales], [3, 5]))
    frozen = hb.QuantAct()
    frozen((x, [s.clone() for s in scales], chans))
    frozen.fix()
    with pytest.raises(NotImplementedError):
        frozen((x, [s.clone() for s in scales], chans))
