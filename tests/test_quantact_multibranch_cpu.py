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


def test_integer_restatement_of_the_multibranch_requant():
    """oracle.int_ref.multibranch_requant (exact integers, one dyadic pair per branch) reproduces the reference's QuantAct on
    concatenated branches: the arithmetic an integer engine needs for the InceptionV3 concatenation edges is pinned."""
    import json
    import numpy as np
    from oracle import int_ref as ir
    g = load_golden("kat_multibranch.npz")
    scales, chans = [0.021, 0.0173, 0.05], [3, 5, 2]           # the generating script's branches (tests/golden/make_golden.py)
    for i, spec in enumerate(json.loads(str(g["specs"]))):
        x, y, sf = g["mb_%d_x" % i], g["mb_%d_y" % i], np.float32(g["mb_%d_sf" % i][0])
        x_int = np.concatenate([np.rint(x[:, c0:c0 + c] / np.float32(s)) for s, c, c0 in zip(scales, chans, np.cumsum([0] + chans[:-1]))], axis=1)
        q = ir.multibranch_requant(x_int.transpose(0, 2, 3, 1).astype(np.int64), scales, chans, sf, spec["bits"], spec["mode"])
        want = np.rint(y / sf).astype(np.int64).transpose(0, 2, 3, 1)
        assert np.array_equal(q, want), (i, spec)
