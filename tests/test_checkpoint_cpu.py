"""SURVEY.md 8(f) rank 1 on CPU: HAWQ checkpoint formats in and out of the engine (hawq_b200/checkpoint.py).

* TVM parameter export against goldens produced by executing the reference's own save_weights / save_bias / pack functions
  (tests/golden/make_tvm_export_golden.py); when /root/reference is present the reference functions are also run live.
* quantized_checkpoint.pth.tar round trip: integers saved from one frozen model become the plan of a model whose float
  weights are different -> identical logits.
* --resume-quantize key filter, and plan invalidation by load_state_dict / unfix (SURVEY 8(b) lifecycle row).
The C-ABI launchers are replaced by the numpy ABI model (tests/abi_model.py)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import hawq_b200 as hb
from hawq_b200 import checkpoint as ck
from hawq_b200.build import build_library
from hawq_b200.synthetic import synthetic_batch
from tests import abi_model

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "kat_tvm_export.npz"))


@pytest.fixture(scope="session", autouse=True)
def _lib():
    build_library()


def _gen():
    spec = importlib.util.spec_from_file_location("make_tvm_export_golden", os.path.join(HERE, "golden", "make_tvm_export_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_int4_pack_matches_reference_kat():
    a = GOLD["kat/pack_in"]
    p = ck.pack_int4_tvm(a)
    assert p.dtype == np.int32 and np.array_equal(p, GOLD["kat/pack_out"])
    assert np.array_equal(ck.unpack_int4_tvm(p), GOLD["kat/unpack_out"])
    assert np.array_equal(ck.unpack_int4_tvm(p), a & 0xF)


@pytest.mark.parametrize("kd,bits", [("int8", 8), ("int4", 4)])
def test_tvm_export_matches_reference_golden(kd, bits):
    gen = _gen()
    cp = gen.fake_checkpoint(seed=7 + bits, bits=bits)
    weights, bias = ck.export_tvm_params(cp, kernel_dtype=kd, num_stages=gen.NUM_STAGES, units=gen.UNITS)
    want_w = {k.split("/", 2)[2]: GOLD[k] for k in GOLD.files if k.startswith(kd + "/w/")}
    want_b = {k.split("/", 2)[2]: GOLD[k] for k in GOLD.files if k.startswith(kd + "/b/")}
    assert set(weights) == set(want_w) and set(bias) == set(want_b)
    for k in want_w:
        assert weights[k].dtype == want_w[k].dtype and weights[k].shape == want_w[k].shape and np.array_equal(weights[k], want_w[k]), k
    for k in want_b:
        assert bias[k].dtype == want_b[k].dtype and bias[k].shape == want_b[k].shape and np.array_equal(bias[k], want_b[k]), k
    if os.path.exists(gen.REF):                     # build container: the reference's functions, live
        ref_w, ref_b = gen.run_reference(cp, kd)
        for k in ref_w:
            assert np.array_equal(ref_w[k], weights[k]), k
        for k in ref_b:
            assert np.array_equal(ref_b[k], bias[k]), k


def test_tvm_export_mixed_precision_and_files(tmp_path):
    gen = _gen()
    cp = gen.fake_checkpoint(seed=1, bits=4)
    names = ["stage%d_unit%d_qconv%d_weight" % (i + 1, j + 1, k + 1) for i in range(2) for j in range(gen.UNITS[i]) for k in range(3)]
    names += ["stage%d_unit1_qsc_weight" % (i + 1) for i in range(2)]
    mixed = {n: ("int4" if idx % 2 else "int8") for idx, n in enumerate(names)}
    weights, bias = ck.save_tvm_params(cp, str(tmp_path), kernel_dtype=mixed, num_stages=2, units=gen.UNITS)
    for n, dt in mixed.items():
        assert weights[n].dtype == (np.int32 if dt == "int4" else np.int8), n
    back = np.load(os.path.join(tmp_path, "weights.npy"), allow_pickle=True).item()
    assert set(back) == set(weights) and all(np.array_equal(back[k], weights[k]) for k in back)
    assert np.load(os.path.join(tmp_path, "bias.npy"), allow_pickle=True).item()["fc_bias"].shape == (10,)


def test_resume_quantize_key_filter():
    sd = {"module.stage1.unit1.quant_convbn1.conv.weight": 1, "module.stage1.unit1.quant_convbn1.bn.num_batches_tracked": 2,
          "module.stage1.unit1.quant_convbn1.weight_integer": 3, "module.quant_output.bias_integer": 4, "module.quant_input.x_min": 5}
    assert ck.filter_resume_state_dict(sd) == {"stage1.unit1.quant_convbn1.conv.weight": 1, "quant_input.x_min": 5}


def test_quantized_checkpoint_round_trip_and_plan_invalidation(monkeypatch, tmp_path):
    abi_model.install_cpu_backend(monkeypatch)
    x = synthetic_batch(1, 3)
    qa = hb.build_synthetic_qresnet("resnet18", "uniform4", calib_batch=1)
    with torch.no_grad():
        la = qa(x)
    path = os.path.join(tmp_path, "quantized_checkpoint.pth.tar")
    saved = ck.save_quantized_checkpoint(qa, path)
    assert set(saved) == {"convbn_scaling_factor", "fc_scaling_factor", "weight_integer", "bias_integer", "act_scaling_factor"}
    assert all(k.startswith("module.") for g in saved.values() for k in g)
    assert len(saved["weight_integer"]) == 21 and len(saved["act_scaling_factor"]) == 27

    # (1) the integer checkpoint is the whole plan: a skeleton with DIFFERENT float weights reproduces the logits bit for bit
    from hawq_b200.synthetic import synthetic_float_resnet
    qb = hb.q_resnet18(synthetic_float_resnet("resnet18", 123))
    hb.stamp_bit_config(qb, hb.get_bit_config("resnet18", "uniform4"))
    qb.eval()
    n = ck.apply_integer_checkpoint(qb, path)
    assert n == 21 + 27
    with torch.no_grad():
        lb = qb(x)
    assert torch.equal(la, lb)

    # (2) --resume-quantize path: float weights + ranges through load_state_dict, integers re-derived by the engine
    qc = hb.q_resnet18(synthetic_float_resnet("resnet18", 123))
    hb.stamp_bit_config(qc, hb.get_bit_config("resnet18", "uniform4"))
    qc.eval()
    with torch.no_grad():
        qc(x)                                         # calibration forward (float path)
    hb.freeze_model(qc)
    with torch.no_grad():
        lc0 = qc(x)                                   # builds a plan from the seed-123 weights
    assert not torch.equal(lc0, la)
    res = ck.load_quantized_checkpoint(qc, {"state_dict": {"module." + k: v for k, v in qa.state_dict().items()}})
    assert not res.unexpected_keys
    with torch.no_grad():
        lc = qc(x)                                    # load_state_dict wrote in place: the cached plan must not be reused
    assert torch.equal(lc, la)

    # (3) unfix() drops the plan, fix() + forward rebuilds it
    assert any("_hawq_cache" in m.__dict__ for m in qc.modules() if isinstance(m, hb.QuantBnConv2d))
    hb.unfreeze_model(qc)
    assert not any("_hawq_cache" in m.__dict__ for m in qc.modules() if isinstance(m, (hb.QuantBnConv2d, hb.QuantLinear)))
    hb.freeze_model(qc)
    with torch.no_grad():
        assert torch.equal(qc(x), la)

    # (4) misuse is reported
    with pytest.raises(KeyError):
        ck.apply_integer_checkpoint(qb, {"weight_integer": {}})
    bad = {g: dict(v) for g, v in saved.items()}
    k0 = "module.stage1.unit1.quant_convbn1.weight_integer"
    bad["weight_integer"][k0] = bad["weight_integer"][k0] * 100
    with pytest.raises(ValueError, match="does not fit"):
        ck.apply_integer_checkpoint(qb, bad)
