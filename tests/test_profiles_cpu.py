"""profiles/ncu_traffic.json is what bench.py's roofline.traffic is read from: its kernel names must be the family names bench.py
reports, and an entry is only used for the build it was measured on (digest of the kernel sources)."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAMILIES = {"conv_tc_kernel", "conv_halo_kernel", "conv1x1_kernel", "conv_dual_kernel", "stem_tc_kernel"}


def test_ncu_traffic_entries_use_the_bench_family_names():
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
        table = json.load(f)
    assert table, "no ncu traffic recorded"
    for key, ent in table.items():
        assert re.fullmatch(r"resnet\d+:[a-z0-9_.]+:\d+", key), key
        assert re.fullmatch(r"[0-9a-f]{16}", ent["build"]), ent["build"]
        assert set(ent["kernels"]) <= FAMILIES, set(ent["kernels"]) - FAMILIES
        for fam in ent["kernels"].values():
            assert fam["traffic_bytes_per_launch"] > 0 and fam["launches"] > 0
    # the family names bench.py maps its per-launch labels to
    src = open(os.path.join(ROOT, "bench.py")).read()
    for name in FAMILIES:
        assert '"%s"' % name in src, name


def test_stale_traffic_is_refused():
    """bench.py compares the entry's digest with the digest of the library it runs (never a number of another build)."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'ent.get("build") == build_digest()' in src
