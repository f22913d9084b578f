"""Extract the per-layer arrays of the reference's bit-allocation notebook (ILP.ipynb cells 4 and 17: Hessian traces, weight
perturbations, parameters, BOPS, T4 latencies) into tests/golden/ilp_data.json.  Data only, read from the unmodified notebook.
Run in the build container:  python tests/golden/make_ilp_data.py"""
import json
import os
import re

REF = "/root/reference/ILP.ipynb"
HERE = os.path.dirname(os.path.abspath(__file__))


def arrays(src):
    out = {}
    for m in re.finditer(r"^(\w+)\s*=\s*np\.array\(\[(.*?)\]\)", src, re.S | re.M):
        out[m.group(1)] = [float(v) for v in m.group(2).replace("\n", " ").split(",") if v.strip()]
    return out


if __name__ == "__main__":
    nb = json.load(open(REF))
    data = {"resnet18": arrays("".join(nb["cells"][4]["source"])), "resnet50": arrays("".join(nb["cells"][17]["source"]))}
    assert {k: len(v) for k, v in data["resnet18"].items()} == {k: 19 for k in data["resnet18"]}
    assert {k: len(v) for k, v in data["resnet50"].items()} == {k: 52 for k in data["resnet50"]}
    json.dump(data, open(os.path.join(HERE, "ilp_data.json"), "w"), indent=0)
    print("wrote ilp_data.json")
