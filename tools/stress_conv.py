#!/usr/bin/env python
"""Stress the fast convolution kernels on ResNet-50 layer shapes: repeated launches must agree with each other and with the
generic IMMA kernel (no ratio promise -> conv_igemm).  usage: python tools/stress_conv.py [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from hawq_b200 import ops  # noqa: E402
from hawq_b200._lib import EPI_REQUANT, EPI_RESIDUAL, dyadic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda:0"
g = torch.Generator().manual_seed(0)
SHAPES = [  # H, Cin, Cout, kind
    (56, 64, 64, "req"), (56, 256, 64, "req"), (56, 64, 256, "res"), (28, 512, 128, "req"), (28, 128, 512, "res"),
    (14, 1024, 256, "req"), (14, 256, 1024, "res"), (7, 2048, 512, "req"), (7, 512, 2048, "res"),
]
bad_total = 0
for a_bits in (8, 4):
    for hh, cin, cout, kind in SHAPES:
        m = B * hh * hh
        if a_bits == 8:
            x = torch.randint(-128, 128, (m * cin,), dtype=torch.int8, generator=g).to(dev)
        else:
            x = torch.randint(0, 256, (m * cin // 2,), dtype=torch.uint8, generator=g).to(dev)
        lim = 128 if a_bits == 8 else 8
        wt = torch.randint(-lim, lim, (cout, 1, 1, cin), dtype=torch.int8, generator=g)
        wd = ops.upload_weights(wt, dev)
        rs = np.random.RandomState(cin + cout)
        me = [dyadic(float(np.exp(rs.uniform(np.log(1e-4), np.log(0.02))))) for _ in range(cout)]
        chan = ops.make_chan(rs.randint(-5000, 5000, size=cout), [a for a, _ in me], [b for _, b in me]).to(dev)
        d = ops.conv_desc(B, hh, hh, cin, cout, 1, 1, 1, 0, a_bits, 1)
        outs = []
        for flags in (1, 1, 1, 1, 0):
            if kind == "req":
                ep = ops.epilogue(EPI_REQUANT, relu=1, out_bits=a_bits, clamp=(0, 15) if a_bits == 4 else (-128, 127), flags=flags)
                out = torch.full((m * cout * a_bits // 8,), 0x55, dtype=torch.uint8, device=dev)
                ops.conv2d(x, d, ep, wd, chan, out=out)
                outs.append((out,))
            else:
                res = torch.randint(0, 30000, (m * cout,), dtype=torch.int16, generator=g).to(dev) if not outs else res
                ep = ops.epilogue(EPI_RESIDUAL, relu=1, res_kind=0, res_bits=16, res_me=dyadic(0.37), y_bits=16, low_bits=a_bits,
                                  low_me=dyadic(0.004), low_clamp=(0, 15) if a_bits == 4 else (-128, 127), flags=flags)
                y = torch.full((m * cout,), 0x5555, dtype=torch.int16, device=dev)
                low = torch.full((m * cout * a_bits // 8,), 0x55, dtype=torch.uint8, device=dev)
                ops.conv2d(x, d, ep, wd, chan, res=res, out=y, out_low=low)
                outs.append((y, low))
        torch.cuda.synchronize()
        ref = outs[-1]
        for i, o in enumerate(outs[:-1]):
            for j, (a, b) in enumerate(zip(o, ref)):
                if not torch.equal(a, b):
                    diff = (a != b).nonzero().flatten()
                    per_row = a.numel() // m
                    rows = (diff // per_row).unique()
                    cols = (diff % per_row).unique()
                    bad_total += 1
                    print("MISMATCH a%d %s H=%d %d->%d run %d out %d: %d elems, %d rows (first %s, last %s), cols %s..%s, row%%128 in %s" % (
                        a_bits, kind, hh, cin, cout, i, j, diff.numel(), rows.numel(), rows[:4].tolist(), rows[-2:].tolist(),
                        int(cols.min()), int(cols.max()), sorted(set((rows % 128).tolist()))[:12]))
        print("a%d %s H=%d %d->%d ok" % (a_bits, kind, hh, cin, cout) if True else "")
print("status word", ops.get_status(0), "mismatching outputs:", bad_total)
