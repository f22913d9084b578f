#!/usr/bin/env python
"""Timeline of one conv1x1 launch (CTA 0): python tools/c1_trace.py N H Cin Cout kind(req|res) a_bits"""
import ctypes as C
import os
import sys

os.environ["HAWQ_B200_HALO_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from hawq_b200 import _lib, ops  # noqa: E402
from hawq_b200._lib import EPI_REQUANT, EPI_RESIDUAL, dyadic  # noqa: E402

n, hh, cin, cout = [int(v) for v in sys.argv[1:5]]
kind, a_bits = sys.argv[5], int(sys.argv[6])
dev = "cuda:0"
m = n * hh * hh
x = (torch.randint(-128, 128, (m * cin,), dtype=torch.int8) if a_bits == 8 else torch.randint(0, 256, (m * cin // 2,), dtype=torch.uint8)).to(dev)
lim = 128 if a_bits == 8 else 8
wd = ops.upload_weights(torch.randint(-lim, lim, (cout, 1, 1, cin), dtype=torch.int8), dev)
me = dyadic(0.003)
chan = ops.make_chan(np.zeros(cout, dtype=np.int64), [me[0]] * cout, [me[1]] * cout).to(dev)
d = ops.conv_desc(n, hh, hh, cin, cout, 1, 1, 1, 0, a_bits, 1)
if kind == "req":
    ep = ops.epilogue(EPI_REQUANT, relu=1, out_bits=a_bits, clamp=(0, 15) if a_bits == 4 else (-128, 127), flags=1)
    kw = dict(out=torch.empty(m * cout * a_bits // 8, dtype=torch.uint8, device=dev))
else:
    ep = ops.epilogue(EPI_RESIDUAL, relu=1, res_kind=0, res_bits=16, res_me=dyadic(0.37), y_bits=16, low_bits=a_bits,
                      low_me=dyadic(0.004), low_clamp=(0, 15) if a_bits == 4 else (-128, 127), flags=1)
    kw = dict(res=torch.randint(0, 30000, (m * cout,), dtype=torch.int16).to(dev), out=torch.empty(m * cout, dtype=torch.int16, device=dev),
              out_low=torch.empty(m * cout * a_bits // 8, dtype=torch.uint8, device=dev))
for _ in range(3):
    ops.conv2d(x, d, ep, wd, chan, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda._sleep(int(2e6)); e0.record(); ops.conv2d(x, d, ep, wd, chan, **kw); e1.record(); torch.cuda.synchronize()
print("%s a%d N=%d H=%d %d->%d: launch %.1f us" % (kind, a_bits, n, hh, cin, cout, e0.elapsed_time(e1) * 1e3))
buf = (C.c_int64 * (4 * 48 * 4))()
got = _lib.load().hawq_debug_c1_trace(C.cast(buf, C.c_void_p), 4 * 48 * 4)
t = np.array(buf[:got], dtype=np.int64).reshape(4, 48, 4)
t0 = t[t > 0].min()
for role, name, evs in ((0, "producer/converter", "wait_data got_data got_stage done"), (1, "mma", "wait_tempty got_tempty got_afull committed"),
                        (2, "epilogue", "wait_res got_acc math_done stores_issued"), (3, "res loader", "wait_empty got_empty - -")):
    print(name, "(cycles since first stamp; %s)" % evs)
    for i in list(range(4)) + list(range(20, 26)):
        print("  %2d " % i + " ".join("%8d" % (v - t0 if v > 0 else -1) for v in t[role, i]))
