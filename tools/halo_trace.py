#!/usr/bin/env python
"""Timeline of one conv_halo launch (CTA 0): HAWQ_B200_HALO_TRACE=1 python tools/halo_trace.py [N H W Cin Cout]"""
import ctypes as C
import os
import sys

os.environ["HAWQ_B200_HALO_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from hawq_b200 import _lib, ops  # noqa: E402
from hawq_b200._lib import EPI_REQUANT  # noqa: E402

n, h, w, cin, cout = [int(v) for v in sys.argv[1:6]] if len(sys.argv) >= 6 else (128, 56, 56, 64, 64)
dev = "cuda:0"
x = torch.randint(-128, 128, (n * h * w * cin,), dtype=torch.int8, device=dev)
wt = torch.randint(-128, 128, (cout, 3, 3, cin), dtype=torch.int8)
wd = ops.upload_weights(wt, dev)
chan = ops.make_chan(np.zeros(cout, dtype=np.int64), [1 << 30] * cout, [40] * cout).to(dev)
out = torch.empty(n * h * w * cout, dtype=torch.int8, device=dev)
d = ops.conv_desc(n, h, w, cin, cout, 3, 3, 1, 1, 8, 1)
ep = ops.epilogue(EPI_REQUANT, relu=1, out_bits=8, clamp=(-128, 127), flags=1)
for _ in range(3):
    ops.conv2d(x, d, ep, wd, chan, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.conv2d(x, d, ep, wd, chan, out=out); e1.record(); torch.cuda.synchronize()
print("launch %.1f us" % (e0.elapsed_time(e1) * 1e3))
buf = (C.c_int64 * (4 * 64 * 4))()
got = _lib.load().hawq_debug_halo_trace(C.cast(buf, C.c_void_p), 4 * 64 * 4)
t = np.array(buf[:got], dtype=np.int64).reshape(4, 64, 4)
t0 = t[t > 0].min()
for role, name, evs in ((0, "producer", "wait_empty got_empty issued -"), (1, "mma", "wait_tempty wait_pfull got_pfull issued"), (2, "epilogue", "wait_tfull got_tfull released stored"),
                        (3, "epilogue detail", "math_done before_bar1 after_bar1 staged(before_bar2)")):
    print(name, "(cycles since first stamp; columns: %s)" % evs)
    for i in range(int(os.environ.get("TRACE_ROWS", "12"))):
        print("  %2d " % i + " ".join("%8d" % (v - t0 if v > 0 else -1) for v in t[role, i]))
