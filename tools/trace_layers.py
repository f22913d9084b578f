"""Timeline trace of the tcgen05 conv kernel (CTA 0) for representative ResNet-50 layer shapes at batch 128."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from hawq_b200 import _lib, ops  # noqa: E402
from hawq_b200._lib import EPI_REQUANT, EPI_RESIDUAL, dyadic  # noqa: E402

dev = "cuda:0"
B = 128
LAYERS = [  # name, H, Cin, Cout, k, stride, mode
    ("s1.conv1 1x1 256->64 REQ", 56, 256, 64, 1, 1, "req"),
    ("s1.conv2 3x3 64->64 REQ", 56, 64, 64, 3, 1, "req"),
    ("s1.conv3 1x1 64->256 RES", 56, 64, 256, 1, 1, "res"),
    ("s3.conv2 3x3 256->256 REQ", 14, 256, 256, 3, 1, "req"),
    ("s3.conv3 1x1 256->1024 RES", 14, 256, 1024, 1, 1, "res"),
]
trace = torch.zeros(3 * 64 * 8, dtype=torch.int64, device=dev)
lib = _lib.load()
r = np.random.RandomState(0)
for name, H, cin, cout, k, s, mode in LAYERS:
    x = torch.from_numpy(r.randint(-128, 128, size=B * H * H * cin).astype(np.int8)).to(dev)
    w = ops.upload_weights(torch.from_numpy(r.randint(-128, 128, size=(cout, k, k, cin)).astype(np.int8)), dev)
    me = [dyadic(1e-3)] * cout
    chan = ops.make_chan([0] * cout, [m for m, _ in me], [e for _, e in me]).to(dev)
    Ho = (H + 2 * (k // 2) - k) // s + 1
    numel = B * Ho * Ho * cout
    d = ops.conv_desc(B, H, H, cin, cout, k, k, s, k // 2, 8, 1)
    if mode == "req":
        ep = ops.epilogue(EPI_REQUANT, relu=1, out_bits=8, clamp=(-128, 127), flags=1)
        args = dict(out=torch.zeros(numel, dtype=torch.int8, device=dev))
    else:
        ep = ops.epilogue(EPI_RESIDUAL, relu=1, res_kind=0, res_bits=16, res_me=dyadic(0.9), y_bits=16, low_bits=8, low_me=dyadic(0.004),
                          low_clamp=(-128, 127), flags=1)
        args = dict(res=torch.zeros(numel, dtype=torch.int16, device=dev), out=torch.zeros(numel, dtype=torch.int16, device=dev),
                    out_low=torch.zeros(numel, dtype=torch.int8, device=dev))
    for _ in range(3):
        ops.conv2d(x, d, ep, w, chan, **args)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.conv2d(x, d, ep, w, chan, **args)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    lib.hawq_debug_set_trace(C.c_void_p(trace.data_ptr()))
    trace.zero_()
    ops.conv2d(x, d, ep, w, chan, **args)
    torch.cuda.synchronize()
    lib.hawq_debug_set_trace(None)
    t = trace.cpu().numpy().reshape(3, 64, 8)
    tiles = ((B * Ho * Ho + 127) // 128) * (cout // (128 if cout % 128 == 0 else 64))
    print("== %s : %.1f us/launch, %d tiles (%.1f per CTA), KT=%d" % (name, us, tiles, tiles / 148, k * k * cin // 64))
    t0 = t[t > 0].min()
    n = min(10, int((t[2, :, 1] > 0).sum()))
    for i in range(2, n):
        p_, m_, e_ = t[0, i], t[1, i], t[2, i]
        if k == 3:
            print("  tile %2d | patch producer (kh=1 row of 3 taps): empty-waits %5d  lds+sts %5d  fence %5d  arrives %5d" % (i, p_[1] - p_[0], p_[2] - p_[1], p_[3] - p_[2], p_[4] - p_[3]))
        print("  tile %2d | prod: start %6d emptywait %5d issue %5d | mma: tempty-wait %5d full-wait(1st) %5d mma+commit %5d | epi: tfull-wait %5d ld/res %5d compute %6d copyout %5d  [epi period %6d]" % (
            i, p_[0] - t0, p_[1] - p_[0], p_[2] - p_[1], m_[1] - m_[0], m_[2] - m_[1], m_[3] - m_[2],
            e_[1] - e_[0], e_[2] - e_[1], e_[3] - e_[2], e_[4] - e_[3], t[2, i, 4] - t[2, i - 1, 4]))
