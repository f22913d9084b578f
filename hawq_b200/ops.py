"""Thin Python layer over the C ABI: torch tensors in, raw device pointers + the current CUDA stream out.

torch is used only as the owner of device memory and streams.  Every function launches asynchronously on
``torch.cuda.current_stream()`` and never synchronises, so sequences of calls can be captured in a CUDA graph.
There is no fallback: tensors must live on a CUDA device and the in-tree library must load.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import EPI_RAW_I32, EPI_REQUANT, EPI_RESIDUAL, hawq_conv_desc, hawq_epilogue_desc

_handles = {}
launch_count = 0   # number of kernels launched through this module (bench reports it as gpu_launches)


def handle(device_index):
    h = _handles.get(device_index)
    if h is None:
        lib = _lib.load()
        out = C.c_void_p()
        _lib.check(lib.hawq_create(int(device_index), C.byref(out)))
        h = _handles[device_index] = out
    return h


def _ctx(t):
    if not t.is_cuda:
        raise RuntimeError("hawq_b200 integer ops need CUDA tensors (got %s); there is no CPU path" % t.device)
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    return handle(idx), C.c_void_p(torch.cuda.current_stream(idx).cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


timer = None   # set to a list to record (kernel, info, start_event, end_event) per launch (bench.py roofline leg)


def _begin():
    if timer is None:
        return None
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


def _count(name="", info=None, ev0=None):
    global launch_count
    launch_count += 1
    if ev0 is not None:
        ev1 = torch.cuda.Event(enable_timing=True)
        ev1.record()
        timer.append((name, info, ev0, ev1))


def conv_work(desc, ep):
    """(MACs, algorithmic HBM bytes) of one fused convolution launch: activations in at their stored width (only the
    sampled pixels for strided 1x1), weights + per-channel parameters once, every output at its stored width, plus the
    residual operand for case-1 epilogues."""
    ho = (desc.H + 2 * desc.pad - desc.kh) // desc.stride + 1
    wo = (desc.W + 2 * desc.pad - desc.kw) // desc.stride + 1
    m = desc.N * ho * wo
    macs = m * desc.Cout * desc.kh * desc.kw * desc.Cin
    in_pix = m if desc.kh == 1 else desc.N * desc.H * desc.W
    b = in_pix * desc.Cin * desc.a_bits // 8 + desc.Cout * desc.kh * desc.kw * desc.Cin + 16 * desc.Cout
    outs = m * desc.Cout
    if ep.mode == EPI_REQUANT:
        b += outs * ep.out_bits // 8
    elif ep.mode == EPI_RESIDUAL:
        b += outs * (4 if ep.res_kind == 1 else ep.res_bits // 8) + outs * ep.y_bits // 8 + outs * ep.low_bits // 8
        if ep.res_kind == 1:
            b += 16 * desc.Cout
    elif ep.mode == EPI_RAW_I32:
        b += outs * 4
    else:
        b += m * ep.cout_store * 4 + 4 * desc.Cout
    return macs, b


def make_chan(bias, m, e):
    """hawq_chan[C] as an int32 [C,4] CPU tensor (m stored by bit pattern)."""
    c = len(bias)
    a = np.zeros((c, 4), dtype=np.int32)
    a[:, 0] = np.asarray(bias, dtype=np.int64).astype(np.int32)
    a[:, 1] = np.asarray(m, dtype=np.uint64).astype(np.uint32).view(np.int32)
    a[:, 2] = np.asarray(e, dtype=np.int32)
    return torch.from_numpy(a)


def conv_desc(N, H, W, Cin, Cout, kh, kw, stride, pad, a_bits, w_layout=0):
    return hawq_conv_desc(N, H, W, Cin, Cout, kh, kw, stride, pad, a_bits, w_layout)


def upload_weights(w_ohwi_cpu, device):
    """int8 OHWI host weights -> device buffer [OHWI | tcgen05 re-tiled copy] (hawq_conv_desc.w_layout = 1)."""
    cout = w_ohwi_cpu.shape[0]
    k = w_ohwi_cpu.numel() // cout
    buf = torch.empty(2 * cout * k, dtype=torch.int8, device=device)
    buf[:cout * k].copy_(w_ohwi_cpu.reshape(-1))
    idx = buf.device.index if buf.device.index is not None else torch.cuda.current_device()
    _lib.check(_lib.load().hawq_retile_weights(handle(idx), C.c_void_p(buf.data_ptr()), cout, k, C.c_void_p(buf.data_ptr() + cout * k),
                                               C.c_void_p(torch.cuda.current_stream(idx).cuda_stream)))
    return buf


def epilogue(mode, relu=0, out_bits=0, clamp=(0, 0), res_kind=0, res_bits=0, res_me=(0, 1), y_bits=0, low_bits=0,
             low_me=(0, 1), low_clamp=(0, 0), cout_store=0, flags=0):
    return hawq_epilogue_desc(mode, int(relu), out_bits, clamp[0], clamp[1], res_kind, res_bits, res_me[0], res_me[1],
                              y_bits, low_bits, low_me[0], low_me[1], low_clamp[0], low_clamp[1], cout_store, flags)


def ratio_flags(*pairs):
    """HAWQ_EP_RATIOS_* promise for a set of (m, e) pairs (or (m list, e list)): LE_ONE when every ratio m * 2^-e <= 1
    (e >= 31 or m == 0), LE_2P20 when every ratio <= 2^20 (e >= 11), else 0.  (Whether the promise is made at all is the
    caller's execution mode: qtensor.EngineConfig.)"""
    min_e = 99
    for m, e in pairs:
        ms = m if isinstance(m, (list, tuple)) else [m]
        es = e if isinstance(e, (list, tuple)) else [e]
        for mi, ei in zip(ms, es):
            if mi != 0:
                min_e = min(min_e, ei)
    if min_e >= 31:
        return _lib.EP_RATIOS_LE_ONE
    if min_e >= 11:
        return _lib.EP_RATIOS_LE_2P20
    return 0


def conv2d(x, desc, ep, w, chan, res=None, res_chan=None, fscale=None, out=None, out_low=None):
    h, s = _ctx(x)
    ev = _begin()
    lib = _lib.load()
    halo0, c10 = (lib.hawq_debug_kernel_count(1), lib.hawq_debug_kernel_count(3)) if ev is not None else (0, 0)
    _lib.check(lib.hawq_conv2d(h, C.byref(desc), C.byref(ep), _p(x), _p(w), _p(chan), _p(res), _p(res_chan),
                               _p(fscale), _p(out), _p(out_low), s))
    if ev is not None:     # per-launch timing (bench.py roofline leg): name the kernel family that took the launch
        name = "conv_halo" if lib.hawq_debug_kernel_count(1) != halo0 else "conv1x1" if lib.hawq_debug_kernel_count(3) != c10 else "conv_tc"
        _count(name, conv_work(desc, ep), ev)
    else:
        _count()


def conv2d_dual(x, desc, ep, w, chan, desc2, x2, w2, chan2, out=None, out_low=None):
    """resize unit: identity 1x1 conv (desc2/x2/w2/chan2) + last 1x1 conv (desc/x/w/chan) + case-1 sum in one kernel."""
    h, s = _ctx(x)
    ev = _begin()
    k0 = _lib.load().hawq_debug_kernel_count(4) if ev is not None else 0
    _lib.check(_lib.load().hawq_conv2d_dual(h, C.byref(desc), C.byref(ep), _p(x), _p(w), _p(chan), C.byref(desc2), _p(x2), _p(w2),
                                            _p(chan2), _p(out), _p(out_low), s))
    work = None
    if ev is not None:
        m = desc.N * desc.H * desc.W
        macs = m * desc.Cout * (desc.Cin + desc2.Cin)
        b = (m * (desc.Cin + desc2.Cin) * desc.a_bits // 8 + desc.Cout * (desc.Cin + desc2.Cin) + 32 * desc.Cout
             + m * desc.Cout * (ep.y_bits + ep.low_bits) // 8)
        work = (macs, b)
    _count("conv_dual" if ev is not None and _lib.load().hawq_debug_kernel_count(4) != k0 else "conv_tc_dual", work, ev)


def linear(x, w, chan, fscale, out, n, k, cout, cout_pad):
    h, s = _ctx(x)
    ev = _begin()
    _lib.check(_lib.load().hawq_linear_i8(h, n, k, cout, cout_pad, _p(x), _p(w), _p(chan), _p(fscale), _p(out), s))
    _count("hawq_linear", (n * k * cout, n * k + cout_pad * k + 20 * cout_pad + n * cout * 4) if ev is not None else None, ev)


def stem_conv(x, w, chan, clamp, out, n, hh, ww):
    h, s = _ctx(x)
    ev = _begin()
    _lib.check(_lib.load().hawq_stem_conv_i8(h, n, hh, ww, _p(x), _p(w), _p(chan), clamp[0], clamp[1], _p(out), s))
    ho, wo = (hh - 1) // 2 + 1, (ww - 1) // 2 + 1
    _count("stem_conv", (n * ho * wo * 64 * 147, n * hh * ww * 3 + 64 * 224 + 1024 + n * ho * wo * 64 * 2) if ev is not None else None, ev)


def stem_pool(x, w256, chan, clamp, n, hh, ww, y_bits, y, low_bits, low_me, low_clamp, out_low):
    """Fused stem (conv 7x7/2 + max-pool 3x3/2 + 16-bit requant + ReLU + low-bit copy) in one kernel; raises HawqError(ERR_UNSUPPORTED)
    for shapes / ratios outside it (callers then use stem_conv + maxpool_requant, same integers)."""
    h, s = _ctx(x)
    ev = _begin()
    _lib.check(_lib.load().hawq_stem_pool_i8(h, n, hh, ww, _p(x), _p(w256), _p(chan), clamp[0], clamp[1], y_bits, _p(y), low_bits, low_me[0],
                                             low_me[1], low_clamp[0], low_clamp[1], _p(out_low), s))
    ho, wo = (hh - 1) // 2 + 1, (ww - 1) // 2 + 1
    po, qo = (ho - 1) // 2 + 1, (wo - 1) // 2 + 1
    _count("stem_tc", (n * ho * wo * 64 * 147, n * hh * ww * 3 + 64 * 256 + 1024 + n * po * qo * 64 * (y_bits + low_bits) // 8) if ev is not None else None, ev)


def maxpool_requant(x, n, hh, ww, c, y_bits, y, low_bits, low_me, low_clamp, out_low):
    h, s = _ctx(x)
    ev = _begin()
    _lib.check(_lib.load().hawq_maxpool_requant(h, n, hh, ww, c, _p(x), y_bits, _p(y), low_bits, low_me[0], low_me[1],
                                                low_clamp[0], low_clamp[1], _p(out_low), s))
    po, qo = (hh - 1) // 2 + 1, (ww - 1) // 2 + 1
    _count("maxpool_requant", (0, n * hh * ww * c * 2 + n * po * qo * c * (y_bits + low_bits) // 8) if ev is not None else None, ev)


def avgpool_requant(x, n, hw, c, x_bits, me, clamp, out):
    h, s = _ctx(x)
    ev = _begin()
    _lib.check(_lib.load().hawq_avgpool_requant(h, n, hw, c, x_bits, _p(x), me[0], me[1], clamp[0], clamp[1], _p(out), s))
    _count("avgpool_requant", (0, n * hw * c * x_bits // 8 + n * c) if ev is not None else None, ev)


def quantize_input(x, scale, clamp, out):
    n, c, hh, ww = x.shape
    h, s = _ctx(x)
    ev = _begin()
    _lib.check(_lib.load().hawq_quantize_input_f32(h, n, c, hh, ww, _p(x), float(scale), clamp[0], clamp[1], _p(out), s))
    _count("quantize_input", (0, n * c * hh * ww * 5) if ev is not None else None, ev)


def quantize_input_u8(x, mean, std, scale, clamp, out):
    """uint8 NHWC images -> int8 NHWC network input (ToTensor + Normalize + QuantAct input branch in one kernel)."""
    n, hh, ww, c = x.shape
    if c != 3:
        raise ValueError("quantize_input_u8 expects NHWC images with 3 channels")
    h, s = _ctx(x)
    ev = _begin()
    m3, s3 = (C.c_float * 3)(*[float(v) for v in mean]), (C.c_float * 3)(*[float(v) for v in std])
    _lib.check(_lib.load().hawq_quantize_input_u8(h, n, hh, ww, _p(x), m3, s3, float(scale), clamp[0], clamp[1], _p(out), s))
    _count("quantize_input_u8", (0, n * hh * ww * 6) if ev is not None else None, ev)


def requant(x, rows, c, x_bits, chan, chan_stride, relu, out_bits, clamp, out):
    h, s = _ctx(x)
    _lib.check(_lib.load().hawq_requant(h, rows, c, x_bits, _p(x), _p(chan), chan_stride, int(relu), out_bits,
                                        clamp[0], clamp[1], _p(out), s))
    _count()


def add_requant(acc, rows, c, chan, ep, res, res_chan, y, out_low):
    h, s = _ctx(acc)
    _lib.check(_lib.load().hawq_add_requant(h, rows, c, _p(acc), _p(chan), C.byref(ep), _p(res), _p(res_chan), _p(y),
                                            _p(out_low), s))
    _count()


def dequant(x, n, hh, ww, c, x_bits, x_signed, scale, out):
    h, s = _ctx(x)
    _lib.check(_lib.load().hawq_dequant_f32(h, n, hh, ww, c, x_bits, int(x_signed), _p(x), float(scale), _p(out), s))
    _count()


def pack_i4(src, dst):
    h, s = _ctx(src)
    _lib.check(_lib.load().hawq_pack_i4(h, src.numel(), _p(src), _p(dst), s))
    _count()


def unpack_i4(src, dst):
    h, s = _ctx(src)
    _lib.check(_lib.load().hawq_unpack_i4(h, dst.numel(), _p(src), _p(dst), s))
    _count()


def reset_status(device_index):
    h = handle(device_index)
    _lib.check(_lib.load().hawq_reset_status(h, C.c_void_p(torch.cuda.current_stream(device_index).cuda_stream)))


def copy_status(device_index, dst):
    h = handle(device_index)
    _lib.check(_lib.load().hawq_copy_status(h, _p(dst), C.c_void_p(torch.cuda.current_stream(device_index).cuda_stream)))


def get_status(device_index):
    h = handle(device_index)
    v = C.c_int32()
    _lib.check(_lib.load().hawq_get_status(h, C.c_void_p(torch.cuda.current_stream(device_index).cuda_stream), C.byref(v)))
    return int(v.value)


def permute_weights_for_i4(w_ohwi_int8):
    """In-place K permutation (host, contiguous int8 [Cout, kh, kw, Cin])."""
    assert w_ohwi_int8.dtype == torch.int8 and w_ohwi_int8.is_contiguous() and not w_ohwi_int8.is_cuda
    cin = w_ohwi_int8.shape[-1]
    _lib.check(_lib.load().hawq_permute_weights_for_i4(C.c_void_p(w_ohwi_int8.data_ptr()),
                                                       w_ohwi_int8.numel() // cin, cin))
    return w_ohwi_int8
