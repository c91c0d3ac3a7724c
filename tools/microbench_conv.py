"""BASELINE.json config 5: conv / deconv microbench -- the tcgen05 implicit-GEMM kernels (forward, data gradient, weight
gradient) next to cuDNN (torch, bf16 channels_last and TF32) on the same GPU.  3x3 stride-2 pad-1 Cin=C -> Cout=2C on N=256
images for (C, H) in the list below, plus K1 (256->256 s1 @32x32) and the decoder transposed convs (K6), N=128.
Burst numbers: 3 warm-up + 10 timed launches per entry (the sustained figures are in profiles/r01_probe_sustained.log).
C=32 is not a shape of the LSPS nets (ch=64 in both YAMLs) and is not supported by the kernels (channels % 64)."""
import ctypes as C
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsps_b200 import _lib  # noqa
from lsps_b200._lib import ConvShape  # noqa

ctx = _lib.context(0)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def case(name, kind, n, h, cin, cout):
    ho = h if kind == 0 else (h // 2 if kind == 1 else 2 * h)
    x = torch.randn(n, h, h, cin, device="cuda").bfloat16()
    dy = torch.randn(n, ho, ho, cout, device="cuda").bfloat16()
    wf = (torch.randn(9, cout, cin, device="cuda") * 0.05).bfloat16()
    wd = wf.transpose(1, 2).contiguous()
    b = torch.zeros(cout, device="cuda")
    y, dx = torch.empty_like(dy), torch.empty_like(x)
    dw = torch.zeros(9, cout, cin, device="cuda")
    sh = C.byref(ConvShape(kind, n, h, h, cin, cout))
    macs = n * (ho * ho if kind != 2 else h * h) * cin * cout * 9
    flop = 2.0 * macs
    io = (x.numel() + dy.numel() + wf.numel()) * 2
    t_f = timeit(lambda: ctx.conv_fwd(sh, x.data_ptr(), wf.data_ptr(), b.data_ptr(), y.data_ptr(), 3, 0.01))
    t_d = timeit(lambda: ctx.conv_dgrad(sh, dy.data_ptr(), wd.data_ptr(), dx.data_ptr(), None, None, 0, 0.01))
    t_w = timeit(lambda: ctx.conv_wgrad(sh, x.data_ptr(), dy.data_ptr(), dw.data_ptr()))
    # cuDNN through torch: bf16 channels_last forward / backward-input / backward-weight, and TF32 forward
    xc = x.permute(0, 3, 1, 2)                                  # NCHW view of NHWC memory = channels_last
    dyc = dy.permute(0, 3, 1, 2)
    if kind == 2:
        wc = wf.reshape(3, 3, cout, cin).permute(3, 2, 0, 1).contiguous(memory_format=torch.channels_last)
        fwd = lambda a, w_: F.conv_transpose2d(a, w_, None, stride=2, padding=1, output_padding=1)
        args = dict(stride=[2, 2], padding=[1, 1], dilation=[1, 1], transposed=True, output_padding=[1, 1], groups=1)
    else:
        wc = wf.reshape(3, 3, cout, cin).permute(2, 3, 0, 1).contiguous(memory_format=torch.channels_last)
        st = 1 if kind == 0 else 2
        fwd = lambda a, w_: F.conv2d(a, w_, None, stride=st, padding=1)
        args = dict(stride=[st, st], padding=[1, 1], dilation=[1, 1], transposed=False, output_padding=[0, 0], groups=1)
    c_f = timeit(lambda: fwd(xc, wc))
    bwd = lambda mask: torch.ops.aten.convolution_backward(dyc, xc, wc, None, args["stride"], args["padding"],
                                                           args["dilation"], args["transposed"], args["output_padding"],
                                                           args["groups"], mask)
    c_d = timeit(lambda: bwd([True, False, False]))
    c_w = timeit(lambda: bwd([False, True, False]))
    torch.backends.cudnn.allow_tf32 = True
    x32, w32 = xc.float(), wc.float()
    c_t = timeit(lambda: fwd(x32, w32))
    tf = lambda ms: flop / ms / 1e9
    print("| %s | %.1f | %.3f / %.0f / %.0f | %.3f / %.0f | %.3f / %.0f | %.3f / %.0f | %.3f / %.0f | %.3f / %.0f | %.3f / %.0f |"
          % (name, flop / 1e9, t_f, tf(t_f), io / t_f / 1e6, t_d, tf(t_d), t_w, tf(t_w), c_f, tf(c_f), c_d, tf(c_d), c_w, tf(c_w),
             c_t, tf(c_t)), flush=True)
    del x, dy, y, dx, x32
    torch.cuda.empty_cache()


print("| shape | GFLOP | ours fwd ms / TFLOP/s / GB/s | ours dgrad ms / TFLOP/s | ours wgrad ms / TFLOP/s | cuDNN bf16 fwd | "
      "cuDNN bf16 dgrad | cuDNN bf16 wgrad | cuDNN TF32 fwd |\n|---|---|---|---|---|---|---|---|---|")
for c, h in ((64, 128), (64, 64), (128, 64), (128, 32), (256, 16)):
    case("s2 %d->%d @%d^2 N=256" % (c, 2 * c, h), 1, 256, h, c, 2 * c)
case("K1 s1 256->256 @32^2 N=128", 0, 128, 32, 256, 256)
case("K6 deconv 256->128 @32^2 N=128", 2, 128, 32, 256, 128)
case("K6 deconv 128->64 @64^2 N=128", 2, 128, 64, 128, 64)
