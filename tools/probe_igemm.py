"""GPU probe: tcgen05 implicit-GEMM conv kernels vs torch (fp32, cuDNN) on bf16-rounded operands.
Usage: python tools/probe_igemm.py [case ...]   -- each case runs independently; prints rel errors."""
import ctypes as C
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "lsps_b200", "csrc", "liblsps_b200.so"))


class Shape(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("kind", "n", "h", "w", "cin", "cout")]


lib.lsps_last_error.restype = C.c_char_p
ctx = C.c_void_p()
assert lib.lsps_ctx_create(C.byref(ctx), 0) == 0
P = lambda t: C.c_void_p(t.data_ptr())
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def chk(rc):
    if rc != 0:
        raise RuntimeError("lsps error %d: %s" % (rc, lib.lsps_last_error(ctx).decode()))


def nhwc(t):  # NCHW fp32 -> NHWC bf16 contiguous
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def from_nhwc(t):
    return t.float().permute(0, 3, 1, 2)


def rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-20)).item(), ((a - b).norm() / (b.norm() + 1e-20)).item()


def run(kind, n, h, w, cin, cout, seed=0, time_it=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    x = torch.randn(n, cin, h, w, device=dev, generator=g).bfloat16().float()
    if kind == 2:
        wt = (torch.randn(cin, cout, 3, 3, device=dev, generator=g) * 0.05).bfloat16().float()
        wf = wt.permute(2, 3, 1, 0).reshape(9, cout, cin).contiguous()
    else:
        wt = (torch.randn(cout, cin, 3, 3, device=dev, generator=g) * 0.05).bfloat16().float()
        wf = wt.permute(2, 3, 0, 1).reshape(9, cout, cin).contiguous()
    wd = wf.transpose(1, 2).contiguous()
    bias = torch.randn(cout, device=dev, generator=g)
    x.requires_grad_(True)
    wt.requires_grad_(True)
    if kind == 0:
        y = F.conv2d(x, wt, bias, stride=1, padding=1)
    elif kind == 1:
        y = F.conv2d(x, wt, bias, stride=2, padding=1)
    else:
        y = F.conv_transpose2d(x, wt, bias, stride=2, padding=1, output_padding=1)
    dy = torch.randn(y.shape, device=dev, generator=g).bfloat16().float()
    y.backward(dy)
    sh = Shape(kind, n, h, w, cin, cout)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    xb, dyb = nhwc(x.detach()), nhwc(dy)
    wfb, wdb = wf.bfloat16(), wd.bfloat16()
    out = {}
    # fwd
    yb = torch.empty(y.shape[0], y.shape[2], y.shape[3], cout, device=dev, dtype=torch.bfloat16)
    chk(lib.lsps_conv_fwd(ctx, C.byref(sh), P(xb), P(wfb), P(bias), P(yb), 1, C.c_float(0.01), st))
    torch.cuda.synchronize()
    out["fwd"] = rel(from_nhwc(yb), y.detach())
    # fwd + lrelu
    chk(lib.lsps_conv_fwd(ctx, C.byref(sh), P(xb), P(wfb), P(bias), P(yb), 3, C.c_float(0.01), st))
    torch.cuda.synchronize()
    out["fwd_lrelu"] = rel(from_nhwc(yb), F.leaky_relu(y.detach(), 0.01))
    # dgrad (+mask +add)
    dxb = torch.empty(n, h, w, cin, device=dev, dtype=torch.bfloat16)
    chk(lib.lsps_conv_dgrad(ctx, C.byref(sh), P(dyb), P(wdb), P(dxb), None, None, 0, C.c_float(0.01), st))
    torch.cuda.synchronize()
    out["dgrad"] = rel(from_nhwc(dxb), x.grad)
    mask = torch.randn(n, cin, h, w, device=dev, generator=g)
    add = torch.randn(n, cin, h, w, device=dev, generator=g).bfloat16().float()
    mb, ab = nhwc(mask), nhwc(add)
    chk(lib.lsps_conv_dgrad(ctx, C.byref(sh), P(dyb), P(wdb), P(dxb), P(mb), P(ab), 12, C.c_float(0.01), st))
    torch.cuda.synchronize()
    ref = (x.grad + add) * torch.where(mb.float().permute(0, 3, 1, 2) > 0, 1.0, 0.01)
    out["dgrad_mask_add"] = rel(from_nhwc(dxb), ref)
    # wgrad
    dw = torch.zeros(9, cout, cin, device=dev)
    chk(lib.lsps_conv_wgrad(ctx, C.byref(sh), P(xb), P(dyb), P(dw), st))
    torch.cuda.synchronize()
    if kind == 2:
        refw = wt.grad.permute(2, 3, 1, 0).reshape(9, cout, cin)
    else:
        refw = wt.grad.permute(2, 3, 0, 1).reshape(9, cout, cin)
    out["wgrad"] = rel(dw, refw)
    msg = " ".join("%s max %.2e l2 %.2e |" % (k, v[0], v[1]) for k, v in out.items())
    ok = all(v[1] < 6e-3 for v in out.values())
    print("%s kind %d n %d h %d w %d cin %d cout %d :: %s" % ("OK  " if ok else "FAIL", kind, n, h, w, cin, cout, msg), flush=True)
    if time_it:
        flops = 2.0 * y.numel() / cout * cout * cin * 9 if kind != 2 else 2.0 * n * h * w * cin * cout * 9
        for name, fn in (("fwd", lambda: lib.lsps_conv_fwd(ctx, C.byref(sh), P(xb), P(wfb), P(bias), P(yb), 3, C.c_float(0.01), st)),
                         ("dgrad", lambda: lib.lsps_conv_dgrad(ctx, C.byref(sh), P(dyb), P(wdb), P(dxb), None, None, 0, C.c_float(0.01), st)),
                         ("wgrad", lambda: lib.lsps_conv_wgrad(ctx, C.byref(sh), P(xb), P(dyb), P(dw), st))):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print("   time %-6s %.3f ms  %.1f TFLOP/s" % (name, ms, flops / ms / 1e9), flush=True)
        # cuDNN bf16 channels_last for comparison
        xc = x.detach().bfloat16().contiguous(memory_format=torch.channels_last)
        wc = wt.detach().bfloat16().contiguous(memory_format=torch.channels_last)
        fn = (lambda: F.conv2d(xc, wc, None, stride=1 if kind == 0 else 2, padding=1)) if kind != 2 else \
            (lambda: F.conv_transpose2d(xc, wc, None, stride=2, padding=1, output_padding=1))
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("   time cudnn-bf16-fwd %.3f ms  %.1f TFLOP/s" % (ms, flops / ms / 1e9), flush=True)
    return ok


CASES = {
    "k1_small": lambda: run(0, 2, 32, 32, 256, 256),
    "k1_c64": lambda: run(0, 3, 32, 32, 64, 64),
    "k1_128": lambda: run(0, 2, 16, 16, 128, 128),
    "s2_64": lambda: run(1, 2, 128, 128, 64, 128),
    "s2_128": lambda: run(1, 3, 64, 64, 128, 256),
    "s2_dis2": lambda: run(1, 5, 16, 16, 256, 512),
    "s2_dis3": lambda: run(1, 9, 8, 8, 512, 1024),
    "s2_dis4": lambda: run(1, 40, 4, 4, 1024, 2048),
    "dc_256": lambda: run(2, 2, 32, 32, 256, 128),
    "dc_128": lambda: run(2, 2, 64, 64, 128, 64),
    "dc_small": lambda: run(2, 3, 8, 8, 64, 64),
    "k1_time": lambda: run(0, 128, 32, 32, 256, 256, time_it=True),
    "s2_time": lambda: run(1, 128, 128, 128, 64, 128, time_it=True),
    "dc_time": lambda: run(2, 128, 64, 64, 128, 64, time_it=True),
    "dis4_time": lambda: run(1, 384, 4, 4, 1024, 2048, time_it=True),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        try:
            CASES[nm]()
        except Exception as e:  # noqa
            print("ERROR %s: %r" % (nm, e), flush=True)
            sys.exit(1)
