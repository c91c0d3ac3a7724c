"""Sustained (power-limited) throughput of the K1 conv kernels next to cuBLAS bf16, same process, same box.
The 10-launch probes in probe_igemm.py are BURST numbers (cool chip, max clocks); a training step runs for seconds
under the 1000 W cap.  Each case loops for ~SEC seconds; per-window TFLOP/s with the SM clock / power sampled via NVML."""
import ctypes as C
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "lsps_b200", "csrc", "liblsps_b200.so"))


class Shape(C.Structure):
    _fields_ = [(k, C.c_int) for k in ("kind", "n", "h", "w", "cin", "cout")]


ctx = C.c_void_p()
assert lib.lsps_ctx_create(C.byref(ctx), 0) == 0
P = lambda t: C.c_void_p(t.data_ptr())
SEC = float(os.environ.get("LSPS_SUSTAIN_SEC", "4"))

try:
    import pynvml
    pynvml.nvmlInit()
    _h = pynvml.nvmlDeviceGetHandleByIndex(0)
except Exception:  # noqa
    pynvml = None


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows, self.stop = [], False

    def run(self):
        while not self.stop and pynvml is not None:
            try:
                self.rows.append((pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM),
                                  pynvml.nvmlDeviceGetPowerUsage(_h) / 1000.0))
            except Exception:  # noqa
                pass
            time.sleep(0.05)


def sustained(name, fn, flops, batch=100):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s = Sampler()
    s.start()
    t_end = time.time() + SEC
    rates = []
    while time.time() < t_end:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(batch):
            fn()
        e1.record()
        torch.cuda.synchronize()
        rates.append(flops * batch / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    s.stop = True
    s.join()
    tail = s.rows[len(s.rows) // 2:] or [(0, 0)]
    clk = sorted(r[0] for r in tail)[len(tail) // 2]
    pw = max(r[1] for r in tail)
    k = max(1, len(rates) // 4)
    print("%-34s first %.0f  last-quarter %.0f TFLOP/s   (SM clock median %d MHz, power max %.0f W, %d windows)"
          % (name, rates[0], sum(rates[-k:]) / k, clk, pw, len(rates)), flush=True)


def conv_case(n):
    dev = "cuda"
    x = torch.randn(n, 32, 32, 256, device=dev).bfloat16()
    w = (torch.randn(9, 256, 256, device=dev) * 0.05).bfloat16()
    b = torch.randn(256, device=dev)
    y = torch.empty_like(x)
    dw = torch.zeros(9, 256, 256, device=dev)
    sh = Shape(0, n, 32, 32, 256, 256)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    fl = 2.0 * n * 1024 * 256 * 2304
    sustained("K1 fwd  N=%d" % n, lambda: lib.lsps_conv_fwd(ctx, C.byref(sh), P(x), P(w), P(b), P(y), 1, C.c_float(0.01), st), fl)
    sustained("K1 dgrad+mask+add N=%d" % n, lambda: lib.lsps_conv_dgrad(ctx, C.byref(sh), P(x), P(w), P(y), P(x), P(x), 12, C.c_float(0.01), st), fl)
    sustained("K1 wgrad N=%d" % n, lambda: lib.lsps_conv_wgrad(ctx, C.byref(sh), P(x), P(x), P(dw), st), fl)


if __name__ == "__main__":
    a = torch.randn(8192, 8192, device="cuda").bfloat16()
    bm = torch.randn(8192, 8192, device="cuda").bfloat16()
    c = torch.empty(8192, 8192, device="cuda", dtype=torch.bfloat16)
    sustained("cuBLAS bf16 8192^3", lambda: torch.matmul(a, bm, out=c), 2.0 * 8192 ** 3, batch=20)
    conv_case(128)
    conv_case(64)
    sustained("cuBLAS bf16 8192^3 (again)", lambda: torch.matmul(a, bm, out=c), 2.0 * 8192 ** 3, batch=20)
