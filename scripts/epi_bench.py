"""Times the backward-data dense layer (mask / add epilogue operands) in isolation on a B200."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from mfm_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
n, H = (int(sys.argv[sys.argv.index("--n") + 1]) if "--n" in sys.argv else 65536), 1024
g = torch.Generator(device=dev); g.manual_seed(0)
A = torch.randn(n, H, generator=g, device=dev); Bt = torch.randn(H, H, generator=g, device=dev) * 0.03
mask = torch.randn(n, H, generator=g, device=dev); add = torch.randn(n, H, generator=g, device=dev)
C = torch.empty(n, H, device=dev)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
ref = (A[:64].double() @ Bt.double().T + add[:64].double()) * (mask[:64] > 0)
for name, m, a in [("plain", None, None), ("mask", mask, None), ("add", None, add), ("mask+add", mask, add)]:
    def run():
        _lib.check(lib.mfm_gemm_tf32x3_gated(n, H, H, A.data_ptr(), H, Bt.data_ptr(), H, m.data_ptr() if m is not None else None, H,
                                             a.data_ptr() if a is not None else None, H, C.data_ptr(), H, st))
    run(); torch.cuda.synchronize()
    if name == "mask+add":
        print("max err", float((C[:64].double() - ref).abs().max() / ref.abs().max()))
    ts = []
    for _ in range(5):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2000 if n > 20000 else 10000):
        run()
    e1.record(); torch.cuda.synchronize()
    if "--timeline" in sys.argv:
        import ctypes
        buf = (ctypes.c_longlong * 64)()
        lib.mfm_debug_gemm_timeline(1, None); run(); torch.cuda.synchronize(); lib.mfm_debug_gemm_timeline(0, buf)
        t = list(buf); t0 = t[0]
        print("   tile: mma_start acc_committed epi_start epi_end (SM clocks since first MMA)")
        for i in range(8):
            print("   ", i, [t[4 * i + k] - t0 for k in range(4)])
        print("    kernel entry", t[62] - t0, "exit", t[63] - t0)
    print(f"{name:9s} burst {np.median(ts):.3f} ms  sustained {e0.elapsed_time(e1) / (2000 if n > 20000 else 10000):.3f} ms  ({2.0 * n * H * H / (e0.elapsed_time(e1) / (2000 if n > 20000 else 10000)) / 1e9:.0f} TFLOP/s)")
