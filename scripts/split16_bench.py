"""Experimental split16 kernel (three bf16 MMAs per 16 k) vs the default (tf32 hi*hi + bf16 cross, weight operand pre-split)."""
import sys
import torch
sys.path.insert(0, ".")
from mfm_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device=dev); g.manual_seed(0)
for (n, N, K, reps) in ((65536, 1024, 1024, 300), (65536, 1600, 1024, 150), (8192, 1024, 1024, 600)):
    A = torch.randn(n, K, generator=g, device=dev); Bt = torch.randn(N, K, generator=g, device=dev) * 0.03
    C = torch.empty(n, N, device=dev); mirror = torch.empty_like(Bt)
    run = lambda: _lib.check(lib.mfm_gemm_tf32x3(n, N, K, A.data_ptr(), K, 1, Bt.data_ptr(), K, 0, None, 0, C.data_ptr(), N, st))
    out = {}
    for mode in (0, 1):
        lib.mfm_set_gemm_split16(mode)
        _lib.check(lib.mfm_gemm_presplit(Bt.data_ptr(), mirror.data_ptr(), N * K, st))
        lib.mfm_gemm_register_mirror(Bt.data_ptr(), N * K, mirror.data_ptr())
        for _ in range(10):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(reps):
            run()
        e1.record(); torch.cuda.synchronize()
        out[mode] = e0.elapsed_time(e1) / reps
        lib.mfm_gemm_register_mirror(None, 0, None)
    lib.mfm_set_gemm_split16(0)
    print(f"{n}x{N}x{K}: default {out[0]:.4f} ms ({2.0*n*N*K/out[0]/1e9:.0f} TFLOP/s)  split16 {out[1]:.4f} ms ({2.0*n*N*K/out[1]/1e9:.0f} TFLOP/s)  x{out[0]/out[1]:.3f}", flush=True)
