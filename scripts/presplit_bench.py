"""Persistent dense-layer GEMM with the B operand's cross tile split in the kernel vs pre-split in global memory (TMA)."""
import sys
import torch
sys.path.insert(0, ".")
from mfm_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device=dev); g.manual_seed(0)
reps = 300
print(f"{'M':>6} {'N':>5} {'K':>5} | split in kernel ms | pre-split ms | speed-up | TFLOP/s (pre-split)")
for n in (65536, 8192):
    for (N, K) in ((1024, 1024), (1024, 1600), (1600, 1024), (1600, 1600), (1024, 2048)):
        A = torch.randn(n, K, generator=g, device=dev); Bt = torch.randn(N, K, generator=g, device=dev) * 0.03
        bias = torch.randn(N, generator=g, device=dev); C = torch.empty(n, N, device=dev); mirror = torch.empty_like(Bt)
        _lib.check(lib.mfm_gemm_presplit(Bt.data_ptr(), mirror.data_ptr(), N * K, st))
        run = lambda: _lib.check(lib.mfm_gemm_tf32x3(n, N, K, A.data_ptr(), K, 1, Bt.data_ptr(), K, 0, bias.data_ptr(), 1, C.data_ptr(), N, st))
        ms = {}
        for rnd in range(2):                    # two rounds each, interleaved, to average out clock drift
            for mode in (0, 1):
                lib.mfm_gemm_register_mirror(Bt.data_ptr() if mode else None, N * K if mode else 0, mirror.data_ptr() if mode else None)
                for _ in range(5):
                    run()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record()
                for _ in range(reps):
                    run()
                e1.record(); torch.cuda.synchronize()
                ms[mode] = ms.get(mode, 0.0) + e0.elapsed_time(e1) / reps / 2
        lib.mfm_gemm_register_mirror(None, 0, None)
        print(f"{n:6d} {N:5d} {K:5d} | {ms[0]:18.4f} | {ms[1]:12.4f} | {ms[0] / ms[1]:8.3f} | {2.0 * n * N * K / ms[1] / 1e9:8.1f}", flush=True)
        del A, Bt, C, mirror
