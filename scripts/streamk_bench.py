"""Dense-layer shapes of the pines MLP at 8 192 / 16 384 / 65 536 chains, remainder round cut along K (stream-K) on / off."""
import sys
import torch
sys.path.insert(0, ".")
from mfm_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device=dev); g.manual_seed(0)
reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 300
print(f"{'M':>6} {'N':>5} {'K':>5} | whole tiles ms | stream-K ms | speed-up | TFLOP/s (stream-K)")
for n in (8192, 16384, 65536):
    for (N, K) in ((1024, 1024), (1024, 1600), (1600, 1024), (1600, 1600), (1024, 256)):
        A = torch.randn(n, K, generator=g, device=dev); Bt = torch.randn(N, K, generator=g, device=dev) * 0.03
        bias = torch.randn(N, generator=g, device=dev); C = torch.empty(n, N, device=dev)
        ms = {}
        for mode in (0, 1):
            lib.mfm_set_gemm_streamk(mode)
            run = lambda: _lib.check(lib.mfm_gemm_tf32x3(n, N, K, A.data_ptr(), K, 1, Bt.data_ptr(), K, 0, bias.data_ptr(), 1, C.data_ptr(), N, st))
            for _ in range(5):
                run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(reps):
                run()
            e1.record(); torch.cuda.synchronize()
            ms[mode] = e0.elapsed_time(e1) / reps
        lib.mfm_set_gemm_streamk(1)
        print(f"{n:6d} {N:5d} {K:5d} | {ms[0]:14.4f} | {ms[1]:11.4f} | {ms[0] / ms[1]:8.3f} | {2.0 * n * N * K / ms[1] / 1e9:8.1f}", flush=True)
        del A, Bt, C
