"""Per-item SM-clock timeline of one CTA pair of the persistent kernel (library built with MFM_TC2_TIMELINE=1 MFM_TL_PAIR=p)."""
import ctypes
import sys
import torch
sys.path.insert(0, ".")
from mfm_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
for (n, N, K) in ((8192, 1024, 1024), (65536, 1024, 1024)):
    A = torch.randn(n, K, device=dev); Bt = torch.randn(N, K, device=dev) * 0.03; bias = torch.randn(N, device=dev); C = torch.empty(n, N, device=dev)
    for mode in (0, 1):
        lib.mfm_set_gemm_streamk(mode)
        run = lambda: _lib.check(lib.mfm_gemm_tf32x3(n, N, K, A.data_ptr(), K, 1, Bt.data_ptr(), K, 0, bias.data_ptr(), 1, C.data_ptr(), N, st))
        for _ in range(20):
            run()
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * 64)()
        lib.mfm_debug_gemm_timeline(1, None); run(); torch.cuda.synchronize(); lib.mfm_debug_gemm_timeline(0, buf)
        t = list(buf); t0 = t[62]
        print(f"--- {n}x{N}x{K} streamk={mode}: clocks since kernel entry; exit {t[63] - t0}")
        for i in range(12):
            row = [t[4 * i + k] - t0 for k in range(4)]
            if any(0 < v < 10**7 for v in row):
                print(f"   item {i}: mma_start {row[0]} acc_committed {row[1]} epi_start {row[2]} epi_end {row[3]}")
        ex = {48: "dump_start", 49: "dump_end", 50: "preload_start", 51: "preload_end", 52: "mma_prefull_ok"}
        print("   ", {v: t[k] - t0 for k, v in ex.items() if 0 < t[k] - t0 < 10**7})
    lib.mfm_set_gemm_streamk(1)
