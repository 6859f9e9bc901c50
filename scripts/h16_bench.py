"""Dense-layer kernels head to head: scaled-fp16 three-pass kernel (default) vs tf32 + bf16-cross (MFM_GEMM_H16=0), K-major operands,
weight operand pre-split, max |A| tracked by the producer.  Burst (10 launches) and sustained (>= 1 s) TFLOP/s, CUDA events."""
import json
import sys

import torch

sys.path.insert(0, ".")
from mfm_b200 import _lib      # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
st = torch.cuda.current_stream().cuda_stream


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


out = []
shapes = [(65536, 1024, 1024), (65536, 1600, 1024), (65536, 1024, 1600), (65536, 1024, 2048), (65536, 1024, 256), (65536, 1600, 1600),
          (8192, 1024, 1024), (8192, 1600, 1024), (8192, 1600, 1600)]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    shapes = shapes[:2] + shapes[6:7]
for n, N, K in shapes:
    A = torch.randn(n, K, device=dev); Bt = torch.randn(N, K, device=dev) / K ** 0.5; C = torch.empty(n, N, device=dev)
    bias = torch.randn(N, device=dev)
    amax = torch.zeros(1, device=dev)
    _lib.check(lib.mfm_absmax(A.data_ptr(), K, n, K, amax.data_ptr(), st))
    mirror = torch.empty(N * K + 16, device=dev)
    rec = {"M": n, "N": N, "K": K}
    for name, mode, groups in [("tf32_bf16x", 0, 2), ("h16_g1", 1, 1), ("h16", 1, 2), ("h16_g4", 1, 4), ("h16_cvt_unpack", 2, 2)][:3 if len(sys.argv) > 2 else 5]:
        lib.mfm_set_gemm_h16(mode)
        lib.mfm_debug_set_h16_groups(groups)
        _lib.check(lib.mfm_gemm_presplit(Bt.data_ptr(), mirror.data_ptr(), N * K, st))
        lib.mfm_gemm_register_mirror(Bt.data_ptr(), N * K, mirror.data_ptr())
        f = lambda: _lib.check(lib.mfm_gemm_dense(n, N, K, A.data_ptr(), K, Bt.data_ptr(), K, bias.data_ptr(), 1, C.data_ptr(), N, amax.data_ptr(), None, None, None, st))
        f(); torch.cuda.synchronize()
        burst = timed(f, 10)
        sus = timed(f, max(10, int(1000.0 / burst)))
        rec[name + "_ms_burst"], rec[name + "_ms_sustained"] = round(burst, 4), round(sus, 4)
        rec[name + "_tflops_sustained"] = round(2.0 * n * N * K / sus / 1e9, 1)
        lib.mfm_gemm_register_mirror(None, 0, None)
        if name == "h16":   # A pre-split by its producer (what the MLP's inner layers see): no splitter work at all
            a_s = torch.empty(n * K + 16, device=dev)
            _lib.check(lib.mfm_gemm_presplit(A.data_ptr(), a_s.data_ptr(), n * K, st))
            lib.mfm_gemm_register_mirror(Bt.data_ptr(), N * K, mirror.data_ptr())
            f3 = lambda: _lib.check(lib.mfm_gemm_dense(n, N, K, A.data_ptr(), K, Bt.data_ptr(), K, bias.data_ptr(), 1, C.data_ptr(), N, amax.data_ptr(), None,
                                                       a_s.data_ptr(), a_s.data_ptr() + 4 * n * K, st))
            f3(); torch.cuda.synchronize()
            b3 = timed(f3, 10); s3 = timed(f3, max(10, int(1000.0 / b3)))
            rec["h16_presplit_a_ms_burst"], rec["h16_presplit_a_ms_sustained"] = round(b3, 4), round(s3, 4)
            rec["h16_presplit_a_tflops_sustained"] = round(2.0 * n * N * K / s3 / 1e9, 1)
            lib.mfm_gemm_register_mirror(None, 0, None)
            del a_s
        if name == "h16":   # the same layer when nobody tracked max |A| (one reduction pass) and when B is split in the kernel
            f2 = lambda: _lib.check(lib.mfm_gemm_dense(n, N, K, A.data_ptr(), K, Bt.data_ptr(), K, bias.data_ptr(), 1, C.data_ptr(), N, None, None, None, None, st))
            lib.mfm_gemm_register_mirror(Bt.data_ptr(), N * K, mirror.data_ptr())
            f2(); rec["h16_untracked_ms_burst"] = round(timed(f2, 10), 4)
            lib.mfm_gemm_register_mirror(None, 0, None)
            f(); rec["h16_kernel_split_b_ms_burst"] = round(timed(f, 10), 4)
    lib.mfm_debug_set_h16_groups(2)
    lib.mfm_set_gemm_h16(1)
    rec["speedup_sustained"] = rec["tf32_bf16x_ms_sustained"] / rec["h16_ms_sustained"]
    print(json.dumps(rec), flush=True)
    out.append(rec)
    del A, Bt, C, mirror
json.dump(out, open("gpurun_out/r02_h16_bench.json", "w"), indent=1)
