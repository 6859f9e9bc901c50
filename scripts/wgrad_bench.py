"""weight-gradient GEMM: scaled-fp16 parts (gemm_tcgen05_wgrad16.cuh) vs 3xTF32, per layer shape of the pines MLP"""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mfm_b200 import _lib
lib = _lib.load(); cuda = torch.device("cuda:0")
fn = lib.mfm_debug_wgrad16; fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 7 + [ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p]
out = []
for n in (65536, 8192):
    for inn, o in ((1024, 1600), (1024, 1024), (2048, 1024), (1600, 1024), (256, 1024)):
        A = torch.relu(torch.randn(n, inn, device=cuda)); G = torch.randn(n, o, device=cuda) * 1e-6
        a_s, g_s = torch.empty_like(A), torch.empty_like(G)
        slots = torch.zeros(2, device=cuda); sb = torch.empty(16 * 2048 * 1024, device=cuda); dW = torch.empty(inn, o, device=cuda)
        st = torch.cuda.current_stream().cuda_stream
        def run(use): _lib.check(fn(n, inn, o, A.data_ptr(), G.data_ptr(), dW.data_ptr(), a_s.data_ptr(), g_s.data_ptr(), slots.data_ptr(), sb.data_ptr(), sb.numel(), use, st))
        run(1)
        res = {}
        for use in (2, 0):
            for _ in range(3): run(use)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20 if n > 10000 else 100
            e0.record()
            for _ in range(reps): run(use)
            e1.record(); torch.cuda.synchronize()
            res[use] = e0.elapsed_time(e1) / reps
        fl = 2.0 * n * inn * o
        row = dict(n=n, inn=inn, out=o, ms_fp16=res[2], ms_tf32=res[0], tflops_fp16=fl / res[2] * 1e-9, tflops_tf32=fl / res[0] * 1e-9, speedup=res[0] / res[2])
        print(json.dumps(row), flush=True); out.append(row)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r02_wgrad_bench.json", "w"), indent=1)
