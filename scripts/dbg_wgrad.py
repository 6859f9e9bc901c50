import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mfm_b200 import _lib
lib = _lib.load(); cuda = torch.device("cuda:0")
n, inn, out = [int(a) for a in sys.argv[1:4]] if len(sys.argv) > 3 else (512, 256, 128)
rng = np.random.default_rng(0)
A = np.maximum(rng.standard_normal((n, inn)), 0).astype(np.float32); G = rng.standard_normal((n, out)).astype(np.float32)
Ad, Gd = torch.from_numpy(A).to(cuda), torch.from_numpy(G).to(cuda)
a_s, g_s = torch.empty_like(Ad), torch.empty_like(Gd)
slots = torch.zeros(2, dtype=torch.float32, device=cuda); sb = torch.empty(16 * inn * out, dtype=torch.float32, device=cuda)
fn = lib.mfm_debug_wgrad16; fn.restype = ctypes.c_int
fn.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 7 + [ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p]
ref = (Ad.double().T @ Gd.double()).cpu().numpy()
for use in (1, 0):
    dW = torch.full((inn, out), float("nan"), dtype=torch.float32, device=cuda)
    rc = fn(n, inn, out, Ad.data_ptr(), Gd.data_ptr(), dW.data_ptr(), a_s.data_ptr(), g_s.data_ptr(), slots.data_ptr(), sb.data_ptr(), sb.numel(), use, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    r = dW.cpu().numpy()
    print("use", use, "rc", rc, "err", np.abs(r - ref).max() / np.abs(ref).max(), "nan", np.isnan(r).sum(), flush=True)
    if use == 1 and np.abs(r - ref).max() / np.abs(ref).max() > 1e-3:
        # which structure? compare a few entries
        print(r[:4, :4]); print(ref[:4, :4])
        e = np.abs(r - ref) / np.abs(ref).max()
        print("bad rows", np.where(e.max(1) > 1e-3)[0][:20], "bad cols", np.where(e.max(0) > 1e-3)[0][:20])
