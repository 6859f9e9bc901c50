"""One dense layer as the FM update launches it (for ncu captures): [n,K] x [N,K]^T, weight mirror registered, max |A| tracked.
usage: python scripts/gemm_one.py [n N K reps]"""
import sys

import torch

sys.path.insert(0, ".")
from mfm_b200 import _lib      # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
st = torch.cuda.current_stream().cuda_stream
n, N, K, reps = (int(v) for v in (sys.argv[1:5] if len(sys.argv) >= 5 else (65536, 1024, 1024, 6)))
A = torch.randn(n, K, device=dev); Bt = torch.randn(N, K, device=dev) / K ** 0.5; C = torch.empty(n, N, device=dev)
bias = torch.randn(N, device=dev)
amax = torch.zeros(1, device=dev)
_lib.check(lib.mfm_absmax(A.data_ptr(), K, n, K, amax.data_ptr(), st))
mirror = torch.empty(N * K + 16, device=dev)
_lib.check(lib.mfm_gemm_presplit(Bt.data_ptr(), mirror.data_ptr(), N * K, st))
lib.mfm_gemm_register_mirror(Bt.data_ptr(), N * K, mirror.data_ptr())
a_s = torch.empty(n * K + 16, device=dev)            # A as the producing layer's epilogue would leave it (pre-split)
_lib.check(lib.mfm_gemm_presplit(A.data_ptr(), a_s.data_ptr(), n * K, st))
for _ in range(reps):
    _lib.check(lib.mfm_gemm_dense(n, N, K, A.data_ptr(), K, Bt.data_ptr(), K, bias.data_ptr(), 1, C.data_ptr(), N, amax.data_ptr(), None,
                                  a_s.data_ptr(), a_s.data_ptr() + 4 * n * K, st))
torch.cuda.synchronize()
print("ok", float(C[0, 0]))
