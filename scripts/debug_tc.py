import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from mfm_b200 import _lib
lib = _lib.load()
cuda = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
def run(M, N, K, akm, bnm, A, B):
    Ad = torch.from_numpy(A if akm else np.ascontiguousarray(A.T)).to(cuda)
    Bd = torch.from_numpy(B if bnm else np.ascontiguousarray(B.T)).to(cuda)
    Cd = torch.full((M, N), float("nan"), dtype=torch.float32, device=cuda)
    _lib.check(lib.mfm_gemm_tf32x3(M, N, K, Ad.data_ptr(), Ad.shape[1], akm, Bd.data_ptr(), Bd.shape[1], bnm, None, 0, Cd.data_ptr(), N, st))
    torch.cuda.synchronize()
    return Cd.cpu().numpy()
for (M, N, K) in [(128, 256, 32), (128, 256, 64), (256, 512, 128)]:
    for akm, bnm in [(1, 0), (1, 1), (0, 0), (0, 1)]:
        A = np.zeros((M, K), np.float32); A[np.arange(M), np.arange(M) % K] = 1.0     # C[m,n] = B[m%K, n]
        B = (np.arange(K)[:, None] * 1000 + np.arange(N)[None, :]).astype(np.float32)
        C = run(M, N, K, akm, bnm, A, B)
        ref = A.astype(np.float64) @ B
        bad = np.argwhere(np.abs(C - ref) > 0.5)
        print(f"M{M} N{N} K{K} akm{akm} bnm{bnm}: bad {len(bad)}/{M*N}", flush=True)
        if len(bad):
            for (m, n) in bad[:6]:
                print("   m", m, "n", n, "got", C[m, n], "ref", ref[m, n])
            print("   bad rows", sorted(set(bad[:, 0]))[:20], "bad cols", sorted(set(bad[:, 1]))[:20])
        # random test
        rng = np.random.default_rng(0)
        A = rng.standard_normal((M, K)).astype(np.float32); B = rng.standard_normal((K, N)).astype(np.float32)
        C = run(M, N, K, akm, bnm, A, B); ref = A.astype(np.float64) @ B
        print("   random max err", np.abs(C - ref).max(), "scale", np.abs(ref).max())
