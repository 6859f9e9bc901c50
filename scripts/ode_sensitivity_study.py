"""CPU study: how the adaptive Dopri5 solve of the pines flow reacts to the dense layers' operand rounding.

The oracle's field evaluation (float64) is re-run with every `activation @ weight` product replaced by an emulation of a
tensor-core scheme acting on fp32 operands (products accumulated in float64, so only OPERAND rounding is modelled):
  fp32      : operands rounded to fp32, exact products (an ideal fp32 GEMM)
  current   : tf32 hi*hi + bf16 cross terms (the shipped kernel)
  bf16x3    : two bf16 parts per operand, hi*hi' + hi*lo' + lo*hi' (the experimental split16 kernel)
  fp16x3s   : two fp16 parts after a power-of-two scale per tensor (the round-2 plan)
Reported per scheme: RK attempts / accepted steps per chain, and the deviation of the pushed samples and log-dets from the
float64 solve, next to the fp32 scheme's own deviation (the floor any fp32 implementation has)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import samplers as OS, targets as OT, threefry as tf, vector_field as VF  # noqa: E402
from scripts.emulation_error_study import bf16, fp16, tf32_trunc                       # noqa: E402

SCHEME = "f64"


def emulate(A, B):          # A [n,k] @ B [k,m], float64 in / out
    f = np.float64
    if SCHEME == "f64":
        return np.asarray(A, f) @ np.asarray(B, f)
    a, b = np.ascontiguousarray(A, np.float32), np.ascontiguousarray(B, np.float32)
    if SCHEME == "fp32":
        return a.astype(f) @ b.astype(f)
    if SCHEME == "current":
        ah, bh = tf32_trunc(a), tf32_trunc(b)
        return ah.astype(f) @ bh.astype(f) + bf16(a - ah).astype(f) @ bf16(b).astype(f) + bf16(a).astype(f) @ bf16(b - bh).astype(f)
    if SCHEME == "bf16x3":
        ah, bh = bf16(a), bf16(b)
        al, bl = bf16(a - ah), bf16(b - bh)
        return ah.astype(f) @ bh.astype(f) + ah.astype(f) @ bl.astype(f) + al.astype(f) @ bh.astype(f)
    if SCHEME == "fp16x3s":
        sa = 2.0 ** np.floor(14 - np.log2(max(np.abs(a).max(), 1e-30))); sb = 2.0 ** np.floor(14 - np.log2(max(np.abs(b).max(), 1e-30)))
        a2, b2 = (a * np.float32(sa)).astype(np.float32), (b * np.float32(sb)).astype(np.float32)
        ah, bh = fp16(a2), fp16(b2)
        al, bl = fp16(a2 - ah), fp16(b2 - bh)
        return (ah.astype(f) @ bh.astype(f) + ah.astype(f) @ bl.astype(f) + al.astype(f) @ bh.astype(f)) / (sa * sb)
    raise ValueError(SCHEME)


class W(np.ndarray):        # weight matrices: `activation @ W` dispatches here (right operand is a subclass overriding __rmatmul__)
    __array_priority__ = 100

    def __rmatmul__(self, left):
        return emulate(np.asarray(left), np.asarray(self))

    def __matmul__(self, right):
        return emulate(np.asarray(self), np.asarray(right))


def main():
    global SCHEME
    d, H, F, n = 1600, 1024, 128, 6
    ot = OT.LogGaussianCoxPines(d)
    rng = np.random.default_rng(0)
    params = VF.init_params(rng, d, H, F, head_scale=0.1, dtype=np.float64)
    for v in params["params"].values():
        v["kernel"] = v["kernel"].astype(np.float32).astype(np.float64).view(W)      # fp32-representable weights, as on the device
    omega = rng.standard_normal(F)
    keys = tf.split(tf.PRNGKey(9), n)
    u = tf.vmap_normal(tf.split(tf.PRNGKey(10), n), d, np.float32).astype(np.float64)
    z = tf.vmap_normal(keys, d, np.float32).astype(np.float64)
    res = {}
    for s in ("f64", "fp32", "current", "bf16x3", "fp16x3s"):
        SCHEME = s
        flow = OS.Flow(params, omega, ot, True, 1e-5, 1e-5, 1000, 1.0, (0.0, 1.0))
        st = {}
        t0 = time.perf_counter()
        y, ldj = flow.transform_and_logdet(keys, u, st, z=z)
        res[s] = (y, ldj, st["n_try"].copy(), st["n_acc"].copy())
        print(f"{s:8s} attempts {st['n_try'].tolist()} accepted {st['n_acc'].tolist()}  ({time.perf_counter() - t0:.0f} s)", flush=True)
    y0, l0 = res["f64"][:2]
    print("deviation from the float64 solve: max |dx| / max |x| , max |d ldj|")
    for s in ("fp32", "current", "bf16x3", "fp16x3s"):
        y, ldj = res[s][:2]
        print(f"  {s:8s} {np.abs(y - y0).max() / np.abs(y0).max():.2e}   {np.abs(ldj - l0).max():.2e}   same step counts as f64: {np.array_equal(res[s][2], res['f64'][2])}")


if __name__ == "__main__":
    main()
