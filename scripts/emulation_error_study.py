"""CPU study (NumPy) of fp32-product emulations on tensor cores: operand-rounding error only (accumulation in float64).

 current  : hi*hi in tf32 (operands truncated to 10 explicit mantissa bits) + cross terms a_lo*b + a*b_lo with bf16 operands
 fp16x3   : a = h + l, h = fp16(a), l = fp16(a - h);  h*h' + h*l' + l*h'   (three kind::f16 MMAs per 16 k-values, -25 % slots)
 fp16x3s  : the same after scaling each operand tensor by a power of two so that its largest magnitude is ~2^14
 bf16x3   : a = h + l with bf16 parts (16 mantissa bits in total), three MMAs

Inputs: one dense layer of the pines MLP fixture (activations after relu of a [n,1600]x[1600,1024] layer, fan-in weights) and
the backward-data form with small deltas (1e-3 scale), where fp16's range matters."""
import numpy as np

rng = np.random.default_rng(0)


def tf32_trunc(x):
    return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def bf16(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16 << 16).astype(np.uint32)
    return r.view(np.float32)


def fp16(x):
    with np.errstate(over="ignore"):
        return x.astype(np.float16).astype(np.float32)


def scheme_current(A, B):
    ah, bh = tf32_trunc(A), tf32_trunc(B)
    al, bl = A - ah, B - bh
    f = np.float64
    return ah.astype(f) @ bh.astype(f).T + bf16(al).astype(f) @ bf16(B).astype(f).T + bf16(A).astype(f) @ bf16(bl).astype(f).T


def scheme_split3(A, B, rnd):
    ah, bh = rnd(A), rnd(B)
    al, bl = rnd(A - ah), rnd(B - bh)
    f = np.float64
    return ah.astype(f) @ bh.astype(f).T + ah.astype(f) @ bl.astype(f).T + al.astype(f) @ bh.astype(f).T


def scaled(A):
    s = 2.0 ** np.floor(14 - np.log2(np.abs(A).max()))
    return (A * np.float32(s)).astype(np.float32), s


def report(name, A, B):
    exact = A.astype(np.float64) @ B.astype(np.float64).T
    scale = np.abs(exact).max()
    rows = np.sqrt((A.astype(np.float64) ** 2).sum(1))[:, None] * np.sqrt((B.astype(np.float64) ** 2).sum(1))[None, :]   # Cauchy-Schwarz scale
    out = {}
    out["current (tf32 + bf16 cross)"] = scheme_current(A, B)
    out["fp16 x3"] = scheme_split3(A, B, fp16)
    As, sa = scaled(A); Bs, sb = scaled(B)
    out["fp16 x3, power-of-two scaled"] = scheme_split3(As, Bs, fp16) / (sa * sb)
    out["bf16 x3"] = scheme_split3(A, B, bf16)
    print(f"--- {name}: A {A.shape} |max| {np.abs(A).max():.3g}, B {B.shape} |max| {np.abs(B).max():.3g}, |C|max {scale:.3g}")
    for k, v in out.items():
        e = np.abs(v - exact)
        print(f"    {k:32s} max err / |C|max = {e.max() / scale:.2e}   max err / (|a||b|) = {(e / rows).max():.2e}")


def main():
    n, K, N = 512, 1600, 1024
    X = (3.88 + 1.4 * rng.standard_normal((n, K))).astype(np.float32)                         # pines positions
    W2 = (rng.standard_normal((1024, K)) / np.sqrt(K)).astype(np.float32)
    H2 = np.maximum(X @ W2.T, 0).astype(np.float32)                                          # activations of Dense_2
    W3 = (rng.standard_normal((N, 1024)) / np.sqrt(1024)).astype(np.float32)
    report("forward Dense_2 (x @ W2)", X, W2)
    report("forward Dense_3 (relu(h2) @ W3)", H2, W3)
    delta = (1e-3 * rng.standard_normal((n, N))).astype(np.float32)                          # small backward signal
    report("backward-data (delta @ W3) with |delta| ~ 1e-3", delta, np.ascontiguousarray(W3.T))
    tiny = (1e-6 * rng.standard_normal((n, N))).astype(np.float32)
    report("backward-data with |delta| ~ 1e-6", tiny, np.ascontiguousarray(W3.T))


if __name__ == "__main__":
    main()
