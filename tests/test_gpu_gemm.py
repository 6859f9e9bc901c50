"""3xTF32 GEMM vs float64 matmul: fp32-level accuracy for every operand layout and ragged shape."""
import numpy as np
import pytest
import torch

from mfm_b200 import _lib

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, 16), (128, 128, 256), (300, 200, 100), (1, 2, 2), (129, 2, 130), (64, 1600, 1600),
          (1000, 128, 2), (257, 1024, 1024), (5, 7, 3), (2048, 256, 64),
          # tcgen05-eligible shapes (M>=128, N>=64, K>=32, 32-aligned MN-major dims), incl. ragged edges
          (512, 512, 512), (130, 96, 40), (384, 320, 1600), (1024, 1600, 1024), (160, 64, 33 * 4),
          # many tiles per CTA pair (persistent kernel: both TMEM buffers and every ring phase), ragged M and N
          (20000, 1088, 96), (40000 + 77, 512, 256)]


# accumulator-truncation bias per unit of K (relative to the largest output), by dense-layer backend:
#   0 auto: persistent CTA-pair kernel where eligible (hi*hi and both cross terms share one TMEM accumulator)
#   3 one-tile CTA-pair kernel / single-CTA kernel (cross terms in their own accumulator); 1 mma.sync (flushed every 32 k)
BIAS_PER_K = {0: 1.2e-8, 3: 2.5e-9, 1: 0.0}


@pytest.mark.parametrize("backend", [0, 3])
@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("akm,bnm", [(1, 1), (1, 0), (0, 1), (0, 0)])
def test_gemm_layouts(cuda, lib, M, N, K, akm, bnm, backend):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    Ad = torch.from_numpy(A if akm else np.ascontiguousarray(A.T)).to(cuda)
    Bd = torch.from_numpy(B if bnm else np.ascontiguousarray(B.T)).to(cuda)
    Cd = torch.full((M, N), float("nan"), dtype=torch.float32, device=cuda)
    bias_d = torch.from_numpy(bias).to(cuda)      # keep alive across the asynchronous launch
    lib.mfm_set_gemm_backend(backend)
    try:
        _lib.check(lib.mfm_gemm_tf32x3(M, N, K, Ad.data_ptr(), Ad.shape[1], akm, Bd.data_ptr(), Bd.shape[1], bnm,
                                       bias_d.data_ptr(), 1, Cd.data_ptr(), N,
                                       torch.cuda.current_stream().cuda_stream))
        got = Cd.cpu().numpy()
    finally:
        lib.mfm_set_gemm_backend(0)
    ref = np.maximum(A.astype(np.float64) @ B.astype(np.float64) + bias, 0)
    assert np.isfinite(got).all()
    # fp32-level: a few ulp of the largest output (3xTF32 drops only the lo*lo term, ~2^-22), plus the
    # tensor core's truncating fp32 accumulator over the whole K (the mma.sync kernel flushes every 32 k)
    tol = (2e-6 + BIAS_PER_K[backend] * K) * max(np.abs(ref).max(), 1.0)
    assert np.abs(got - ref).max() <= tol, (np.abs(got - ref).max(), tol)


@pytest.mark.parametrize("M,N,K", [(512, 512, 512), (384, 320, 1600), (1024, 1600, 1024), (20000, 1088, 96), (40000 + 77, 512, 256),
                                   (4096, 1024, 2048), (300, 200, 100)])
def test_gemm_bf16_cross_terms(cuda, lib, M, N, K):
    """K-major x K-major persistent kernel with the cross terms taken from one bf16 MMA per k-step."""
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((M, K)).astype(np.float32) * np.exp(rng.standard_normal((M, 1))).astype(np.float32)
    Bt = rng.standard_normal((N, K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    Ad, Bd, bias_d = torch.from_numpy(A).to(cuda), torch.from_numpy(Bt).to(cuda), torch.from_numpy(bias).to(cuda)
    Cd = torch.full((M, N), float("nan"), dtype=torch.float32, device=cuda)
    lib.mfm_set_gemm_cross_bf16(1)
    lib.mfm_set_gemm_h16(0)                     # this test is about the tf32 + bf16-cross kernel (the fallback of the h16 kernel)
    try:
        _lib.check(lib.mfm_gemm_tf32x3(M, N, K, Ad.data_ptr(), K, 1, Bd.data_ptr(), K, 0, bias_d.data_ptr(), 0, Cd.data_ptr(), N,
                                       torch.cuda.current_stream().cuda_stream))
        got = Cd.cpu().numpy()
    finally:
        lib.mfm_set_gemm_cross_bf16(1)         # the library defaults
        lib.mfm_set_gemm_h16(1)
    ref = A.astype(np.float64) @ Bt.astype(np.float64).T + bias
    # row-wise scale: the rows of A span two orders of magnitude
    scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1.0)
    err = (np.abs(got - ref) / scale).max()
    assert np.isfinite(got).all() and err <= 3e-6 + 8e-9 * K, err


@pytest.mark.parametrize("M,N,K", [(1024, 512, 256), (40000 + 77, 1024, 128), (700, 1600, 1024), (130, 96, 40)])
@pytest.mark.parametrize("use_mask,use_add,inplace", [(1, 0, 0), (0, 1, 0), (1, 1, 1)])
def test_gemm_gated_epilogue(cuda, lib, M, N, K, use_mask, use_add, inplace):
    """Backward-data form: C = (A Bt^T + add) * [mask > 0]; add may alias C (in-place accumulate)."""
    rng = np.random.default_rng(M + 3 * N + K + use_mask + 2 * use_add)
    A = rng.standard_normal((M, K)).astype(np.float32)
    Bt = rng.standard_normal((N, K)).astype(np.float32)
    mask = rng.standard_normal((M, N)).astype(np.float32)
    add = rng.standard_normal((M, N)).astype(np.float32)
    Ad, Bd = torch.from_numpy(A).to(cuda), torch.from_numpy(Bt).to(cuda)
    md, ad = torch.from_numpy(mask).to(cuda), torch.from_numpy(add).to(cuda)
    Cd = ad.clone() if inplace else torch.full((M, N), float("nan"), dtype=torch.float32, device=cuda)
    add_ptr = (Cd if inplace else ad).data_ptr() if use_add else None
    _lib.check(lib.mfm_gemm_tf32x3_gated(M, N, K, Ad.data_ptr(), K, Bd.data_ptr(), K, md.data_ptr() if use_mask else None, N,
                                         add_ptr, N, Cd.data_ptr(), N, torch.cuda.current_stream().cuda_stream))
    ref = A.astype(np.float64) @ Bt.astype(np.float64).T
    if use_add:
        ref = ref + add
    if use_mask:
        ref = np.where(mask > 0, ref, 0.0)
    got = Cd.cpu().numpy()
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() <= (3e-6 + 1.2e-8 * K) * max(np.abs(ref).max(), 1.0)


# shapes whose last round of 256 x 256 tiles is partially filled on 74 CTA pairs (tests/test_gemm_plan.py checks the
# work lists on the host): two-item pairs, integer split factors, fewer tiles than pairs, ragged M / N / K
STREAMK_SHAPES = [(8192, 1600, 1024), (1024, 1024, 1600), (4096 + 77, 1088, 520), (256, 256, 4096), (16384, 1600, 1600),
                  (5000 - 256, 1600, 1600), (19200, 1024, 1024), (8192, 1024, 1024)]


@pytest.mark.parametrize("M,N,K", STREAMK_SHAPES)
def test_gemm_streamk_matches_whole_tiles(cuda, lib, M, N, K):
    """Remainder round cut along K (partials through the L2 scratch, fixed summation order) vs whole tiles vs float64."""
    rng = np.random.default_rng(M + 5 * N + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    Bt = rng.standard_normal((N, K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    Ad, Bd, bias_d = torch.from_numpy(A).to(cuda), torch.from_numpy(Bt).to(cuda), torch.from_numpy(bias).to(cuda)
    st = torch.cuda.current_stream().cuda_stream
    outs = {}
    try:
        for mode in (1, 0, 1):
            lib.mfm_set_gemm_streamk(mode)
            for rep in range(2 if mode else 1):
                Cd = torch.full((M, N), float("nan"), dtype=torch.float32, device=cuda)
                _lib.check(lib.mfm_gemm_tf32x3(M, N, K, Ad.data_ptr(), K, 1, Bd.data_ptr(), K, 0, bias_d.data_ptr(), 1, Cd.data_ptr(), N, st))
                outs.setdefault(mode, []).append(Cd.cpu().numpy())
    finally:
        lib.mfm_set_gemm_streamk(1)
    ref = np.maximum(A.astype(np.float64) @ Bt.astype(np.float64).T + bias, 0)
    scale = max(np.abs(ref).max(), 1.0)
    for mode, lst in outs.items():
        for got in lst:
            assert np.isfinite(got).all()
            assert np.abs(got - ref).max() <= (3e-6 + 8e-9 * K) * scale, (mode, np.abs(got - ref).max() / scale)
    # deterministic: four launches (epochs 1, 2, then 3, 4 after a whole-tile launch) give the same bits
    for got in outs[1][1:]:
        assert np.array_equal(got, outs[1][0])
    # the two differ by the accumulator-truncation bias (partials are added with round-to-nearest)
    assert np.abs(outs[1][0] - outs[0][0]).max() <= (3e-6 + 8e-9 * K) * scale


@pytest.mark.parametrize("M,N,K", [(8192, 1600, 1024), (1024, 1024, 1024)])
def test_gemm_streamk_gated_epilogue(cuda, lib, M, N, K):
    """Stream-K under the backward-data epilogue (mask and in-place residual operands are read by the finishing pair only)."""
    rng = np.random.default_rng(M + N)
    A = rng.standard_normal((M, K)).astype(np.float32)
    Bt = rng.standard_normal((N, K)).astype(np.float32)
    mask = rng.standard_normal((M, N)).astype(np.float32)
    add = rng.standard_normal((M, N)).astype(np.float32)
    Ad, Bd, md = torch.from_numpy(A).to(cuda), torch.from_numpy(Bt).to(cuda), torch.from_numpy(mask).to(cuda)
    Cd = torch.from_numpy(add).to(cuda)
    lib.mfm_set_gemm_streamk(1)
    _lib.check(lib.mfm_gemm_tf32x3_gated(M, N, K, Ad.data_ptr(), K, Bd.data_ptr(), K, md.data_ptr(), N, Cd.data_ptr(), N, Cd.data_ptr(), N,
                                         torch.cuda.current_stream().cuda_stream))
    ref = np.where(mask > 0, A.astype(np.float64) @ Bt.astype(np.float64).T + add, 0.0)
    got = Cd.cpu().numpy()
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() <= (3e-6 + 1.2e-8 * K) * max(np.abs(ref).max(), 1.0)


@pytest.mark.parametrize("rows", [8192, 5000, 4800, 700, 256, 1, 0])
def test_gemm_streamk_device_row_count(cuda, lib, rows):
    """Active-row count read on the device (ODE compaction): the kernel derives its work list from the live M."""
    M, N, K = 9000, 1024, 1024
    rng = np.random.default_rng(rows)
    A = rng.standard_normal((M, K)).astype(np.float32)
    Bt = rng.standard_normal((N, K)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    Ad, Bd, bias_d = torch.from_numpy(A).to(cuda), torch.from_numpy(Bt).to(cuda), torch.from_numpy(bias).to(cuda)
    Cd = torch.full((M, N), -7.0, dtype=torch.float32, device=cuda)
    nrows = torch.tensor([rows], dtype=torch.int32, device=cuda)
    _lib.check(lib.mfm_gemm_tf32x3_rows(M, N, K, Ad.data_ptr(), K, Bd.data_ptr(), K, bias_d.data_ptr(), Cd.data_ptr(), N, nrows.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream))
    got = Cd.cpu().numpy()
    ref = A[:rows].astype(np.float64) @ Bt.astype(np.float64).T + bias
    assert (got[rows:] == -7.0).all()
    if rows:
        assert np.abs(got[:rows] - ref).max() <= (3e-6 + 8e-9 * K) * np.abs(ref).max()


@pytest.mark.parametrize("M,N,K", [(512, 512, 512), (8192, 1600, 1024), (4096 + 77, 1088, 520), (1024, 1024, 1600), (300, 256, 64)])
def test_gemm_presplit_b_operand(cuda, lib, M, N, K):
    """tf32 + bf16-cross kernel (h16 switched off): B operand's bf16 cross tile pre-split in global memory and loaded by TMA:
    the same bits reach the tensor core, so the result is bit-identical to splitting B in the kernel."""
    rng = np.random.default_rng(M + N + K + 1)
    A = rng.standard_normal((M, K)).astype(np.float32)
    Bt = (rng.standard_normal((N, K)) * np.exp(rng.standard_normal((N, 1)))).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    Ad, Bd, bias_d = torch.from_numpy(A).to(cuda), torch.from_numpy(Bt).to(cuda), torch.from_numpy(bias).to(cuda)
    mirror = torch.empty(N * K + 16, dtype=torch.float32, device=cuda)
    st = torch.cuda.current_stream().cuda_stream
    lib.mfm_set_gemm_h16(0)
    try:
        _lib.check(lib.mfm_gemm_presplit(Bd.data_ptr(), mirror.data_ptr(), N * K, st))
        # mirror layout: per 8 floats, 8 bf16 of the values then 8 bf16 of (value - tf32 truncation)
        mb = mirror[:N * K].view(torch.bfloat16).view(N, K // 8, 2, 8).float().cpu().numpy()
        vals = Bt.reshape(N, K // 8, 8)
        trunc = (vals.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
        assert np.array_equal(mb[:, :, 0], torch.from_numpy(vals).bfloat16().float().numpy())
        assert np.array_equal(mb[:, :, 1], torch.from_numpy(vals - trunc).bfloat16().float().numpy())
        outs = []
        for use in (0, 1):
            Cd = torch.full((M, N), float("nan"), dtype=torch.float32, device=cuda)
            if use:
                lib.mfm_gemm_register_mirror(Bd.data_ptr(), N * K, mirror.data_ptr())
            try:
                _lib.check(lib.mfm_gemm_tf32x3(M, N, K, Ad.data_ptr(), K, 1, Bd.data_ptr(), K, 0, bias_d.data_ptr(), 1, Cd.data_ptr(), N, st))
                outs.append(Cd.cpu().numpy())
            finally:
                lib.mfm_gemm_register_mirror(None, 0, None)
    finally:
        lib.mfm_set_gemm_h16(1)
    ref = np.maximum(A.astype(np.float64) @ Bt.astype(np.float64).T + bias, 0)
    assert np.abs(outs[1] - ref).max() <= (3e-6 + 8e-9 * K) * max(np.abs(ref).max(), 1.0)
    assert np.array_equal(outs[0], outs[1])


def _h16_split_ref(x, amax):
    """numpy restatement of the scaled fp16 split (gemm_tcgen05_h16.cuh): s = 2^(14 - floor(log2 amax)), hi = fp16(x s), lo = fp16(x s - hi)"""
    s = np.float32(2.0) ** np.float32(14 - np.floor(np.log2(np.float32(amax))))
    xs = (x * s).astype(np.float32)
    hi = xs.astype(np.float16)
    return hi, (xs - hi.astype(np.float32)).astype(np.float16), s


@pytest.mark.parametrize("M,N,K", [(512, 512, 512), (8192, 1024, 1024), (8192, 1600, 1024), (4096 + 77, 1088, 528), (65536, 1024, 1024),
                                   (300, 256, 64), (333, 128, 128), (333, 128, 256), (1024, 128, 128), (65536, 128, 128)])
@pytest.mark.parametrize("a_scale,b_scale", [(1.0, 1.0), (3e-7, 1.0), (2.5e4, 1e-3)])
def test_gemm_h16(cuda, lib, M, N, K, a_scale, b_scale):
    """Default dense-layer kernel: operands scaled per tensor and split into two fp16 parts, three kind::f16 MMAs per 16 k.
    fp32-level accuracy for O(1) data, for back-propagation-sized (3e-7) and for large (2.5e4) operands alike - the per-tensor
    power-of-two scale is what makes fp16's 5 exponent bits sufficient - with the weight operand pre-split or split in the
    kernel, and with the maximum of A tracked by its producer or reduced on the fly."""
    rng = np.random.default_rng(M + N + K + 2)
    A = (rng.standard_normal((M, K)) * a_scale).astype(np.float32)
    Bt = (rng.standard_normal((N, K)) / np.sqrt(K) * b_scale).astype(np.float32)
    bias = (rng.standard_normal(N) * a_scale * b_scale).astype(np.float32)
    Ad, Bd, bias_d = torch.from_numpy(A).to(cuda), torch.from_numpy(Bt).to(cuda), torch.from_numpy(bias).to(cuda)
    mirror = torch.empty(N * K + 16, dtype=torch.float32, device=cuda)
    st = torch.cuda.current_stream().cuda_stream
    ref = np.maximum(A.astype(np.float64) @ Bt.astype(np.float64).T + bias, 0)
    tol = (1.5e-6 + 6e-9 * K) * np.abs(ref).max()
    a_amax = torch.zeros(1, dtype=torch.float32, device=cuda)
    _lib.check(lib.mfm_absmax(Ad.data_ptr(), K, M, K, a_amax.data_ptr(), st))
    assert a_amax.item() == np.abs(A).max()
    outs = {}
    try:
        for name, use_mirror, use_amax in [("kernel-split B, reduced max", 0, 0), ("mirror, tracked max", 1, 1), ("mirror, reduced max", 1, 0)]:
            if use_mirror:
                _lib.check(lib.mfm_gemm_presplit(Bd.data_ptr(), mirror.data_ptr(), N * K, st))
                if K % 16 == 0 and name.startswith("mirror, tracked"):
                    hi, lo, s = _h16_split_ref(Bt.reshape(N, K // 16, 16), np.abs(Bt).max())
                    mb = mirror[:N * K].view(torch.float16).view(N, K // 16, 2, 16).cpu().numpy()
                    assert np.array_equal(mb[:, :, 0], hi) and np.array_equal(mb[:, :, 1], lo)
                    assert mirror[N * K].item() == np.abs(Bt).max()
                lib.mfm_gemm_register_mirror(Bd.data_ptr(), N * K, mirror.data_ptr())
            Cd = torch.full((M, N), float("nan"), dtype=torch.float32, device=cuda)
            c_amax = torch.zeros(1, dtype=torch.float32, device=cuda)
            _lib.check(lib.mfm_gemm_dense(M, N, K, Ad.data_ptr(), K, Bd.data_ptr(), K, bias_d.data_ptr(), 1, Cd.data_ptr(), N,
                                          a_amax.data_ptr() if use_amax else None, c_amax.data_ptr(), None, None, st))
            got = Cd.cpu().numpy()
            lib.mfm_gemm_register_mirror(None, 0, None)
            assert np.isfinite(got).all(), name
            assert np.abs(got - ref).max() <= tol, (name, np.abs(got - ref).max(), tol)
            assert c_amax.item() == got.max(), name          # the producer-side maximum the next layer would consume (relu: max = max |.|)
            outs[name] = got
    finally:
        lib.mfm_gemm_register_mirror(None, 0, None)
    # the same parts reach the tensor core whichever way they were made (narrow layers without a tracked maximum are not worth
    # the reduction pass and run on the tf32 + bf16-cross kernel instead: only the tolerance above applies to them)
    if N * K >= 512 * 512:
        assert np.array_equal(outs["mirror, tracked max"], outs["mirror, reduced max"])
        assert np.array_equal(outs["mirror, tracked max"], outs["kernel-split B, reduced max"])

def test_gemm_strided_views(cuda, lib):
    """ld > logical width (writing into a column block of a concatenated buffer)."""
    rng = np.random.default_rng(0)
    M, N, K = 200, 96, 72
    Abig = rng.standard_normal((M, 160)).astype(np.float32)
    B = rng.standard_normal((K, N)).astype(np.float32)
    Ad = torch.from_numpy(Abig).to(cuda)
    Bd = torch.from_numpy(B).to(cuda)
    Cd = torch.zeros((M, 256), dtype=torch.float32, device=cuda)
    off = 40
    _lib.check(lib.mfm_gemm_tf32x3(M, N, K, Ad.data_ptr() + 4 * off, 160, 1, Bd.data_ptr(), N, 1, None, 0,
                                   Cd.data_ptr() + 4 * 128, 256, torch.cuda.current_stream().cuda_stream))
    ref = Abig[:, off:off + K].astype(np.float64) @ B
    got = Cd.cpu().numpy()
    assert np.abs(got[:, 128:128 + N] - ref).max() < 5e-5
    assert (got[:, :128] == 0).all() and (got[:, 128 + N:] == 0).all()


def test_gemm_backends_agree(cuda, lib):
    """tcgen05 and mma.sync kernels implement the same 3xTF32 arithmetic."""
    rng = np.random.default_rng(3)
    M, N, K = 640, 768, 1024
    A = torch.from_numpy(rng.standard_normal((M, K)).astype(np.float32)).to(cuda)
    B = torch.from_numpy(rng.standard_normal((K, N)).astype(np.float32)).to(cuda)
    outs = []
    for backend in (0, 1):
        lib.mfm_set_gemm_backend(backend)
        C = torch.empty((M, N), dtype=torch.float32, device=cuda)
        _lib.check(lib.mfm_gemm_tf32x3(M, N, K, A.data_ptr(), K, 1, B.data_ptr(), N, 1, None, 0, C.data_ptr(), N,
                                       torch.cuda.current_stream().cuda_stream))
        outs.append(C.cpu().numpy().astype(np.float64))
    lib.mfm_set_gemm_backend(0)
    ref = A.cpu().numpy().astype(np.float64) @ B.cpu().numpy().astype(np.float64)
    e_tc, e_mma = np.abs(outs[0] - ref).max(), np.abs(outs[1] - ref).max()
    scale = np.abs(ref).max()
    assert e_tc < (2e-6 + BIAS_PER_K[0] * K) * scale and e_mma < 2e-6 * scale, (e_tc / scale, e_mma / scale)


@pytest.mark.parametrize("M,K,N1,N2", [(512, 512, 512, 256), (8192, 1024, 1024, 1600), (4096 + 77, 528, 1088, 512), (65536, 256, 1024, 1024),
                                       (333, 256, 128, 128), (1024, 128, 128, 128)])
@pytest.mark.parametrize("a_scale", [1.0, 2e-6])
def test_gemm_presplit_handover(cuda, lib, M, K, N1, N2, a_scale):
    """Two chained dense layers as the MLP runs them: the first epilogue writes its result as fp32 AND as the pre-split
    (scaled fp16 hi | lo) A operand of the second, scaled by the power of two of its output BOUND
    max|A| * max_n sum_k |B1[n][k]| + max|b1| (the exact maximum does not exist before the last tile is written); the second
    layer's TMA loads that copy and nothing is split in shared memory.  Checks the copy bit for bit against a numpy restatement,
    that the bound really bounds, and both results against float64."""
    import ctypes
    rng = np.random.default_rng(M + K + N1)
    A = (rng.standard_normal((M, K)) * a_scale).astype(np.float32)
    B1t = (rng.standard_normal((N1, K)) / np.sqrt(K)).astype(np.float32)
    b1 = (rng.standard_normal(N1) * 0.1 * a_scale).astype(np.float32)
    B2t = (rng.standard_normal((N2, N1)) / np.sqrt(N1)).astype(np.float32)
    dev = lambda a: torch.from_numpy(a).to(cuda)
    Ad, B1d, b1d, B2d = dev(A), dev(B1t), dev(b1), dev(B2t)
    C1 = torch.full((M, N1), float("nan"), dtype=torch.float32, device=cuda)
    C1s = torch.zeros((M, N1), dtype=torch.float32, device=cuda)
    C2 = torch.full((M, N2), float("nan"), dtype=torch.float32, device=cuda)
    wn = np.float32(np.abs(B1t).sum(axis=1, dtype=np.float32).max())
    wnorm, bmax = dev(np.array([wn], np.float32)), dev(np.array([np.abs(b1).max()], np.float32))
    slots = torch.zeros(4, dtype=torch.float32, device=cuda)
    mirrors = []
    try:
        for Bd in (B1d, B2d):         # weight mirrors, as the MLP's layers have them
            m = torch.empty(Bd.numel() + 16, dtype=torch.float32, device=cuda)
            _lib.check(lib.mfm_gemm_presplit(Bd.data_ptr(), m.data_ptr(), Bd.numel(), torch.cuda.current_stream().cuda_stream))
            lib.mfm_gemm_register_mirror(Bd.data_ptr(), Bd.numel(), m.data_ptr())
            mirrors.append(m)
        fn = lib.mfm_debug_dense_chain
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p] * 10 + [ctypes.c_void_p]
        _lib.check(fn(M, K, N1, N2, Ad.data_ptr(), B1d.data_ptr(), b1d.data_ptr(), B2d.data_ptr(), C1.data_ptr(), C1s.data_ptr(), C2.data_ptr(),
                      wnorm.data_ptr(), bmax.data_ptr(), slots.data_ptr(), torch.cuda.current_stream().cuda_stream))
        c1, c2, sl = C1.cpu().numpy(), C2.cpu().numpy(), slots.cpu().numpy()
    finally:
        lib.mfm_gemm_register_mirror(None, 0, None)
    ref1 = np.maximum(A.astype(np.float64) @ B1t.astype(np.float64).T + b1, 0)
    assert np.abs(c1 - ref1).max() <= (1.5e-6 + 6e-9 * K) * np.abs(ref1).max()
    assert sl[0] == np.abs(A).max() and sl[1] == c1.max()
    assert sl[2] >= sl[1] and sl[2] <= 1.01 * (np.abs(A).max() * wn + np.abs(b1).max())          # a bound, and the one documented
    hi, lo, s = _h16_split_ref(c1.reshape(M, N1 // 16, 16), sl[2])
    got = C1s.view(torch.float16).view(M, N1 // 16, 2, 16).cpu().numpy()
    assert np.array_equal(got[:, :, 0], hi) and np.array_equal(got[:, :, 1], lo)
    assert np.isfinite(hi.astype(np.float32)).all()                                             # no overflow, by construction
    ref2 = c1.astype(np.float64) @ B2t.astype(np.float64).T
    assert np.abs(c2 - ref2).max() <= (1.5e-6 + 6e-9 * N1) * np.abs(ref2).max()


@pytest.mark.parametrize("n,inn,out", [(512, 256, 128), (4096, 1024, 1024), (8192, 1600, 1024), (8192 + 77, 1024, 1600), (65536, 2048, 1024),
                                       (65536, 256, 1024), (1000, 272, 144)])
@pytest.mark.parametrize("scales", [(1.0, 1.0), (1.0, 3e-7)])
def test_wgrad_fp16_parts(cuda, lib, n, inn, out, scales):
    """Weight gradient dW = A^T G from the split16 copies of both operands (csrc/gemm_tcgen05_wgrad16.cuh: MN-major fp16 tiles
    picked out of the K-major copies by a 4-D TMA map, three tensor-core passes, split-K over the batch) against float64, and
    against the 3xTF32 kernel it replaces.  Back-propagated signals are tiny (1 / (n d) factors): the second scale set."""
    import ctypes
    rng = np.random.default_rng(n + inn + out)
    A = (np.maximum(rng.standard_normal((n, inn)), 0) * scales[0]).astype(np.float32)       # activations: half zeros
    G = (rng.standard_normal((n, out)) * scales[1]).astype(np.float32)
    G[rng.random((n, out)) < 0.3] = 0.0
    dev = lambda a: torch.from_numpy(a).to(cuda)
    Ad, Gd = dev(A), dev(G)
    a_s, g_s = torch.empty_like(Ad), torch.empty_like(Gd)
    slots = torch.zeros(2, dtype=torch.float32, device=cuda)
    sb = torch.empty(16 * inn * out, dtype=torch.float32, device=cuda)
    fn = lib.mfm_debug_wgrad16
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 7 + [ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p]
    res = {}
    for use in (1, 0):
        dW = torch.full((inn, out), float("nan"), dtype=torch.float32, device=cuda)
        _lib.check(fn(n, inn, out, Ad.data_ptr(), Gd.data_ptr(), dW.data_ptr(), a_s.data_ptr(), g_s.data_ptr(), slots.data_ptr(),
                      sb.data_ptr(), sb.numel(), use, torch.cuda.current_stream().cuda_stream))
        res[use] = dW.cpu().numpy()
    ref = (Ad.double().T @ Gd.double()).cpu().numpy()
    scale = np.abs(ref).max()
    e16, e32 = np.abs(res[1] - ref).max() / scale, np.abs(res[0] - ref).max() / scale
    assert np.isfinite(res[1]).all()
    # operand rounding 2^-22 per part pair, plus the accumulator's truncation (2.5e-9 per accumulated k, main and cross kept apart)
    assert e16 <= 1e-6 + 2.5e-9 * min(n, 8192), (e16, e32)
    assert e16 <= max(2.0 * e32, 2.5e-6), (e16, e32)      # (small shapes: the 3xTF32 reference runs on mma.sync, which rounds to nearest)
