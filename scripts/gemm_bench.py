"""GEMM microbenchmark + accuracy check for the three 3xTF32 dense-layer kernels (run on a B200).

For every (shape, layout) of the pines hot path: max error against a float64 matmul, then CUDA-event
timing (L2 flushed between launches) of the persistent CTA-pair kernel, the one-tile CTA-pair kernel,
the single-CTA tcgen05 kernel and, for reference, the mma.sync kernel.
usage: python scripts/gemm_bench.py [--quick] [--json out.json]
"""
import argparse
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from mfm_b200 import _lib  # noqa: E402


def run(lib, M, N, K, akm, bnm, A, B, bias, C, stream):
    _lib.check(lib.mfm_gemm_tf32x3(M, N, K, A.data_ptr(), A.shape[1], akm, B.data_ptr(), B.shape[1], bnm,
                                   bias.data_ptr() if bias is not None else None, 1 if bias is not None else 0,
                                   C.data_ptr(), N, stream))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--json", default=None)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--timeline", action="store_true", help="needs a MFM_TC2_TIMELINE=1 build")
    ap.add_argument("--sustained", type=float, default=0.0, help="also run each kernel back to back for this many seconds, sampling nvidia-smi clocks/power")
    ap.add_argument("--only", default=None, help="comma-separated shape labels")
    ap.add_argument("--kernels", default="persist,persist_bf16x,pair,tc1,mma")
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    n = 8192 if args.quick else 65536
    # (M, N, K, a_kmajor, b_nmajor, label)
    shapes = [
        (n, 1024, 1024, 1, 1, "fwd H->H"),
        (n, 1024, 256, 1, 1, "fwd 2F->H"),
        (n, 1024, 1600, 1, 1, "fwd d->H"),
        (n, 1600, 1024, 1, 1, "fwd H->d"),
        (n, 1024, 2048, 1, 1, "fwd 2H->H"),
        (n, 1600, 1600, 1, 1, "pines Kinv"),
        (n, 1024, 1024, 1, 0, "dgrad H<-H"),
        (n, 1024, 1600, 1, 0, "dgrad H<-d"),
        (n, 1600, 1600, 1, 0, "Kinv KxK"),
        (n, 1024, 2048, 1, 0, "2H->H KxK"),
        (n, 1024, 256, 1, 0, "2F->H KxK"),
        (1024, 1024, n, 0, 1, "wgrad HxH"),
        (1024, 1600, n, 0, 1, "wgrad Hxd"),
        (1600, 1024, n, 0, 1, "wgrad dxH"),
    ]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    out = []
    g = torch.Generator(device=dev); g.manual_seed(0)
    only = set(args.only.split(",")) if args.only else None
    kernels = set(args.kernels.split(","))
    for (M, N, K, akm, bnm, label) in shapes:
        if only and label not in only:
            continue
        A = torch.randn((M, K) if akm else (K, M), generator=g, device=dev, dtype=torch.float32)
        B = torch.randn((K, N) if bnm else (N, K), generator=g, device=dev, dtype=torch.float32)
        bias = torch.randn(N, generator=g, device=dev, dtype=torch.float32)
        C = torch.empty((M, N), dtype=torch.float32, device=dev)
        # float64 reference on a row/column sample (the full product would take too long in fp64)
        rows = torch.randint(0, M, (64,), generator=g, device=dev)
        Am = (A if akm else A.t())[rows].double()
        Bm = (B if bnm else B.t()).double()
        ref = torch.relu(Am @ Bm + bias.double())
        rec = {"label": label, "M": M, "N": N, "K": K, "akm": akm, "bnm": bnm}
        for name, backend, raw in [("persist", 0, 1), ("persist_bf16x", 0, 1), ("pair", 3, 1), ("pair_rnsplit", 3, 0), ("tc1", 2, 0), ("mma", 1, 0)]:
            if name not in kernels or (name == "mma" and not args.quick and M * N * K > 2e11):
                continue
            if name == "persist_bf16x" and not (akm == 1 and bnm == 0):
                continue
            lib.mfm_set_gemm_backend(backend)
            lib.mfm_set_gemm_raw_hi(raw)
            lib.mfm_set_gemm_cross_bf16(1 if name == "persist_bf16x" else 0)
            C.fill_(float("nan"))
            run(lib, M, N, K, akm, bnm, A, B, bias, C, st)
            torch.cuda.synchronize()
            err = (C[rows].double() - ref).abs().max().item() / max(ref.abs().max().item(), 1.0)
            finite = bool(torch.isfinite(C).all().item())
            ts = []
            for _ in range(args.reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); run(lib, M, N, K, akm, bnm, A, B, bias, C, st); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            if args.timeline and name == "pair":
                import ctypes
                buf = (ctypes.c_longlong * 16)()
                lib.mfm_debug_gemm_timeline(1, None)
                run(lib, M, N, K, akm, bnm, A, B, bias, C, st)
                torch.cuda.synchronize()
                lib.mfm_debug_gemm_timeline(0, buf)
                t = list(buf)
                names = ["entry", "prologue done", "first TMA landed", "first split done", "first MMA issue",
                         "last MMA issued", "accumulator ready", "epilogue done", "after final cluster sync"]
                print("   timeline (SM clocks since entry): " + ", ".join(f"{nm}={t[i] - t[0]}" for i, nm in enumerate(names)))
            rec[name] = {"rel_err": err, "finite": finite, "ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9}
            if args.sustained > 0:
                import subprocess, time
                mon = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits",
                                        "-lms", "100", "-i", "0"], stdout=subprocess.PIPE, text=True)
                n_it = max(10, int(args.sustained * 1e3 / ms))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); time.sleep(0.3)
                e0.record()
                for _ in range(n_it):
                    run(lib, M, N, K, akm, bnm, A, B, bias, C, st)
                e1.record(); torch.cuda.synchronize()
                mon.terminate(); lines = mon.stdout.read().strip().splitlines()
                vals = [l.split(",") for l in lines if l.count(",") == 2]
                clk = [float(v[0]) for v in vals][3:]; pw = [float(v[1]) for v in vals][3:]
                sms = e0.elapsed_time(e1) / n_it
                rec[name]["sustained"] = {"ms": sms, "tflops": 2.0 * M * N * K / sms / 1e9, "sm_mhz_median": float(np.median(clk)) if clk else None,
                                          "power_w_median": float(np.median(pw)) if pw else None}
                print(f"      sustained {n_it} launches: {sms:8.3f} ms {2.0 * M * N * K / sms / 1e9:7.1f} TFLOP/s  sm_mhz median "
                      f"{np.median(clk) if clk else -1:.0f} min {min(clk) if clk else -1:.0f}  power median {np.median(pw) if pw else -1:.0f} W  "
                      f"power_cap_active {sum(1 for v in vals if 'Active' in v[2] and 'Not' not in v[2])}/{len(vals)}", flush=True)
            print(f"{label:12s} {name:10s} M={M} N={N} K={K} err={err:.2e} finite={finite} {ms:8.3f} ms "
                  f"{rec[name]['tflops']:7.1f} TFLOP/s", flush=True)
        out.append(rec)
    lib.mfm_set_gemm_backend(0); lib.mfm_set_gemm_raw_hi(1); lib.mfm_set_gemm_cross_bf16(0)
    if args.json:
        with open(args.json, "w") as fh:
            json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
