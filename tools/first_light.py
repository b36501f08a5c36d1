"""GPU bring-up diagnostics for the tcgen05 engine: each case runs in its own subprocess (a trap or hang in
one cannot take the others down) and reports an error PATTERN, not just pass/fail. Writes gpurun_out/first_light.log.
Usage (on the GPU box): python tools/first_light.py"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASE_SRC = r'''
import sys, json, numpy as np
sys.path.insert(0, %r)
import multipathnet_b200 as mpn
kind, impl, args = sys.argv[1], int(sys.argv[2]), json.loads(sys.argv[3])
ctx = mpn.Context(0)
rng = np.random.default_rng(0)
def bf16r(x):
    import torch
    return torch.from_numpy(x).bfloat16().float().numpy()
if kind == "gemm":
    M, N, K = args
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    got = ctx.gemm_check(A, B, None, relu=False, impl=impl)
    ref = (A.astype(np.float64) @ B.astype(np.float64).T)
    Ah, Bh = bf16r(A), bf16r(B)
    hh = Ah.astype(np.float64) @ Bh.astype(np.float64).T
    e = np.abs(got - ref); s = np.abs(ref).max()
    out = dict(rel=float(e.max() / s), rel_vs_hihi=float(np.abs(got - hh).max() / s), nan=int(np.isnan(got).sum()),
               zero_frac=float((got == 0).mean()))
    if out["rel"] > 1e-4:
        bad = e > 1e-3 * s
        out["bad_rows"] = np.nonzero(bad.any(1))[0][:40].tolist()
        out["bad_cols"] = np.nonzero(bad.any(0))[0][:40].tolist()
        out["bad_frac"] = float(bad.mean())
        out["sample_got"] = got[:2, :6].tolist(); out["sample_ref"] = ref[:2, :6].tolist()
    print("RESULT " + json.dumps(out))
elif kind == "conv":
    import torch, torch.nn.functional as F
    N, Cin, H, W, Cout, k, s, p = args
    x = rng.standard_normal((N, Cin, H, W)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin, k, k)) / np.sqrt(Cin * k * k)).astype(np.float32)
    got = ctx.conv_check(x, w, None, stride=s, pad=p, relu=False, impl=impl)
    ref = F.conv2d(torch.from_numpy(x).double(), torch.from_numpy(w).double(), None, stride=s, padding=p).numpy()
    e = np.abs(got - ref); sc = np.abs(ref).max()
    out = dict(rel=float(e.max() / sc), nan=int(np.isnan(got).sum()), zero_frac=float((got == 0).mean()))
    if out["rel"] > 1e-4:
        bad = e > 1e-3 * sc
        out["bad_frac"] = float(bad.mean())
        out["bad_h"] = np.nonzero(bad.any((0, 1, 3)))[0][:30].tolist()
        out["bad_w"] = np.nonzero(bad.any((0, 1, 2)))[0][:30].tolist()
        out["bad_c"] = np.nonzero(bad.any((0, 2, 3)))[0][:30].tolist()
    print("RESULT " + json.dumps(out))
''' % ROOT

CASES = [
    ("gemm", 0, [256, 64, 64]), ("gemm", 0, [256, 256, 64]), ("gemm", 0, [256, 128, 256]), ("gemm", 0, [384, 256, 128]),
    ("gemm", 0, [1000, 512, 256]), ("gemm", 0, [1000, 4096, 1024]), ("gemm", 0, [40000, 64, 576]), ("gemm", 0, [300, 84, 4096]),
    ("gemm", 0, [128, 64, 64]), ("gemm", 0, [100, 21, 256]),
    ("conv", 0, [1, 64, 16, 16, 64, 3, 1, 1]), ("conv", 0, [1, 128, 37, 53, 256, 3, 1, 1]), ("conv", 0, [3, 64, 7, 7, 64, 3, 1, 1]),
    ("conv", 0, [2, 64, 14, 14, 128, 3, 2, 1]), ("conv", 0, [1, 128, 28, 36, 256, 1, 2, 0]), ("conv", 0, [1, 512, 38, 50, 512, 3, 1, 1]),
]


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "first_light.log"), "w")
    src = os.path.join(ROOT, "gpurun_out", "_case.py")
    open(src, "w").write(CASE_SRC)
    for kind, impl, args in CASES:
        try:
            r = subprocess.run([sys.executable, src, kind, str(impl), json.dumps(args)], capture_output=True, text=True, timeout=180)
            res = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            tail = (r.stdout[-600:] + r.stderr[-1200:]) if not res else ""
            line = f"{kind} impl={impl} {args}: rc={r.returncode} {res[0] if res else 'NO RESULT'} {tail}"
        except subprocess.TimeoutExpired:
            line = f"{kind} impl={impl} {args}: TIMEOUT"
        print(line); log.write(line + "\n"); log.flush()


if __name__ == "__main__":
    main()
