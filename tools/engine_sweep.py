"""tcgen05 engine microbenchmark sweep (GPU box): per-K-block cost vs BN / CTA group / K, to separate MMA pacing from
operand ingest and from fixed per-launch cost. Prints one line per case."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CASE = r'''
import sys, json
sys.path.insert(0, %r)
import multipathnet_b200 as mpn
ctx = mpn.Context(0)
out = []
for (M, N, K) in json.loads(sys.argv[1]):
    ms, bn, cg, sk = ctx.gemm_bench(M, N, K, 20)
    tiles_m = (M + 127) // 128; tn = (N + bn - 1) // bn
    units = ((tiles_m + cg - 1) // cg) * tn * sk
    slots = 148 // cg
    rounds = (units + slots - 1) // slots
    kb = K // 64
    cyc_per_kb = ms * 1e-3 * 1.965e9 / rounds / (kb / sk)
    tf = 2.0 * M * N * K * 3 / (ms * 1e-3) / 1e12
    out.append(dict(M=M, N=N, K=K, ms=round(ms, 4), BN=bn, CG=cg, splitk=sk, units=units, rounds=rounds, cyc_per_kblock=round(cyc_per_kb), issued_TF=round(tf)))
print("RESULT " + json.dumps(out))
''' % ROOT
def run(env, cases):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, "-c", CASE, json.dumps(cases)], capture_output=True, text=True, env=e, timeout=600)
    for l in r.stdout.splitlines():
        if l.startswith("RESULT "):
            for d in json.loads(l[7:]): print(env, d)
            return
    print(env, "FAILED", r.stdout[-300:], r.stderr[-500:])
if __name__ == "__main__":
    full = 148 * 128
    cases = []
    for N in (64, 128, 256):
        for K in (576, 1152, 2304, 4608, 18432):
            cases.append((full * (4 if N == 64 else 2), N, K))          # several rounds, every SM busy
    cases += [(1000, 4096, 25088), (1000, 4096, 4096), (7500, 512, 4608), (30000, 256, 2304), (1900, 512, 4608)]
    run({}, cases)
    run({"MPN_TC_CTA_GROUP": "1"}, cases)
