"""Developer check of the tall-skinny QR path (tsqr.cuh) against the CPU oracle.  Run with QTT_TS_MIN_ROWS=330 so that
moderate sizes reach it (default threshold 1024 rows)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import tt_oracle as O
from oracle.make_golden import random_tt
from laplacesolvermps_b200.batched import BatchedTT, DeviceMPO, Engine


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def main():
    dev = "cuda:0"
    eng = Engine(dev)
    rng = np.random.default_rng(5)
    ok = True
    for (L, rk, N) in ((7, 100, 3), (7, 150, 2), (8, 70, 2)):
        insts = [random_tt(rng, [(4,)] * L, [rk] * (L - 1), scale=0.3) for _ in range(N)]
        X = BatchedTT.from_host(insts, dev)
        n2 = eng.norm2(X).cpu().numpy()
        ref = np.array([O.norm_squared(t) for t in insts])
        print(f"L={L} r={rk}: norm2 rel {rel(n2, ref):.2e}")
        ok &= rel(n2, ref) < 1e-11
        Q, _ = eng.orthogonalize([(None, X)])
        Qh = Q.to_host(); e = 0; eo = 0
        for i in range(N):
            e = max(e, rel(O.dense(Qh[i], squeeze_result=False), O.dense(insts[i], squeeze_result=False)))
            for c in Qh[i][1:]:
                m = c.reshape(c.shape[0], -1)
                eo = max(eo, np.abs(m @ m.T - np.eye(m.shape[0])).max())
        print(f"   orthogonalize: dense rel {e:.2e}, orthonormality {eo:.2e}, ranks {Q.ranks_host()[0].tolist()}")
        ok &= e < 1e-11 and eo < 1e-11
        for kw in (dict(caps=5, rel_eps=1e-16), dict(caps=None, rel_eps=1e-8)):
            R = eng.round_sum([(None, X)], **kw)
            Rh = R.to_host(); e = 0; same = True
            for i in range(N):
                refr = O.reapprox(O.clone(insts[i]), ranks_new=kw["caps"], rel_error=kw["rel_eps"])
                same &= [c.shape for c in Rh[i]] == [c.shape for c in refr]
                e = max(e, rel(O.dense(Rh[i], squeeze_result=False), O.dense(refr, squeeze_result=False)))
            print(f"   round {kw}: dense rel {e:.2e}, same ranks {same}")
            ok &= e < 1e-9 and same
    # operator term (product never materialised)
    L = 7
    op = random_tt(rng, [(4, 4)] * L, [30] * (L - 1), scale=0.2)
    xs = [random_tt(rng, [(4,)] * L, [4] * (L - 1), scale=0.5) for _ in range(3)]
    A, X = DeviceMPO(op, dev), BatchedTT.from_host(xs, dev)
    Y = eng.round_sum([(A, X)], caps=4, rel_eps=1e-14).to_host()
    e = 0
    for i in range(3):
        refr = O.reapprox(O.matmul(op, xs[i]), ranks_new=4, rel_error=1e-14)
        e = max(e, rel(O.dense(Y[i], squeeze_result=False), O.dense(refr, squeeze_result=False)))
    print(f"round(A@x) rank 120 product: dense rel {e:.2e}")
    ok &= e < 1e-9
    print("launches", eng.counters())
    print("ALL OK" if ok else "FAILED")


if __name__ == "__main__":
    main()
