"""Developer smoke check of the kernels against the CPU oracle (run on a GPU box)."""
import sys, os, time
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
    rng = np.random.default_rng(3)
    ok = True
    # ---- ragged small batch
    L = 6
    insts = []
    for i in range(5):
        rk = [int(rng.integers(1, 6)) for _ in range(L - 1)]
        insts.append(random_tt(rng, [(2,)] * L, rk))
    insts2 = []
    for i in range(5):
        rk = [int(rng.integers(1, 4)) for _ in range(L - 1)]
        insts2.append(random_tt(rng, [(2,)] * L, rk))
    X = BatchedTT.from_host(insts, dev); Y = BatchedTT.from_host(insts2, dev)
    # roundtrip
    back = X.to_host()
    print("roundtrip", max(rel(a, b) for tt, tt2 in zip(back, insts) for a, b in zip(tt, tt2)))
    # inner
    got = eng.inner(X, Y).cpu().numpy()
    ref = np.array([O.inner(a, b) for a, b in zip(insts, insts2)])
    print("inner", rel(got, ref)); ok &= rel(got, ref) < 1e-12
    # axpby
    al = torch.tensor(rng.standard_normal(5), device=dev); be = torch.tensor(rng.standard_normal(5), device=dev)
    Z = eng.axpby(al, X, be, Y).to_host()
    e = 0
    for i in range(5):
        refd = O.dense(O.add(O.scale(insts[i], float(al[i])), O.scale(insts2[i], float(be[i]))), squeeze_result=False)
        e = max(e, rel(O.dense(Z[i], squeeze_result=False), refd))
    print("axpby", e); ok &= e < 1e-13
    # norm2 / orthogonalize
    n2 = eng.norm2(X).cpu().numpy()
    ref = np.array([O.norm_squared(t) for t in insts])
    print("norm2", rel(n2, ref)); ok &= rel(n2, ref) < 1e-12
    Q, n2b = eng.orthogonalize([(None, X)])
    Qh = Q.to_host(); e = 0; eo = 0
    for i in range(5):
        e = max(e, rel(O.dense(Qh[i], squeeze_result=False), O.dense(insts[i], squeeze_result=False)))
        for c in Qh[i][1:]:
            m = c.reshape(c.shape[0], -1)
            eo = max(eo, np.abs(m @ m.T - np.eye(m.shape[0])).max())
    print("orth dense", e, "orthonormality", eo, "ranks", Q.ranks_host()[0], [c.shape for c in O.left_orthogonalize(O.clone(insts[0]))]); ok &= e < 1e-12 and eo < 1e-12
    # round: cap, eps
    for kw in (dict(caps=2, rel_eps=1e-16), dict(caps=None, rel_eps=1e-10), dict(caps=3, rel_eps=1e-12)):
        S = eng.axpby(None, X, None, Y)
        R = eng.round_sum([(None, S)], **kw)
        Rh = R.to_host(); e = 0; rk_ok = True
        for i in range(5):
            s = O.add(insts[i], insts2[i])
            ref = O.reapprox(O.clone(s), ranks_new=kw["caps"], rel_error=kw["rel_eps"])
            e = max(e, rel(O.dense(Rh[i], squeeze_result=False), O.dense(ref, squeeze_result=False)))
            rk_ok &= [c.shape for c in Rh[i]] == [c.shape for c in ref]
        print("round", kw, e, "ranks equal", rk_ok); ok &= e < 1e-10 and rk_ok
        # same through the unmaterialised sum
        R2 = eng.round_sum([(None, X), (None, Y)], **kw).to_host()
        e2 = max(rel(O.dense(R2[i], squeeze_result=False), O.dense(Rh[i], squeeze_result=False)) for i in range(5))
        print("  round_sum (two terms)", e2); ok &= e2 < 1e-10
    # operator term
    L = 5
    op = random_tt(rng, [(3, 2)] * L, [3, 4, 4, 2])
    xs = [random_tt(rng, [(2,)] * L, [2, 3, 3, 2]) for _ in range(4)]
    bs = [random_tt(rng, [(3,)] * L, [2, 2, 2, 2]) for _ in range(4)]
    A = DeviceMPO(op, dev); XB = BatchedTT.from_host(xs, dev); BB = BatchedTT.from_host(bs, dev)
    Ym = eng.matmul(A, XB, [2] * L, [(3,)] * L).to_host()
    e = max(rel(O.dense(Ym[i], squeeze_result=False), O.dense(O.matmul(op, xs[i]), squeeze_result=False)) for i in range(4))
    print("matmul", e); ok &= e < 1e-13
    for kw in (dict(caps=4, rel_eps=1e-16), dict(caps=None, rel_eps=1e-10)):
        R = eng.round_sum([(A, XB)], **kw).to_host(); e = 0; rk_ok = True
        for i in range(4):
            ref = O.reapprox(O.matmul(op, xs[i]), ranks_new=kw["caps"], rel_error=kw["rel_eps"])
            e = max(e, rel(O.dense(R[i], squeeze_result=False), O.dense(ref, squeeze_result=False)))
            rk_ok &= [c.shape for c in R[i]] == [c.shape for c in ref]
        print("apply_round", kw, e, rk_ok); ok &= e < 1e-10 and rk_ok
        coef = torch.tensor(rng.standard_normal(4), device=dev)
        R = eng.round_sum([(None, BB), (A, XB, coef, -1.0)], **kw).to_host(); e = 0; rk_ok = True
        for i in range(4):
            ref = O.reapprox(O.sub(bs[i], O.scale(O.matmul(op, xs[i]), float(coef[i]))), ranks_new=kw["caps"], rel_error=kw["rel_eps"])
            e = max(e, rel(O.dense(R[i], squeeze_result=False), O.dense(ref, squeeze_result=False)))
            rk_ok &= [c.shape for c in R[i]] == [c.shape for c in ref]
        print("b - c*A@x round", kw, e, rk_ok); ok &= e < 1e-10 and rk_ok
    # larger ranks: exercise outer panels (> 32 columns) and multi-sub-panel leaves
    L = 8
    xs = [random_tt(rng, [(2,)] * L, [2, 4, 8, 16, 8, 4, 2]) for _ in range(3)]
    op = random_tt(rng, [(2, 2)] * L, [5] * (L - 1))
    A = DeviceMPO(op, dev); XB = BatchedTT.from_host(xs, dev)
    R = eng.round_sum([(A, XB)], caps=12, rel_eps=1e-14).to_host(); e = 0; rk_ok = True
    for i in range(3):
        ref = O.reapprox(O.matmul(op, xs[i]), ranks_new=12, rel_error=1e-14)
        e = max(e, rel(O.dense(R[i], squeeze_result=False), O.dense(ref, squeeze_result=False)))
        rk_ok &= [c.shape for c in R[i]] == [c.shape for c in ref]
    print("apply_round big", e, rk_ok, [c.shape for c in R[0]]); ok &= e < 1e-10 and rk_ok
    torch.cuda.synchronize()
    print("flops, launches", eng.counters())
    print("ALL OK" if ok else "FAILURES")
    return 0 if ok else 1

if __name__ == "__main__":
    sys.exit(main())
