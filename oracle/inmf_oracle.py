"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy / scipy, float64) of pyliger's iNMF hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.  Nothing under ``pyliger_b200/`` does: the product path is CUDA only.

What is restated (reference file:line, relative to /root/reference/src/pyliger/factorization/):

* ``bpp`` / ``nnlsm_blockpivot``      -- _utilities.py:181-301 (block principal pivoting NNLS) and
                                         _utilities.py:304-358 (grouped normal-equation solves)
* ``als``                             -- _iNMF_ANLS.py:79-185  (optimize_ALS)
* ``update_A_B``                      -- _online_iNMF.py:563-601
* ``hals_W`` / ``hals_V``             -- _utilities.py:102-129
* ``minibatch_rows``                  -- _online_iNMF.py:511-552 + pyliger/_utilities.py:100-116
* ``online_cal_W_V`` / ``online_cal_H`` / ``online`` / ``projection`` / ``online_refine``
                                      -- _online_iNMF.py:125-508
* ``init_W`` / ``init_V_online``      -- _utilities.py:16-26, 69-88
* ``hals_H`` / ``inmf_hals``          -- _utilities.py:132-148, _iNMF_HALS.py:60-163 (batch HALS factorisation)

The restatement works on *products* (G = AᵀA, R = AᵀB built from scipy CSR) instead of the
reference's dense stacked matrices, so it also runs at shapes where the reference's ``.toarray()``
does not fit.  Parity pin: ``tests/test_oracle_golden.py`` checks every function here against
vectors produced by executing the unmodified reference (``tests/golden/make_golden.py``) and against
the known-answer vectors of SURVEY.md Appendix C.  Parity is therefore PINNED by executed-reference
goldens (the reference ships no tests of its own).
"""
import numpy as np
import scipy.sparse as sps

EPS_NONNEG = 1e-16  # _utilities.py:10-13 clamps to 1e-16, not 0
ZERO_TOL = 1e-16  # _utilities.py:290,292


# ----------------------------------------------------------------------------------------------
# Block principal pivoting NNLS
# ----------------------------------------------------------------------------------------------
def _pack_masks(P):
    """bool (n, cols), n <= 64 -> uint64 key per column."""
    n = P.shape[0]
    w = (np.uint64(1) << np.arange(n, dtype=np.uint64))[:, None]
    return (P.astype(np.uint64) * w).sum(axis=0, dtype=np.uint64)


def _solve_grouped(G, R, P):
    """Solve G[P,P] z = R[P] for every column, one factorisation per distinct passive set.

    Mirrors the counting quirks of _utilities.py:318-358: the all-passive shortcut counts one
    factorisation and ``cols`` solves; the single-column path counts (1, 1); the grouped path
    counts each group once and its size twice.
    """
    n, cols = R.shape
    Z = np.zeros((n, cols))
    n_fact = 0
    n_eq = 0
    if cols == 0:
        return Z, 0, 0
    if P.all():
        return np.linalg.solve(G, R), 1, cols
    if cols == 1:
        idx = np.flatnonzero(P[:, 0])
        if idx.size:
            Z[idx, 0] = np.linalg.solve(G[np.ix_(idx, idx)], R[idx, 0])
            return Z, 1, 1
        return Z, 0, 0
    if n <= 64:
        keys = _pack_masks(P)
        order = np.argsort(keys, kind="stable")
        sk = keys[order]
        starts = np.flatnonzero(np.r_[True, sk[1:] != sk[:-1]])
        ends = np.r_[starts[1:], sk.size]
        groups = [order[s:e] for s, e in zip(starts, ends)]
    else:  # generic fallback, not used on the iNMF path (n is the factor count k)
        _, inv = np.unique(P.T, axis=0, return_inverse=True)
        groups = [np.flatnonzero(inv == u) for u in range(inv.max() + 1)]
    for grp in groups:
        idx = np.flatnonzero(P[:, grp[0]])
        if idx.size == 0:
            continue
        Z[np.ix_(idx, grp)] = np.linalg.solve(G[np.ix_(idx, idx)], R[np.ix_(idx, grp)])
        n_fact += 1
        n_eq += 2 * len(grp)
    return Z, n_fact, n_eq


def bpp(G, R, init=None, max_rounds=None):
    """Block principal pivoting on the normal equations: min ||Ax-b|| s.t. x>=0 per column.

    G = AᵀA (n x n), R = AᵀB (n x cols).  Returns X, (success, Y, n_fact, n_eq, n_backup, rounds).
    Follows _utilities.py:219-301 rule for rule: exchange of all infeasible indices while the
    infeasible count improves (and for up to ``p_bar`` = 3 non-improving rounds), then the backup
    rule (flip the largest infeasible index).
    """
    G = np.asarray(G, dtype=np.float64)
    R = np.asarray(R, dtype=np.float64)
    n, cols = R.shape
    limit = 5 * n if max_rounds is None else max_rounds
    if init is not None:
        P = np.asarray(init) > 0
        X, n_fact, n_eq = _solve_grouped(G, R, P)
        Y = G @ X - R
    else:
        X = np.zeros((n, cols))
        Y = -R.copy()
        P = np.zeros((n, cols), dtype=bool)
        n_fact = n_eq = 0
    p_bar = 3
    p_vec = np.full(cols, p_bar, dtype=np.int64)
    ninf = np.full(cols, n + 1, dtype=np.int64)
    bad_y = (Y < 0) & ~P
    bad_x = (X < 0) & P
    bad = bad_y.sum(0) + bad_x.sum(0)
    active = bad > 0
    rounds = 0
    n_backup = 0
    success = True
    while active.any():
        rounds += 1
        if limit > 0 and rounds > limit:
            success = False
            break
        improved = active & (bad < ninf)
        stalled = active & ~improved
        retry = stalled & (p_vec >= 1)
        backup = stalled & (p_vec < 1)
        if improved.any():
            p_vec[improved] = p_bar
            ninf[improved] = bad[improved]
        if retry.any():
            p_vec[retry] -= 1
        full = improved | retry
        if full.any():
            P[bad_y & full[None, :]] = True
            P[bad_x & full[None, :]] = False
        for c in np.flatnonzero(backup):
            j = np.flatnonzero(bad_y[:, c] | bad_x[:, c]).max()
            P[j, c] = ~P[j, c]
            n_backup += 1
        cols_now = np.flatnonzero(active)
        Z, f, e = _solve_grouped(G, R[:, cols_now], P[:, cols_now])
        X[:, cols_now] = Z
        n_fact += f
        n_eq += e
        X[np.abs(X) < ZERO_TOL] = 0
        Y[:, cols_now] = G @ X[:, cols_now] - R[:, cols_now]
        Y[np.abs(Y) < ZERO_TOL] = 0
        bad_y = (Y < 0) & ~P & active[None, :]
        bad_x = (X < 0) & P & active[None, :]
        bad = bad_y.sum(0) + bad_x.sum(0)
        active = bad > 0
    return X, (success, Y, n_fact, n_eq, n_backup, rounds)


def nnlsm_blockpivot(A, B, is_input_prod=False, init=None):
    """Same signature and return tuple as _utilities.py:181-207."""
    if is_input_prod:
        G, R = A, B
    else:
        G = A.T.dot(A)
        R = np.asarray(B.T.dot(A)).T if sps.issparse(B) else A.T.dot(B)
    X, (ok, Y, n_fact, n_eq, n_backup, _) = bpp(G, R, init=init)
    return X, (ok, Y, n_fact, n_eq, n_backup)


def kkt_violation(G, R, X):
    """Max violation of the NNLS optimality conditions, scaled by |R|_max (oracle-free check)."""
    Y = G @ X - R
    s = max(1.0, float(np.abs(R).max()))
    return max(float(np.maximum(-X, 0).max()), float(np.maximum(-Y, 0).max()) / s,
               float(np.abs(X * Y).max()) / s)


# ----------------------------------------------------------------------------------------------
# optimize_ALS
# ----------------------------------------------------------------------------------------------
def als_init(ns, g, k, rand_seed, rep=0):
    """Seeded init in the reference's draw order (_iNMF_ANLS.py:103-108): W, V_0.., H_0.. ."""
    np.random.seed(seed=rand_seed + rep - 1)
    W = np.abs(np.random.uniform(0, 2, (k, g)))
    V = [np.abs(np.random.uniform(0, 2, (k, g))) for _ in ns]
    H = [np.abs(np.random.uniform(0, 2, (n, k))) for n in ns]
    return W, V, H


def objective_dense(Xs, H, W, V, lam):
    """_iNMF_ANLS.py:157-162 verbatim in meaning: residual norms on densified X."""
    a = 0.0
    p = 0.0
    for Xi, Hi, Vi in zip(Xs, H, V):
        D = Xi.toarray() if sps.issparse(Xi) else Xi
        a += np.linalg.norm(D - Hi @ (W + Vi)) ** 2
        p += np.linalg.norm(Hi @ Vi) ** 2
    return a + lam * p


def objective_stats(xnorm2, S, HtH, W, V, lam):
    """Same objective from sufficient statistics: no n x g temporaries."""
    obj = 0.0
    for x2, Si, Ci, Vi in zip(xnorm2, S, HtH, V):
        M = W + Vi
        obj += x2 - 2.0 * np.sum(Si * M) + np.sum(Ci * (M @ M.T)) + lam * np.sum(Ci * (Vi @ Vi.T))
    return obj


def als(Xs, k, value_lambda=5.0, thresh=1e-6, max_iters=30, nrep=1, H_init=None, W_init=None,
        V_init=None, rand_seed=1, dense_objective=None, callback=None):
    """optimize_ALS on a list of scipy CSR matrices (cells x genes).

    Returns dict(H=[n_i x k], W=k x g, V=[k x g], objectives=[obj0, obj1, ...], iters_run=int).
    Update order as _iNMF_ANLS.py:135-152: every H_i, then every V_i with the OLD W, then W with
    the NEW V_i.  The loop does not stop early, it just stops doing work (``:166-167``).
    """
    Xs = [sps.csr_matrix(X, dtype=np.float64) for X in Xs]
    ns = [X.shape[0] for X in Xs]
    g = Xs[0].shape[1]
    N = len(Xs)
    if k >= min(ns):
        raise ValueError(
            "Select k lower than the number of cells in smallest dataset: {}".format(min(ns)))
    if k >= g:
        raise ValueError("Select k lower than the number of variable genes: {}".format(g))
    if dense_objective is None:
        dense_objective = sum(n * g for n in ns) <= 3e7
    xnorm2 = [float(X.multiply(X).sum()) for X in Xs]
    lam = float(value_lambda)
    best = None
    for rep in range(nrep):
        W, V, H = als_init(ns, g, k, rand_seed, rep)
        if W_init is not None:
            W = np.array(W_init, dtype=np.float64)
        if V_init is not None:
            V = [np.array(v, dtype=np.float64) for v in V_init]
        if H_init is not None:
            H = [np.array(h, dtype=np.float64) for h in H_init]

        def objective():
            if dense_objective:
                return objective_dense(Xs, H, W, V, lam)
            S = [np.asarray(Xi.T.dot(Hi)).T for Xi, Hi in zip(Xs, H)]
            C = [Hi.T @ Hi for Hi in H]
            return objective_stats(xnorm2, S, C, W, V, lam)

        obj = objective()
        objs = [obj]
        delta = 1.0
        iters_run = 0
        for it in range(max_iters):
            if not delta > thresh:
                continue
            for i in range(N):
                M = W + V[i]
                G = M @ M.T + lam * (V[i] @ V[i].T)
                R = np.asarray(Xs[i].dot(M.T)).T  # k x n_i
                H[i] = np.ascontiguousarray(bpp(G, R)[0].T)
            S = [np.asarray(Xs[i].T.dot(H[i])).T for i in range(N)]  # k x g
            C = [H[i].T @ H[i] for i in range(N)]
            for i in range(N):
                V[i] = bpp((1.0 + lam) * C[i], S[i] - C[i] @ W)[0]
            Gw = sum(C)
            Rw = sum(S[i] - C[i] @ V[i] for i in range(N))
            W = bpp(Gw, Rw)[0]
            prev = obj
            obj = objective_dense(Xs, H, W, V, lam) if dense_objective else objective_stats(
                xnorm2, S, C, W, V, lam)
            objs.append(obj)
            delta = abs(prev - obj) / ((prev + obj) / 2.0)
            iters_run += 1
            if callback is not None:
                callback(it, obj, H, W, V)
        if best is None or obj < best["objective"]:
            best = dict(H=H, W=W, V=V, objective=obj, objectives=objs, iters_run=iters_run)
    return best


# ----------------------------------------------------------------------------------------------
# online_iNMF
# ----------------------------------------------------------------------------------------------
def chunk_ranges(chunk_size, n):
    """pyliger/_utilities.py:100-116: [0,c), [c,2c), ..., last one clipped at n."""
    out = []
    lo = 0
    while lo < n:
        out.append((lo, min(lo + chunk_size, n)))
        lo += chunk_size
    return out


def minibatch_rows(num_iter, mb, max_epochs, chunk_size, n):
    """_online_iNMF.py:511-552: per step, the list of (left, right) row ranges.

    Chunks are taken in order (the permutation is commented out in the reference, :547) for
    ``max_epochs + 1`` epochs and cut greedily into mini-batches of ``mb`` rows.
    """
    queue = []
    for _ in range(max_epochs + 1):
        queue.extend(chunk_ranges(chunk_size, n))
    queue = [list(q) for q in queue]
    pos = 0
    steps = []
    for _ in range(num_iter):
        need = int(mb)
        cur = []
        while need > 0:
            lo, hi = queue[pos]
            take = min(hi - lo, need)
            cur.append((lo, lo + take))
            need -= take
            if lo + take == hi:
                pos += 1
            else:
                queue[pos][0] = lo + take
        steps.append(cur)
    return steps


def init_W(g, k, rand_seed):
    """_utilities.py:16-26."""
    np.random.seed(seed=rand_seed)
    W = np.abs(np.random.uniform(0, 2, (g, k)))
    return W / np.sqrt(np.sum(np.square(W), axis=0))


def init_V_online(n, k, X, rand_seed):
    """_utilities.py:69-88: k cells sampled with replacement (sorted), as unit-norm columns."""
    np.random.seed(seed=rand_seed)
    idx = np.sort(np.random.choice(list(range(n)), k))
    V = X[idx, :].T.toarray()
    return V / np.sqrt(np.sum(np.square(V), axis=0))


def update_A_B(A, B, A_old, B_old, H_mb, X_mb, mb, it, epoch, epoch_prev):
    """_online_iNMF.py:563-601.  H_mb is k x mb, X_mb is g x mb (dense or sparse)."""
    if it == 0:
        s = 0
    elif it == 1:
        s = 1 / mb
    else:
        s = (it - 1) / it
    if epoch > 0 and epoch - epoch_prev == 1:
        A = A - A_old
        A_old = s * A
        B = B - B_old
        B_old = s * B
    else:
        A_old = s * A_old
        B_old = s * B_old
    A = s * A + (H_mb @ H_mb.T) / mb
    XH = np.asarray(X_mb.dot(H_mb.T)) if sps.issparse(X_mb) else X_mb @ H_mb.T
    B = s * B + XH / mb
    d = np.diagonal(A).copy()
    d[d == 0] = 1e-15
    np.fill_diagonal(A, d)
    return A, B, A_old, B_old


def hals_W(A, B, W, V):
    """_utilities.py:102-115; in place on W (g x k), sequential over factors j."""
    for j in range(W.shape[1]):
        num = np.zeros(W.shape[0])
        den = 0.0
        for Ai, Bi, Vi in zip(A, B, V):
            num = num + Bi[:, j] - (W + Vi) @ Ai[:, j]
            den += Ai[j, j]
        col = W[:, j] + num / den
        col[col < EPS_NONNEG] = EPS_NONNEG
        W[:, j] = col
    return W


def hals_V(A, B, W, V, lam):
    """_utilities.py:118-129; in place on every V_i (g x k), j outer / dataset inner."""
    for j in range(W.shape[1]):
        for Ai, Bi, Vi in zip(A, B, V):
            col = Vi[:, j] + (Bi[:, j] - (W + (1 + lam) * Vi) @ Ai[:, j]) / ((1 + lam) * Ai[j, j])
            col[col < EPS_NONNEG] = EPS_NONNEG
            Vi[:, j] = col
    return V


def solve_H_block(X_rows, W, V, lam):
    """H (k x m) for a block of cells; W, V are g x k (online layout). _online_iNMF.py:419-426."""
    M = W + V
    G = M.T @ M + lam * (V.T @ V)
    R = np.asarray(X_rows.dot(M)).T
    return bpp(G, R)[0]


def online_cal_W_V(Xs, k, value_lambda=5.0, max_epochs=5, miniBatch_max_iters=1,
                   miniBatch_size=5000, h5_chunk_size=1000, rand_seed=1, W_init=None,
                   V_init=None, A_init=None, B_init=None, prev_info=None, trace=None):
    """_online_iNMF.py:334-462."""
    Xs = [sps.csr_matrix(X, dtype=np.float64) for X in Xs]
    N = len(Xs)
    ns = [X.shape[0] for X in Xs]
    g = Xs[0].shape[1]
    lam = float(value_lambda)
    np.random.seed(seed=rand_seed)
    W = init_W(g, k, rand_seed) if W_init is None else W_init
    Vs = [init_V_online(ns[i], k, Xs[i], rand_seed) for i in range(N)] if V_init is None else V_init
    A = [np.zeros((k, k)) for _ in range(N)] if A_init is None else A_init
    B = [np.zeros((g, k)) for _ in range(N)] if B_init is None else B_init
    A_old = [np.zeros((k, k)) for _ in range(N)]
    B_old = [np.zeros((g, k)) for _ in range(N)]
    mbs = np.asarray([np.round((ns[i] / np.sum(ns)) * miniBatch_size).astype(int) for i in range(N)])
    total_iters = int(np.floor(ns[0] * max_epochs / mbs[0]))
    plan = [minibatch_rows(total_iters, mbs[i], max_epochs, h5_chunk_size, ns[i]) for i in range(N)]
    epoch = 0
    for it in range(total_iters):
        epoch_prev = epoch
        epoch = ((it + 1) * mbs[0]) // ns[0]
        X_mb = [sps.vstack([Xs[i][lo:hi] for lo, hi in plan[i][it]]).tocsr() for i in range(N)]
        H_mb = [solve_H_block(X_mb[i], W, Vs[i], lam) for i in range(N)]
        for i in range(N):
            A[i], B[i], A_old[i], B_old[i] = update_A_B(
                A[i], B[i], A_old[i], B_old[i], H_mb[i], X_mb[i].T.tocsr(), mbs[i], it, epoch,
                epoch_prev)
        for _ in range(miniBatch_max_iters):
            if prev_info is None:
                W = hals_W(A, B, W, Vs)
            else:
                W = hals_W(prev_info["A_prev"] + A, prev_info["B_prev"] + B, W,
                           prev_info["V_prev"] + Vs)
            Vs = hals_V(A, B, W, Vs, lam)
        if trace is not None:
            trace(it, W, Vs, A, B)
    return W, Vs, A, B


def online_cal_H(Xs, W, Vs, value_lambda, miniBatch_size):
    """_online_iNMF.py:465-508 -> list of k x n_i."""
    Hs = []
    for X, V in zip(Xs, Vs):
        X = sps.csr_matrix(X, dtype=np.float64)
        n = X.shape[0]
        nb = int(np.ceil(n / miniBatch_size))
        Hs.append(np.hstack([
            solve_H_block(X[b * miniBatch_size:(b + 1) * miniBatch_size], W, V, float(value_lambda))
            for b in range(nb)]))
    return Hs


def online(Xs, k=20, value_lambda=5.0, max_epochs=5, miniBatch_max_iters=1, miniBatch_size=5000,
           h5_chunk_size=1000, rand_seed=1, W_init=None, V_init=None, A_init=None, B_init=None):
    """Scenario 1 (_online_iNMF.py:125-178). Returns dict(H=[n_i x k], W, V, A, B) in the
    written-back layouts (W, V_i, B_i are g x k; A_i is k x k)."""
    W, Vs, A, B = online_cal_W_V(Xs, k, value_lambda, max_epochs, miniBatch_max_iters,
                                 miniBatch_size, h5_chunk_size, rand_seed, W_init, V_init, A_init,
                                 B_init)
    Hs = online_cal_H(Xs, W, Vs, value_lambda, miniBatch_size)
    return dict(H=[h.T for h in Hs], W=W, V=Vs, A=A, B=B)


def projection(X_new, W, k, miniBatch_size=5000):
    """Scenario 3 (_online_iNMF.py:271-331): H = BPP(A=W, B=X_mb); V = 0."""
    G = W.T @ W
    out_H = []
    out_V = []
    for X in X_new:
        X = sps.csr_matrix(X, dtype=np.float64)
        n, g = X.shape
        nb = int(np.ceil(n / miniBatch_size))
        Hs = [bpp(G, np.asarray(X[b * miniBatch_size:(b + 1) * miniBatch_size].dot(W)).T)[0]
              for b in range(nb)]
        out_H.append(np.hstack(Hs).T)
        out_V.append(np.zeros((g, k)))
    return dict(H=out_H, W=W, V=out_V)


def online_refine(X_old, prev, X_new, k=20, value_lambda=5.0, max_epochs=5, miniBatch_max_iters=1,
                  miniBatch_size=5000, h5_chunk_size=1000, rand_seed=1, W_init=None, V_init=None,
                  A_init=None, B_init=None):
    """Scenario 2 (_online_iNMF.py:181-268).  ``prev`` = dict(W, V=[...], A=[...], B=[...]) of the
    already factorised datasets ``X_old``.  Returns dict(new=[per new dataset dict(H,W,V,A,B)],
    H_old=[n_i x k], W)."""
    W0 = prev["W"] if W_init is None else W_init
    prev_info = dict(A_prev=list(prev["A"]), B_prev=list(prev["B"]), V_prev=list(prev["V"]))
    news = []
    W = W0
    for X in X_new:
        W, Vs, A, B = online_cal_W_V([X], k, value_lambda, max_epochs, miniBatch_max_iters,
                                     miniBatch_size, h5_chunk_size, rand_seed, W, V_init, A_init,
                                     B_init, prev_info=prev_info)
        Hs = online_cal_H([X], W, Vs, value_lambda, miniBatch_size)
        news.append(dict(H=Hs[0].T, W=W, V=Vs[0], A=A[0], B=B[0]))
    H_old = [online_cal_H([X], W, [V], value_lambda, miniBatch_size)[0].T
             for X, V in zip(X_old, prev["V"])]
    return dict(new=news, H_old=H_old, W=W)


# ----------------------------------------------------------------------------------------------
# iNMF_HALS (batch HALS factorisation)
# ----------------------------------------------------------------------------------------------
def hals_H(H, V, W, Xs, lam):
    """_utilities.py:132-148 in product form; in place on every H_i (k x n_i).  W, V_i are g x k, Xs[i] cells x genes:
    H[j,:] = nonneg(H[j,:] + ((W+V)[:,j]' X' - G1[j,:] H - lam G2[j,:] H) / (G1[j,j] + lam G2[j,j]))."""
    for Hi, Vi, X in zip(H, V, Xs):
        M = W + Vi
        G1 = M.T @ M
        G2 = Vi.T @ Vi
        R = np.asarray(X.dot(M)).T  # k x n
        for j in range(W.shape[1]):
            row = Hi[j, :] + (R[j, :] - G1[j, :] @ Hi - lam * (G2[j, :] @ Hi)) / (G1[j, j] + lam * G2[j, j])
            row[row < EPS_NONNEG] = EPS_NONNEG
            Hi[j, :] = row
    return H


def inmf_hals(Xs, k=20, value_lambda=5.0, thresh=1e-4, max_iters=25, nrep=1, rand_seed=1):
    """_iNMF_HALS.py:60-161 (without the failing attribute assignment at :156).  Xs: cells x genes CSR."""
    N = len(Xs)
    ns = [X.shape[0] for X in Xs]
    g = Xs[0].shape[1]
    np.random.seed(seed=rand_seed)
    final = None
    all_objs = []
    for _ in range(nrep):
        V = [np.asarray(Xs[i][np.random.choice(list(range(ns[i])), k), :].todense()).T for i in range(N)]
        V = [v / np.sqrt(np.sum(np.square(v), axis=0)) for v in V]
        W = init_W(g, k, rand_seed)
        H = [np.random.uniform(0, 2, (k, ns[i])) for i in range(N)]
        xn = [float(X.multiply(X).sum()) for X in Xs]

        def objective():
            tot = 0.0
            for i in range(N):
                M = W + V[i]
                S = np.asarray(Xs[i].T.dot(H[i].T))  # g x k
                HHt = H[i] @ H[i].T
                tot += xn[i] - 2.0 * np.sum(S * M) + np.sum(HHt * (M.T @ M)) + value_lambda * np.sum(HHt * (V[i].T @ V[i]))
            return tot

        obj = objective()
        objs = [obj]
        it = 1
        delta = np.inf
        while delta > thresh and it <= max_iters:
            H = hals_H(H, V, W, Xs, value_lambda)
            XHt = [np.asarray(Xs[i].T.dot(H[i].T)) for i in range(N)]
            HHt = [H[i] @ H[i].T for i in range(N)]
            W = hals_W(HHt, XHt, W, V)
            V = hals_V(HHt, XHt, W, V, value_lambda)
            prev = obj
            obj = objective()
            objs.append(obj)
            delta = np.absolute(prev - obj) / ((prev + obj) / 2)
            if obj <= prev:
                final = ([h.copy() for h in H], W.copy(), [v.copy() for v in V])
            it += 1
        all_objs.append(objs)
    return {"H": [h.T for h in final[0]], "W": final[1], "V": final[2], "objectives": all_objs}
