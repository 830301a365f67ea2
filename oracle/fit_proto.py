"""numpy PROTOTYPE of the K3 batched minimiser (test infrastructure, lives with the oracle).

Lock-step projected Newton.  Every fit of the batch advances one trial evaluation per round
(one batched logpdf+grad call).  Curvature comes from forward differences of the *analytic*
gradient (a batch of perturbed points through the same logpdf+grad kernel):

  * round 0 uses one shared matrix, the Hessian of 2NLL on the Asimov dataset at the start point;
  * afterwards each fit owns a Hessian taken at its current point, refreshed when the predicted
    decrease stops shrinking fast enough (modified Newton in between).

Bounds: active-set projection (a variable on a bound whose gradient pushes outward is frozen for
the step).  Globalisation: Levenberg-Marquardt -- a rejected trial (Armijo test on the projected
step) raises a per-fit damping added to the Hessian diagonal and the direction is recomputed at
the same point; accepted steps relax it.  (Plain backtracking failed on ~1e-4 of the C3 toys whose
exact Hessian is nearly singular along a degenerate normsys direction.)

The CUDA implementation (pyhf_b200/csrc/hf_fit.cu + the host loop in pyhf_b200/fit.py) is a
transcription of this file; tests compare both against the reference's scipy SLSQP results.
"""
import numpy as np

from . import hf_oracle as O


def fd_hessian(pl, x, data, free, g0=None, h=1e-5):
    """Hessian of 2NLL at x [N,P] (data [N,D]) by forward differences of the analytic gradient."""
    N, P = x.shape
    idx = np.nonzero(free)[0]
    if g0 is None:
        g0 = -2.0 * O.logpdf_grad(pl, x, data)[1]
    Hm = np.zeros((N, P, P))
    for j in idx:
        xp = x.copy()
        step = h * np.maximum(1.0, np.abs(x[:, j]))
        xp[:, j] += step
        gj = -2.0 * O.logpdf_grad(pl, xp, data)[1]
        Hm[:, :, j] = (gj - g0) / step[:, None]
    Hm = 0.5 * (Hm + np.swapaxes(Hm, 1, 2))
    return Hm


def solve_reduced(Hm, g, act, damp=None, max_step=None):
    """Newton direction on the non-active variables: (H_FF + tau I) d = -g_F, d_A = 0.
    tau starts at the fit's Levenberg-Marquardt damping and grows until the Cholesky
    factorisation succeeds (modified Newton)."""
    N, P = g.shape
    d = np.zeros((N, P))
    edm = np.zeros(N)
    for i in range(N):
        f = np.nonzero(~act[i])[0]
        if not len(f):
            continue
        A = Hm[i][np.ix_(f, f)]
        tau = 0.0 if damp is None else float(damp[i])
        scale = max(np.max(np.abs(np.diag(A))), 1e-300)
        for _ in range(60):
            try:
                L = np.linalg.cholesky(A + tau * np.eye(len(f)))
                break
            except np.linalg.LinAlgError:
                tau = max(1e-10 * scale, tau * 10)
        z = np.linalg.solve(L, -g[i, f])
        di = np.linalg.solve(L.T, z)
        if max_step is not None:
            # trust-region flavour: a matrix that is barely positive definite after the shift gives
            # a huge step along its soft direction; raise the shift until the step is bounded
            for _ in range(40):
                if np.max(np.abs(di)) <= max_step:
                    break
                tau = max(4.0 * tau, 1e-6 * scale)
                L = np.linalg.cholesky(A + tau * np.eye(len(f)))
                di = np.linalg.solve(L.T, np.linalg.solve(L, -g[i, f]))
            if damp is not None:
                damp[i] = max(damp[i], 0.0 if tau < 1e-6 * scale else tau)
        d[i, f] = di
        edm[i] = -g[i, f] @ di
    return d, edm


def fit_batch(pl, data, init, lo, hi, fixed, max_iter=100, tol_edm=1e-10, tol_step=1e-7,
              refresh_ratio=0.05, exact_below=10.0, verbose=False, rho_damping=False, damp0=0.0,
              damp0_edm=np.inf, rho_hi=np.inf, rho_hi_fac=0.2, max_step=None):
    data = np.atleast_2d(data)
    init = np.atleast_2d(init)
    N = max(data.shape[0], init.shape[0])
    P = pl.n_pars
    fixed = np.asarray(fixed, bool)
    free = ~fixed
    x = np.clip(np.broadcast_to(init, (N, P)).copy(), lo, hi)
    x[:, fixed] = np.broadcast_to(init, (N, P))[:, fixed]
    dat = np.broadcast_to(data, (N, data.shape[1]))

    def fg(xx, rows):
        v, g = O.logpdf_grad(pl, xx, dat[rows])
        return -2.0 * v, -2.0 * g

    # shared start-up curvature: Asimov data at the (first) start point
    dA = O.expected_data(pl, x[:1])
    H0 = fd_hessian(pl, x[:1], dA, free)[0]
    Hm = np.broadcast_to(H0, (N, P, P)).copy()
    own = np.zeros(N, bool)        # fit owns a Hessian taken at one of its own points
    n_hess = 0
    f, g = fg(x, slice(None))
    status = np.full(N, 1)
    niter = np.zeros(N, int)
    done = np.zeros(N, bool)
    damp = np.zeros(N)
    nrej = np.zeros(N, int)
    fresh = np.ones(N, bool)   # direction needed after an ACCEPTED step (curvature may refresh)
    d = np.zeros((N, P))
    need_dir = np.ones(N, bool)
    edm_prev = np.full(N, np.inf)
    edm_last = np.full(N, np.inf)
    nevals = 1
    for it in range(max_iter):
        rows = np.nonzero(need_dir & ~done)[0]
        if len(rows):
            xr, gr = x[rows], g[rows]
            act = fixed[None, :] | ((xr <= lo) & (gr > 0)) | ((xr >= hi) & (gr < 0))
            dmp = damp[rows].copy()
            dd, edm = solve_reduced(Hm[rows], gr, act, dmp, max_step)
            if max_step is not None:
                damp[rows] = dmp
            if it == 0 and damp0 > 0.0:
                # start-up: a huge predicted decrease means the start point is far outside the
                # quadratic regime; begin with a damped step instead of discovering it by rejections
                far = edm > damp0_edm
                if np.any(far):
                    damp[rows[far]] = damp0
                    dd[far], edm[far] = solve_reduced(Hm[rows[far]], gr[far], act[far], damp[rows[far]])
            # curvature refresh: first own Hessian after the start-up step, later when the
            # predicted decrease did not shrink by 1/refresh_ratio
            slow = edm > refresh_ratio * edm_prev[rows]
            refresh = fresh[rows] & (niter[rows] >= 1) & (~own[rows] | slow) & (edm > tol_edm)
            if np.any(refresh):
                rr = rows[refresh]
                # far from the minimum the exact Hessian is often indefinite (log-sum-exp of the
                # normsys products): use the Fisher matrix = Hessian on the Asimov data of the
                # CURRENT point (always PSD); switch to the exact Hessian once edm is small
                exact = edm[refresh] < exact_below
                dd_rows = dat[rr].copy()
                if np.any(~exact):
                    dd_rows[~exact] = O.expected_data(pl, x[rr[~exact]])
                Hm[rr] = fd_hessian(pl, x[rr], dd_rows, free)
                own[rr] = True
                n_hess += len(rr)
                dmp = damp[rr].copy()
                dd2, edm2 = solve_reduced(Hm[rr], g[rr], act[refresh], dmp, max_step)
                if max_step is not None:
                    damp[rr] = dmp
                # a nearly singular / indefinite exact Hessian shows up as an exploding step:
                # fall back to the Fisher matrix for this iteration
                blow = exact & (~(edm2 <= 100.0 * edm[refresh] + 1.0) | (np.max(np.abs(dd2), axis=1) > 100.0))
                if max_step is not None:
                    blow &= False
                if np.any(blow):
                    rb = rr[blow]
                    Hm[rb] = fd_hessian(pl, x[rb], O.expected_data(pl, x[rb]), free)
                    n_hess += len(rb)
                    dd3, edm3 = solve_reduced(Hm[rb], g[rb], act[refresh][blow], damp[rb])
                    dd2[blow], edm2[blow] = dd3, edm3
                dd[refresh], edm[refresh] = dd2, edm2
            step = np.max(np.abs(dd), axis=1)
            conv = (edm < tol_edm) & (step < tol_step)
            edm_prev[rows] = edm
            edm_last[rows] = edm
            done[rows[conv]] = True
            status[rows[conv]] = 0
            d[rows] = dd
            need_dir[rows] = False
            if verbose:
                print(it, "edm", edm[:4], "f", f[rows][:4], "refresh", refresh[:4])
        live = np.nonzero(~done)[0]
        if not len(live):
            break
        xt = np.clip(x[live] + d[live], lo, hi)
        ft, gt = fg(xt, live)
        nevals += 1
        dec = np.einsum("ij,ij->i", g[live], xt - x[live])
        slack = 4e-15 * np.maximum(1.0, np.abs(f[live]))
        ok = np.isfinite(ft) & (ft <= f[live] + 1e-4 * dec + slack)
        acc = live[ok]
        f_old = f[live].copy()
        x[acc], f[acc], g[acc] = xt[ok], ft[ok], gt[ok]
        need_dir[acc] = True
        fresh[acc] = True
        niter[acc] += 1
        nrej[acc] = 0
        if rho_damping:
            # trust-region style: compare the achieved decrease with the one the quadratic model
            # predicted (>= -dec/2 for a damped Newton step) and move the damping accordingly
            rho = (f_old[ok] - ft[ok]) / np.maximum(-0.5 * dec[ok], 1e-300)
            fac = np.where(rho > rho_hi, rho_hi_fac, np.where(rho > 0.75, 0.2, np.where(rho > 0.25, 1.0, 4.0)))
            nd = np.where((fac > 1.0) & (damp[acc] < 1e-3), 1e-2, fac * damp[acc])
            damp[acc] = np.where(nd < 1e-6, 0.0, nd)
        else:
            damp[acc] = np.where(0.2 * damp[acc] < 1e-6, 0.0, 0.2 * damp[acc])
        rej = live[~ok]
        # rejected: converged if the predicted decrease is below the rounding floor of f,
        # otherwise raise the damping and recompute the direction at the same point
        floor = edm_last[rej] < 1e-7
        done[rej[floor]] = True
        status[rej[floor]] = 0
        rej = rej[~floor]
        nrej[rej] += 1
        dead = rej[nrej[rej] > 60]
        done[dead] = True
        status[dead] = 2
        damp[rej] = np.maximum(8.0 * damp[rej], 1e-2)
        need_dir[rej] = True
        fresh[rej] = False
    return x, f, status, niter, dict(rounds=nevals, hessians=n_hess)
