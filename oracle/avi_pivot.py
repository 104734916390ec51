"""ORACLE (test infrastructure, never shipped): readable Python statement of the
bounded-variable complementary-pivoting AVI solve.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this package.  The product path (quadraticprogramnetworks.jl_b200 + libqpn_cuda)
must never route through it.

What it restates
----------------
Reference call site: `solve_avi` (/root/reference/src/avi.jl:63-77) hands
(M, q=N*w+o, l, u, z0) to PATHSolver.solve_mcp -- closed-source PATH 5.x (compat
"1.7" in Project.toml:28, no Manifest, so unpinned and absent here).  For an
*affine* problem PATH's major iteration is one pivotal solve of the normal-map
path (Cao & Ferris, "A pivotal method for affine variational inequalities",
Math. Oper. Res. 21, 1996; Dirkse & Ferris, "The PATH solver", 1995).  The only
in-repo statement of that pivotal method is the unfinished
/root/reference/src/deprecated/avi_scratch.jl:2-134, which this file follows
with the defects listed in SURVEY.md row A9 repaired:

  * tableau rows  M z - w + r t = r - q   (avi_scratch.jl:23, with w = u - v)
  * r = M*proj(z0) + q + z0 - proj(z0)    (avi_scratch.jl:20-21, the normal map)
  * initial basis by bound activity of z0  (avi_scratch.jl:30-50)
  * ratio test over finite bounds of basics (avi_scratch.jl:65-77)
  * rank-1 pivot                           (avi_scratch.jl:2-7)
  * complementary entering rule            (avi_scratch.jl:105-131)
  * success when t reaches 1               (repairs avi_scratch.jl:78,84-90,105-106)

Additions the scratch file lacks (needed because every GAVI the package builds
has structurally singular start bases): a crash phase that
brings interior/free variables into the basis by degenerate exchanges and, for
dependent columns, walks to an extreme point (Cao-Ferris stage 2).

PARITY UNPINNED against PATH itself (cannot run here); pinned on KAT-0..8 of
SURVEY.md 8c, on check_avi_solution residuals and on uniqueness for strongly
monotone instances.  The C oracle (oracle/qpn_oracle.c) and the CUDA kernel
implement exactly this procedure, operation for operation.
"""
import math
import numpy as np

SUCCESS, RAY_TERM, MAX_ITERS, FAILURE = 1, 2, 3, 4

PIV_TOL = 1e-9     # smallest |pivot| accepted in a crash exchange
D_TOL = 1e-10      # |direction entry| treated as zero in ratio tests
TIE_TOL = 1e-10    # ratios within this (relative) of the minimum are ties
INF = math.inf

AT_L, AT_U, FLOAT, BASIC = 0, 1, 2, 3


from ._fma import fma  # noqa: E402


class _Tab:
    """Compact tableau  x_B = beta - T * (x_N - current x_N)."""

    def __init__(self, M, q, l, u, z0):
        n = self.n = len(q)
        self.l, self.u = l, u
        zb = np.minimum(np.maximum(z0, l), u)
        r = np.empty(n)
        for i in range(n):
            acc = 0.0
            for j in range(n):
                acc = fma(M[i, j], zb[j], acc)
            r[i] = ((acc + q[i]) + z0[i]) - zb[i]
        self.r = r
        self.T = np.empty((n, n + 1))
        self.T[:, :n] = -M
        self.T[:, n] = -r
        self.rowvar = [n + i for i in range(n)]          # w_i basic
        self.colvar = [j for j in range(n)] + [2 * n]    # z_j, t nonbasic
        self.beta = np.array([zb[i] - z0[i] for i in range(n)])
        self.nbval = np.array(list(zb) + [0.0])
        self.zst = []
        for i in range(n):
            if z0[i] <= l[i]:
                self.zst.append(AT_L)
            elif z0[i] >= u[i]:
                self.zst.append(AT_U)
            else:
                self.zst.append(FLOAT)
        self.pivots = 0

    # --- bookkeeping -------------------------------------------------------
    def col_of(self, var):
        return self.colvar.index(var)

    def row_of(self, var):
        return self.rowvar.index(var)

    def is_basic(self, var):
        return var in self.rowvar

    def bounds(self, var):
        n = self.n
        if var == 2 * n:
            return 0.0, 1.0
        if var < n:
            return self.l[var], self.u[var]
        k = var - n
        if self.l[k] == self.u[k]:
            return -INF, INF
        st = self.zst[k]
        if st == AT_L:
            return 0.0, INF
        if st == AT_U:
            return -INF, 0.0
        return 0.0, 0.0          # z_k basic or floating: w_k is an artificial fixed at 0

    def artificial_row(self, i):
        v = self.rowvar[i]
        if v < self.n or v == 2 * self.n:
            return False
        lo, up = self.bounds(v)
        return lo == 0.0 and up == 0.0

    # --- primitive operations ---------------------------------------------
    def pivot(self, rho, c):
        T, n = self.T, self.n
        p = T[rho, c]
        prow = np.empty(n + 1)
        for j in range(n + 1):
            prow[j] = (1.0 / p) if j == c else T[rho, j] / p
        for i in range(n):
            if i == rho:
                continue
            d = T[i, c]
            if d == 0.0:
                T[i, c] = 0.0
                continue
            for j in range(n + 1):
                tij = 0.0 if j == c else T[i, j]
                if prow[j] != 0.0:
                    T[i, j] = fma(-d, prow[j], tij)
                else:
                    T[i, j] = tij
        T[rho, :] = prow
        ev, lv = self.colvar[c], self.rowvar[rho]
        self.rowvar[rho], self.colvar[c] = ev, lv
        self.beta[rho], self.nbval[c] = self.nbval[c], self.beta[rho]
        self.pivots += 1

    def best_artificial_row(self, c):
        best, brow = 0.0, -1
        for i in range(self.n):
            if self.artificial_row(i):
                a = abs(self.T[i, c])
                if a > best:
                    best, brow = a, i
        return brow if best > PIV_TOL else -1

    def ratio_test(self, c, sigma):
        """Largest step theta of entering column c in direction sigma.
        Returns (theta, rho, bound) with rho=-1 when no basic blocks."""
        n = self.n
        ratios = [INF] * n
        which = [0] * n
        for i in range(n):
            d = sigma * self.T[i, c]
            lo, up = self.bounds(self.rowvar[i])
            if d > D_TOL and lo > -INF:
                ratios[i] = max((self.beta[i] - lo) / d, 0.0)
                which[i] = -1
            elif d < -D_TOL and up < INF:
                ratios[i] = max((up - self.beta[i]) / (-d), 0.0)
                which[i] = +1
        theta = min(ratios) if n else INF
        if theta == INF:
            return INF, -1, 0
        cut = theta + TIE_TOL * (1.0 + theta)
        # among ties: t first (so the path terminates), then largest |d|, then lowest row
        best, rho = -1.0, -1
        for i in range(n):
            if ratios[i] <= cut:
                if self.rowvar[i] == 2 * n:
                    rho = i
                    break
                a = abs(self.T[i, c])
                if a > best:
                    best, rho = a, i
        return ratios[rho], rho, which[rho]

    def move(self, c, sigma, theta):
        if theta == 0.0:
            return
        for i in range(self.n):
            d = self.T[i, c]
            if d != 0.0:
                self.beta[i] = fma(-(sigma * theta), d, self.beta[i])
        self.nbval[c] = fma(sigma, theta, self.nbval[c])

    def leave_at(self, rho, bound):
        """Snap the blocking basic onto its bound before it leaves."""
        lo, up = self.bounds(self.rowvar[rho])
        self.beta[rho] = lo if bound < 0 else up

    # --- phase 1: crash ----------------------------------------------------
    def try_exchange(self, var):
        """Bring nonbasic var in through an artificial row (no movement)."""
        c = self.col_of(var)
        if c < 0:                                   # the variable is basic: nothing to bring in
            return False
        rho = self.best_artificial_row(c)
        if rho < 0:
            return False
        self.pivot(rho, c)
        return True

    def is_free(self, k):
        return self.l[k] == -INF and self.u[k] == INF

    def best_free_row(self, c):
        """Largest |T[i,c]| over rows still holding the slack w_k of a FREE variable (such a
        row is artificial whatever the start point is)."""
        best, brow = 0.0, -1
        for i in range(self.n):
            v = self.rowvar[i]
            if self.n <= v < 2 * self.n and self.is_free(v - self.n):
                a = abs(self.T[i, c])
                if a > best:
                    best, brow = a, i
        return brow if best > PIV_TOL else -1

    def recompute_tcol(self):
        """T[:, t] = B^-1 r from the slack columns: column of a nonbasic w_k is -B^-1 e_k, a basic
        w_k sitting in row rho means B^-1 e_k = -e_rho.  Sequential fma over k."""
        n = self.n
        tc = self.col_of(2 * n)
        cols = [self.colvar.index(n + k) if (n + k) in self.colvar else -1 for k in range(n)]
        rows = [self.rowvar.index(n + k) if (n + k) in self.rowvar else -1 for k in range(n)]
        for i in range(n):
            acc = 0.0
            for k in range(n):
                pik = -self.T[i, cols[k]] if cols[k] >= 0 else (-1.0 if rows[k] == i else 0.0)
                if pik != 0.0:
                    acc = fma(pik, self.r[k], acc)
            self.T[i, tc] = acc

    def freeze(self):
        """Rows that hold a FREE variable after phase 0 are frozen.  A free basic never blocks a ratio test, never
        leaves the basis and is no candidate row of the crash, so no later decision reads its row: only the final z
        does.  The row as it stands now is a valid equation between the variables whatever pivots follow,
            x_B[i] = beta0[i] - sum_j T0[i, j] * (x(colvar0[j]) - nbval0[j]),
        and `solution` evaluates exactly that (sequential fma over the columns) instead of carrying the row through
        every later pivot.  Engines therefore leave these rows out of their pivot sweeps altogether."""
        n = self.n
        self.frozen = [i for i in range(n) if self.rowvar[i] < n and self.is_free(self.rowvar[i])]
        self.T0 = self.T[self.frozen].copy()
        self.beta0 = self.beta[self.frozen].copy()
        self.colvar0 = list(self.colvar)
        self.nbval0 = self.nbval.copy()

    def crash(self):
        n = self.n
        # phase 0: free variables exchange against rows of free variables only.  Nothing here
        # depends on the start point or on q, so for a matrix shared by a batch it is done once
        # (the GPU engine precomputes it per matrix); the homotopy column is then rebuilt from r.
        piv0 = self.pivots
        for i in range(n):
            if not self.is_free(i):
                continue
            c = self.col_of(i)
            rho = self.best_free_row(c)
            if rho >= 0:
                self.pivot(rho, c)
                self.zst[i] = BASIC
        if self.pivots > piv0:
            self.recompute_tcol()
        self.freeze()
        # phase 1: everything still floating, against any artificial row
        for i in range(n):
            if self.zst[i] != FLOAT:
                continue
            c = self.col_of(i)
            rho = self.best_artificial_row(c)
            if rho >= 0:
                self.pivot(rho, c)
                self.zst[i] = BASIC
                continue
            # dependent column: walk towards an extreme point
            cand = []
            for sigma in (+1.0, -1.0):
                th, rb, wb = self.ratio_test(c, sigma)
                own = (self.u[i] - self.nbval[c]) if sigma > 0 else (self.nbval[c] - self.l[i])
                cand.append((min(th, own), sigma, th, rb, wb, own))
            (ta, sa, *_), (tb, sb, *_) = cand
            pick = cand[0] if ta <= tb else cand[1]
            step, sigma, th, rb, wb, own = pick
            if step == INF:
                continue                           # lineality direction: stays parked
            if own <= th:
                self.move(c, sigma, own)
                self.nbval[c] = self.u[i] if sigma > 0 else self.l[i]
                self.zst[i] = AT_U if sigma > 0 else AT_L
                continue
            self.move(c, sigma, th)
            self.leave_at(rb, wb)
            lv = self.rowvar[rb]
            self.pivot(rb, c)
            self.zst[i] = BASIC
            k = lv if lv < n else lv - n
            if lv < n:
                self.zst[k] = AT_L if wb < 0 else AT_U
                if not self.is_basic(n + k):       # else w_k just turned from artificial to proper
                    self.try_exchange(n + k)
            else:
                if not self.try_exchange(k):
                    self.try_exchange(n + k)
                else:
                    self.zst[k] = BASIC

    def neither_pairs(self):
        """Pairs with z_k nonbasic at a bound and w_k nonbasic too."""
        n = self.n
        bas = set(self.rowvar)
        return [k for k in range(n) if self.zst[k] in (AT_L, AT_U) and self.l[k] != self.u[k]
                and k not in bas and (n + k) not in bas]

    def repair(self):
        """Exchange members of 'neither' pairs (and parked columns that have become
        independent) into remaining artificial rows until nothing changes."""
        n = self.n
        progress = True
        while progress:
            progress = False
            if not any(self.artificial_row(i) for i in range(n)):
                return
            for k in self.neither_pairs():
                if self.try_exchange(n + k):
                    progress = True
                elif self.try_exchange(k):
                    self.zst[k] = BASIC
                    progress = True
            bas = set(self.rowvar)
            for k in range(n):
                if self.zst[k] == FLOAT and k not in bas:
                    if self.try_exchange(k):
                        self.zst[k] = BASIC
                        progress = True

    # --- phase 2: complementary pivoting ----------------------------------
    def lemke(self, max_pivots):
        n = self.n
        ent, sigma = 2 * n, +1.0
        while True:
            if self.pivots > max_pivots:
                return MAX_ITERS
            c = self.col_of(ent)
            th, rb, wb = self.ratio_test(c, sigma)
            if ent == 2 * n:
                own = 1.0 - self.nbval[c]
            elif ent < n:
                own = (self.u[ent] - self.l[ent])
            else:
                own = INF
            if own == INF and th == INF:
                return RAY_TERM
            if own <= th:
                self.move(c, sigma, own)
                if ent == 2 * n:
                    self.nbval[c] = 1.0
                    return SUCCESS
                # entering z flips to its opposite bound
                self.nbval[c] = self.u[ent] if sigma > 0 else self.l[ent]
                self.zst[ent] = AT_U if sigma > 0 else AT_L
                ent, sigma = n + ent, (-1.0 if sigma > 0 else +1.0)
                continue
            self.move(c, sigma, th)
            self.leave_at(rb, wb)
            lv = self.rowvar[rb]
            was_art = self.artificial_row(rb)
            self.pivot(rb, c)
            if ent < n:
                self.zst[ent] = BASIC
            if lv == 2 * n:
                return SUCCESS if wb > 0 else RAY_TERM
            if lv < n:
                self.zst[lv] = AT_L if wb < 0 else AT_U
                if self.is_basic(n + lv):
                    return FAILURE             # no complementary column to continue with
                ent, sigma = n + lv, (+1.0 if wb < 0 else -1.0)
            else:
                k = lv - n
                if was_art or self.is_basic(k) or self.zst[k] not in (AT_L, AT_U):
                    return FAILURE             # blocked by an equation that cannot move
                ent, sigma = k, (+1.0 if self.zst[k] == AT_L else -1.0)

    def value(self, var):
        if var in self.rowvar:
            return self.beta[self.rowvar.index(var)]
        return self.nbval[self.colvar.index(var)]

    def solution(self):
        n = self.n
        z = np.empty(n)
        for i in range(n):
            z[i] = self.value(i)
        # frozen rows (see freeze): their free variables from the phase-0 equations.  Every variable that was
        # nonbasic then is nonbasic or basic in an unfrozen row now, so `value` never reads a frozen row.
        for idx, i in enumerate(self.frozen):
            acc = self.beta0[idx]
            for j in range(n + 1):
                acc = fma(-self.T0[idx, j], self.value(self.colvar0[j]) - self.nbval0[j], acc)
            z[self.rowvar[i]] = acc
        return z

    def basis_codes(self):
        """1 = at lower, 2 = interior/basic, 3 = at upper, 4 = fixed."""
        out = np.empty(self.n, dtype=np.int8)
        for i in range(self.n):
            if self.l[i] == self.u[i]:
                out[i] = 4
            elif i in self.rowvar or self.zst[i] == FLOAT:
                out[i] = 2
            else:
                out[i] = 1 if self.zst[i] == AT_L else 3
        return out


def check_avi_solution(M, q, l, u, z, tol=1e-6):
    """/root/reference/src/avi.jl:148-156 with q = N*w + o already formed."""
    r = M @ z + q
    bad = 0
    for i in range(len(z)):
        if r[i] > tol and abs(z[i] - l[i]) > tol:
            bad += 1
        if r[i] < -tol and abs(z[i] - u[i]) > tol:
            bad += 1
        if z[i] - l[i] < -tol:
            bad += 1
        if z[i] - u[i] > tol:
            bad += 1
    return bad > 0, bad, r


def solve_avi(M, q, l, u, z0, max_pivots=None):
    """Returns (z, status, pivots, basis_codes).  Mirrors the return of
    /root/reference/src/avi.jl:63-77 (status FAILURE when the final check fails)."""
    M = np.asarray(M, dtype=float)
    q = np.asarray(q, dtype=float)
    l = np.asarray(l, dtype=float)
    u = np.asarray(u, dtype=float)
    z0 = np.asarray(z0, dtype=float)
    n = len(q)
    if max_pivots is None:
        max_pivots = 50 * n + 100
    tab = _Tab(M, q, l, u, z0)
    tab.crash()
    tab.repair()
    status = tab.lemke(max_pivots)
    z = tab.solution()
    if status == SUCCESS:
        bad, _, _ = check_avi_solution(M, q, l, u, z)
        if bad:
            status = FAILURE
    return z, status, tab.pivots, tab.basis_codes()
