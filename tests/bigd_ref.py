"""NumPy statement of the STRUCTURED algebra the large-dimension t-target path (csrc/bigd.cuh) runs on the device, written
out so that tests/test_bigd_math.py can check every identity against the dense oracle (oracle/targets.py: dense metric, `eigh`
softabs, `inv`, `cholesky`).  TEST INFRASTRUCTURE -- the product never imports this file.

Target (examples/t_distribution/src/GAMC/simulations.jl:17-33): x ~ MvT_nu(0, Sigma_t), Sigma_t = (nu-2) Sigma / nu, Sigma_ij = c^|i-j|.
    T = inv(Sigma_t) is TRIDIAGONAL and persymmetric:  T = nu/((nu-2)(1-c^2)) tridiag(-c; 1, 1+c^2, ..., 1+c^2, 1; -c)
    v = T x, q = x'v, a = (nu+n)/(nu+q), beta = 2/(nu+q):  tensor = -Hessian = M = a (T - beta v v')      (rank-one modification)
softabs (Klara, :39):  G = Q diag(s(lambda)) Q', s(l) = l coth(alpha l).  s(l) - l = 2 l / (exp(2 alpha l) - 1) vanishes to working
precision for alpha*l >= 21, so  G = M + sum_{j in J} (s(l_j) - l_j) q_j q_j',  J = {j : a mu_j < 21/alpha}, (mu_j, q_j) the lowest
eigenpairs of A = T - beta v v'.  Those follow from T alone (no dense matrix): with the LDL' recurrence of T - mu,
    #eig(A) < mu = #negative pivots + [f(mu) < 0],   f(mu) = 1 - beta v'(T-mu)^-1 v   (secular function; interlacing)
(bisection on the count), and q_j is parallel to (T - mu_j)^-1 v.  Hence
    G = a T + V S V',   V = [x_J ..., v],  S = diag(c_j/|x_j|^2 ..., -a beta)                       (tridiagonal + low rank)
and its Cholesky factor is the product  L = sqrt(a) L_T K_1 ... K_m  of the bidiagonal factor of T and factors
K = chol(I + W S W') computed by a row sweep with an r x r state (r <= 32 columns per factor):
    t = M_i w_i, delta_i = 1 + w_i't, b_i = t/sqrt(delta_i), M_{i+1} = M_i - t t'/delta_i;   K_ii = sqrt(delta_i), K_ij = w_i'b_j (i > j).
Everything the SMMALA step needs -- G^-1 grad, chol(G^-1) z, logdet, quadratic forms -- is a recurrence over the n rows."""
import math

import numpy as np

CHUNK = 32


def tri_constants(n, nu=30.0, c=0.9):
    """diag / off-diagonal of T = inv(Sigma_t), its lower bidiagonal Cholesky factor (ld, le) and the log-pdf constant."""
    s = nu / ((nu - 2.0) * (1.0 - c * c))
    td = np.full(n, s * (1.0 + c * c))
    td[0] = td[-1] = s
    if n == 1:
        td[0] = nu / (nu - 2.0)
    te = np.full(max(n - 1, 0), -s * c)
    ld, le = np.empty(n), np.empty(max(n - 1, 0))
    ld[0] = math.sqrt(td[0])
    for i in range(n - 1):
        le[i] = te[i] / ld[i]
        ld[i + 1] = math.sqrt(td[i + 1] - le[i] * le[i])
    logdet_T = 2.0 * float(np.sum(np.log(ld)))
    logconst = math.lgamma((nu + n) / 2.0) - math.lgamma(nu / 2.0) - 0.5 * n * math.log(nu * math.pi) + 0.5 * logdet_T
    return td, te, ld, le, logconst


def tmatvec(td, te, x):
    y = td * x
    y[:-1] += te * x[1:]
    y[1:] += te * x[:-1]
    return y


def evaluate(td, te, logconst, nu, x):
    n = len(x)
    v = tmatvec(td, te, x)
    q = float(x @ v)
    a = (nu + n) / (nu + q)
    lt = logconst - 0.5 * (nu + n) * math.log1p(q / nu)
    return lt, -a * v, v, q, a, 2.0 / (nu + q)


def count_and_f(td, te, v, beta, mu, pivmin):
    """(#eigenvalues of T - beta v v' below mu, secular function f(mu)) from one LDL' sweep of T - mu."""
    n = len(td)
    D = td[0] - mu
    if abs(D) < pivmin:
        D = -pivmin
    z = v[0]
    neg = 1 if D < 0 else 0
    acc = z * z / D
    for i in range(1, n):
        l = te[i - 1] / D
        D = (td[i] - mu) - l * te[i - 1]
        if abs(D) < pivmin:
            D = -pivmin
        z = v[i] - l * z
        neg += 1 if D < 0 else 0
        acc += z * z / D
    f = 1.0 - beta * acc
    return neg + (1 if f < 0 else 0), f


def lowest_roots(td, te, v, beta, upper, nbisect=64):
    """All eigenvalues of A = T - beta v v' below `upper`, ascending, by bisection on the count."""
    gl = float(np.min(td) - 2.0 * np.max(np.abs(te))) if len(te) else float(td[0])
    gu = float(np.max(td) + 2.0 * np.max(np.abs(te))) if len(te) else float(td[0])
    lo0 = gl - beta * float(v @ v) - 1e-3 * max(1.0, abs(gl))
    pivmin = 1e-290
    r, _ = count_and_f(td, te, v, beta, upper, pivmin)
    roots = []
    for j in range(1, r + 1):
        lo, hi = lo0, min(upper, gu)
        for _ in range(nbisect):
            mid = 0.5 * (lo + hi)
            cnt, _ = count_and_f(td, te, v, beta, mid, pivmin)
            if cnt >= j:
                hi = mid
            else:
                lo = mid
        roots.append(0.5 * (lo + hi))
    return np.array(roots)


def eigvec_unnormalised(td, te, v, mu, pivmin=1e-290):
    """x = (T - mu)^-1 v by the LDL' factorisation (no pivoting; |D| floored): parallel to the eigenvector of A for the root mu."""
    n = len(td)
    D, l, y = np.empty(n), np.zeros(n), np.empty(n)
    D[0] = td[0] - mu
    if abs(D[0]) < pivmin:
        D[0] = -pivmin
    z = v[0]
    y[0] = z / D[0]
    for i in range(1, n):
        l[i] = te[i - 1] / D[i - 1]
        D[i] = (td[i] - mu) - l[i] * te[i - 1]
        if abs(D[i]) < pivmin:
            D[i] = -pivmin
        z = v[i] - l[i] * z
        y[i] = z / D[i]
    x = np.empty(n)
    x[n - 1] = y[n - 1]
    for i in range(n - 2, -1, -1):
        x[i] = y[i] - l[i + 1] * x[i + 1]
    return x


def softabs_gap(lam, alpha):
    """s(lam) - lam = lam coth(alpha lam) - lam, accurately for every sign and size of lam (1/alpha at 0)."""
    x = alpha * lam
    if abs(x) < 1e-8:
        return 1.0 / alpha - lam
    if x > 0:
        return 2.0 * lam / math.expm1(2.0 * x) if x < 350.0 else 0.0
    # lam < 0: s = |lam| coth(alpha |lam|) = |lam| + 2|lam|/(exp(2 alpha |lam|) - 1);  s - lam = 2|lam| + 2|lam|/expm1(2 alpha |lam|)
    return -2.0 * lam + ((-2.0 * lam) / math.expm1(-2.0 * x) if x > -350.0 else 0.0)


def low_rank_terms(td, te, v, a, beta, alpha):
    """Columns V (n x r, natural order) and weights S of G = a T + V diag(S) V'; the last column is v itself."""
    thr = 21.0 / alpha
    roots = lowest_roots(td, te, v, beta, thr / a)
    cols, S = [], []
    for mu in roots:
        x = eigvec_unnormalised(td, te, v, mu)
        x = x / np.max(np.abs(x))                       # scale only (keeps the squares far from overflow)
        cols.append(x)
        S.append(softabs_gap(a * mu, alpha) / float(x @ x))
    cols.append(v.copy())
    S.append(-a * beta)
    return np.array(cols).T, np.array(S), roots


# ---- the structured factor in "M space" (index-reversed order; T is persymmetric, so T' = T) ---------------------------------
class Factor:
    """Lm = sqrt(a) L_T K_1 ... K_m with Lm Lm' = P G P (P = index reversal): the reverse Cholesky factor of G."""

    def __init__(self, ld, le, a, V, S):
        self.ld, self.le, self.sa = ld, le, math.sqrt(a)
        self.n = len(ld)
        self.chunks = []          # (W, B, delta)
        Vm = V[::-1, :]           # M space
        r = Vm.shape[1]
        for c0 in range(0, r, CHUNK):
            U = np.column_stack([self._apply_prev(self._bidiag_solve(Vm[:, k])) for k in range(c0, min(r, c0 + CHUNK))])
            self.chunks.append(self._sweep(U, S[c0:c0 + CHUNK]))

    def _bidiag_solve(self, b):                 # (sqrt(a) L_T) u = b
        u = np.empty(self.n)
        u[0] = b[0] / (self.sa * self.ld[0])
        for i in range(1, self.n):
            u[i] = (b[i] - self.sa * self.le[i - 1] * u[i - 1]) / (self.sa * self.ld[i])
        return u

    def _bidiag_solve_t(self, b):               # (sqrt(a) L_T)' y = b
        y = np.empty(self.n)
        y[-1] = b[-1] / (self.sa * self.ld[-1])
        for i in range(self.n - 2, -1, -1):
            y[i] = (b[i] - self.sa * self.le[i] * y[i + 1]) / (self.sa * self.ld[i])
        return y

    @staticmethod
    def _sweep(W, S):
        n, r = W.shape
        M = np.diag(np.asarray(S, dtype=np.float64))
        B, delta = np.empty((n, r)), np.empty(n)
        for i in range(n):
            t = M @ W[i]
            d = 1.0 + float(W[i] @ t)
            if not d > 0.0:
                raise np.linalg.LinAlgError(f"structured factor: pivot {i} not positive")
            delta[i] = d
            B[i] = t / math.sqrt(d)
            M = M - np.outer(t, t) / d
        return W, B, delta

    @staticmethod
    def _k_solve(chunk, u):                     # K y = u
        W, B, delta = chunk
        y, sig = np.empty(len(u)), np.zeros(W.shape[1])
        for i in range(len(u)):
            y[i] = (u[i] - float(W[i] @ sig)) / math.sqrt(delta[i])
            sig += B[i] * y[i]
        return y

    @staticmethod
    def _k_solve_t(chunk, b):                   # K' w = b
        W, B, delta = chunk
        w, tau = np.empty(len(b)), np.zeros(W.shape[1])
        for j in range(len(b) - 1, -1, -1):
            w[j] = (b[j] - float(B[j] @ tau)) / math.sqrt(delta[j])
            tau += W[j] * w[j]
        return w

    def _apply_prev(self, u):
        for ch in self.chunks:
            u = self._k_solve(ch, u)
        return u

    def solve(self, b):                         # Lm y = b
        return self._apply_prev(self._bidiag_solve(b))

    def solve_t(self, b):                       # Lm' y = b
        for ch in reversed(self.chunks):
            b = self._k_solve_t(ch, b)
        return self._bidiag_solve_t(b)

    def sumlog(self):                           # sum_i log Lm_ii
        return self.n * math.log(self.sa) + float(np.sum(np.log(self.ld))) + 0.5 * sum(float(np.sum(np.log(d))) for _, _, d in self.chunks)

    def dense_inverse(self):                    # Lm^-1 (lower), column by column
        X = np.zeros((self.n, self.n))
        for c in range(self.n):
            e = np.zeros(self.n)
            e[c] = 1.0
            X[:, c] = self.solve(e)
        return X


def structured_point(td, te, ld, le, logconst, nu, alpha, x):
    """Everything the SMMALA step needs at the point x, from the structure alone."""
    lt, grad, v, q, a, beta = evaluate(td, te, logconst, nu, x)
    V, S, roots = low_rank_terms(td, te, v, a, beta, alpha)
    F = Factor(ld, le, a, V, S)
    gm = grad[::-1]
    fterm = F.solve_t(F.solve(gm))[::-1]        # G^-1 grad
    return {"lt": lt, "grad": grad, "v": v, "a": a, "beta": beta, "V": V, "S": S, "roots": roots, "F": F, "fterm": fterm, "sumlog": F.sumlog()}


def g_times(td, te, pt, d):
    """G d from the structure: a T d + V (S .* (V'd))."""
    return pt["a"] * tmatvec(td, te, d) + pt["V"] @ (pt["S"] * (pt["V"].T @ d))


def chol_inv_times(pt, z):
    """chol(inv(G)) z  (the lower Cholesky factor of the inverse metric applied to z) = P Lm^-T P z."""
    return pt["F"].solve_t(z[::-1])[::-1]
