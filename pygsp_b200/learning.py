"""``pygsp/learning.py`` on the CUDA engine: Tikhonov regression / classification on graphs.

``regression_tikhonov`` (learning.py:255-365) minimises ``|Mx - y|^2 + tau x'Lx``.  The
reference solves ``(M + tau L) x = M y`` with ``scipy.sparse.linalg.cg``, one column at a
time in a Python loop, every product a SciPy SpMV (learning.py:326-337); for ``tau = 0`` it
solves the harmonic extension ``L_uu x_u = -L_ul y_l`` with a sparse direct solver (:350-365).
Here both are ONE block conjugate-gradient run on the device (``gsp_cg_*``, csrc/cg.cu): all
columns advance together, the product with L is the SpMM step kernel of the filter path.
"""
import numpy as np

from . import _native as nat
from . import utils

logger = utils.build_logger(__name__)


def _to_logits(x):
    logits = np.zeros([len(x), np.max(x) + 1])
    logits[range(len(x)), x] = 1
    return logits


def classification_tikhonov(G, y, M, tau=0):
    r"""Classification by Tikhonov regression on the one-hot logits (learning.py:176-252)."""
    y = np.array(y, copy=True)
    y[np.asarray(M) == False] = 0  # noqa: E712
    Y = _to_logits(y.astype(int))
    return regression_tikhonov(G, Y, M, tau)


def _block_cg(G, tau, row_scale, diag, B, tol, maxiter):
    """Solve (diag(row_scale) tau L + diag(diag)) X = B on the device; B (N, nsig) tensor."""
    torch = nat.require_cuda()
    L = G.L
    n, nsig = B.shape
    cap = int(maxiter)
    X = torch.empty_like(B)
    R, P, Q = torch.empty_like(B), torch.empty_like(B), torch.empty_like(B)
    scal = torch.zeros((cap + 1 + 2048) * nsig, dtype=torch.float64, device=B.device)
    done, batch = 0, 25
    best, stall = None, 0
    while done < cap:
        nxt = min(cap, done + batch)
        with torch.cuda.device(B.device):
            nat.call("gsp_cg_" + nat.suffix(B.dtype), nat.i64(n), nat.i64(L.nnz), L.indptr, L.indices,
                     L.data, nat.f64(tau), row_scale, diag, B, X, R, P, Q, nat.i64(nsig),
                     nat.i32(done), nat.i32(nxt), nat.i32(cap), scal, nat.stream_ptr(B.device))
        done = nxt
        rr = scal[:(done + 1) * nsig].reshape(done + 1, nsig).cpu().numpy()
        rel = np.sqrt(rr[-1] / np.maximum(rr[0], 1e-300))
        worst = float(rel.max())
        if worst <= tol:
            return X, done, worst
        if best is None or worst < 0.5 * best:             # float32 stagnates above tiny tols
            best, stall = worst, 0
        else:
            stall += 1
            if stall >= 8:
                break
    logger.warning("conjugate gradients stopped at relative residual %.2e after %d iterations",
                   worst, done)
    return X, done, worst


def regression_tikhonov(G, y, M, tau=0, *, tol=None, maxiter=None):
    r"""Solve a regression problem on a graph via Tikhonov minimisation (learning.py:255-365).

    ``argmin_x |Mx - y|^2 + tau x'Lx`` for ``tau > 0``;
    ``argmin_x x'Lx  s.t.  y = Mx`` otherwise.  ``y``: (N,) or (N, Nv) measurements (entries where
    ``M`` is False are ignored and may be NaN), ``M``: boolean mask of length N.  The inputs are
    not modified.  ``tol``: relative residual at which CG stops (default 1e-6 for a float32
    graph, 1e-10 for float64; the reference's SciPy default is 1e-5).  NumPy in -> NumPy out,
    CUDA tensor in -> CUDA tensor out.
    """
    torch = nat.require_cuda()
    is_tensor = torch.is_tensor(y)
    M_host = np.asarray(M.cpu() if torch.is_tensor(M) else M)
    if M_host.size != G.n_vertices:
        raise ValueError("M should be of size [G.n_vertices,]")
    mask = torch.as_tensor(M_host.astype(bool).ravel(), device=G.device)
    yt = (y if is_tensor else torch.as_tensor(np.asarray(y, dtype=np.float64))).to(
        device=G.device, dtype=G.dtype)
    one_d = yt.dim() == 1
    Y = yt.reshape(G.N, -1).clone()
    Y[~mask] = 0                                            # learning.py:321-322 / :358 (NaNs dropped)
    if tol is None:
        tol = 1e-6 if G.dtype == torch.float32 else 1e-10
    if maxiter is None:
        maxiter = int(min(10 * G.N, 4000))                 # SciPy's default is 10 N
    m = mask.to(G.dtype)
    out = torch.empty_like(Y)
    for lo in range(0, Y.shape[1], 256):                    # block CG: <= 256 columns at a time
        B = Y[:, lo:lo + 256].contiguous()
        if tau > 0:
            X, _, _ = _block_cg(G, float(tau), None, m, B, tol, maxiter)
            out[:, lo:lo + 256] = X
        else:
            # harmonic extension: CG on L restricted to the unlabelled vertices, written on
            # full-length vectors (identity on the labelled ones)
            rhs = -(G.L.dot(B)) * (1 - m)[:, None]
            X, _, _ = _block_cg(G, 1.0, (1 - m).contiguous(), m, rhs.contiguous(), tol, maxiter)
            out[:, lo:lo + 256] = B + X * (1 - m)[:, None]
    if one_d:
        out = out[:, 0]
    if is_tensor:
        return out
    return out.cpu().numpy().astype(np.float64 if G.dtype == torch.float64 else np.float32)
