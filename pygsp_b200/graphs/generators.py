"""Graph models that produce the inputs of the BASELINE configurations.

Each class builds an adjacency ``W`` and hands it to :class:`Graph`; what they
generate follows the reference generators (file:line cited per class) but the
construction is vectorised -- the reference's per-vertex Python loops
(nngraph.py:221-226) make it unusable beyond ~1e6 vertices.  They are input
fabrication for the filtering path, not part of the timed hot path.
"""
import os

import numpy as np
from scipy import sparse, spatial

from .. import _native as nat
from .. import utils
from .csr import DeviceCSR
from .graph import Graph, _torch_dtype, symmetrize_average_device

_DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")


def morton_order(coords, bits=None):
    """Permutation that sorts points along a Z-order (Morton) curve.

    Vertex ids that are close in memory are then close in space, which is what
    makes the neighbour gather of the SpMM hit L1/L2 (SURVEY.md section 7,
    "gather locality").  Works for 2-D and 3-D coordinates.
    """
    coords = np.asarray(coords, dtype=np.float64)
    n, d = coords.shape
    if bits is None:
        bits = 21 if d <= 3 else 64 // d
    lo = coords.min(axis=0)
    span = np.maximum(coords.max(axis=0) - lo, 1e-300)
    q = np.minimum(((coords - lo) / span * (1 << bits)).astype(np.uint64), (1 << bits) - 1)
    code = np.zeros(n, dtype=np.uint64)
    for b in range(bits):
        for k in range(d):
            code |= ((q[:, k] >> np.uint64(b)) & np.uint64(1)) << np.uint64(b * d + k)
    return np.argsort(code, kind="stable")


class Logo(Graph):
    r"""GSP logo graph, N = 1130 (pygsp/graphs/logo.py:21-33).

    The adjacency and coordinates are the ones of the reference's
    ``data/pointclouds/logogsp.mat``, stored as ``pygsp_b200/data/logo.npz``.
    """

    def __init__(self, **kwargs):
        z = np.load(os.path.join(_DATA, "logo.npz"))
        n = len(z["indptr"]) - 1
        W = sparse.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=(n, n))
        self.info = {k: z[k] for k in ("idx_g", "idx_s", "idx_p")}
        plotting = {"limits": np.array([0, 640, -400, 0])}
        super().__init__(W, coords=z["coords"], plotting=plotting, **kwargs)


class Ring(Graph):
    r"""Ring graph: vertex i is linked to i +- 1..k (pygsp/graphs/ring.py)."""

    def __init__(self, N=64, k=1, **kwargs):
        if N < 3:
            raise ValueError("There should be at least 3 vertices.")
        if 2 * k > N:
            raise ValueError("Too many neighbors requested.")
        self.k = k
        rows, cols = [], []
        idx = np.arange(N)
        for s in range(1, k + 1):
            rows.append(idx)
            cols.append((idx + s) % N)
        rows, cols = np.concatenate(rows), np.concatenate(cols)
        W = sparse.coo_matrix((np.ones(rows.size), (rows, cols)), shape=(N, N)).tocsr()
        W = W + W.T
        W.data[:] = 1.0                             # unit weights (antipodal edge met twice)
        theta = 2 * np.pi * idx / N
        super().__init__(W, coords=np.stack([np.cos(theta), np.sin(theta)], axis=1), **kwargs)


class Grid2d(Graph):
    r"""N1 x N2 grid with 4-neighbour (5-point stencil) connectivity, unit weights.

    Same graph and row-major vertex numbering as pygsp/graphs/grid2d.py:40-89.
    """

    def __init__(self, N1=16, N2=None, backend="device", **kwargs):
        if N2 is None:
            N2 = N1
        self.N1, self.N2 = N1, N2
        N = N1 * N2
        if backend == "device":                     # stencil written straight into HBM
            W = grid2d_adjacency_device(N1, N2, kwargs.get("dtype"), kwargs.get("device"))
        else:
            right = np.ones(N - 1)
            right[N2 - 1::N2] = 0                   # no edge across a row end
            W = sparse.diags([right, np.ones(N - N2)], [1, N2], shape=(N, N), format="csr")
            W.eliminate_zeros()
            W = (W + W.T).tocsr()
        x = np.tile(np.arange(N2) / float(N2), N1)
        y = np.repeat(np.arange(N1)[::-1] / float(N1), N2)
        super().__init__(W, coords=np.stack([x, y], axis=1), **kwargs)


def _device_of(device):
    torch = nat.require_cuda()
    return torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())


def grid2d_adjacency_device(N1, N2, dtype=None, device=None):
    """Grid2d adjacency built by ``gsp_grid2d_*`` (no host matrix)."""
    torch = nat.require_cuda()
    dev, dt = _device_of(device), _torch_dtype(torch, dtype)
    n = N1 * N2
    indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nat.call("gsp_grid2d_count", nat.i64(N1), nat.i64(N2), indptr, nat.stream_ptr(dev))
        nnz = int(indptr[-1].item())
        indices = torch.empty(nnz, dtype=torch.int32, device=dev)
        data = torch.empty(nnz, dtype=dt, device=dev)
        nat.call("gsp_grid2d_fill_" + nat.suffix(dt), nat.i64(N1), nat.i64(N2), indptr, indices,
                 data, nat.stream_ptr(dev))
    return DeviceCSR(indptr, indices, data, (n, n))


def knn_device(points, k, device=None, points_per_cell=3.0):
    """k nearest neighbours of every point (self excluded) on the GPU.

    Stands in for ``scipy.spatial.KDTree(X).query(X, k + 1)`` (nngraph.py:213-216):
    returns (nn, dist), both (N, k) CUDA tensors, ascending distance.  2-D / 3-D, k <= 32.
    """
    torch = nat.require_cuda()
    dev = _device_of(device)
    pts = torch.as_tensor(np.asarray(points, dtype=np.float64) if not torch.is_tensor(points)
                          else points).to(device=dev, dtype=torch.float64).contiguous()
    n, dim = pts.shape
    lo = pts.min(dim=0).values.cpu().numpy().astype(np.float64)
    hi = pts.max(dim=0).values.cpu().numpy().astype(np.float64)
    span = np.maximum(hi - lo, 1e-300)
    # cubic cells holding ~points_per_cell points on average
    h = (np.prod(span) * points_per_cell / n) ** (1.0 / dim)
    cells = np.maximum(np.floor(span / h), 1).astype(np.int32)
    while np.prod(cells.astype(np.int64)) >= 2 ** 31:
        cells = np.maximum(cells // 2, 1)
    nn = torch.empty((n, k), dtype=torch.int32, device=dev)
    dist = torch.empty((n, k), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        nat.call("gsp_knn_grid", nat.i64(n), nat.i32(dim), pts, nat.i32(k),
                 np.ascontiguousarray(lo), np.ascontiguousarray(hi), np.ascontiguousarray(cells),
                 nn, dist, nat.stream_ptr(dev))
    return nn, dist


def knn_adjacency_device(points, k, sigma=None, dtype=None, device=None):
    """Symmetric Gaussian k-NN adjacency (nngraph.py:213-226,289-297) built on the GPU."""
    torch = nat.require_cuda()
    dev, dt = _device_of(device), _torch_dtype(torch, dtype)
    nn, dist = knn_device(points, k, dev)
    n = nn.shape[0]
    if sigma is None:
        sigma = float(dist.mean().item())
    indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    indices = torch.empty(n * k, dtype=torch.int32, device=dev)
    data = torch.empty(n * k, dtype=dt, device=dev)
    with torch.cuda.device(dev):
        nat.call("gsp_knn_to_csr_" + nat.suffix(dt), nat.i64(n), nat.i32(k), nn, dist,
                 nat.f64(sigma), indptr, indices, data, nat.stream_ptr(dev))
    return symmetrize_average_device(DeviceCSR(indptr, indices, data, (n, n))), sigma


class NNGraph(Graph):
    r"""k-nearest-neighbour graph of a point cloud with Gaussian weights.

    ``w_ij = exp(-d_ij^2 / sigma)`` for the k nearest neighbours j of i, sigma =
    mean neighbour distance, then ``W <- (W + W^T)/2`` -- the 'knn' branch of
    pygsp/graphs/nngraphs/nngraph.py:147-226,289-297 with its default
    ``center`` / ``rescale`` preprocessing (:127-136).  ``backend='device'`` (default for
    2-D / 3-D clouds, k <= 32) searches and assembles in HBM (csrc/generate.cu);
    ``backend='host'`` uses scipy's cKDTree.  Both give the same adjacency.
    """

    def __init__(self, Xin, k=10, sigma=None, center=True, rescale=True, order=None,
                 backend=None, **kwargs):
        Xin = np.asarray(Xin, dtype=np.float64)
        N, d = Xin.shape
        if k >= N:
            raise ValueError("The number of neighbors (k={}) must be smaller than the number "
                             "of nodes ({}).".format(k, N))
        X = Xin - Xin.mean(axis=0) if center else Xin.copy()
        if rescale:
            radius = 0.5 * np.linalg.norm(X.max(axis=0) - X.min(axis=0), 2)
            X *= (np.power(N, 1.0 / float(min(d, 3))) / 10.0) / radius
        if order == "morton":
            X = X[morton_order(X)]
        elif order is not None:
            X = X[np.asarray(order)]
        if backend is None:     # the device search covers 2-D / 3-D clouds and k <= 32
            backend = "device" if (d in (2, 3) and k <= 32) else "host"
        if backend == "device":                     # grid-hash k-NN + symmetrisation in HBM
            W, sigma = knn_adjacency_device(X, k, sigma, kwargs.get("dtype"), kwargs.get("device"))
            self.k, self.sigma = k, sigma
        else:
            D, NN = spatial.cKDTree(X).query(X, k=k + 1, workers=-1)
            if sigma is None:
                sigma = np.mean(D[:, 1:])
            self.k, self.sigma = k, sigma
            rows = np.repeat(np.arange(N), k)
            W = sparse.csr_matrix((np.exp(-D[:, 1:].ravel() ** 2 / float(sigma)),
                                   (rows, NN[:, 1:].ravel())), shape=(N, N))
            W = utils.symmetrize(W, "average").tocsr()
        super().__init__(W, coords=X, **kwargs)


class Sensor(NNGraph):
    r"""Random sensor network: N uniform points in the unit square, k-NN graph.

    pygsp/graphs/nngraphs/sensor.py:50-75 (non-distributed variant):
    ``coords = default_rng(seed).uniform(0, 1, (N, 2))``, ``NNGraph(k=k,
    rescale=False, center=False)``.  ``order='morton'`` renumbers the vertices
    along a Z-curve (an isomorphic graph with gather-friendly numbering).
    """

    def __init__(self, N=64, k=6, seed=None, order=None, backend=None, **kwargs):
        self.seed = seed
        kwargs["backend"] = backend
        coords = np.random.default_rng(seed).uniform(0, 1, (N, 2))
        kwargs.setdefault("plotting", {"limits": np.array([0, 1, 0, 1])})
        super().__init__(coords, k=k, center=False, rescale=False, order=order, **kwargs)


class SensorStrips:
    r"""Row block of a Sensor-type k-NN graph on the strip domain [0, P) x [0, 1).

    Weak-scaling input of the partitioned path: strip q holds ``n_per`` uniform
    points (``default_rng(seed + q)``, Morton-numbered inside the strip, global ids
    q*n_per ...), every strip is one rank's row block.  A rank regenerates its two
    neighbour strips, so no point data crosses ranks; the graph is exactly the k-NN
    graph of the union of all strips with NNGraph's weights and 'average'
    symmetrisation (nngraph.py:218-226,289-297), as ``tests`` check against a
    directly built global graph.  Host-side input fabrication (scipy cKDTree).
    """

    def __init__(self, rank, parts, n_per, k=10, seed=0):
        self.rank, self.parts, self.n_per, self.k = rank, parts, n_per, k
        self.strips = [q for q in (rank - 1, rank, rank + 1) if 0 <= q < parts]
        pts = []
        for q in self.strips:
            p = np.random.default_rng(seed + q).uniform(0, 1, (n_per, 2))
            p = p[morton_order(p)]
            p[:, 0] += q
            pts.append(p)
        self.points = np.concatenate(pts)
        self.own_lo = self.strips.index(rank) * n_per
        own = np.arange(self.own_lo, self.own_lo + n_per)
        margin = 8.0 * np.sqrt(k / (np.pi * n_per))
        x = self.points[:, 0]
        near = (np.abs(x - rank) < margin) | (np.abs(x - (rank + 1)) < margin)
        near[own] = False
        self.sel = np.concatenate([own, np.flatnonzero(near)])
        tree = spatial.cKDTree(self.points)
        self.D, self.NN = tree.query(self.points[self.sel], k=k + 1, workers=-1)
        if self.D[:, -1].max() * 2 >= margin:
            raise RuntimeError("strip margin too small for this density")
        self.coords = self.points[own]

    def distance_sum(self):
        """(sum, count) of the own points' neighbour distances: sigma = global mean."""
        d = self.D[:self.n_per, 1:]
        return float(d.sum()), int(d.size)

    def adjacency_rows(self, sigma):
        """W[rows of this rank, :] as CSR with GLOBAL column ids."""
        m, k = self.points.shape[0], self.k
        src = np.repeat(self.sel, k)
        A = sparse.csr_matrix((np.exp(-self.D[:, 1:].ravel() ** 2 / float(sigma)),
                               (src, self.NN[:, 1:].ravel())), shape=(m, m))
        S = ((A + A.T) / 2).tocsr()[self.own_lo:self.own_lo + self.n_per]
        S.sort_indices()
        local = S.indices.astype(np.int64)
        strip = np.asarray(self.strips, dtype=np.int64)[local // self.n_per]
        gcol = strip * self.n_per + local % self.n_per
        n_global = self.parts * self.n_per
        # global ids keep the local order inside a strip and strips are ascending,
        # so the columns stay sorted
        return sparse.csr_matrix((S.data, gcol, S.indptr), shape=(self.n_per, n_global))


def morton_order_device(coords, bits=None):
    """:func:`morton_order` on the GPU: the same integer codes (float64 quantisation, bit
    interleave) and a stable sort, so the permutation is identical to the host one."""
    torch = nat.require_cuda()
    n, d = coords.shape
    if bits is None:
        bits = 21 if d <= 3 else 64 // d
    c = coords.to(torch.float64)
    lo = c.min(dim=0).values
    span = torch.clamp(c.max(dim=0).values - lo, min=1e-300)
    q = torch.clamp(((c - lo) / span * float(1 << bits)).to(torch.int64), max=(1 << bits) - 1)
    code = torch.zeros(n, dtype=torch.int64, device=coords.device)
    for b in range(bits):
        for k in range(d):
            code |= ((q[:, k] >> b) & 1) << (b * d + k)
    return torch.sort(code, stable=True).indices


def _unit_ball(dim):
    return {2: np.pi, 3: 4.0 * np.pi / 3.0}[dim]


class KnnSlabs:
    r"""Row block of ONE k-NN graph of ``parts * n_per`` points in the unit square / cube.

    Input of the partitioned path for BASELINE configs[4] (synthetic k-NN point cloud, k = 16,
    3-D, 5e7 points on 8 GPUs) and, in 2-D, the strong-scaling twin of ``Sensor``.  Slab q =
    ``{x_0 in [q/P, (q+1)/P)}`` holds ``n_per`` uniform points (``default_rng(seed + q)``,
    Morton-numbered inside the slab, global ids ``q * n_per ...``) and is rank q's row block.
    A rank regenerates its two neighbour slabs and keeps their points within ``2 m`` of its
    faces, searches the k nearest neighbours of every kept point on the GPU (``gsp_knn_grid``)
    and uses the lists of its own points and of the neighbour points within ``m`` (their
    search balls lie inside the kept set when every k-th distance is <= m, which is checked).
    Weights, sigma = global mean neighbour distance and the 'average' symmetrisation are
    NNGraph's (pygsp/graphs/nngraphs/nngraph.py:213-226,289-297); the result is exactly the row
    block of the graph of the union of all slabs (tests compare with ``NNGraph``).
    ``backend='host'`` runs the same plan with scipy's cKDTree (CPU tests of the host logic).
    """

    def __init__(self, rank, parts, n_per, dim=3, k=16, seed=0, backend="device",
                 margin_factor=2.5, device=None):
        self.rank, self.parts, self.n_per, self.dim, self.k = rank, parts, n_per, dim, k
        self.backend = backend
        n_global = parts * n_per
        self.n_global = n_global
        r_mean = (k / (_unit_ball(dim) * n_global)) ** (1.0 / dim)
        self.margin = m = margin_factor * r_mean
        if parts > 1 and 2 * m > 1.0 / parts:
            raise ValueError("slabs thinner than the search margin: fewer parts or more points")
        strips = [q for q in (rank - 1, rank, rank + 1) if 0 <= q < parts]
        x_lo, x_hi = rank / parts, (rank + 1) / parts
        pts, gid, near = [], [], []
        for q in strips:
            p = np.random.default_rng(seed + q).uniform(0, 1, (n_per, dim))
            p[:, 0] = (p[:, 0] + q) / parts
            if backend == "device":
                torch = nat.require_cuda()
                dev = _device_of(device)
                pt = torch.from_numpy(p).to(dev)
                pt = pt[morton_order_device(pt)]
                ids = torch.arange(q * n_per, (q + 1) * n_per, device=dev)
                if q != rank:
                    dist = (x_lo - pt[:, 0]) if q < rank else (pt[:, 0] - x_hi)
                    keep = dist < 2 * m
                    pt, ids, dist = pt[keep], ids[keep], dist[keep]
                    near.append(dist < m)
                else:
                    near.append(torch.ones(n_per, dtype=torch.bool, device=dev))
                    self.own_lo = int(sum(x.shape[0] for x in pts))
            else:
                p = p[morton_order(p)]
                ids = np.arange(q * n_per, (q + 1) * n_per)
                if q != rank:
                    dist = (x_lo - p[:, 0]) if q < rank else (p[:, 0] - x_hi)
                    keep = dist < 2 * m
                    p, ids, dist = p[keep], ids[keep], dist[keep]
                    near.append(dist < m)
                else:
                    near.append(np.ones(n_per, dtype=bool))
                    self.own_lo = int(sum(x.shape[0] for x in pts))
                pt = p
            pts.append(pt)
            gid.append(ids)
        if backend == "device":
            torch = nat.require_cuda()
            self.points = torch.cat(pts)
            self.gid = torch.cat(gid)
            self.used = torch.cat(near)
            self.coords = self.points[self.own_lo:self.own_lo + n_per]
            self.NN, self.D = knn_device(self.points, k, self.points.device)
            kth = float(self.D[self.used, -1].max().item())
        else:
            self.points = np.concatenate(pts)
            self.gid = np.concatenate(gid)
            self.used = np.concatenate(near)
            self.coords = self.points[self.own_lo:self.own_lo + n_per]
            D, NN = spatial.cKDTree(self.points).query(self.points, k=k + 1, workers=-1)
            self.D, self.NN = D[:, 1:], NN[:, 1:]
            kth = float(self.D[self.used, -1].max())
        if parts > 1 and kth > m:
            raise RuntimeError("search margin too small for this density (k-th distance %g > %g): "
                               "raise margin_factor" % (kth, m))

    def distance_sum(self):
        """(sum, count) of the own points' neighbour distances: sigma = global mean."""
        d = self.D[self.own_lo:self.own_lo + self.n_per]
        if self.backend == "device":
            return float(d.sum(dtype=d.dtype).item()), int(d.numel())
        return float(d.sum()), int(d.size)

    def adjacency_rows(self, sigma):
        """W[rows of this rank, :] as a host CSR with GLOBAL column ids (host backend)."""
        m, k = self.points.shape[0], self.k
        D, NN, used = self.D, self.NN, self.used
        if self.backend == "device":
            D, NN, used = D.cpu().numpy(), NN.cpu().numpy().astype(np.int64), used.cpu().numpy()
        gid = self.gid if self.backend != "device" else self.gid.cpu().numpy()
        src = np.repeat(np.flatnonzero(used), k)
        A = sparse.csr_matrix((np.exp(-D[used].ravel() ** 2 / float(sigma)),
                               (src, NN[used].ravel())), shape=(m, m))
        S = ((A + A.T) / 2).tocsr()[self.own_lo:self.own_lo + self.n_per]
        S.sort_indices()
        # kept points are in ascending global order, so renaming keeps the rows sorted
        return sparse.csr_matrix((S.data, gid[S.indices], S.indptr),
                                 shape=(self.n_per, self.n_global))

    def laplacian_rows_device(self, sigma, dtype=None):
        """Rows of L = D - W of this rank, built in HBM: (indptr int32, indices int32 GLOBAL
        ids, data) tensors in canonical CSR order (sorted rows, diagonal in place), and dw."""
        torch = nat.require_cuda()
        if self.backend != "device":
            raise ValueError("laplacian_rows_device needs backend='device'")
        dev, dt = self.points.device, _torch_dtype(torch, dtype)
        m, k, n = int(self.points.shape[0]), self.k, self.n_per
        # unused rows (kept points farther than `margin` from the slab) get zero-weight lists:
        # an infinite distance gives exp(-inf) = 0 and the zeros are dropped below
        dist = torch.where(self.used[:, None], self.D, torch.full_like(self.D, float("inf")))
        indptr = torch.empty(m + 1, dtype=torch.int32, device=dev)
        indices = torch.empty(m * k, dtype=torch.int32, device=dev)
        data = torch.empty(m * k, dtype=dt, device=dev)
        with torch.cuda.device(dev):
            nat.call("gsp_knn_to_csr_" + nat.suffix(dt), nat.i64(m), nat.i32(k), self.NN, dist,
                     nat.f64(sigma), indptr, indices, data, nat.stream_ptr(dev))
        W = symmetrize_average_device(DeviceCSR(indptr, indices, data, (m, m)))
        del indptr, indices, data, dist
        ptr = W.indptr[self.own_lo:self.own_lo + n + 1].long()
        a, b = int(ptr[0].item()), int(ptr[-1].item())
        cols, vals = W.indices[a:b].long(), W.data[a:b]
        counts = ptr[1:] - ptr[:-1]
        rows = torch.repeat_interleave(torch.arange(n, device=dev), counts)
        keep = vals != 0
        rows, vals = rows[keep], vals[keep]
        gcol = self.gid[cols[keep]]
        del cols, keep, W
        counts = torch.bincount(rows, minlength=n)
        dw = torch.segment_reduce(vals.double(), "sum", lengths=counts)
        # L row i = -W row i with dw_i inserted at the diagonal's sorted position
        gdiag = self.gid[self.own_lo:self.own_lo + n]
        before = gcol < gdiag[rows]
        n_before = torch.segment_reduce(before.double(), "sum", lengths=counts).long()
        has_diag = dw != 0                                  # isolated vertex: empty row (graph.py:620)
        l_counts = counts + has_diag.long()
        l_ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(l_counts, 0, out=l_ptr[1:])
        nnz = int(l_ptr[-1].item())
        if nnz >= 2 ** 31:
            raise ValueError("the rank's rows of L must fit int32 offsets (nnz = %d)" % nnz)
        w_ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts, 0, out=w_ptr[1:])
        pos = torch.arange(rows.numel(), device=dev) - w_ptr[rows] + l_ptr[rows] + \
            (~before & has_diag[rows]).long()
        l_idx = torch.empty(nnz, dtype=torch.int32, device=dev)
        l_val = torch.empty(nnz, dtype=dt, device=dev)
        l_idx[pos] = gcol.int()
        l_val[pos] = -vals
        dpos = (l_ptr[:-1] + n_before)[has_diag]
        l_idx[dpos] = gdiag[has_diag].int()
        l_val[dpos] = dw[has_diag].to(dt)
        return l_ptr.int(), l_idx, l_val, dw


def laplacian_rows(W_rows, row_offset):
    """Rows of the combinatorial Laplacian D - W from rows of a SYMMETRIC adjacency.

    Host helper of the partitioned path: dw of a row block is its row sums, so the
    rows of L are built without the rest of the matrix (graph.py:618-620)."""
    W_rows = W_rows.tocsr()
    n_local, n = W_rows.shape
    dw = np.asarray(W_rows.sum(axis=1)).ravel()
    D = sparse.csr_matrix((dw, (np.arange(n_local), np.arange(n_local) + row_offset)),
                          shape=(n_local, n))
    L = (D - W_rows).tocsr()
    L.eliminate_zeros()
    L.sort_indices()
    return L, dw


def sbm_adjacency(N=1024, k=5, z=None, p=0.7, q=None, seed=None):
    r"""Adjacency of an undirected, loop-free stochastic block model, vectorised.

    Same model as pygsp/graphs/stochasticblockmodel.py:61-144 with its defaults
    (``z = sort(rng.integers(0, k, N))``, edge (i, j), i != j, present with probability
    ``M[z_i, z_j]``, ``M`` = ``q`` off the diagonal and ``p`` on it, unit weights).  The
    reference draws one uniform per vertex pair in a Python loop (N^2 iterations: unusable
    beyond N ~ 3e3); here every block pair draws its edge COUNT from the binomial law and then
    that many distinct pairs uniformly -- the same distribution, O(edges) work.  The
    random stream necessarily differs from the reference's, so parity is statistical
    (tests/test_generators_cpu.py).
    """
    rng = np.random.default_rng(seed)
    if z is None:
        z = np.sort(rng.integers(0, k, N))
    z = np.asarray(z)
    pv = np.asarray(p, dtype=np.float64)
    pv = pv * np.ones(k) if pv.size == 1 else pv
    if pv.shape != (k,):
        raise ValueError("Optional parameter p is neither a scalar nor a vector of length k.")
    if q is None:
        q = 0.3 / k
    M = np.asarray(q, dtype=np.float64)
    M = M * np.ones((k, k)) if M.size == 1 else M.copy()
    if M.shape != (k, k):
        raise ValueError("Optional parameter q is neither a scalar nor a matrix of size k x k.")
    M.flat[::k + 1] = pv
    if (M < 0).any() or (M > 1).any():
        raise ValueError("Probabilities should be in [0, 1].")
    if np.any(np.diff(z) < 0):
        raise ValueError("z must be sorted (blocks contiguous) for the vectorised sampler")
    start = np.searchsorted(z, np.arange(k), side="left")
    size = np.searchsorted(z, np.arange(k), side="right") - start

    def distinct(n_pairs, m):
        """m distinct integers from range(n_pairs), uniformly."""
        if m == 0:
            return np.zeros(0, dtype=np.int64)
        if m > n_pairs // 3:
            return rng.choice(n_pairs, size=m, replace=False).astype(np.int64)
        got = np.unique(rng.integers(0, n_pairs, int(m * 1.05) + 16))
        while got.size < m:
            got = np.unique(np.concatenate([got, rng.integers(0, n_pairs, m - got.size + 16)]))
        return rng.permutation(got)[:m]

    rows, cols = [], []
    for a in range(k):
        for b in range(a + 1):
            if a == b:
                n_pairs = int(size[a]) * (int(size[a]) - 1) // 2
            else:
                n_pairs = int(size[a]) * int(size[b])
            if n_pairs == 0 or M[a, b] == 0:
                continue
            idx = distinct(n_pairs, int(rng.binomial(n_pairs, M[a, b])))
            if a == b:            # index -> (i > j) of the strict lower triangle
                i = np.floor((1 + np.sqrt(1 + 8 * idx.astype(np.float64))) / 2).astype(np.int64)
                i -= (i * (i - 1) // 2 > idx)
                i += ((i + 1) * i // 2 <= idx)
                j = idx - i * (i - 1) // 2
            else:
                i, j = idx // int(size[b]), idx % int(size[b])
            rows.append(start[a] + i)
            cols.append(start[b] + j)
    if rows:
        r = np.concatenate(rows)
        c = np.concatenate(cols)
    else:
        r = c = np.zeros(0, dtype=np.int64)
    W = sparse.coo_matrix((np.ones(2 * r.size), (np.concatenate([r, c]), np.concatenate([c, r]))),
                          shape=(N, N)).tocsr()
    W.sort_indices()
    return W, z


class StochasticBlockModel(Graph):
    r"""Stochastic block model graph (undirected, no self-loops); see :func:`sbm_adjacency`."""

    def __init__(self, N=1024, k=5, z=None, p=0.7, q=None, seed=None, **kwargs):
        self.k, self.p, self.q, self.seed = k, p, q, seed
        W, self.z = sbm_adjacency(N, k, z, p, q, seed)
        self.info = {"node_com": self.z, "comm_sizes": np.bincount(self.z, minlength=k),
                     "world_rad": np.sqrt(N)}
        super().__init__(W, **kwargs)
