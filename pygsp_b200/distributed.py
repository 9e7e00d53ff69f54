"""Vertex-partitioned Chebyshev filtering: one process per GPU, halo exchange per step.

The reference is single-process; this is the B200 scale-out of
``approximations.cheby_op`` (pygsp/filters/approximations.py:58-114).  The graph's
vertices are split into P contiguous row blocks.  Rank p owns rows
[bounds[p], bounds[p+1]) of L and the matching slices of T_{k-2}, T_{k-1} and r;
the only data a recurrence step needs from other ranks are the rows of T_{k-1}
its stored columns reference -- the *halo*.  Per step and per rank:

    boundary rows (those with a remote column)   -> fused step kernel
    pack their new values per peer, all-to-all-v -> halo of T_k      (NCCL, NVLink)
    interior rows                                -> fused step kernel, overlapped
                                                    with the exchange on a second stream

Row sums are accumulated in the stored order of the *global* CSR rows (column ids
are renamed, never re-sorted), so the partitioned result equals the single-GPU
result bit for bit.

Host-side planning is NumPy + ``torch.distributed`` (NCCL on GPUs, gloo in the CPU
tests); the compute is the same C-ABI step kernel as the single-GPU path.
"""
import numpy as np

from . import _native as nat


def even_bounds(n, parts):
    """Contiguous 1-D partition: bounds[p] = floor(n * p / parts)."""
    return (np.arange(parts + 1, dtype=np.int64) * n) // parts


class HaloPlan:
    """Everything rank ``rank`` needs to know about its row block, on the host.

    Parameters
    ----------
    rows : scipy.sparse.csr_matrix, shape (n_local, N)
        Rows [bounds[rank], bounds[rank+1]) of the global Laplacian with GLOBAL
        column ids, entries in the global CSR order.
    bounds : array of P + 1 ints
    rank, group : this process and its ``torch.distributed`` group (None = world)

    Attributes (all NumPy, local ids are in the *local order* below)
    ----------
    perm : local order -> original local row.  Boundary rows first (padded to a
        multiple of 4 rows with interior rows), interior rows after.
    n_boundary : rows in the first launch
    indptr, indices, data : local CSR, shape (n_local, n_local + n_halo); column
        j < n_local is local row j (local order), column n_local + h is halo slot h
    halo_ids : global ids of the halo slots, ascending (hence grouped by owner)
    recv_counts[q], send_counts[q], send_idx : the all-to-all-v of one step
    """

    def __init__(self, rows, bounds, rank, group=None, exchange_ids=None):
        bounds = np.asarray(bounds, dtype=np.int64)
        P = len(bounds) - 1
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        rows = rows.tocsr()
        n_local = hi - lo
        if rows.shape[0] != n_local:
            raise ValueError("rows must hold exactly the rank's row block")
        self.rank, self.parts, self.bounds = rank, P, bounds
        self.n_local, self.n_global = n_local, int(bounds[-1])
        indptr = rows.indptr.astype(np.int64)
        cols = rows.indices.astype(np.int64)
        owned = (cols >= lo) & (cols < hi)

        # halo slots: distinct remote columns, ascending => contiguous per owner
        self.halo_ids = np.unique(cols[~owned])
        owner = np.searchsorted(bounds, self.halo_ids, side="right") - 1
        self.recv_counts = np.bincount(owner, minlength=P).astype(np.int64)
        self.n_halo = int(self.halo_ids.size)

        # local order: boundary rows first
        row_of = np.repeat(np.arange(n_local), np.diff(indptr))
        is_boundary = np.bincount(row_of[~owned], minlength=n_local) > 0
        boundary = np.flatnonzero(is_boundary)
        interior = np.flatnonzero(~is_boundary)
        pad = min((-boundary.size) % 4, interior.size)
        self.perm = np.concatenate([boundary, interior]).astype(np.int64)
        self.n_boundary = int(boundary.size + pad)
        self.n_true_boundary = int(boundary.size)
        inv = np.empty(n_local, dtype=np.int64)
        inv[self.perm] = np.arange(n_local)
        self.inv_perm = inv

        # local CSR: permute rows, rename columns, keep the within-row order
        counts = np.diff(indptr)[self.perm]
        new_ptr = np.zeros(n_local + 1, dtype=np.int64)
        np.cumsum(counts, out=new_ptr[1:])
        gather = np.repeat(indptr[self.perm] - new_ptr[:-1], counts) + np.arange(new_ptr[-1])
        c = cols[gather]
        mine = (c >= lo) & (c < hi)
        local = np.empty_like(c)
        local[mine] = inv[c[mine] - lo]
        local[~mine] = n_local + np.searchsorted(self.halo_ids, c[~mine])
        if new_ptr[-1] >= 2 ** 31 or n_local + self.n_halo >= 2 ** 31:
            raise ValueError("local block must fit int32 indices")
        self.indptr = new_ptr.astype(np.int32)
        self.indices = local.astype(np.int32)
        self.data = rows.data[gather]

        # who needs which of my rows: exchange the halo id lists once
        if exchange_ids is None:
            exchange_ids = _exchange_ids_torch
        requested, self.send_counts = exchange_ids(self.halo_ids, self.recv_counts, rank, P, group)
        self.send_idx = inv[requested - lo]              # local-order rows to pack, grouped by peer

    @property
    def nnz(self):
        return int(self.indices.size)


def _exchange_ids_torch(halo_ids, recv_counts, rank, P, group):
    """All-to-all-v of the id lists: I tell owner q which of its rows I need."""
    import torch
    import torch.distributed as dist
    if P == 1:
        return np.zeros(0, dtype=np.int64), np.zeros(1, dtype=np.int64)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" \
        else torch.device("cpu")
    mine = torch.from_numpy(recv_counts.copy()).to(dev)
    theirs = torch.empty_like(mine)
    dist.all_to_all_single(theirs, mine, group=group)
    send_counts = theirs.cpu().numpy().astype(np.int64)
    out = torch.empty(int(send_counts.sum()), dtype=torch.int64, device=dev)
    dist.all_to_all_single(out, torch.from_numpy(halo_ids.copy()).to(dev),
                           output_split_sizes=send_counts.tolist(),
                           input_split_sizes=recv_counts.tolist(), group=group)
    return out.cpu().numpy(), send_counts


class PartitionedCheby:
    """``cheby_op`` on one rank's row block of a partitioned Laplacian.

    ``step`` / ``gather_rows`` default to the CUDA kernels behind the C ABI; the CPU
    (gloo) tests of the host logic inject NumPy stand-ins.  There is no automatic
    fallback: without an injected backend a CUDA device is required.
    """

    def __init__(self, plan, dtype=None, device=None, group=None, backend=None, overlap=None):
        import torch
        self.plan, self.group = plan, group
        self.backend = backend if backend is not None else _CudaBackend(device)
        self.device = self.backend.device
        self.dtype = dtype if dtype is not None else torch.float32
        # Splitting a step into boundary + interior launches costs ~30 us; it pays
        # only when the exchange itself is long (measured on 2 x B200: a 0.4 MB
        # halo is 3 % faster unsplit).  None = decide per call from the halo size.
        self.overlap = overlap
        self.overlap_min_bytes = 16 << 20
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(device=self.device, dtype=dt)
        self.indptr = t(plan.indptr, torch.int32)
        self.indices = t(plan.indices, torch.int32)
        self.data = t(plan.data, self.dtype)
        self.perm = t(plan.perm, torch.int64)
        self.send_idx = t(plan.send_idx, torch.int64)
        self.in_splits = [int(v) for v in plan.send_counts]
        self.out_splits = [int(v) for v in plan.recv_counts]
        self._tile_plans = {}
        self.bytes_sent_per_step = 0

    # ------------------------------------------------------------------ pieces
    def _exchange(self, buf, nsig):
        """Fill the halo rows of ``buf`` (n_local + n_halo, nsig) from the owners."""
        import torch.distributed as dist
        p = self.plan
        if p.parts == 1:
            return
        send = self.backend.gather_rows(buf, self.send_idx, nsig)
        self.bytes_sent_per_step = send.numel() * send.element_size()
        dist.all_to_all_single(buf[p.n_local:], send, output_split_sizes=self.out_splits,
                               input_split_sizes=self.in_splits, group=self.group)

    def _tile_plan(self, nsig, nscales):
        key = (nsig, nscales)
        if key not in self._tile_plans:
            self._tile_plans[key] = self.backend.tile_plan(self, nsig, nscales)
        return self._tile_plans[key]

    # ---------------------------------------------------------------- operator
    def cheby_op(self, lmax, c, x, local_order=False):
        """r = cheby_op(L, c, x) restricted to this rank's rows.

        x : (n_local, nsig) tensor on ``self.device`` (original local row order
            unless ``local_order``); returns (nscales, n_local, nsig) in the same order.
        """
        import torch
        p = self.plan
        c = np.atleast_2d(np.asarray(c, dtype=np.float64))
        nscales, M = c.shape
        if M < 2:
            raise TypeError("The coefficients have an invalid shape")
        if x.shape[0] != p.n_local:
            raise ValueError("First dimension must be the number of local vertices")
        nsig = int(x.shape[1])
        n, nb = p.n_local, p.n_boundary
        ext = n + p.n_halo
        bufs = [torch.empty((ext, nsig), dtype=self.dtype, device=self.device) for _ in range(2)]
        xin = x.to(self.dtype)
        bufs[0][:n] = xin if local_order else xin.index_select(0, self.perm)
        r = torch.empty((nscales, n, nsig), dtype=self.dtype, device=self.device)
        plan = self._tile_plan(nsig, nscales)
        be = self.backend
        halo_bytes = p.n_halo * nsig * bufs[0].element_size()
        overlap = self.overlap if self.overlap is not None else halo_bytes >= self.overlap_min_bytes
        overlap = bool(overlap) and be.has_streams and p.parts > 1
        if overlap and p.send_idx.size and int(p.send_idx.max()) >= p.n_boundary:
            overlap = False          # non-symmetric pattern: sent rows are not all boundary rows
        self._exchange(bufs[0], nsig)                       # halo of T_0
        cur, old = 0, 1
        for k in range(1, M):
            first = k == 1
            ck = np.ascontiguousarray(c[:, k])
            c0 = np.ascontiguousarray(c[:, 0])
            coef = (2.0 / lmax, -1.0, 0.0) if first else (4.0 / lmax, -2.0, -1.0)
            # T_k overwrites T_{k-2} (row-local); for k == 1 it goes to the spare buffer
            x_cur, x_new = bufs[cur], bufs[old]
            args = (self, first, x_cur, x_new, x_new, r, nsig, nscales, ck, c0, coef, plan)
            last = k == M - 1
            if overlap and not last:
                be.step(*args, rows=(0, nb))
                be.fork_exchange(lambda: self._exchange(x_new, nsig))
                be.step(*args, rows=(nb, n))
                be.join_exchange()
            else:
                be.step(*args, rows=(0, n))
                if not last:
                    self._exchange(x_new, nsig)
            cur, old = old, cur
        if local_order:
            return r
        out = torch.empty_like(r)
        out[:, self.perm] = r
        return out


class _CudaBackend:
    has_streams = True

    def __init__(self, device=None):
        torch = nat.require_cuda()
        self.device = torch.device(device if device is not None
                                   else "cuda:%d" % torch.cuda.current_device())
        self.comm_stream = torch.cuda.Stream(device=self.device)
        self._evt = None

    def tile_plan(self, op, nsig, nscales):
        torch = nat.require_cuda()
        if op.dtype != torch.float32:
            return None
        plan = nat.TilePlan()
        with torch.cuda.device(self.device):
            nat.call("gsp_cheby_tile_plan", nat.i64(op.plan.n_local), op.indptr, nat.i64(nsig),
                     nat.i32(nscales), plan, nat.stream_ptr(self.device))
        return plan if plan.rows_per_tile > 0 else None

    def gather_rows(self, buf, idx, nsig):
        torch = nat.require_cuda()
        out = torch.empty((idx.numel(), nsig), dtype=buf.dtype, device=self.device)
        with torch.cuda.device(self.device):
            nat.call("gsp_gather_rows_" + nat.suffix(buf.dtype), nat.i64(idx.numel()), idx, buf,
                     nat.i64(nsig), out, nat.stream_ptr(self.device))
        return out

    def step(self, op, first, x_cur, x_old, x_new, r, nsig, nscales, ck, c0, coef, plan, rows):
        torch = nat.require_cuda()
        if rows[1] <= rows[0]:
            return
        with torch.cuda.device(self.device):
            nat.call("gsp_cheby_step_" + nat.suffix(op.dtype), nat.i32(1 if first else 0),
                     nat.i64(rows[0]), nat.i64(rows[1]), nat.i64(op.plan.nnz), op.indptr,
                     op.indices, op.data, x_cur, None if first else x_old, x_new, r,
                     nat.i64(op.plan.n_local), nat.i64(nsig), nat.i32(nscales), ck, c0,
                     nat.f64(coef[0]), nat.f64(coef[1]), nat.f64(coef[2]), plan,
                     nat.stream_ptr(self.device))

    def fork_exchange(self, fn):
        """Run the pack + all-to-all on the side stream, after what is queued so far."""
        torch = nat.require_cuda()
        main = torch.cuda.current_stream(self.device)
        self.comm_stream.wait_stream(main)
        with torch.cuda.stream(self.comm_stream):
            fn()
            self._evt = self.comm_stream.record_event()

    def join_exchange(self):
        torch = nat.require_cuda()
        if self._evt is not None:
            torch.cuda.current_stream(self.device).wait_event(self._evt)
            self._evt = None
