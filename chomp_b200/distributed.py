"""Multi-GPU evaluation of one parameter batch: shard, evaluate, one all-gather.

Every parameter point is independent, so the batch is cut
into contiguous shards, one per rank of a ``torch.distributed`` process group (one
process per GPU, NCCL over NVLink), every rank runs the four stages on its shard,
and ONE ``all_gather_into_tensor`` assembles the ``[B, n_theta]`` result table (the
per-point status flags ride along as an extra column, so a step costs a single
collective).  Results do not depend on the number of ranks: the kernels reduce in a
fixed order, and a point's result does not depend on the batch it sits in.

    sharded = ShardedEngine(survey)              # after dist.init_process_group("nccl")
    w, status = sharded.wtheta(cosmo, halo, hod) # global [B, .] host arrays in, full table out on every rank

The MCMC recipe this replaces evaluates one point at a time
(examples/example_script.py:141-143); the reference has no multi-process path.
"""
import numpy as np

from . import _lib, design


def shard_bounds(n_points, world_size):
    """[(start, stop)] of every rank's contiguous shard (same rule as design.shard)."""
    out = []
    for r in range(world_size):
        sl = design.shard(n_points, r, world_size)
        out.append((sl.start, sl.stop))
    return out


def _dist():
    import torch.distributed as dist
    return dist


class ShardedEngine(object):
    """One `engine.Engine` per rank behind a shard / all-gather front.

    `evaluate(cosmo, halo, hod) -> (table [rows, n_cols] tensor, status [rows] int32 tensor)` may be
    supplied instead of a survey: the host-side logic (sharding, padding of uneven shards, the single
    collective, the double-buffered pipeline) is then exercised without a GPU (tests, gloo)."""

    def __init__(self, survey=None, evaluate=None, group=None, device=None, which=None, tri_moment=None):
        import torch
        self.torch = torch
        dist = _dist()
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        self.survey = survey
        self.engine = None
        if evaluate is None:
            if survey is None:
                raise ValueError("give a Survey or an evaluate callable")
            from . import engine as _engine
            if device is None:
                device = torch.cuda.current_device()
            self.engine = _engine.Engine(survey, device=device)
            if tri_moment is not None:          # the covariance path needs the 1-halo trispectrum's node list
                cfg = survey.config()
                cfg.tri_moment = int(tri_moment)
                self.engine.configure(cfg)
            self.which = _lib.POWER_SPEC[survey.power_spec] if which is None else int(which)
            self._theta = torch.as_tensor(np.ascontiguousarray(survey.theta), device="cuda:%d" % int(device))
            evaluate = self._evaluate_engine
        self.evaluate = evaluate
        self._bufs = {}

    # -- the GPU evaluation of one shard ----------------------------------------------------
    def _evaluate_engine(self, cosmo, halo, hod):
        t = self.torch
        eng = self.engine
        rows = cosmo.shape[0]
        status = t.zeros(rows, dtype=t.int32, device=self._theta.device)
        w = eng.wtheta(cosmo, halo, hod, self._theta, self.which, status=status)
        return w, status

    # -- one step --------------------------------------------------------------------------------
    def local_shard(self, n_points, rank=None):
        return design.shard(n_points, self.rank if rank is None else rank, self.world)

    def _packed_local(self, cosmo, halo, hod, slot=0):
        """Evaluate this rank's shard into a [max_rows, n_cols + 1] buffer (last column: status)."""
        t = self.torch
        n = cosmo.shape[0]
        bounds = shard_bounds(n, self.world)
        max_rows = max(b - a for a, b in bounds)
        a, b = bounds[self.rank]
        table, status = self.evaluate(cosmo[a:b], halo[a:b], hod[a:b])
        n_cols = table.shape[1]
        local, full = self._gather_buffers(slot, max_rows, n_cols, table)
        rows = b - a
        if rows:
            local[:rows, :n_cols] = table
            local[:rows, n_cols] = status.to(table.dtype)
        return local, full, bounds, n_cols

    def _gather_buffers(self, slot, max_rows, n_cols, like):
        t = self.torch
        key = (slot, max_rows, n_cols, like.device, like.dtype)
        if key not in self._bufs:
            self._bufs[key] = (t.zeros((max_rows, n_cols + 1), dtype=like.dtype, device=like.device),
                               t.empty((self.world, max_rows, n_cols + 1), dtype=like.dtype, device=like.device))
        return self._bufs[key]

    def _unpack(self, full, bounds, n_cols):
        t = self.torch
        if all(b - a == full.shape[1] for a, b in bounds):
            flat = full.reshape(-1, n_cols + 1)
        else:
            flat = t.cat([full[r, :b - a] for r, (a, b) in enumerate(bounds)], 0)
        return flat[:, :n_cols], flat[:, n_cols].to(t.int32)

    def wtheta(self, cosmo, halo, hod):
        """Global batch in (host arrays or tensors, identical on every rank), full [B, n_theta] table and
        [B] status flags out on every rank: shard -> four stages -> one all-gather."""
        local, full, bounds, n_cols = self._packed_local(cosmo, halo, hod, slot="call")   # not a pipeline slot
        if self.world > 1:
            _dist().all_gather_into_tensor(full.view(-1, n_cols + 1), local, group=self.group)
        else:
            full = local.unsqueeze(0)
        w, st = self._unpack(full, bounds, n_cols)
        return w.clone(), st                # the staging buffers are reused by the next call

    def wtheta_host(self, cosmo, halo, hod):
        """The same step end to end with HOST buffers: global numpy arrays in, numpy (w [B, n_theta],
        status [B]) out on every rank.  One rank: the C-ABI host call (pinned staging, H2D, stages, D2H).
        Several ranks: the shard's parameters go through pinned memory to the device, the stages run, one
        all-gather assembles the table on the device, one D2H copy brings it back."""
        t = self.torch
        if self.engine is None:
            w, st = self.wtheta(cosmo, halo, hod)
            return w.cpu().numpy(), st.cpu().numpy()
        if self.world == 1:
            return self.engine.wtheta_host(cosmo, halo, hod, self.survey.theta, self.which)
        n = cosmo.shape[0]
        bounds = shard_bounds(n, self.world)
        a, b = bounds[self.rank]
        rows = b - a
        widths = (_lib.N_COSMO, _lib.N_HALO, _lib.N_HOD)
        key = ("host", rows, n)
        if key not in self._bufs:
            self._bufs[key] = (t.empty(max(rows, 1)*sum(widths), dtype=t.float64).pin_memory(),
                               t.empty((n, len(self.survey.theta) + 1), dtype=t.float64).pin_memory())
        pin_in, pin_out = self._bufs[key]
        off, views = 0, []
        for arr, wd in zip((cosmo, halo, hod), widths):
            blk = pin_in[off:off + rows*wd].view(rows, wd)
            blk.copy_(t.from_numpy(np.ascontiguousarray(arr[a:b], dtype=np.float64)))
            views.append((off, wd))
            off += rows*wd
        dev = pin_in.to(self._theta.device, non_blocking=True)
        c, h, g = (dev[o:o + rows*wd].view(rows, wd) for o, wd in views)
        max_rows = max(y - x for x, y in bounds)
        table, status = self.evaluate(c, h, g)
        n_cols = table.shape[1]
        local, full = self._gather_buffers("host", max_rows, n_cols, table)
        if rows:
            local[:rows, :n_cols] = table
            local[:rows, n_cols] = status.to(table.dtype)
        _dist().all_gather_into_tensor(full.view(-1, n_cols + 1), local, group=self.group)
        if all(y - x == max_rows for x, y in bounds):
            pin_out.copy_(full.view(-1, n_cols + 1), non_blocking=True)
        else:
            pin_out.copy_(t.cat([full[r, :y - x] for r, (x, y) in enumerate(bounds)], 0), non_blocking=True)
        t.cuda.current_stream(self._theta.device).synchronize()
        out = pin_out.numpy()
        return out[:, :n_cols].copy(), out[:, n_cols].astype(np.int32)

    # -- config 5: w(theta) and its covariance, one all-gather ---------------------------------------------
    def _evaluate_covariance(self, setup):
        """evaluate callable: [rows, n_theta + n_bins^2] = w(theta) | covariance (row-major) per point.  The
        covariance call leaves the handle in the state of the w(theta) path at z_bar, so the correlation
        function costs one more Hankel launch (include/chomp_b200.h, chomp_b200_covariance)."""
        t = self.torch
        eng = self.engine

        def evaluate(cosmo, halo, hod):
            rows = cosmo.shape[0]
            status = t.zeros(rows, dtype=t.int32, device=self._theta.device)
            cov = eng.covariance(cosmo, halo, hod, setup, status=status)
            w = eng.wtheta_stage(rows, self.which, self._theta, status=status)
            return t.cat([w, cov.reshape(rows, -1)], 1), status
        return evaluate

    def covariance(self, cosmo, halo, hod, setup):
        """Global batch in, (w [B, n_theta], cov [B, n_bins, n_bins], status [B]) out on every rank: shard ->
        Limber / K_NG / trispectrum / halo tables / P + G + NG terms / Hankel -> ONE all-gather of the
        [B/G, n_theta + n_bins^2 + 1] rows (covariance.py:276-321 per point; the [B, n_k, n_k] trispectrum tables
        never leave the owning rank)."""
        saved = self.evaluate
        self.evaluate = self._evaluate_covariance(setup)
        try:
            local, full, bounds, n_cols = self._packed_local(cosmo, halo, hod, slot="cov")
        finally:
            self.evaluate = saved
        if self.world > 1:
            _dist().all_gather_into_tensor(full.view(-1, n_cols + 1), local, group=self.group)
        else:
            full = local.unsqueeze(0)
        table, st = self._unpack(full, bounds, n_cols)
        n_theta, nb = self._theta.numel(), setup.bins.shape[0]
        return table[:, :n_theta].clone(), table[:, n_theta:].reshape(-1, nb, nb).clone(), st

    def covariance_host(self, cosmo, halo, hod, setup):
        """The same end to end with HOST buffers: numpy in (pinned staging, H2D of the shard), numpy out (one D2H
        of the gathered table)."""
        t = self.torch
        n = cosmo.shape[0]
        a, b = shard_bounds(n, self.world)[self.rank]
        rows = b - a
        widths = (_lib.N_COSMO, _lib.N_HALO, _lib.N_HOD)
        n_theta, nb = self._theta.numel(), setup.bins.shape[0]
        key = ("cov_host", rows, n)
        if key not in self._bufs:
            self._bufs[key] = (t.empty(max(rows, 1)*sum(widths), dtype=t.float64).pin_memory(),
                               t.empty((n, n_theta + nb*nb + 1), dtype=t.float64).pin_memory())
        pin_in, pin_out = self._bufs[key]
        off, views = 0, []
        for arr, wd in zip((cosmo, halo, hod), widths):
            pin_in[off:off + rows*wd].view(rows, wd).copy_(t.from_numpy(np.ascontiguousarray(arr[a:b], dtype=np.float64)))
            views.append((off, wd))
            off += rows*wd
        dev = pin_in.to(self._theta.device, non_blocking=True)
        c, h, g = (dev[o:o + rows*wd].view(rows, wd) for o, wd in views)
        # the sharded front works on global arrays: hand it this rank's rows in place
        evaluate = self._evaluate_covariance(setup)
        table, status = evaluate(c, h, g)
        bounds = shard_bounds(n, self.world)
        max_rows = max(y - x for x, y in bounds)
        n_cols = table.shape[1]
        local, full = self._gather_buffers("cov_host", max_rows, n_cols, table)
        if rows:
            local[:rows, :n_cols] = table
            local[:rows, n_cols] = status.to(table.dtype)
        if self.world > 1:
            _dist().all_gather_into_tensor(full.view(-1, n_cols + 1), local, group=self.group)
        else:
            full = local.unsqueeze(0)
        if all(y - x == max_rows for x, y in bounds):
            pin_out.copy_(full.reshape(-1, n_cols + 1), non_blocking=True)
        else:
            pin_out.copy_(t.cat([full[r, :y - x] for r, (x, y) in enumerate(bounds)], 0), non_blocking=True)
        t.cuda.current_stream(self._theta.device).synchronize()
        out = pin_out.numpy()
        return (out[:, :n_theta].copy(), out[:, n_theta:n_cols].reshape(n, nb, nb).copy(),
                out[:, n_cols].astype(np.int32))

    def pipeline(self, batches):
        """Double-buffered steps: the all-gather of step s runs (asynchronously, on the collective's own
        stream) while the stages of step s + 1 compute.  Yields (w, status) per batch, in order."""
        pending = None
        for s, (cosmo, halo, hod) in enumerate(batches):
            local, full, bounds, n_cols = self._packed_local(cosmo, halo, hod, slot=s & 1)
            work = None
            if self.world > 1:
                work = _dist().all_gather_into_tensor(full.view(-1, n_cols + 1), local, group=self.group, async_op=True)
            else:
                full = local.unsqueeze(0)
            if pending is not None:
                yield self._finish(*pending)
            pending = (work, full, bounds, n_cols)
        if pending is not None:
            yield self._finish(*pending)

    def _finish(self, work, full, bounds, n_cols):
        if work is not None:
            work.wait()
        w, st = self._unpack(full, bounds, n_cols)
        return w.clone(), st            # the buffer is reused two steps later
