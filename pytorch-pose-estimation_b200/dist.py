"""Image-sharded multi-GPU glue: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

The hot path shards by image with no exchange inside render / loss / decode.  Two collectives remain
(SURVEY.md 8 e), both latency-bound and both enqueued stream-ordered with no host synchronisation:

* loss: all-reduce(SUM) of the two fp64 un-normalised numerators each rank's fused kernel produced;
  every rank then holds the global-batch loss.  dlogits are never exchanged.  Their normalisation is the caller's
  choice and must match how PARAMETER gradients are combined: under DDP (mean over ranks) keep the local batch
  (`global_batch=None`, what the reference does -- mean of local-batch gradients == global-batch gradient); pass
  `global_batch=B*world` only where the ranks' gradients are summed (or not exchanged at all).
* predictions: all-gather of the fixed-size [B_local, K, 3] rows (+ scores, image ids) so that OKS/AP is
  evaluated over the whole validation set on every rank.  The reference never gathers (each rank writes
  its own results.json, utils/sbp_utils.py:167-169); this fixes that race.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, world_size, rank):
    """Contiguous image shard [lo, hi) of rank `rank`; sizes differ by at most one, earlier ranks get the extras."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _active(group):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def global_sbp_loss(loss_num, num_keypoints, global_batch, lambda_pos=5.0, lambda_neg=1.0, group=None, local_loss=None):
    """loss_num: fp64[2] local (S_pos, S_neg) -> 0-dim fp32 global loss (same on every rank).  In place on loss_num.
    With a single rank the kernel's own `local_loss` (already normalised by the global batch) is returned untouched."""
    if not _active(group):
        if local_loss is not None:
            return local_loss
    else:
        dist.all_reduce(loss_num, op=dist.ReduceOp.SUM, group=group)
    return ((lambda_pos * loss_num[0] + lambda_neg * loss_num[1]) / (2.0 * num_keypoints * global_batch)).to(torch.float32)


def global_spm_loss(loss_num, global_batch, lambda_root=1.0, lambda_disp=0.1, group=None):
    if _active(group):
        dist.all_reduce(loss_num, op=dist.ReduceOp.SUM, group=group)
    return ((lambda_root * loss_num[0] + lambda_disp * loss_num[1]) / float(global_batch)).to(torch.float32)


def gather_packed(packed, ids, group=None):
    """All-gather equal-sized shards of the packed prediction rows [B, 3K+1] (fp32) and ids [B, 2] (int64), ordered by
    rank (= by image for contiguous shards).  Two collectives, no packing kernels, no host synchronisation."""
    if not _active(group):
        return packed, ids
    world = dist.get_world_size(group)
    out_p = torch.empty((world * packed.size(0), packed.size(1)), dtype=packed.dtype, device=packed.device)
    out_i = torch.empty((world * ids.size(0), ids.size(1)), dtype=ids.dtype, device=ids.device)
    dist.all_gather_into_tensor(out_p, packed, group=group)
    dist.all_gather_into_tensor(out_i, ids, group=group)
    return out_p, out_i


class ShardExchange:
    """ONE all-gather per step for everything the ranks exchange.

    Per-rank send buffer = [ packed prediction rows [B,3K+1] fp32 | loss numerators [2] fp64 | ids [B,2] int64 ].
    The fused kernel's epilogue writes rows and numerators straight into it (`out_views()`), so no packing kernel runs;
    after `exchange()` every rank holds all shards in rank (= image) order and `global_loss()` reduces the gathered
    numerators in a fixed order on the device (bit-identical on every rank, no second collective).
    """

    def __init__(self, batch_local, num_keypoints, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if _active(group) else 1
        self.b, self.k = batch_local, num_keypoints
        row_bytes = batch_local * (3 * num_keypoints + 1) * 4
        self.off_num = (row_bytes + 15) // 16 * 16
        self.off_ids = self.off_num + 16
        self.nbytes = self.off_ids + batch_local * 16
        self.send = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self.recv = torch.zeros((self.world, self.nbytes), dtype=torch.uint8, device=device) if self.world > 1 else self.send.view(1, -1)
        self.packed = self.send[:row_bytes].view(torch.float32).view(batch_local, 3 * num_keypoints + 1)
        self.loss_num = self.send[self.off_num:self.off_num + 16].view(torch.float64)
        self.ids = self.send[self.off_ids:].view(torch.int64).view(batch_local, 2)
        self._row_bytes = row_bytes
        self.loss = torch.zeros((), dtype=torch.float32, device=device)

    def out_views(self):
        """kwargs for sbp_fused(out=...): write rows and numerators directly into the send buffer."""
        return {"packed": self.packed, "loss_num": self.loss_num}

    def exchange(self):
        if self.world > 1:
            dist.all_gather_into_tensor(self.recv.view(-1), self.send, group=self.group)
        return self

    def gathered_packed(self):
        """[world*B, 3K+1] fp32 rows in image order (a copy for world > 1: the per-rank blocks are strided in `recv`)."""
        if self.world == 1:
            return self.packed
        return self.recv[:, :self._row_bytes].contiguous().view(torch.float32).view(self.world * self.b, 3 * self.k + 1)

    def gathered_ids(self):
        if self.world == 1:
            return self.ids
        return self.recv[:, self.off_ids:].contiguous().view(torch.int64).view(self.world * self.b, 2)

    def global_loss(self, global_batch, lambda_pos=5.0, lambda_neg=1.0, local_loss=None):
        """0-dim fp32 loss of the global batch from the gathered numerators (call after exchange())."""
        if self.world == 1 and local_loss is not None:
            return local_loss
        from ._cabi import check, lib, ptr, stream_ptr
        nums = self.recv.view(-1)[self.off_num:]
        with torch.cuda.device(self.send.device):
            check(lib().pose_loss_reduce(ptr(nums), self.world, self.nbytes // 8, float(lambda_pos), float(lambda_neg),
                                         1.0 / (2.0 * self.k * global_batch), ptr(self.loss), None, stream_ptr(self.send.device)),
                  "pose_loss_reduce")
        return self.loss


class PeerExchange:
    """The per-step exchange WITHOUT a collective call: peer-mapped symmetric memory + the fused epilogue kernel.

    Every rank allocates one exchange buffer with torch's symmetric-memory allocator and maps all peers' buffers.
    `sbp_fused(..., exchange=self)` makes the back-projection epilogue store its rows, the loss numerators and the ids
    directly into every rank's receive region (NVLink peer stores) and raise a per-rank flag; `finish()` launches the
    one-CTA kernel that waits for all flags and reduces the numerators in rank order.  Receive regions form a ring of
    4 steps; `self.steps` (host) mirrors the device step counter so the views of the completed step can be handed out
    without a synchronisation -- call `advance(n)` after replaying a captured step n times.

    `defer=0` (lock-step): `finish()` launches a one-CTA kernel that publishes this rank's step, waits for everybody's
    and reduces the global loss of that step.  `defer=1` (in-band): no extra launch per step -- the NEXT step's fused
    kernel publishes a step and the next step's epilogue waits for / reduces it, so `finish()` of step s returns the loss
    of step s-1 (written by that epilogue), ranks may drift up to two steps apart, and `flush()` completes the last step.
    """

    def __init__(self, batch_local, num_keypoints, device, image_ids, category_ids, group=None, multicast=None, defer=0, _connect=True):
        import ctypes

        import torch.distributed._symmetric_memory as symm_mem

        from ._cabi import ExchangeDesc, lib
        group = group if group is not None else dist.group.WORLD
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.b, self.k, self.device = batch_local, num_keypoints, device
        # ---- local phase: nothing here talks to another rank
        d = ExchangeDesc()
        d.world, d.rank, d.batch_local, d.num_keypoints = self.world, self.rank, batch_local, num_keypoints
        d.defer = self.defer = int(defer)
        nbytes = int(lib().pose_exchange_layout(ctypes.byref(d)))
        if nbytes == 0:
            raise ValueError("bad exchange shape")
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
        self.ids = torch.stack([image_ids.to(device, torch.int64), category_ids.to(device, torch.int64)], dim=1).contiguous()
        d.ids_local = self.ids.data_ptr()
        self.steps = 0
        self.loss = torch.zeros((), dtype=torch.float32, device=device)
        d.loss_prev = self.loss.data_ptr()
        self.desc = d
        self._multicast_wanted = multicast
        self.multicast = False
        self.handle = None
        if _connect:
            self._connect()

    def _connect(self):
        """Collective phase (every rank of the group must call it): map the peers' buffers, zero, barrier."""
        import os

        import torch.distributed._symmetric_memory as symm_mem
        d = self.desc
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        self.buf.zero_()
        for r, p in enumerate(self.handle.buffer_ptrs):
            d.peer_base[r] = int(p)
        # NVLS: one multimem.st reaches every rank's copy through the switch (egress independent of the world size)
        multicast = self._multicast_wanted
        if multicast is None:
            multicast = os.environ.get("POSE_B200_MULTICAST", "1") == "1"
        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        self.multicast = bool(multicast and mc)
        d.multicast_base = mc if self.multicast else None
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)            # every rank's buffer is zeroed before anyone's first peer store

    def _view(self, off, dtype, shape):
        n = 1
        for s in shape:
            n *= s
        return self.buf[off:off + n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(*shape)

    def _wait(self, fn, global_batch, lambda_pos, lambda_neg):
        import ctypes

        from ._cabi import check, lib, ptr, stream_ptr
        with torch.cuda.device(self.device):
            check(getattr(lib(), fn)(ctypes.byref(self.desc), float(lambda_pos), float(lambda_neg),
                                     1.0 / (2.0 * self.k * global_batch), ptr(self.loss), stream_ptr(self.device)), fn)
        return self.loss

    def finish(self, global_batch, lambda_pos=5.0, lambda_neg=1.0):
        """Publish this rank's step, wait for every rank's rows of step (current - defer) and reduce its global loss
        (stream-ordered, no host sync)."""
        loss = self._wait("pose_exchange_finish", global_batch, lambda_pos, lambda_neg) if self.defer == 0 else self.loss
        self.steps += 1
        return loss

    def flush(self, global_batch, lambda_pos=5.0, lambda_neg=1.0, check=False):
        """defer=1 only: complete the last published step (its rows / loss are what gathered_*() / the result refer to).
        `check=True` also synchronises and raises if any wait of this exchange has timed out (end of an epoch)."""
        if self.defer == 0:
            loss = self.loss
        else:
            loss = self._wait("pose_exchange_flush", global_batch, lambda_pos, lambda_neg)
            self._flushed = self.steps
        if check:
            self.raise_on_error()
        return loss

    def _completed(self):
        """Index of the newest step whose rows are complete on this rank (host mirror)."""
        if self.defer and getattr(self, "_flushed", -1) == self.steps:
            return self.steps
        return self.steps - self.defer

    def advance(self, n=1):
        """Tell the host mirror that a captured step was replayed n more times."""
        self.steps += n

    def raise_on_error(self):
        """Host check of the sticky device error flag (one small D2H copy + sync).  A wait that timed out (a peer never
        published its step) poisoned that step's loss with NaN; the rows of that step are stale.  Raises PoseB200Error."""
        e = self.error()
        if e:
            from ._cabi import PoseB200Error
            raise PoseB200Error(f"PeerExchange: rank {self.rank} gave up waiting for rank {e - 1} (exchange time-out): the gathered rows / "
                                "losses since then are not valid")

    def gathered_padded(self, check=False):
        """[world*B, row_stride] fp32: the receive region of the last finished step as it lies in memory (contiguous).
        `check=True`: synchronise and raise if a wait timed out (use where the rows are consumed, e.g. once per validation)."""
        if check:
            self.raise_on_error()
        return self._view(int(self.desc.off_rows[self._completed() % 4]), torch.float32, (self.world * self.b, int(self.desc.row_stride)))

    def gathered_packed(self, check=False):
        """[world*B, 3K+1] fp32 rows of the last finished step, image order (a strided view: rows are 16-byte padded)."""
        return self.gathered_padded(check)[:, :3 * self.k + 1]

    def gathered_ids(self, check=False):
        if check:
            self.raise_on_error()
        return self._view(int(self.desc.off_ids[self._completed() % 4]), torch.int64, (self.world * self.b, 2))

    def error(self):
        """Non-zero if a wait timed out (costs a host sync; for tests / diagnostics)."""
        return int(self._view(int(self.desc.off_ctrl) + 12, torch.int32, (1,)).item())


def make_exchange(batch_local, num_keypoints, device, image_ids, category_ids, group=None, prefer_p2p=True, defer=0):
    """PeerExchange when symmetric memory can be set up across the group, else the NCCL ShardExchange.  Returns (exchange, kind).

    Failure-symmetric: the constructor's local phase (allocation, layout) runs under try on every rank, then ONE
    all_reduce(MIN) agrees on whether everybody got that far, and only then do the ranks enter the collective phase
    (rendezvous, barrier) together -- a rank that fails early can no longer leave the others inside a barrier."""
    if not _active(group):
        ex = ShardExchange(batch_local, num_keypoints, device, group)
        ex.ids.copy_(torch.stack([image_ids.to(device, torch.int64), category_ids.to(device, torch.int64)], dim=1))
        return ex, "single"
    ex = None
    if prefer_p2p:
        ok = torch.zeros(1, device=device)
        try:
            ex = PeerExchange(batch_local, num_keypoints, device, image_ids, category_ids, group, defer=defer, _connect=False)
            ok.fill_(1)
        except Exception as e:          # noqa: BLE001 -- any local set-up failure (no VMM, old driver) means: use NCCL
            import sys
            print(f"[pose_b200] symmetric-memory exchange unavailable ({type(e).__name__}: {e}); using NCCL all-gather", file=sys.stderr)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if float(ok.item()) == 1.0:
            # collective phase: every rank is here.  A failure inside it is a failure on all ranks (rendezvous is collective).
            try:
                ex._connect()
                ok.fill_(1)
            except Exception as e:      # noqa: BLE001
                import sys
                print(f"[pose_b200] symmetric-memory rendezvous failed ({type(e).__name__}: {e}); using NCCL all-gather", file=sys.stderr)
                ok.fill_(0)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if float(ok.item()) == 1.0:
                return ex, "p2p"
    ex = ShardExchange(batch_local, num_keypoints, device, group)
    ex.ids.copy_(torch.stack([image_ids.to(device, torch.int64), category_ids.to(device, torch.int64)], dim=1))
    return ex, "nccl"


def gather_rows(rows, score, image_ids, category_ids, group=None):
    """rows [B,K,3], score [B], ids [B] -> the same tensors for the global batch (equal-sized shards)."""
    if not _active(group):
        return rows, score, image_ids, category_ids
    k = rows.size(1)
    packed = torch.cat([rows.reshape(rows.size(0), -1), score[:, None]], dim=1).contiguous()
    ids = torch.stack([image_ids.to(packed.device, torch.int64), category_ids.to(packed.device, torch.int64)], dim=1).contiguous()
    out_p, out_i = gather_packed(packed, ids, group)
    return out_p[:, :-1].reshape(-1, k, 3), out_p[:, -1], out_i[:, 0], out_i[:, 1]


def gather_rows_ragged(rows, score, image_ids, category_ids, group=None):
    """Same for shards whose sizes differ by at most one (shard_bounds): pad to the max, gather, strip."""
    if not _active(group):
        return rows, score, image_ids, category_ids
    n = torch.tensor([rows.size(0)], dtype=torch.int64, device=rows.device)
    sizes = [torch.zeros_like(n) for _ in range(dist.get_world_size(group))]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s) for s in sizes]
    m = max(sizes)
    pad = m - rows.size(0)
    if pad:
        rows = torch.cat([rows, rows.new_zeros((pad,) + tuple(rows.shape[1:]))])
        score = torch.cat([score, score.new_zeros(pad)])
        image_ids = torch.cat([image_ids, image_ids.new_full((pad,), -1)])
        category_ids = torch.cat([category_ids, category_ids.new_full((pad,), -1)])
    r, s, i, c = gather_rows(rows, score, image_ids, category_ids, group)
    keep = torch.cat([torch.arange(m, device=r.device) < sz for sz in sizes])
    return r[keep], s[keep], i[keep], c[keep]


def gather_spm_people(kps, counts, image_ids, category_ids, image_w, image_h, group=None):
    """SPM predictions of the global batch: kps [B_local,Pmax,K,3] fp32 (first counts[i] rows of image i valid) + counts
    [B_local] i32 + per-image ids / sizes [B_local] -> the same tensors with leading dimension B_global, in image order.

    The number of persons per image is data dependent, so the fixed-size decode buffers are what travels (SURVEY.md 8 e):
    three collectives -- an all-gather of the shard sizes (one int64 per rank), one of the [Pmax*K*3] fp32 rows and one of
    the [counts, ids, sizes] int64 records; every rank must use the same Pmax.  Shards may differ in size
    by at most one image (shard_bounds): they are padded to the longest shard with count 0 and the padding is stripped.
    """
    if not _active(group):
        return kps, counts, image_ids, category_ids, image_w, image_h
    dev = kps.device
    world = dist.get_world_size(group)
    n = torch.tensor([kps.size(0)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s) for s in sizes]
    m = max(sizes)
    row = kps[0].numel() if kps.size(0) else int(kps.shape[1] * kps.shape[2] * kps.shape[3])
    meta = torch.stack([counts.to(dev, torch.int64), image_ids.to(dev, torch.int64), category_ids.to(dev, torch.int64),
                        image_w.to(dev, torch.int64), image_h.to(dev, torch.int64)], dim=1)          # [B,5]
    send_k = kps.new_zeros((m, row))
    send_k[:kps.size(0)] = kps.reshape(kps.size(0), -1)
    send_m = meta.new_zeros((m, 5))
    send_m[:meta.size(0)] = meta
    out_k = kps.new_empty((world * m, row))
    out_m = meta.new_empty((world * m, 5))
    dist.all_gather_into_tensor(out_k, send_k, group=group)
    dist.all_gather_into_tensor(out_m, send_m, group=group)
    keep = torch.cat([torch.arange(m, device=dev) < sz for sz in sizes])
    out_k, out_m = out_k[keep], out_m[keep]
    return (out_k.view(-1, *kps.shape[1:]), out_m[:, 0].to(torch.int32), out_m[:, 1], out_m[:, 2], out_m[:, 3], out_m[:, 4])
