"""World-size-2 gloo test (CPU) of the multi-GPU host logic: contiguous image shards, the loss-numerator all-reduce
and the prediction all-gather.  The per-rank numerators / rows are produced by the ORACLE here (the CUDA kernels need a
GPU); what is under test is pose_b200.dist -- sharding, normalisation by the global batch, ordering of gathered rows."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sbp_oracle as so


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pose_b200 import dist as pd
        kp, logits, bbox, iid, cid = so.make_config1_inputs(batch, 5, 16, 12, seed=3, torch_seed=3)
        lo, hi = pd.shard_bounds(batch, world, rank)
        t = torch.from_numpy(so.sbp_render(kp[lo:hi], 16, 12, 1))
        s = torch.sigmoid(logits[lo:hi].double())
        pos = t > 0
        s_pos = torch.where(pos, (s - t) ** 2, t.double() ** 2).sum()
        s_neg = torch.where(pos, torch.zeros_like(s), (s - t) ** 2).sum()
        loss = pd.global_sbp_loss(torch.stack([s_pos, s_neg]), 5, batch)
        joints = so.sbp_decode(logits[lo:hi], 48, 0.25, True)
        img = so.sbp_backproject(joints, bbox[lo:hi], (64, 48))
        rows = torch.where(img[..., 2:3] < 0, torch.zeros_like(img), torch.cat([img[..., :2], torch.ones_like(img[..., 2:3])], -1))
        score = torch.where(img[..., 2] < 0, torch.zeros_like(img[..., 2]), img[..., 2]).sum(1) / 5
        r, sc, gi, gc = pd.gather_rows_ragged(rows, score, iid[lo:hi], cid[lo:hi])
        q.put((rank, float(loss), r.numpy(), sc.numpy(), gi.numpy(), (lo, hi)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7])
def test_two_rank_loss_allreduce_and_prediction_allgather(batch):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    kp, logits, bbox, iid, cid = so.make_config1_inputs(batch, 5, 16, 12, seed=3, torch_seed=3)
    want_loss, _ = so.sbp_loss_closed_form_f64(logits, torch.from_numpy(so.sbp_render(kp, 16, 12, 1)))
    assert got[0][5] == (0, (batch + 1) // 2) and got[1][5] == ((batch + 1) // 2, batch)
    for rank, loss, rows, score, ids, _ in got:
        assert abs(loss - float(want_loss)) <= 1e-6 * float(want_loss)        # global-batch loss on every rank
        assert rows.shape == (batch, 5, 3) and np.array_equal(ids, iid.numpy())  # gathered in image order
    assert np.array_equal(got[0][2], got[1][2]) and np.array_equal(got[0][3], got[1][3])


def test_shard_bounds_cover_and_balance():
    from pose_b200.dist import shard_bounds
    for n in (0, 1, 7, 8, 4096, 32768, 32771):
        for w in (1, 2, 3, 4, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _spm_worker(rank, world, port, batch, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pose_b200 import dist as pd
        g = torch.Generator().manual_seed(11)
        pmax, k = 4, 3
        kps = torch.rand((batch, pmax, k, 3), generator=g)
        counts = torch.randint(0, pmax + 1, (batch,), generator=g, dtype=torch.int32)
        iid, cid = torch.arange(batch) + 100, torch.ones(batch, dtype=torch.int64)
        w, h = torch.arange(batch) + 640, torch.arange(batch) + 480
        nums = torch.rand((batch, 2), generator=g, dtype=torch.float64)
        lo, hi = pd.shard_bounds(batch, world, rank)
        out = pd.gather_spm_people(kps[lo:hi], counts[lo:hi], iid[lo:hi], cid[lo:hi], w[lo:hi], h[lo:hi])
        loss = pd.global_spm_loss(nums[lo:hi].sum(0), batch)
        q.put((rank, [t.numpy() for t in out], float(loss)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [6, 5])
def test_two_rank_spm_people_gather_and_loss(batch):
    """SPM (SURVEY.md 8 e): fixed-size [B,Pmax,K,3] rows + counts travel in one all-gather, even and ragged shards."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_spm_worker, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(11)
    kps = torch.rand((batch, 4, 3, 3), generator=g)
    counts = torch.randint(0, 5, (batch,), generator=g, dtype=torch.int32)
    nums = torch.rand((batch, 2), generator=g, dtype=torch.float64)
    want_loss = float((nums[:, 0].sum() + 0.1 * nums[:, 1].sum()) / batch)
    for rank, out, loss in got:
        assert np.array_equal(out[0], kps.numpy()) and np.array_equal(out[1], counts.numpy())
        assert np.array_equal(out[2], np.arange(batch) + 100) and np.array_equal(out[4], np.arange(batch) + 640)
        assert np.array_equal(out[5], np.arange(batch) + 480) and out[1].dtype == np.int32
        assert abs(loss - want_loss) <= 1e-6 * want_loss
