#!/usr/bin/env python
"""Multi-GPU parity of the per-step exchange, run on the real ranks (bench.py calls it at N > 1; also stand-alone:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/exchange_check.py

What the reference does at this point: every rank appends its own shard's rows to its own `result_list` and writes the
same ./results.json (utils/sbp_utils.py:148-169) -- no exchange at all.  What is checked here:

parity   ONE step on the same inputs through the three implementations -- peer-memory epilogue in lock-step (defer 0) and
         in-band (defer 1) mode, and the NCCL all-gather (`ShardExchange`): the gathered `[world*B, 3K+1]` rows and the
         ids must be bit-identical between them and rank r's block must equal rank r's local `packed` (every rank checks
         every block: the local rows of all ranks are all-gathered with NCCL for that); the global loss must be
         bit-identical on all ranks and equal the fp64 all-reduce of the numerators.
stress   `steps` replays of a captured 4-step graph of the in-band mode (the ring has 4 slots) with a tag that changes
         every step: bbox.x += 1 and category_id += 1 inside the graph, bbox w/h = the input size so that
         x_img = 4*col + tag exactly.  After every step, still inside the graph, the rows / ids of the newest COMPLETED step
         of EVERY rank are compared on the device with what that rank must have produced (this rank's own rows of that
         step shifted by the rank offset of the tag): a flag that became visible before its data, a stale ring slot or a
         torn row shows up as a mismatch count.  One rank is slowed down periodically so the ranks drift.
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K, H, W, SIGMA, IN_H, IN_W, THR = 17, 64, 48, 2, 256, 192, 0.25
RANK_TAG = 4096.0          # bbox.x of rank r starts at r * RANK_TAG; the step tag is added to it (all exact in fp32)


def _inputs(dev, b, seed):
    gen = torch.Generator(device=dev).manual_seed(seed)
    logits = torch.randn(b, K, H, W, device=dev, generator=gen) * 3.0
    kp = torch.stack([torch.rand(b, K, device=dev, generator=gen, dtype=torch.float64) * W,
                      torch.rand(b, K, device=dev, generator=gen, dtype=torch.float64) * H], dim=-1)
    kp[torch.rand(b, K, device=dev, generator=gen) >= 0.85] = -1.0
    return logits, kp


def parity(pb, pd, dev, b, world, rank):
    """-> dict of mismatch counts (all zero when the three exchanges agree)."""
    gb = b * world
    logits, kp = _inputs(dev, b, 100 + rank)
    gen = torch.Generator(device=dev).manual_seed(7 + rank)
    bbox = torch.stack([torch.rand(b, device=dev, generator=gen, dtype=torch.float64) * 400,
                        torch.rand(b, device=dev, generator=gen, dtype=torch.float64) * 400,
                        torch.rand(b, device=dev, generator=gen, dtype=torch.float64) * 260 + 40,
                        torch.rand(b, device=dev, generator=gen, dtype=torch.float64) * 340 + 60], dim=-1)
    iid = torch.arange(b, device=dev, dtype=torch.int64) + rank * b
    cid = torch.full((b,), 1 + rank, device=dev, dtype=torch.int64)
    base = dict(keypoints=kp, sigma=SIGMA, want_grad=False, decode=True, conf_threshold=THR, coord_scale=IN_W / W, global_batch=gb,
                bbox=bbox, input_size=(IN_H, IN_W))
    local = pb.sbp_fused(logits, **base)                               # no exchange: this rank's own rows and numerators
    # reference for every block: NCCL all-gather of the local rows, fp64 all-reduce of the numerators
    all_rows = torch.empty((world * b, 3 * K + 1), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(all_rows, local["packed"].contiguous())
    num = local["loss_num"].clone()
    dist.all_reduce(num, op=dist.ReduceOp.SUM)
    want_loss = ((5.0 * num[0] + 1.0 * num[1]) / (2.0 * K * gb)).to(torch.float32)
    want_ids = torch.stack([torch.arange(world * b, device=dev), 1 + torch.arange(world * b, device=dev) // b], dim=1)

    out = {}
    results = {}
    # NCCL exchange
    sx = pd.ShardExchange(b, K, dev)
    sx.ids.copy_(torch.stack([iid, cid], dim=1))
    r = pb.sbp_fused(logits, out=sx.out_views(), **base)
    sx.exchange()
    results["nccl"] = (sx.gathered_packed().clone(), sx.gathered_ids().clone(), sx.global_loss(gb, local_loss=r["loss"]).clone())
    for defer in (0, 1):
        try:
            px = pd.PeerExchange(b, K, dev, iid, cid, defer=defer)
        except Exception as e:      # noqa: BLE001
            out["p2p_unavailable"] = f"{type(e).__name__}: {e}"
            break
        pb.sbp_fused(logits, exchange=px, **base)
        loss = px.finish(gb)
        loss = px.flush(gb)                                            # defer 1: complete the step just produced (no-op for 0)
        torch.cuda.synchronize()
        results[f"p2p_defer{defer}"] = (px.gathered_packed().clone(), px.gathered_ids().clone(), loss.clone())
        out[f"p2p_defer{defer}_multicast"] = bool(px.multicast)
        out[f"p2p_defer{defer}_error_flag"] = px.error()
    torch.cuda.synchronize()
    bits = lambda t: t.contiguous().view(torch.int32)                  # noqa: E731
    for name, (rows, ids, loss) in results.items():
        out[f"{name}_rows_vs_allgathered_local"] = int((bits(rows) != bits(all_rows)).sum())
        out[f"{name}_ids"] = int((ids != want_ids).sum())
        out[f"{name}_loss_rel_err_vs_f64_allreduce"] = abs(float(loss) - float(want_loss)) / abs(float(want_loss))
        # the loss must be the same bits on every rank
        lg = [torch.zeros_like(loss) for _ in range(world)]
        dist.all_gather(lg, loss.reshape(()))
        out[f"{name}_loss_differs_between_ranks"] = int(sum(int(bits(x.reshape(1)) != bits(lg[0].reshape(1))) for x in lg))
    if "p2p_defer0" in results:
        for a in ("p2p_defer1", "nccl"):
            if a in results:
                out[f"p2p_defer0_vs_{a}_rows"] = int((bits(results["p2p_defer0"][0]) != bits(results[a][0])).sum())
                out[f"p2p_defer0_vs_{a}_loss_bits"] = int(bits(results["p2p_defer0"][2].reshape(1)) != bits(results[a][2].reshape(1)))
    return out


def stress(pb, pd, dev, b, world, rank, steps):
    """-> dict(steps, mismatches, ...) for `steps` in-band (defer 1) steps replayed from a 4-step CUDA graph."""
    gb = b * world
    logits, kp = _inputs(dev, b, 5)                                    # the SAME logits on every rank: rows differ only by the tag
    bbox = torch.zeros(b, 4, dtype=torch.float64, device=dev)
    bbox[:, 0] = rank * RANK_TAG
    bbox[:, 2], bbox[:, 3] = IN_W, IN_H                                # ratio exactly 1: x_img = 4*col + bbox.x, y_img = 4*row
    iid = torch.arange(b, device=dev, dtype=torch.int64) + rank * b
    cid = torch.zeros(b, device=dev, dtype=torch.int64)
    try:
        px = pd.PeerExchange(b, K, dev, iid, cid, defer=1)
    except Exception as e:          # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}
    packed = torch.zeros(b, 3 * K + 1, dtype=torch.float32, device=dev)
    prev = torch.zeros_like(packed)                                    # this rank's rows of the newest completed step
    prev_tag = torch.zeros((), dtype=torch.int64, device=dev)
    bad = torch.zeros((), dtype=torch.int64, device=dev)
    checked = torch.zeros((), dtype=torch.int64, device=dev)
    loss_log = torch.zeros(steps + 8, dtype=torch.float32, device=dev)
    cursor = torch.zeros((), dtype=torch.int64, device=dev)
    outs = dict(joints=torch.empty(b, K, 3, device=dev), loss=torch.empty((), device=dev), packed=packed)
    offs = ((torch.arange(world, device=dev) - rank).float() * RANK_TAG).view(world, 1, 1)
    want_iid = torch.arange(world * b, device=dev).view(world, b)

    def one_step(slow):
        bbox[:, 0] += 1.0
        px.ids[:, 1] += 1
        if slow:
            torch.cuda._sleep(400_000)                                 # ~0.2 ms: this rank falls behind, the others run ahead
        pb.sbp_fused(logits, keypoints=kp, sigma=SIGMA, want_grad=False, decode=True, conf_threshold=THR, coord_scale=IN_W / W,
                     global_batch=gb, bbox=bbox, input_size=(IN_H, IN_W), out=outs, exchange=px)
        loss = px.finish(gb)                                           # in-band: no launch; `loss` = global loss of the previous step
        if px._completed() >= 1:                                       # the newest completed step is the previous one
            g = px.gathered_padded().view(world, b, -1)
            gi = px.gathered_ids().view(world, b, 2)
            flag = prev[:, 2:3 * K:3]
            want_x = prev[:, 0:3 * K:3].unsqueeze(0) + offs * flag.unsqueeze(0)
            n = (g[:, :, 0:3 * K:3] != want_x).sum() + (g[:, :, 1:3 * K:3] != prev[:, 1:3 * K:3].unsqueeze(0)).sum() \
                + (g[:, :, 2:3 * K:3] != flag.unsqueeze(0)).sum() + (g[:, :, 3 * K] != prev[:, 3 * K].unsqueeze(0)).sum() \
                + (gi[:, :, 1] != prev_tag).sum() + (gi[:, :, 0] != want_iid).sum()
            bad.add_(n)
            checked.add_(1)
            loss_log.index_copy_(0, cursor.reshape(1), loss.reshape(1))
            cursor.add_(1)
        prev.copy_(packed)
        prev_tag.copy_(px.ids[0, 1])

    for i in range(4):                                                 # warm-up, eager: one turn of the ring
        one_step(False)
    torch.cuda.synchronize()
    dist.barrier()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    host_steps = px.steps
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(4):
            one_step(slow=(i == 1 and rank == world - 1))
    px.steps = host_steps                                              # capture ran nothing on the device
    reps = max(1, steps // 4)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        graph.replay()
    t1.record()
    px.advance(4 * reps)
    px.flush(gb)
    torch.cuda.synchronize()
    n_checked, n_bad, err = int(checked.item()), int(bad.item()), px.error()
    # the logged global losses must be the same bits on every rank
    n_log = int(cursor.item())
    mine = loss_log[:n_log].contiguous()
    ref = mine.clone()
    dist.broadcast(ref, src=0)
    loss_diff = int((mine.view(torch.int32) != ref.view(torch.int32)).sum())
    tot = torch.tensor([n_bad, loss_diff, err], device=dev, dtype=torch.int64)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    return {"steps": 4 * reps, "steps_checked_per_rank": n_checked, "row_or_id_mismatches_all_ranks": int(tot[0]),
            "loss_bit_differences_vs_rank0_all_ranks": int(tot[1]), "error_flags_all_ranks": int(tot[2]),
            "nan_losses": int(torch.isnan(mine).sum()), "multicast": bool(px.multicast), "ms_per_step_with_checks": t0.elapsed_time(t1) / (4 * reps)}


def run(pb, pd, dev, b, world, rank, steps):
    res = {"parity": parity(pb, pd, dev, b, world, rank)}
    res["stress_defer1"] = stress(pb, pd, dev, b, world, rank, steps) if "p2p_unavailable" not in res["parity"] else {"skipped": "no peer exchange"}
    p = res["parity"]
    fails = [k for k, v in p.items() if (k.endswith(("_rows_vs_allgathered_local", "_ids", "_rows", "_loss_bits", "_loss_differs_between_ranks", "_error_flag")) and v != 0)
             or (k.endswith("_loss_rel_err_vs_f64_allreduce") and v > 1e-6)]
    s = res["stress_defer1"]
    if "steps" in s and (s["row_or_id_mismatches_all_ranks"] or s["loss_bit_differences_vs_rank0_all_ranks"] or s["error_flags_all_ranks"] or s["nan_losses"]):
        fails.append("stress_defer1")
    res["ok"] = not fails
    res["failed"] = fails
    return res


def main():
    import json
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import pose_b200 as pb
    from pose_b200 import dist as pd
    steps = int(os.environ.get("POSE_B200_STRESS_STEPS", "20000"))
    res = run(pb, pd, dev, int(os.environ.get("POSE_B200_CHECK_BATCH", "4096")), world, rank, steps)
    allok = torch.tensor([int(res["ok"])], device=dev)
    dist.all_reduce(allok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"exchange_check": res, "all_ranks_ok": bool(allok.item())}))
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if bool(allok.item()) else 1)


if __name__ == "__main__":
    main()
