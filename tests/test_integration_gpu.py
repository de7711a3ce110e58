"""Drop-in integration: a stand-in for module/sbp_detector.py's SBPDetector (training_step / validation_step /
on_validation_epoch_end wiring, minus Lightning) driven once with the reference-shaped oracle on the CPU and once with
the pose_b200 drop-ins on the GPU.  Same weights, same batches -> same loss trajectory, same COCO rows.

Also a randomized shape sweep of every SBP kernel against the oracle (odd sizes, non-multiple-of-4 widths, tiny maps)."""
import numpy as np
import pytest
import torch
from torch import nn

from helpers import allclose, assert_joints, assert_rows, close
from oracle import sbp_oracle as so

pytestmark = pytest.mark.gpu
REL = 1e-5


def _tiny_model(k):
    torch.manual_seed(3)
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, k, 1))


class _Detector:
    """Mirrors module/sbp_detector.py:8-45: holds loss_fn + map_metric, calls them exactly like the reference."""

    def __init__(self, model, loss_fn, map_metric):
        self.model, self.loss_fn, self.map_metric = model, loss_fn, map_metric
        self.opt = torch.optim.SGD(model.parameters(), lr=0.05)

    def training_step(self, batch):
        img, target = batch
        pred = self.model(img)
        loss = self.loss_fn(pred, target['heatmaps'])
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return loss.detach()

    def validation_step(self, batch):
        img, target = batch
        pred = self.model(img)
        loss = self.loss_fn(pred, target['heatmaps'])
        self.map_metric.update_state(target, pred)
        return loss.detach()


class _OracleLoss(nn.Module):
    def forward(self, x, t):
        return so.sbp_loss(x, t)


class _OracleMetric:
    def __init__(self, input_size, thr):
        self.input_size, self.thr, self.result_list = input_size, thr, []

    def update_state(self, target, y_pred):
        j = so.sbp_decode(y_pred.detach(), self.input_size[1], self.thr, True)
        img = so.sbp_backproject(j, target['bbox'], self.input_size)
        self.result_list.extend(so.sbp_result_rows(img, target['image_id'], target['category_id']))


def test_detector_dropin_training_and_validation():
    import pose_b200 as pb
    dev = torch.device("cuda", 0)
    k, h, w, in_size = 17, 64, 48, (256, 192)
    kp, _, bbox, iid, cid = so.make_config1_inputs(8, k, h, w, seed=5)
    gen = torch.Generator().manual_seed(0)
    imgs = torch.randn(8, 3, h, w, generator=gen)
    heat = torch.from_numpy(so.sbp_render(kp, h, w, 2))

    ref = _Detector(_tiny_model(k), _OracleLoss(), _OracleMetric(in_size, 0.25))
    dut = _Detector(_tiny_model(k).to(dev), pb.SBPLoss(), pb.SBPmAPCOCO(None, list(in_size), 0.25))
    fused = _Detector(_tiny_model(k).to(dev), pb.SBPLoss(sigma=2), pb.SBPmAPCOCO(None, list(in_size), 0.25))   # keypoints-only hand-off

    for step in range(4):
        lr = ref.training_step((imgs, {'heatmaps': heat}))
        ld = dut.training_step((imgs.to(dev), {'heatmaps': heat.to(dev)}))
        lf = fused.training_step((imgs.to(dev), {'heatmaps': torch.from_numpy(kp).to(dev)}))
        # cuDNN vs CPU convolutions differ at 1e-6; the loss modules themselves are within 1e-5
        assert close(ld.item(), lr.item(), 2e-4), (step, ld.item(), lr.item())
        assert close(lf.item(), ld.item(), 1e-5)
    tgt = {'bbox': bbox, 'image_id': iid, 'category_id': cid}
    with torch.no_grad():
        vr = ref.validation_step((imgs, {**tgt, 'heatmaps': heat}))
        vd = dut.validation_step((imgs.to(dev), {**tgt, 'heatmaps': heat.to(dev)}))
    assert close(vd.item(), vr.item(), 2e-4)
    assert len(dut.map_metric.result_list) == len(ref.map_metric.result_list) == 8
    # the rows agree wherever the two (slightly different) models picked the same pixel; at least the layout is identical
    for a, b in zip(dut.map_metric.result_list, ref.map_metric.result_list):
        assert a["image_id"] == b["image_id"] and len(a["keypoints"]) == len(b["keypoints"]) == 3 * k
    # feeding the GPU model's own logits to both paths must give identical rows
    with torch.no_grad():
        pred = dut.model(imgs.to(dev))
    m_cpu = _OracleMetric(in_size, 0.25)
    m_cpu.update_state(tgt, pred.cpu())
    m_gpu = pb.SBPmAPCOCO(None, list(in_size), 0.25)
    m_gpu.update_state(tgt, pred)
    assert_rows(m_gpu.result_list, m_cpu.result_list, REL)


@pytest.mark.parametrize("seed", range(12))
def test_randomized_shape_sweep_against_oracle(seed):
    import pose_b200 as pb
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(100 + seed)
    k = int(rng.integers(1, 20))
    h = int(rng.integers(5, 80))
    w = int(rng.integers(4, 80))
    sigma = [1, 1.5, 2, 3, -1][int(rng.integers(0, 5))] if h >= 24 and w >= 24 else 1
    if sigma == -1 and h / 64 < 0.3:
        sigma = 1
    b = int(rng.integers(1, 6))
    kp, logits, bbox, iid, cid = so.make_config1_inputs(b, k, h, w, seed=seed, torch_seed=seed)
    kp[0, 0] = [w + 2.0, h + 2.0]
    logits = logits * float(rng.choice([0.5, 3.0, 9.0]))
    want_t = so.sbp_render(kp, h, w, sigma)
    gen = pb.SBPHeatmapGenerator([h, w], k, sigma)
    assert np.array_equal(gen.render_batch(kp).cpu().numpy(), want_t), (k, h, w, sigma)
    wl, wg = so.sbp_loss_closed_form_f64(logits, torch.from_numpy(want_t))
    scale = 4.0
    for tma in (False, True):
        r = pb.sbp_fused(logits.to(dev), keypoints=kp, sigma=sigma, want_grad=True, decode=True, conf_threshold=0.25, coord_scale=scale,
                         want_target=not tma, bbox=bbox, input_size=(4 * h, 4 * w), tma=tma)
        assert close(r["loss"].item(), float(wl), REL), (k, h, w, sigma, tma)
        assert allclose(r["dlogits"], wg, REL)
        wj = so.sbp_decode(logits, 4 * w, 0.25, True)
        assert_joints(r["joints"], wj, REL)
        assert_rows(pb.packed_to_results(r["packed"], iid, cid), so.sbp_result_rows(so.sbp_backproject(wj, bbox, (4 * h, 4 * w)), iid, cid), REL)
        if not tma:
            assert np.array_equal(r["target"].cpu().numpy(), want_t)
    d = pb.sbp_fused(logits.to(dev), target=torch.from_numpy(want_t).to(dev), want_grad=True)
    assert close(d["loss"].item(), float(wl), REL) and allclose(d["dlogits"], wg, REL)
    for pred in (True, False):
        assert_joints(pb.decode_batch(logits.to(dev), 0.25, scale, pred), so.sbp_decode(logits, 4 * w, 0.25, pred), REL)
        # the reference's own ops on this GPU: rows identical bit for bit (no scalar-tail caveat on CUDA)
        assert torch.equal(pb.decode_batch(logits.to(dev), 0.25, scale, pred, sigmoid_ref="cuda").cpu(), so.sbp_decode(logits.to(dev), 4 * w, 0.25, pred))
