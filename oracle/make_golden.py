"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE on seeded inputs.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_golden

The oracle is NOT used to produce any stored output: every array written here
comes out of a reference function (utils/sbp_utils.py, utils/spm_utils.py,
utils/sbp_pis_utils.py, models/loss/*.py).  Inputs come from oracle/cases.py
(seeded); a sha256 of the inputs is stored next to the outputs.
"""
import json
import os
import sys

import numpy as np
import torch

from . import cases
from .reference_loader import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _metric_stub(ref_cls, input_size, thr, **extra):
    """Build the reference's mAP helper without touching pycocotools (its ctor only stores fields)."""
    m = object.__new__(ref_cls)
    return m


def golden_sbp(ref, name):
    kp, logits, bbox, iid, cid, meta = cases.sbp_case(name)
    k, h, w, sigma, in_size = meta["k"], meta["h"], meta["w"], meta["sigma"], meta["input_size"]
    su = ref.sbp_utils

    gen = su.SBPHeatmapGenerator([h, w], k, sigma)
    target = np.stack([gen(kp[b]) for b in range(kp.shape[0])])

    x = logits.clone().requires_grad_(True)
    loss = ref.SBPLoss()(x, torch.from_numpy(target))
    loss.backward()

    out = dict(
        inputs_sha=np.array(cases.digest(kp, logits, bbox)),
        template=gen.g, target=target,
        loss=loss.detach().numpy(), dlogits=x.grad.numpy(),
    )
    for thr in (0.25, 0.99):
        dec = su.DecodeSBP(list(in_size), thr, True)
        out[f"joints_pred_thr{thr}"] = np.stack([dec(logits[b:b + 1]).numpy() for b in range(logits.size(0))])
    dec_t = su.DecodeSBP(list(in_size), 0.99, False)
    tt = torch.from_numpy(target)
    out["joints_target_thr0.99"] = np.stack([dec_t(tt[b:b + 1]).numpy() for b in range(tt.size(0))])

    # update_state rows (COCO + PIS variants)
    for cls, tag in ((su.SBPmAPCOCO, "coco"), (ref.sbp_pis_utils.SBPmAPPIS, "pis")):
        m = object.__new__(cls)
        m.input_size = list(in_size)
        m.decoder = su.DecodeSBP(list(in_size), 0.25, True)
        m.result_list = []
        m.update_state({"bbox": bbox, "image_id": iid, "category_id": cid}, logits)
        out[f"rows_{tag}"] = np.array(json.dumps(m.result_list))
    np.savez_compressed(os.path.join(OUT, f"sbp_{name}.npz"), **out)
    print("sbp", name, "B", kp.shape[0], "loss", float(loss))


def golden_sbp_adversarial(ref):
    maps = cases.sbp_adversarial_maps()
    su = ref.sbp_utils
    out = dict(inputs_sha=np.array(cases.digest(maps)))
    for thr in (0.25, 0.5):
        dec = su.DecodeSBP([256, 192], thr, True)
        out[f"joints_pred_thr{thr}"] = np.stack([dec(maps[b:b + 1]).numpy() for b in range(maps.size(0))])
    dec = su.DecodeSBP([256, 192], 0.99, False)
    out["joints_raw_thr0.99"] = np.stack([dec(maps[b:b + 1]).numpy() for b in range(maps.size(0))])
    # near-ties: top logits 1-4 ulp apart -- the pick depends on the last bit of torch.sigmoid (CPU tensors here)
    near = cases.sbp_neartie_maps()
    out["neartie_sha"] = np.array(cases.digest(near))
    dec = su.DecodeSBP([256, 192], 0.25, True)
    out["joints_neartie_thr0.25"] = np.stack([dec(near[b:b + 1]).numpy() for b in range(near.size(0))])
    np.savez_compressed(os.path.join(OUT, "sbp_adversarial.npz"), **out)
    print("sbp adversarial ok")


def golden_sbp_config1(ref):
    """BASELINE.json configs[0] (B=32): scalars + checksums only (inputs are regenerated from seeds)."""
    from . import sbp_oracle as so
    kp, logits, bbox, iid, cid = so.make_config1_inputs(32)
    su = ref.sbp_utils
    gen = su.SBPHeatmapGenerator([64, 48], 17, 2)
    target = np.stack([gen(kp[b]) for b in range(32)])
    x = logits.clone().requires_grad_(True)
    loss = ref.SBPLoss()(x, torch.from_numpy(target))
    loss.backward()
    dec = su.DecodeSBP([256, 192], 0.25, True)
    joints = np.stack([dec(logits[b:b + 1]).numpy() for b in range(32)])
    np.savez_compressed(
        os.path.join(OUT, "sbp_config1.npz"),
        inputs_sha=np.array(cases.digest(kp, logits, bbox)),
        loss=loss.detach().numpy(),
        target_sum=np.float64(target.astype(np.float64).sum()),
        target_nnz=np.int64((target > 0).sum()),
        grad_sum=np.float64(x.grad.double().sum()), grad_abs_sum=np.float64(x.grad.double().abs().sum()),
        grad_slice=x.grad[:2].numpy(),
        joints=joints)
    print("sbp config1 loss", float(loss))


def golden_spm(ref, name, n_images):
    people, target_o, logits, meta = cases.spm_case(name, n_images)
    k, res, sigma, in_size = meta["k"], meta["res"], meta["sigma"], meta["input_size"]
    pu = ref.spm_utils
    hg = pu.SPMHeatmapGenerator(res, 1, sigma)
    mg = pu.SPMMaskGenerator(res, sigma)
    dg = pu.SPMDisplacementGenerator(res, k)
    targets = []
    for c, j in people:
        hm = hg(c)
        masks = mg(c)
        disp = dg(j, masks)
        targets.append(np.concatenate([hm, disp], axis=0))
    target = np.stack(targets)

    tt = torch.from_numpy(target)
    x = logits.clone().requires_grad_(True)
    loss = ref.SPMLoss()(x, tt)
    loss.backward()

    out = dict(inputs_sha=np.array(cases.digest(logits, *[a for p in people for a in p])),
               target=target, loss=loss.detach().numpy())
    if name == "small":
        out["dlogits"] = x.grad.numpy()
    else:
        out["grad_sum"] = np.float64(x.grad.double().sum())
        out["grad_abs_sum"] = np.float64(x.grad.double().abs().sum())
        out["grad_slice"] = x.grad[:1, :, 40:88, 40:88].numpy()

    for tag, src, pred, thr in (("pred", logits, True, 0.5), ("target", tt, False, 0.99)):
        dec = pu.DecodeSPM(in_size, sigma, thr, pred)
        for b in range(src.size(0)):
            r, kj = dec(src[b:b + 1].clone())
            out[f"roots_{tag}_{b}"] = r.numpy()
            out[f"kps_{tag}_{b}"] = kj.numpy()

    m = object.__new__(pu.SPMmAPCOCO)
    m.input_size = in_size
    m.conf_threshold = 0.5
    m.decoder = pu.DecodeSPM(in_size, sigma, 0.5, True)
    m.result_list = []
    rng = np.random.default_rng(99)
    widths = torch.from_numpy(rng.integers(300, 700, n_images))
    heights = torch.from_numpy(rng.integers(300, 700, n_images))
    m.update_state({"image_size": [widths, heights], "image_id": torch.arange(n_images) + 7,
                    "category_id": torch.ones(n_images, dtype=torch.int64)}, logits.clone())
    out["rows"] = np.array(json.dumps(m.result_list))
    out["image_w"] = widths.numpy()
    out["image_h"] = heights.numpy()
    np.savez_compressed(os.path.join(OUT, f"spm_{name}.npz"), **out)
    print("spm", name, "loss", float(loss), "rows", len(m.result_list))


def main():
    ref = load_reference()
    if ref is None:
        sys.exit("reference tree not found (set POSE_REF or run in the build container)")
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    for name in cases.SBP_SHAPES:
        golden_sbp(ref, name)
    golden_sbp_adversarial(ref)
    golden_sbp_config1(ref)
    golden_spm(ref, "small", 6)
    golden_spm(ref, "coco", 4)


if __name__ == "__main__":
    main()
