"""Property tests of the oracle against itself (CPU): loop forms vs vectorised forms, closed forms vs op-order forms,
render -> decode round trips, on hypothesis-generated shapes and keypoints (edge coordinates included)."""
import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from helpers import chained_field as _chained_field
from oracle import sbp_oracle as so
from oracle import spm_oracle as po

shapes = st.sampled_from([(3, 16, 12, 1), (2, 20, 28, 1.5), (4, 32, 24, 2), (1, 9, 7, 1), (2, 64, 48, 3)])
coords = st.one_of(st.floats(-3, 80, allow_nan=False), st.sampled_from([-1.0, -0.0, 0.0, 0.999999999, 11.5, 47.9, 63.9, 1e6]))


@settings(max_examples=40, deadline=None)
@given(shape=shapes, data=st.data())
def test_sbp_render_loop_equals_vectorised_and_roundtrips(shape, data):
    k, h, w, sigma = shape
    kp = np.array(data.draw(st.lists(st.tuples(coords, coords), min_size=2 * k, max_size=2 * k))).reshape(2, k, 2)
    vec = so.sbp_render(kp, h, w, sigma)
    loop = np.stack([so.sbp_render_loop(kp[b], h, w, sigma) for b in range(2)])
    assert vec.dtype == np.float32 and np.array_equal(vec, loop)
    if sigma != int(sigma):
        return          # half-integer template centre: no pixel reaches 1.0, the known-answer round trip does not apply
    j = so.sbp_decode(torch.from_numpy(vec), 4 * w, 0.99, False).numpy()
    vis = ~((kp[..., 0] < 0) | (kp[..., 1] < 0))
    cx = np.clip(np.trunc(np.where(vis, kp[..., 0], 0)), 0, w - 1)
    cy = np.clip(np.trunc(np.where(vis, kp[..., 1], 0)), 0, h - 1)
    assert np.array_equal(j[vis][:, 0], 4.0 * cx[vis]) and np.array_equal(j[vis][:, 1], 4.0 * cy[vis])
    assert np.all(j[~vis] == np.array([-4.0, -4.0, -1.0], dtype=np.float32))


@settings(max_examples=15, deadline=None)
@given(shape=shapes, seed=st.integers(0, 10_000), scale=st.sampled_from([0.5, 3.0, 8.0]))
def test_sbp_loss_closed_form_and_decode_forms_agree(shape, seed, scale):
    k, h, w, sigma = shape
    kp, logits, *_ = so.make_config1_inputs(3, k, h, w, seed=seed, torch_seed=seed)
    logits = logits * scale
    t = torch.from_numpy(so.sbp_render(kp, h, w, sigma))
    loss, grad = so.sbp_loss_and_grad(logits, t)
    l64, g64 = so.sbp_loss_closed_form_f64(logits, t)
    assert abs(float(loss) - float(l64)) <= 2e-6 * abs(float(l64))
    assert float((grad.double() - g64).abs().max()) <= 2e-6 * float(g64.abs().max()) + 1e-12
    for thr, pred in ((0.25, True), (0.9, True), (0.5, False)):
        a = so.sbp_decode(logits, 4 * w, thr, pred)
        b = torch.stack([so.sbp_decode_loop(logits[i:i + 1], 4 * w, thr, pred) for i in range(3)])
        assert torch.equal(a, b)


@settings(max_examples=10, deadline=None)
@given(seed=st.integers(0, 10_000), res=st.sampled_from([24, 32, 48]), k=st.integers(1, 4))
def test_spm_forms_agree(seed, res, k):
    people = po.make_config4_people(2, k=k, res=res, max_people=4, seed=seed)
    target = np.stack([po.spm_render(c, j, res, 1) for c, j in people])
    logits = po.spm_logits_from_target(target, seed=seed)
    tt = torch.from_numpy(target)
    loss, grad = po.spm_loss_and_grad(logits, tt)
    l64, g64 = po.spm_loss_closed_form_f64(logits, tt)
    assert abs(float(loss) - float(l64)) <= 2e-6 * abs(float(l64)) + 1e-12
    assert float((grad.double() - g64).abs().max()) <= 1e-5 * float(g64.abs().max()) + 1e-12
    # decoding the target itself finds every non-overlapping centre exactly
    for b, (c, j) in enumerate(people):
        roots, kps = po.spm_decode(tt[b:b + 1], res, 1, 0.99, False)
        found = {(int(r[0]), int(r[1])) for r in roots} if roots.dim() == 2 else set()
        assert found <= {(int(x), int(y)) for x, y in c[:, 0]}


def test_spm_chained_keypoints_restatement():
    """PARITY UNPINNED helper (hierarchical chaining is not in the reference): with every parent = -1 it IS the reference's single
    hop (bit for bit against the pinned restatement), and a field encoded joint-to-parent decodes to the skeleton it encodes."""
    people = po.make_config4_people(1, k=4, res=48, max_people=3, seed=11)
    target = torch.from_numpy(np.stack([po.spm_render(c, j, 48, 1) for c, j in people]))
    roots = po.spm_nms(target[0, 0:1], 0.99, 4.0)
    assert roots.shape[0] >= 1
    single = po.spm_keypoints(roots, target[0, 1:], 4.0)
    assert torch.equal(po.spm_keypoints_chained(roots, target[0, 1:], [-1] * 4, 4.0), single)
    r, disp, parents, joints = _chained_field()
    got = po.spm_keypoints_chained(r, disp, parents, 4.0)
    assert np.allclose(got[..., :2].numpy(), joints, atol=1e-3) and torch.all(got[..., 2] == 0.9)
    # single hop on the same field reads every displacement at the root: joints deeper than one hop come out wrong / absent
    flat = po.spm_keypoints(r, disp, 4.0)
    assert not np.allclose(flat[:, 2, :2].numpy(), joints[:, 2], atol=0.5)
    # a joint that decodes off the map is reported as the reference reports it (no clamp) but takes its descendants with it;
    # a cycle is rejected
    far = disp.clone()
    far[0] = 5.0
    gone = po.spm_keypoints_chained(r, far, parents, 4.0)
    assert torch.all(gone[:, 0, 0] > 64) and torch.all(gone[:, 1] == 0) and torch.all(gone[:, 2] == 0) and torch.all(gone[:, 3, 2] == 0.9)
    assert torch.all(po.spm_keypoints_chained(r, disp, [1, 0, -1, -1, 3], 4.0)[:, :2] == 0)


def test_flip_average_of_a_mirrored_copy_is_the_identity():
    """PARITY UNPINNED helper: mirroring + swapping a prediction and averaging it back returns the prediction itself
    ((h + h) * 0.5 is exact in fp32), so decode(flip-averaged) == decode(plain)."""
    import torch
    from oracle import sbp_oracle as so
    torch.manual_seed(0)
    x = torch.randn(3, 17, 64, 48)
    pairs = [[1, 2], [3, 4], [5, 6], [7, 8], [9, 10], [11, 12], [13, 14], [15, 16]]
    perm = list(range(17))
    for a, b in pairs:
        perm[a], perm[b] = b, a
    xf = x[:, perm].flip(-1)
    # (CPU torch.sigmoid is not bit-reproducible across memory positions -- vector body vs scalar tail -- hence 1e-6, not equality)
    from helpers import allclose, assert_joints
    assert allclose(so.sbp_flip_average(x, xf, pairs, True), torch.sigmoid(x), 1e-6)
    assert_joints(so.sbp_decode(so.sbp_flip_average(x, xf, pairs, True), 192, 0.25, pred=False), so.sbp_decode(x, 192, 0.25, True), 1e-6)
    h = torch.rand(2, 17, 8, 12)
    assert torch.equal(so.sbp_flip_average(h, h[:, perm].flip(-1), pairs, False), h)          # without the activation it is exact


def test_cpu_pipeline_serial_and_worker_processes_agree():
    """bench.py's CPU arm: the forked-worker form computes exactly what the single-process loop form does."""
    from oracle.cpu_pipeline import ReferencePipeline
    sample = so.make_config1_inputs(6, 17, 64, 48)
    one = ReferencePipeline(sample, 64, 48, 2, 256, 192, 0.25, procs=1)
    two = ReferencePipeline(sample, 64, 48, 2, 256, 192, 0.25, procs=2)
    try:
        a, b = one.run_pass(), two.run_pass()
    finally:
        two.close()
    assert a == b and a[1] == 6
    want = so.sbp_loss(sample[1], torch.from_numpy(so.sbp_render(sample[0], 64, 48, 2)))
    assert abs(a[0] - float(want)) <= 1e-6 * float(want)


def test_c_restatement_of_aten_cpu_sigmoid_matches_torch():
    """oracle/csrc/aten_sigmoid.c (Sleef expf_u10 + IEEE add / divide) IS torch.sigmoid on contiguous CPU tensors: 2^24 random
    bit patterns, every float in [-20, 20) at stride 16, and the saturation / underflow edges.  (`python -m
    oracle.check_aten_sigmoid` repeats it for all 2^32 inputs: 0 mismatches on torch 2.11, AVX2 and AVX512 dispatch.)"""
    import numpy as np
    import torch
    from oracle.build_native import aten_sigmoid
    rng = np.random.default_rng(0)
    rnd = rng.integers(0, 2 ** 32, 1 << 24, dtype=np.uint64).astype(np.uint32).view(np.float32)
    lo, hi = int(np.float32(1e-30).view(np.uint32)), int(np.float32(20).view(np.uint32))
    pos = np.arange(lo, hi, 16, dtype=np.uint32).view(np.float32)
    edge = np.array([0.0, -0.0, 16.6, 16.7, 17.0, 88.0, 89.0, 100.0, 104.0, 105.0, -87.0, -88.0, -88.8, -100.0, -104.0, -105.0, np.inf, -np.inf],
                    dtype=np.float32)
    from oracle.sbp_oracle import torch_sigmoid_vector_body
    x = np.ascontiguousarray(np.concatenate([rnd, pos, -pos, edge]))
    want = torch_sigmoid_vector_body(torch.from_numpy(x)).numpy()     # (chunk tails go through glibc expf in ATen: avoided)
    got = aten_sigmoid(x)
    both_nan = np.isnan(want) & np.isnan(got)
    assert np.array_equal(got.view(np.uint32)[~both_nan], want.view(np.uint32)[~both_nan])
