"""GPU: the deepim/operator_py mirror driven exactly like MXNet drives CustomOps
(create(op_type, **string_attrs) -> forward(is_train, req, in_data, out_data, aux) / backward(...)),
plus the Render_Py / RT_transform / gpu_flow call-compatible shims.  These mirror the reference's own
self-tests: transform3d.py:311-539 (forward < 1e-4 vs RT_transform, finite-difference gradient check,
thresh 5e-3), zoom_trans.py:106-154 (zoom / inverse-zoom round trip)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from oracle import oracle as O  # noqa: E402
from deepim_b200 import operator_py as ops  # noqa: E402
from deepim_b200 import synth  # noqa: E402
from deepim_b200.context import Context  # noqa: E402
from deepim_b200.RT_transform import RT_transform  # noqa: E402
from deepim_b200.gpu_flow import gpu_flow  # noqa: E402
from deepim_b200.render_py_multi import Render_Py  # noqa: E402

K = synth.K_LINEMOD
KSTR = "[" + " ".join("%r" % float(v) for v in K.reshape(-1)) + "]"
MEANS_ATTR = "[123.68 116.779 103.939]"  # cfg.network.PIXEL_MEANS.flatten() as the symbol passes it
DEV = torch.device("cuda", 0)
H, W = 480, 640


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture(scope="module")
def env():
    meshes = [synth.make_cube(), synth.make_blob()]
    ctx = Context(0, max_batch=4, max_classes=2, max_verts=6000, max_faces=11000)
    for i, m in enumerate(meshes):
        ctx.upload_mesh(i, m)
    ops.set_default_context(ctx)
    yield ctx, meshes
    ctx.close()


def _run(op, in_data, out_shapes, req=None):
    out = [torch.full(s, -7.0, device=DEV) for s in out_shapes]
    op.forward(False, req or ["write"] * len(out), in_data, out, [])
    return out


def test_test_graph_ops_chain_like_the_symbol(env):
    """ZoomMask -> ZoomImageWithFactor -> (net) -> ZoomTrans exactly as get_test_symbol_share wires them
    (deepIM_flownet.py:575-606,716-725), attrs as strings."""
    ctx, meshes = env
    B = 2
    obs, ini = synth.sample_pose_pairs(B, 91)
    cls = np.array([0, 1], np.int32)
    ren = [O.render(meshes[cls[b]], ini[b], K, means_rgb=synth.PIXEL_MEANS_RGB) for b in range(B)]
    mr = np.stack([r["mask"] for r in ren])[:, None]
    mo = np.stack([O.box_mask(r["bbox"], H, W) for r in ren])[:, None]
    img_r = np.stack([r["image"] for r in ren])
    img_o = np.stack([synth.transform_image(synth.composite_observed(
        O.render(meshes[cls[b]], obs[b], K)["bgr"], O.render(meshes[cls[b]], obs[b], K)["mask"], b)) for b in range(B)])
    pose32 = ini.astype(np.float32)

    zm = ops.create("ZoomMask", K=KSTR, height="480", width="640")
    zo, zg, zr, zf = _run(zm, [dev(mo), dev(mo), dev(mr), dev(pose32)], [(B, 1, H, W)] * 3 + [(B, 4)])
    ozo, _, ozr, ozf, obb = O.zoom_mask(mo, mo, mr, pose32, K)
    assert np.array_equal(zm.bbox.cpu().numpy(), obb)
    assert np.array_equal(zf.cpu().numpy(), ozf) and np.array_equal(zo.cpu().numpy(), ozo) and np.array_equal(zr.cpu().numpy(), ozr)

    zi = ops.create("ZoomImageWithFactor", height="480", width="640", pixel_means=MEANS_ATTR)
    zio, zir = _run(zi, [zf, dev(img_o), dev(img_r)], [(B, 3, H, W)] * 2)
    ozio, ozir = O.zoom_image_with_factor(ozf, img_o, img_r, synth.PIXEL_MEANS_RGB.astype(np.float32))
    assert np.array_equal(zio.cpu().numpy(), ozio) and np.array_equal(zir.cpu().numpy(), ozir)

    # req = "add" and "null" are honoured (MXNet CustomOp.assign semantics)
    acc = torch.ones(B, 3, H, W, device=DEV)
    untouched = torch.full((B, 3, H, W), 5.0, device=DEV)
    zi.forward(False, ["add", "null"], [zf, dev(img_o), dev(img_r)], [acc, untouched], [])
    assert np.array_equal(acc.cpu().numpy(), ozio + 1.0) and float(untouched.min()) == 5.0

    zt = ops.create("ZoomTrans", b_inv_zoom="True")
    t = np.array([[0.1, -0.2, 0.3], [0.01, 0.02, -0.03]], np.float32)
    (out,) = _run(zt, [zf, dev(t)], [(B, 3)])
    assert np.array_equal(out.cpu().numpy(), O.zoom_trans(ozf, t, True))
    # zoom_trans.py:136-154 round trip: zoom in then inverse zoom gives the input back
    zt_in = ops.create("ZoomTrans", b_inv_zoom="False")
    (z1,) = _run(zt_in, [zf, dev(t)], [(B, 3)])
    (z2,) = _run(zt, [zf, z1], [(B, 3)])
    np.testing.assert_allclose(z2.cpu().numpy(), t, rtol=2e-7)
    # backward: zero grad to zoom_factor, (un)scaled grad to trans (b_zoom_grad False -> pass-through)
    gz, gt = torch.full((B, 4), 9.0, device=DEV), torch.zeros(B, 3, device=DEV)
    zt.backward(["write", "write"], [dev(t)], [zf, dev(t)], [out], [gz, gt], [])
    assert float(gz.abs().max()) == 0.0 and np.array_equal(gt.cpu().numpy(), t)

    with pytest.raises(ValueError):  # empty observed mask: the reference dies in np.min; here a ValueError
        _run(zm, [dev(np.zeros_like(mo)), dev(np.zeros_like(mo)), dev(mr), dev(pose32)], [(B, 1, H, W)] * 3 + [(B, 4)])


def test_label_zoom_ops(env):
    ctx, meshes = env
    B = 2
    rng = np.random.default_rng(5)
    zf = np.array([[0.31, 0.31, 0.05, -0.1], [0.55, 0.55, -0.2, 0.15]], np.float32)
    flow = (rng.normal(size=(B, 2, H, W)) * 4).astype(np.float32)
    fw = (rng.uniform(size=(B, 1, H, W)) > 0.4).astype(np.float32)
    mask = (rng.uniform(size=(B, 1, H, W)) > 0.5).astype(np.float32) * 0.9
    depth = rng.uniform(0.5, 1.0, size=(B, 1, H, W)).astype(np.float32)
    zfl = ops.create("ZoomFlow", height="480", width="640", b_inv_zoom="False")
    o1, o2 = _run(zfl, [dev(zf), dev(flow), dev(fw)], [(B, 2, H, W), (B, 1, H, W)])
    e1, e2 = O.zoom_flow(zf, flow, fw, False)
    assert np.array_equal(o1.cpu().numpy(), e1) and np.array_equal(o2.cpu().numpy(), e2)
    izfl = ops.create("ZoomFlow", height="480", width="640", b_inv_zoom="True")
    assert izfl and len(ops.REGISTRY["ZoomFlow"](b_inv_zoom="True").list_outputs()) == 1
    (o3,) = _run(izfl, [dev(zf), dev(flow)], [(B, 2, H, W)])
    assert np.array_equal(o3.cpu().numpy(), O.zoom_flow(zf, flow, None, True)[0])
    for inv in ("False", "True"):
        zmf = ops.create("ZoomMaskWithFactor", height="480", width="640", b_inv_zoom=inv)
        (o4,) = _run(zmf, [dev(zf), dev(mask)], [(B, 1, H, W)])
        assert np.array_equal(o4.cpu().numpy(), O.zoom_mask_with_factor(zf, mask, inv == "True"))
    zd = ops.create("ZoomDepth", height="480", width="640")
    d1, d2 = _run(zd, [dev(zf), dev(depth), dev(depth)], [(B, 1, H, W)] * 2)
    assert np.array_equal(d1.cpu().numpy(), O.zoom_depth(zf, depth)) and torch.equal(d1, d2)


def test_transform3d_like_the_reference_selftest(env):
    """transform3d.py:311-539: forward vs RT_transform < 1e-4; finite-difference gradient check with
    thresh 5e-3 on rotation and translation."""
    ctx, meshes = env
    rng = np.random.default_rng(1)
    B, N = 2, 3000
    pts = (rng.normal(size=(B, 3, N)) * 0.05).astype(np.float32)
    w = rng.uniform(0.5, 1.5, size=(B, 3, N)).astype(np.float32)      # d loss / d output
    q = rng.normal(size=(B, 4)) * 0.1 + np.array([1.0, 0, 0, 0])
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    t = (rng.normal(size=(B, 3)) * 0.05).astype(np.float32)
    ps = np.zeros((B, 3, 4), np.float32)
    ps[:] = np.array([[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 1, 1]], np.float32)   # v_pose_src of the self-test
    op = ops.create("Transform3D", T_means="[0 0 0]", T_stds="[1 1 1]", rot_coord="CAMERA")
    (out,) = _run(op, [dev(pts), dev(q), dev(t), dev(ps)], [(B, 3, N)])
    for b in range(B):
        P = O.rt_transform(ps[b].astype(np.float64), q[b], t[b], (0, 0, 0), (1, 1, 1), "camera")
        assert np.abs(out[b].cpu().numpy() - (P[:, :3] @ pts[b] + P[:, 3:4])).max() < 1e-4
    gr = [torch.zeros(B, 3, N, device=DEV), torch.zeros(B, 4, device=DEV), torch.zeros(B, 3, device=DEV),
          torch.zeros(B, 3, 4, device=DEV)]
    op.backward(["write"] * 4, [dev(w)], [dev(pts), dev(q), dev(t), dev(ps)], [out], gr, [])
    assert float(gr[0].abs().max()) == 0.0 and float(gr[3].abs().max()) == 0.0

    def loss(qq, tt):  # float64 forward on the CPU (through the normalised quaternion, as RT_transform does)
        tot = 0.0
        for b in range(B):
            P = O.rt_transform(ps[b].astype(np.float64), qq[b], tt[b], (0, 0, 0), (1, 1, 1), "camera")
            tot += float(((P[:, :3] @ pts[b].astype(np.float64) + P[:, 3:4]) * w[b]).sum())
        return tot
    step = 1e-4
    for b in range(B):
        for k in range(3):
            tp, tm = t.astype(np.float64).copy(), t.astype(np.float64).copy()
            tp[b, k] += step; tm[b, k] -= step
            fd = (loss(q.astype(np.float64), tp) - loss(q.astype(np.float64), tm)) / (2 * step)
            assert abs(fd - float(gr[2][b, k])) < 5e-3 * max(1.0, abs(fd))
        for k in range(4):
            qp, qm = q.astype(np.float64).copy(), q.astype(np.float64).copy()
            qp[b, k] += step; qm[b, k] -= step
            fd = (loss(qp, t.astype(np.float64)) - loss(qm, t.astype(np.float64))) / (2 * step)
            assert abs(fd - float(gr[1][b, k])) < 5e-3 * max(1.0, abs(fd))


def test_flow_updater_and_shims(env):
    ctx, meshes = env
    B = 2
    obs, ini = synth.sample_pose_pairs(B, 93)
    d_src = np.stack([O.render(meshes[1], ini[b], K)["depth"] for b in range(B)])[:, None]
    d_tgt = np.stack([O.render(meshes[1], obs[b], K)["depth"] for b in range(B)])[:, None]
    fu = ops.create("FlowUpdater", K=KSTR, thresh="3e-3", batch_size=str(B), height="480", width="640")
    fl, va = _run(fu, [dev(d_src), dev(d_tgt), dev(ini.astype(np.float32)), dev(obs.astype(np.float32))],
                  [(B, 2, H, W), (B, 2, H, W)])      # flow_weights = the validity plane tiled to 2 channels (flow_updater.py:98-99)
    K64 = K.astype(np.float64)
    KT = np.zeros((B, 3, 4), np.float32)
    for b in range(B):
        p32s, p32t = ini[b].astype(np.float32).astype(np.float64), obs[b].astype(np.float32).astype(np.float64)
        R = p32t[:, :3] @ p32s[:, :3].T
        KT[b] = (K64 @ np.hstack([R, (p32t[:, 3] - R @ p32s[:, 3])[:, None]])).astype(np.float32)
    Kinv = np.linalg.inv(K64).astype(np.float32)
    ofl, ova = O.flow(d_src, d_tgt, KT, Kinv)
    assert ova.sum() > 1000
    assert np.array_equal(va.cpu().numpy(), np.tile(ova, (1, 2, 1, 1))) and np.array_equal(fl.cpu().numpy(), ofl)
    # gpu_flow shim (lib/flow_c/gpu_flow.pyx signature): numpy in / numpy out
    f2, v2 = gpu_flow(d_src, d_tgt, KT, Kinv, 0)
    assert np.array_equal(f2, ofl) and np.array_equal(v2, ova)
    # RT_transform shim (lib/pair_matching/RT_transform.py:127 signature)
    q, t = np.array([0.99, 0.02, -0.03, 0.01]), np.array([0.01, -0.02, 0.03])
    np.testing.assert_allclose(RT_transform(ini[0], q, t, np.zeros(3), np.ones(3), "CAMERA"),
                               O.rt_transform(ini[0], q.astype(np.float32), t.astype(np.float32)), atol=1e-12)
    # Render_Py shim (render_py_multi.py:101 signature): BGR float image in [0,255] + metric depth
    rm = Render_Py(meshes, ["cube", "blob"], K, 640, 480, 0.25, 6.0, ctx=ctx)
    bgr, depth = rm.render(1, obs[0][:, :3], obs[0][:, 3], r_type="mat", K=K)
    r = O.render(meshes[1], obs[0], K, trunc_u8=False)
    assert bgr.shape == (480, 640, 3) and np.array_equal(bgr, r["bgr"]) and np.array_equal(depth, r["depth"])


def test_zoom_image_and_group_picker(env):
    """ZoomImage (zoom_image.py:26-107, the INPUT_MASK: False front end) and GroupPicker (group_picker.py:22-56,
    whose reference self-test checks forward against slicing and backward against scatter)."""
    ctx, meshes = env
    B = 2
    obs, ini = synth.sample_pose_pairs(B, 97)
    img_o = np.stack([O.render(meshes[b % 2], obs[b], K, means_rgb=synth.PIXEL_MEANS_RGB)["image"] for b in range(B)])
    img_r = np.stack([O.render(meshes[b % 2], ini[b], K, means_rgb=synth.PIXEL_MEANS_RGB)["image"] for b in range(B)])
    pose32 = ini.astype(np.float32)
    zi = ops.create("ZoomImage", K=KSTR, height="480", width="640", pixel_means=MEANS_ATTR)
    zo, zr, zf = _run(zi, [dev(img_o), dev(img_r), dev(pose32)], [(B, 3, H, W)] * 2 + [(B, 4)])
    eo, er, ezf, ebb = O.zoom_image(img_o, img_r, pose32, K, synth.PIXEL_MEANS_RGB.astype(np.float32))
    assert np.array_equal(zi.bbox.cpu().numpy(), ebb) and (ebb[:, 1] > ebb[:, 0]).all()
    assert np.array_equal(zf.cpu().numpy(), ezf)
    assert np.array_equal(zo.cpu().numpy(), eo) and np.array_equal(zr.cpu().numpy(), er)
    # the image-derived boxes coincide with the mask-derived ones of ZoomMask on clean renders
    mo = np.stack([O.render(meshes[b % 2], obs[b], K)["mask"] for b in range(B)])[:, None]
    mr = np.stack([O.render(meshes[b % 2], ini[b], K)["mask"] for b in range(B)])[:, None]
    assert np.array_equal(O.zoom_mask(mo, mo, mr, pose32, K)[4], ebb)

    rng = np.random.default_rng(3)
    G, cg = 13, 4
    x = rng.normal(size=(B, G * cg)).astype(np.float32)
    gi = np.array([[5.0], [12.0]], np.float32)
    gp = ops.create("GroupPicker", group_num=str(G))
    (picked,) = _run(gp, [dev(x), dev(gi)], [(B, cg)])
    want = np.stack([x[b, int(gi[b, 0]) * cg:(int(gi[b, 0]) + 1) * cg] for b in range(B)])
    assert np.array_equal(picked.cpu().numpy(), want)
    og = rng.normal(size=(B, cg)).astype(np.float32)
    gr = [torch.full((B, G * cg), 3.0, device=DEV), torch.full((B, 1), 3.0, device=DEV)]
    gp.backward(["write", "write"], [dev(og)], [dev(x), dev(gi)], [picked], gr, [])
    wantg = np.zeros((B, G * cg), np.float32)
    for b in range(B):
        wantg[b, int(gi[b, 0]) * cg:(int(gi[b, 0]) + 1) * cg] = og[b]
    assert np.array_equal(gr[0].cpu().numpy(), wantg) and float(gr[1].abs().max()) == 0.0
    # 4-D input (feature maps), as the per-class regressor graphs use it
    x4 = rng.normal(size=(B, 2 * 3, 5, 7)).astype(np.float32)
    (p4,) = _run(ops.create("GroupPicker", group_num="2"), [dev(x4), dev(np.array([1.0, 0.0], np.float32))], [(B, 3, 5, 7)])
    assert np.array_equal(p4.cpu().numpy(), np.stack([x4[0, 3:6], x4[1, 0:3]]))


def test_pose_error_add_adi_against_reference_goldens(env):
    """Device ADD / ADI (dim_pose_error) against the fixtures generated from the live lib/utils/pose_error.py
    (tests/golden/ref_pose_error.npz), and the evaluate_pose_add bookkeeping (LM6D_REFINE.py:372-512)."""
    import os
    from deepim_b200 import pose_eval
    ctx, meshes = env
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_pose_error.npz"))
    pe = np.hstack([d["R_est"], d["t_est"].reshape(3, 1)])[None]   # the fixture holds one pose pair
    pg = np.hstack([d["R_gt"], d["t_gt"].reshape(3, 1)])[None]
    add = pose_eval.pose_errors(ctx, pe, pg, d["pts"], False).cpu().numpy()
    adi = pose_eval.pose_errors(ctx, pe, pg, d["pts"], True).cpu().numpy()
    assert add.shape == (1,) and abs(add[0] - float(d["add"])) < 1e-14 and abs(adi[0] - float(d["adi"])) < 1e-14
    assert abs(pose_eval.add(ctx, d["R_est"], d["t_est"], d["R_gt"], d["t_gt"], d["pts"]) - float(d["add"])) < 1e-14
    assert abs(pose_eval.adi(ctx, d["R_est"], d["t_est"], d["R_gt"], d["t_gt"], d["pts"]) - float(d["adi"])) < 1e-14
    # a larger cloud than one tile, ADI vs the oracle's cKDTree restatement
    pts = meshes[1].verts.astype(np.float32)
    obs, ini = synth.sample_pose_pairs(6, 5)
    e = pose_eval.pose_errors(ctx, ini, obs, pts, True).cpu().numpy()
    for i in range(6):
        assert abs(e[i] - O.adi_metric(ini[i][:, :3], ini[i][:, 3], obs[i][:, :3], obs[i][:, 3], pts)) < 1e-12
    # evaluator: two classes (second one scored with ADI), two "iterations"
    cls = np.array([0, 1, 0, 1, 1, 0])
    est = np.stack([ini, 0.5 * (ini + obs)])          # iteration 2 = closer to the target (not a rotation; irrelevant here)
    ptsc = [meshes[0].verts.astype(np.float32), pts]
    diam = [0.17, 0.10]
    res = pose_eval.evaluate_pose_add(ctx, est, obs, cls, ptsc, diam, [False, True])
    for c in (0, 1):
        sel = cls == c
        for it in range(2):
            f = O.adi_metric if c == 1 else O.add_metric
            err = np.array([f(est[it, i][:, :3], est[it, i][:, 3], obs[i][:, :3], obs[i][:, 3], ptsc[c]) for i in np.nonzero(sel)[0]])
            np.testing.assert_allclose(res["classes"][c]["errors"][it], err, rtol=1e-10)
            assert res["classes"][c]["0.10"][it] == 100.0 * (err < np.float32(0.10 * diam[c])).sum() / sel.sum()
    assert res["mean"]["auc"][1] >= res["mean"]["auc"][0]
    # Simpson rule of the AUC = scipy.integrate.simps(even='avg') of the reference's era: exact for odd sample counts,
    # first/last-trapezoid average for the 1000-sample threshold grid
    y = np.linspace(0, 1, 1001) ** 2
    assert abs(pose_eval.simpson(y, 1e-3) - 1.0 / 3.0) < 1e-12


def test_lit_renderer_bit_exact(env):
    """dim_render_lit (Lambert shading of render_py_light_modelnet_multi.py:36-79) against the CPU checker: colours
    (8-bit quantised), depth, mask and bbox bit-exact; brightness_ratio = 0 falls back to the unlit colours."""
    from deepim_b200.render_py_light import Render_Py_Light_ModelNet_Multi
    ctx, meshes = env
    normals = [synth.vertex_normals(m) for m in meshes]
    for i, n in enumerate(normals):
        ctx.upload_normals(i, n)
    B = 2
    obs, ini = synth.sample_pose_pairs(B, 99)
    cls = np.array([0, 1], np.int32)
    rng = np.random.default_rng(4)
    lp = (rng.normal(size=(B, 3)) * np.array([0.5, 0.5, 0.3]) + np.array([0, 0, 0.2])).astype(np.float32)   # around the camera (GL frame)
    li = rng.uniform(0.8, 1.3, size=(B, 3)).astype(np.float32)
    out = ctx.render_lit(dev(cls).int(), dev(obs.astype(np.float32)), K, dev(lp), dev(li), 0.7,
                         pixel_means_rgb=synth.PIXEL_MEANS_RGB, want=("bgr", "depth", "image", "mask"))
    for b in range(B):
        r = O.render_lit(meshes[cls[b]], normals[cls[b]], obs[b], K, lp[b], li[b], 0.7, means_rgb=synth.PIXEL_MEANS_RGB)
        assert np.array_equal(out["bgr"][b].cpu().numpy(), r["bgr"]) and r["bgr"].max() > 50
        assert np.array_equal(out["depth"][b, 0].cpu().numpy(), r["depth"])
        assert np.array_equal(out["mask"][b, 0].cpu().numpy(), r["mask"])
        assert np.array_equal(out["image"][b].cpu().numpy(), r["image"])
        assert np.array_equal(out["bbox"][b].cpu().numpy(), r["bbox"])
        lit_vals = r["bgr"][r["mask"] > 0]
        assert len(np.unique(lit_vals)) > 20 and np.array_equal(lit_vals, np.round(lit_vals))      # shaded, 8-bit levels
    # ratio 0 and unit intensity: colour = round(texel/255 * 255) = the unlit render
    one = np.ones((B, 3), np.float32)
    flat = ctx.render_lit(dev(cls).int(), dev(obs.astype(np.float32)), K, dev(lp), dev(one), 0.0, want=("bgr",))
    unlit = ctx.render(dev(cls).int(), dev(obs.astype(np.float32)), K, trunc_u8=True, want=("bgr",))
    assert np.array_equal(flat["bgr"].cpu().numpy(), unlit["bgr"].cpu().numpy())
    # call-compatible shim
    class M:  # mesh with normals
        pass
    ms = []
    for m, n in zip(meshes, normals):
        mm = synth.Mesh(m.verts, m.uvs, m.faces, m.tex)
        mm.normals = n
        ms.append(mm)
    rm = Render_Py_Light_ModelNet_Multi(ms, K, brightness_ratios=[0.7], ctx=ctx)
    bgr, depth = rm.render(1, obs[1][:, :3], obs[1][:, 3], lp[1], li[1], brightness_k=0, r_type="mat")
    r = O.render_lit(meshes[1], normals[1], obs[1], K, lp[1], li[1], 0.7)
    assert bgr.dtype == np.uint8 and np.array_equal(bgr.astype(np.float32), r["bgr"]) and np.array_equal(depth, r["depth"])
