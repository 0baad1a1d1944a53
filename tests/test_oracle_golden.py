"""Pin the CPU oracle (oracle/render_oracle.py) against outputs of the unmodified reference.

The fixtures under tests/golden/ were produced by tests/golden/make_golden.py, which imports and
runs /root/reference itself.  Integer work (sample_pdf indices) must be bit-exact; floating point
is compared at 2e-6 (same torch ops, possibly a different BLAS blocking on another host).
"""
import numpy as np
import torch

from oracle import render_oracle as O

T = torch.from_numpy


def close(a, b, tol=2e-6):
    a = a.detach().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    err = float(np.max(np.abs(a.astype(np.float64) - np.asarray(b, np.float64)))) if a.size else 0.0
    assert err <= tol, f"max-abs {err:.3e} > {tol:.1e}"


def test_tables(golden):
    g = golden("tables")
    assert np.array_equal(torch.linspace(0., 1., 64).numpy(), g["t_vals64"])
    assert np.array_equal(torch.linspace(0., 1., 128).numpy(), g["u128"])


def test_get_rays_and_embed(golden):
    g = golden("rays_embed")
    ro, rd = O.get_rays(6, 5, 9.5, T(g["c2w"]))
    close(ro, g["rays_o"], 0); close(rd, g["rays_d"], 1e-7)
    ro, rd = O.get_rays(6, 5, 9.5, T(g["c2w"]), 2.25, 3.5)
    close(rd, g["rays_d_c"], 1e-7)
    cam = O.synthetic_camera()
    ro, rd = O.get_rays(cam["H"], cam["W"], cam["focal"], cam["c2w"], cam["cx"], cam["cy"])
    close(rd.reshape(-1, 3)[g["frame_pick"]], g["frame_rays_d"], 0)
    close(rd.double().sum((0, 1)), g["frame_rays_d_sum"], 1e-9)
    close(O.positional_encoding(T(g["x"]), 10), g["embed10"], 0)
    close(O.positional_encoding(T(g["x"]), 4), g["embed4"], 0)


def test_face_nerf(golden):
    g = golden("face_nerf")
    out = O.face_nerf_forward(O.init_face_nerf(int(g["seed_head"])), T(g["x"]), T(g["aud"]), T(g["expr"]), T(g["latent"]))
    close(out, g["out_head"])
    out = O.face_nerf_forward(O.init_face_nerf(int(g["seed_torso"]), 106, 0, 0), T(g["x"]), T(g["aud_torso"]))
    close(out, g["out_torso"])


def test_raw2outputs(golden):
    g = golden("raw2outputs")
    for tag in ("s64", "s192", "s7"):
        a = [T(g[f"{tag}_{k}"]) for k in ("raw", "z", "d", "bc")]
        rgb, disp, acc, w, depth, fg = O.raw2outputs(*a, with_fg=True)
        close(rgb, g[f"{tag}_rgb"]); close(acc, g[f"{tag}_acc"]); close(w, g[f"{tag}_w"])
        close(depth, g[f"{tag}_depth"]); close(fg, g[f"{tag}_rgb_fg"]); close(rgb, g[f"{tag}_torso_rgb"])
        close(disp, g[f"{tag}_disp"], 1e-5)
        close(O.raw2outputs(*a, white_bkgd=True)[0], g[f"{tag}_rgb_white"])
        rn = O.raw2outputs(*a, noise=T(g[f"{tag}_noise"]))
        close(rn[0], g[f"{tag}_rgb_noise"]); close(rn[3], g[f"{tag}_w_noise"])


def test_raw2outputs_backward(golden):
    g = golden("raw2outputs")
    raw = T(g["s64_raw"][:48]).clone().requires_grad_(True)
    t = O.raw2outputs(raw, T(g["s64_z"][:48]), T(g["s64_d"][:48]), T(g["s64_bc"][:48]), with_fg=True)
    ((t[0] * T(g["bwd_g_rgb"])).sum() + (t[5] * T(g["bwd_g_fg"])).sum() + (t[3] * T(g["bwd_g_w"])).sum()
     + (t[2] * T(g["bwd_g_acc"])).sum() + (t[4] * T(g["bwd_g_depth"])).sum() + (t[1] * T(g["bwd_g_disp"])).sum()).backward()
    close(raw.grad, g["bwd_d_raw"], 1e-5)


def test_rowsum_model_matches_torch():
    torch.manual_seed(3)
    for n in (62, 30, 126, 14, 64, 8, 3):   # n in 5..7 takes another ATen path; N_samples-2 >= 8 in practice
        x = torch.rand(20000, n) ** 3 + 1e-5
        assert np.array_equal(O.torch_cpu_rowsum_f32(x.numpy()), torch.sum(x, -1).numpy()) or \
            torch.backends.cpu.get_cpu_capability() not in ("AVX512", "AVX2"), n


def test_sample_pdf_exact_bitwise(golden):
    g = golden("sample_pdf")
    s, inds, cdf = O.sample_pdf_exact(g["bins"], g["weights"], torch.linspace(0., 1., 128).numpy())
    assert np.array_equal(cdf, g["cdf"])
    assert np.array_equal(inds, g["inds_det"].astype(np.int64))
    assert np.array_equal(s, g["samples_det"])
    s, inds, _ = O.sample_pdf_exact(g["bins"], g["weights"], g["u_rnd"])
    assert np.array_equal(inds, g["inds_rnd"].astype(np.int64))
    assert np.array_equal(s, g["samples_rnd"])


def test_sample_pdf_torch_form(golden):
    g = golden("sample_pdf")
    u = torch.linspace(0., 1., 128).expand(g["bins"].shape[0], 128)
    s, inds = O.sample_pdf(T(g["bins"]), T(g["weights"]), u)
    close(s, g["samples_det"], 1e-5)   # an index flip moves a sample by ~1 ulp of the bin edge only


def _presets(g):
    rays, aud, expr, lat = T(g["rays"]), T(g["aud"]), T(g["expr"]), T(g["latent"])
    out = {}
    for tag in ("init", "dense"):
        c, f = O.init_face_nerf(1), O.init_face_nerf(2)
        c["alpha_linear.weight"], c["alpha_linear.bias"] = T(g[f"{tag}_alpha_w_c"]), T(g[f"{tag}_alpha_b_c"])
        f["alpha_linear.weight"], f["alpha_linear.bias"] = T(g[f"{tag}_alpha_w_f"]), T(g[f"{tag}_alpha_b_f"])
        out[tag] = (c, f)
    return rays, aud, expr, lat, out


def test_synthetic_batch_is_reproducible(golden):
    g = golden("render_3072")
    b = O.synthetic_train_batch(0)
    assert np.array_equal(b["pixel_index"].numpy(), g["pixel_index"])
    assert np.array_equal(b["rays"].numpy(), g["rays"])
    assert np.array_equal(b["bc_rgb"].numpy(), g["bc_rgb"])
    assert np.array_equal(b["aud"].numpy(), g["aud"])


def test_render_rays_stages_and_outputs(golden):
    g, st = golden("render_3072"), golden("render_stages")
    rays, aud, expr, lat, presets = _presets(g)
    sub = T(st["sub"])
    for tag, (c, f) in presets.items():
        with torch.no_grad():
            r = O.render_rays(rays[sub], T(g["bc_rgb"])[sub], c, f, aud, expr, lat, retraw=True)
        close(r["raw"], st[f"{tag}_raw1"], 2e-4)          # sigma is scaled by ~x100 in the dense preset
        close(r["_z_vals"], st[f"{tag}_z1"], 1e-6)
        close(r["_weights"], st[f"{tag}_w1"], 2e-5)
        for k in ("rgb_map", "acc_map", "rgb0", "acc0", "z_std", "last_weight"):
            close(r[k], g[f"{tag}_{k}"][st["sub"]], 2e-5)
        close(r["disp_map"], g[f"{tag}_disp_map"][st["sub"]], 1e-4)
    # the dense preset is the non-trivial one: background weight must be spread out, not 0 or 1
    lw = g["dense_last_weight"]
    assert 0.02 < np.quantile(lw, 0.1) and np.quantile(lw, 0.9) < 0.98 and g["dense_z_std"].std() > 1e-3


def test_render_perturb(golden):
    g, p = golden("render_3072"), golden("render_perturb")
    rays, aud, expr, lat, presets = _presets(g)
    c, f = presets["dense"]
    idx = T(p["idx"])
    with torch.no_grad():
        r = O.render_rays(rays[idx], T(g["bc_rgb"])[idx], c, f, aud, expr, lat,
                          t_rand=T(p["t_rand"]), u_rand=T(p["u_rand"]))
    close(r["_z_vals"], p["z1"], 1e-6)
    for k in ("rgb_map", "acc_map", "rgb0", "z_std", "last_weight"):
        close(r[k], p[k], 2e-5)


def test_train_step_grads(golden):
    g, tr = golden("render_3072"), golden("train_step")
    rays, aud, expr, lat, presets = _presets(g)
    c, f = presets["dense"]
    c = {k: v.clone().requires_grad_(True) for k, v in c.items()}
    f = {k: v.clone().requires_grad_(True) for k, v in f.items()}
    aud, expr, lat = (t.clone().requires_grad_(True) for t in (aud, expr, lat))
    idx = T(tr["idx"])
    r = O.render_rays(rays[idx], T(g["bc_rgb"])[idx], c, f, aud, expr, lat)
    loss = O.head_loss(r["rgb_map"], r["rgb0"], T(g["target"])[idx], lat)
    loss.backward()
    close(loss, tr["loss"], 1e-6)
    close(aud.grad, tr["d_aud"], 1e-6); close(expr.grad, tr["d_expr"], 1e-6); close(lat.grad, tr["d_latent"], 1e-6)
    for key in tr:
        kind, _, name = key.partition(":")
        if kind not in ("norm", "grad", "samp"):
            continue
        net, _, pname = name.partition(".")
        gr = (c if net == "c" else f)[pname].grad
        if kind == "norm":
            assert abs(float(gr.double().norm()) - float(tr[key])) <= 1e-4 * max(1e-6, float(tr[key])) + 1e-9, key
        elif kind == "grad":
            close(gr, tr[key], 1e-5 * max(1.0, float(np.abs(tr[key]).max())))
        else:
            close(gr.reshape(-1)[::97], tr[key], 1e-5 * max(1.0, float(np.abs(tr[key]).max())))
    assert c["feature_linear.weight"].grad is None      # constructed but never applied (face_nerf.py:34)


def test_pose_to_euler_trans(golden):
    g = golden("torso_misc")
    close(O.pose_to_euler_trans(T(g["poses"])), g["euler_trans"], 1e-6)


def test_dense_preset_rounding_floor(golden):
    """How much the REFERENCE algorithm itself moves when only the MLP's rounding changes (fp32 -> fp64 FaceNeRF,
    everything else identical): the yardstick for the fp32-mode GPU tolerances on the dense preset.  The inverse CDF
    divides by bin masses ~1e-4 and gamma_10 multiplies depth changes by 2^9, so last-bit changes of the coarse
    weights become ~1e-4..1e-3 in the fine outputs."""
    g = golden("render_3072")
    rays, aud, expr, lat, presets = _presets(g)
    c, f = presets["dense"]
    sub = torch.arange(0, 3072, 12)
    orig = O.run_network

    def run_network_fp64(sd, pts, viewdirs, a, e, l, netchunk=65536):
        sd64 = {k: v.double() for k, v in sd.items()}
        return orig(sd64, pts.double(), viewdirs.double(), a.double(), e.double(), l.double(), netchunk).float()

    with torch.no_grad():
        r32 = O.render_rays(rays[sub], T(g["bc_rgb"])[sub], c, f, aud, expr, lat)
        O.run_network = run_network_fp64
        try:
            r64 = O.render_rays(rays[sub], T(g["bc_rgb"])[sub], c, f, aud, expr, lat)
        finally:
            O.run_network = orig
    d_rgb = float((r32["rgb_map"] - r64["rgb_map"]).abs().max())
    d_lw = float((r32["last_weight"] - r64["last_weight"]).abs().max())
    d_rgb0 = float((r32["rgb0"] - r64["rgb0"]).abs().max())
    print(f"reference fp32-vs-fp64 MLP on 256 dense-preset rays: rgb {d_rgb:.2e}, last_weight {d_lw:.2e}, rgb0 {d_rgb0:.2e}")
    assert d_rgb0 < 5e-6                     # the coarse pass is quiet ...
    assert 1e-5 < d_rgb < 1e-3 and 1e-5 < d_lw < 2e-3      # ... the fine pass amplifies rounding by 2-3 orders of magnitude


def test_philox_restatement_known_answers():
    """oracle/philox_ref.py against the Random123 known-answer vectors of philox4x32-10 (kat_vectors: counter, key -> output)."""
    from oracle import philox_ref as P
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, want in kats:
        got = tuple(int(x) for x in P.philox4x32_10(*c, *k))
        assert got == want, (c, k, [hex(g) for g in got])
    u = P.draws_u01(1234, 5, 1, 4097)
    assert u.dtype.name == "float32" and u.shape == (4097,) and float(u.min()) >= 0.0 and float(u.max()) < 1.0
    assert abs(float(u.mean()) - 0.5) < 0.02
