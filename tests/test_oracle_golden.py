"""The oracle (oracle/stylex_oracle.py) against the committed golden vectors that were produced by
running the UNMODIFIED reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import stylex_b200
from stylex_b200 import synthetic, classifiers
from oracle import stylex_oracle as O
from helpers import state_from_npz, tiny_cnn_from

torch.set_grad_enabled(False)


def test_generator_small_matches_reference(golden):
    z = golden("gen_small.npz")
    sd = state_from_npz(z, "G.")
    w = torch.from_numpy(z["latents"])
    noise = torch.from_numpy(z["noise"])
    L = len(O.generator_layout(sd))
    img, sc = O.generator_forward(sd, O.styles_def_to_tensor([(w, L)]), noise, get_style_coords=True)
    # same torch ops in the same order as the reference: tolerance only covers CPU-kernel differences
    assert np.abs(img.numpy() - z["image"]).max() <= 2e-6
    assert np.abs(sc.numpy() - z["style_coords"]).max() <= 2e-6
    img2, sc2 = O.generator_forward(sd, torch.from_numpy(z["styles_per_layer"]), noise, get_style_coords=True)
    assert np.abs(img2.numpy() - z["image_per_layer"]).max() <= 2e-6
    assert np.abs(sc2.numpy() - z["style_coords_per_layer"]).max() <= 2e-6


@pytest.mark.parametrize("size", [64, 256])
def test_generator_full_shapes_match_reference(golden, size):
    z = golden("gen_full.npz")
    sd = synthetic.make_generator_state(size, seed=42)
    n = z[f"image_{size}"].shape[0]
    w = synthetic.make_latents(n, 42)
    noise = synthetic.make_noise(size, 42)
    fp = np.array([float(sd["blocks.0.conv1.weight"].double().sum()),
                   float(sd[f"blocks.{len(O.generator_layout(sd)) - 1}.conv2.weight"].double().abs().sum()),
                   float(w.double().sum()), float(noise.double().sum())])
    assert np.allclose(fp, z[f"fp_{size}"], rtol=1e-9, atol=1e-6), "torch CPU RNG drifted: seeded inputs differ"
    L = len(O.generator_layout(sd))
    img, sc = O.generator_forward(sd, O.styles_def_to_tensor([(w, L)]), noise, get_style_coords=True)
    assert np.abs(img.numpy() - z[f"image_{size}"]).max() <= 5e-6
    assert np.abs(sc.numpy() - z[f"style_coords_{size}"]).max() <= 5e-6
    assert sc.shape[1] == {64: 2464, 256: 4512}[size]


@pytest.mark.parametrize("kind", ["mobilenet", "resnet"])
def test_attfind_sweep_matches_verbatim_notebook(golden, kind):
    z = golden("attfind_small.npz")
    sd = state_from_npz(z, "G.")
    model = tiny_cnn_from(z, f"{kind}.clf.")
    clf = classifiers.make_classifier(kind, model, int(z["image_size"]))
    latents = torch.from_numpy(z[f"{kind}.latents"])
    noise = torch.from_numpy(z["noise"])
    res = O.attfind_sweep(sd, clf.classify_images, latents, noise, shift_size=1.0)
    assert np.abs(res["style_coordinates"].numpy() - z[f"{kind}.style_coordinates"]).max() <= 2e-6
    assert np.abs(res["base_prob"].numpy() - z[f"{kind}.base_prob"]).max() <= 1e-5
    assert np.abs(res["minima"].numpy() - z[f"{kind}.minima"][0]).max() <= 2e-6
    assert np.abs(res["maxima"].numpy() - z[f"{kind}.maxima"][0]).max() <= 2e-6
    # the notebook patches biases in place (quirk Q2: fp32 residue accumulates); the oracle is drift-free
    err = np.abs(res["style_change"].numpy() - z[f"{kind}.style_change"]).max()
    assert err <= 2e-5, err
    picks, merged, _ = O.attfind_select(res["style_change"].numpy(), res["base_prob"].numpy(), 5, 0.5)
    assert [tuple(p) for p in z[f"{kind}.picks0"]] == picks[0]
    assert [tuple(p) for p in z[f"{kind}.picks1"]] == picks[1]
    assert [tuple(p) for p in z[f"{kind}.merged"]] == merged


def test_selection_golden_picks_from_golden_effects(golden):
    """selection alone, on the reference's own effects (no float noise in between): must be exact."""
    z = golden("attfind_small.npz")
    for kind in ("mobilenet", "resnet"):
        picks, merged, _ = O.attfind_select(z[f"{kind}.style_change"], z[f"{kind}.base_prob"], 5, 0.5)
        assert [tuple(p) for p in z[f"{kind}.picks0"]] == picks[0]
        assert [tuple(p) for p in z[f"{kind}.picks1"]] == picks[1]
        assert [tuple(p) for p in z[f"{kind}.merged"]] == merged


def test_find_significant_styles_cases(golden):
    z = golden("select_cases.npz")
    for name in z["cases"]:
        eff = z[f"{name}.effects"]
        mie, k = z[f"{name}.params"]
        for c in (0, 1):
            got = O.find_significant_styles(eff.copy(), int(k), c, max_image_effect=float(mie))
            assert got == [tuple(p) for p in z[f"{name}.picks{c}"]], (name, c)


@pytest.mark.parametrize("kind", ["mobilenet", "resnet"])
def test_counterfactual_rendering_matches_verbatim_notebook(golden, kind):
    """oracle restatement of NB cells 17-19 vs the executed reference (tests/golden/make_golden.py::counterfactual_small)."""
    z = golden("attfind_small.npz")
    c = golden("counterfactual_small.npz")
    sd = state_from_npz(z, "G.")
    clf = classifiers.make_classifier(kind, tiny_cnn_from(z, f"{kind}.clf."), int(z["image_size"]))
    latents = torch.from_numpy(z[f"{kind}.latents"])
    noise = torch.from_numpy(z["noise"])
    smin, smax = z[f"{kind}.minima"][0], z[f"{kind}.maxima"][0]
    for i, (n, sindex, d, cls, shift) in enumerate(c[f"{kind}.cases"]):
        n, sindex, d, cls = int(n), int(sindex), int(d), int(cls)
        img, prob = O.generate_change_image_given_dlatent(sd, clf.classify_images, latents[n:n + 1], cls, sindex,
                                                          smin[sindex], smax[sindex], d, float(shift), noise)
        assert np.abs(img[0].numpy() - c[f"{kind}.images"][i]).max() <= 5e-6
        assert abs(float(prob[0]) - c[f"{kind}.change_prob"][i]) <= 1e-5
        panel, cp, bp = O.generate_images_given_dlatent(sd, clf.classify_images, latents[n:n + 1], cls, sindex, smin[sindex],
                                                        smax[sindex], d, noise, shift_size=float(shift))
        assert abs(bp - c[f"{kind}.base_prob"][i]) <= 1e-5 and abs(cp - c[f"{kind}.change_prob"][i]) <= 1e-5
        diff = np.abs(panel.astype(np.int32) - c[f"{kind}.panels"][i].astype(np.int32))
        assert diff.max() <= 1 and (diff > 0).mean() <= 0.01      # uint8 truncation at an integer boundary


def test_sindex_mapping():
    pairs = synthetic.generator_pairs(64)
    assert O.sindex_to_block_idx_and_index(pairs, 0) == (0, 0)
    assert O.sindex_to_block_idx_and_index(pairs, 1023) == (0, 1023)
    assert O.sindex_to_block_idx_and_index(pairs, 1024) == (1, 0)
    assert O.sindex_to_block_idx_and_index(pairs, 2463) == (4, 95)


# ---- phase A front end ------------------------------------------------------
def _frontend_states(z):
    size, cap = int(z["image_size"]), int(z["network_capacity"])
    enc_sd = synthetic.make_discriminator_state(size, seed=21, network_capacity=cap, encoder=True)
    dis_sd = synthetic.make_discriminator_state(size, seed=22, network_capacity=cap)
    g_sd = synthetic.make_generator_state(size, seed=23, network_capacity=cap)
    fp = np.array([float(enc_sd["fc.weight"].double().sum()), float(dis_sd["blocks.0.net.0.weight"].double().abs().sum()),
                   float(g_sd["blocks.0.conv1.weight"].double().sum())])
    assert np.allclose(fp, z["fp"], rtol=1e-9, atol=1e-6), "torch CPU RNG drifted: seeded weights differ"
    return size, cap, enc_sd, dis_sd, g_sd


def test_encoder_discriminator_match_reference(golden):
    z = golden("frontend_small.npz")
    _, _, enc_sd, dis_sd, _ = _frontend_states(z)
    images = torch.from_numpy(z["images"])
    assert np.abs(O.discriminator_forward(enc_sd, images).numpy() - z["enc_batch"]).max() <= 2e-6
    assert np.abs(O.discriminator_forward(dis_sd, images).numpy() - z["dis_batch"]).max() <= 2e-6
    single = O.discriminator_forward(enc_sd, images[:1])
    assert single.shape == z["enc_single"].shape == (512,)                 # the squeeze of ST:909
    assert np.abs(single.numpy() - z["enc_single"]).max() <= 2e-6
    assert O.discriminator_forward(dis_sd, images[:1]).shape == z["dis_single"].shape == ()


def test_phase_a_matches_verbatim_notebook(golden):
    z = golden("frontend_small.npz")
    size, cap, enc_sd, dis_sd, g_sd = _frontend_states(z)
    model = tiny_cnn_from(z, "clf.")
    clf = classifiers.make_classifier("mobilenet", model, size)
    images = torch.from_numpy(z["images"])
    lat, logits = O.encode_images(enc_sd, clf.classify_images, images)
    assert np.abs(lat.numpy() - z["nb.latents"]).max() <= 2e-6
    L = len(O.generator_layout(g_sd))
    noise = torch.from_numpy(z["noise"])
    gen, sc = O.generator_forward(g_sd, O.styles_def_to_tensor([(lat, L)]), noise, get_style_coords=True)
    assert np.abs(sc.numpy() - z["nb.style_coordinates"]).max() <= 5e-6
    assert np.abs(clf.classify_images(gen).numpy() - z["nb.base_prob"]).max() <= 1e-5
    d = torch.stack([O.discriminator_forward(dis_sd, gen[i: i + 1]) for i in range(gen.shape[0])])
    assert np.abs(d.numpy() - z["nb.discriminator"][:, 0]).max() <= 1e-5


# ---- Conv2DMod gradients ------------------------------------------------------
@pytest.mark.parametrize("case", ["k3_demod", "k1_rgb", "k3_wide"])
def test_conv2dmod_gradients_match_reference_autograd(golden, case):
    z = golden("conv2dmod_grad.npz")
    b, ci, co, hw, k, demod = (int(v) for v in z[f"{case}.cfg"])
    t = {key: torch.from_numpy(z[f"{case}.{key}"]) for key in ("x", "y", "w", "go", "out", "gx", "gy", "gw")}
    assert t["w"].shape == (co, ci, k, k) and t["x"].shape == (b, ci, hw, hw)
    out, gx, gy, gw = O.modconv_grads(t["x"], t["w"], t["y"], t["go"], demod=bool(demod))
    for got, ref in ((out, t["out"]), (gx, t["gx"]), (gy, t["gy"]), (gw, t["gw"])):
        assert float((got.float() - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))
    # fp32 evaluation of the same restatement agrees too (the reference ran in fp32)
    _, gx32, gy32, gw32 = O.modconv_grads(t["x"], t["w"], t["y"], t["go"], demod=bool(demod), dtype=torch.float32)
    assert float((gw32 - t["gw"]).abs().max()) <= 2e-5 * float(t["gw"].abs().max())


def test_generator_gradients_match_reference_autograd(golden):
    z = golden("generator_grad.npz")
    size, cap = int(z["image_size"]), int(z["network_capacity"])
    sd = synthetic.make_generator_state(size, seed=31, network_capacity=cap)
    L = len(O.generator_layout(sd))
    b = z["go"].shape[0]
    styles = synthetic.make_latents(b * L, 32).reshape(b, L, -1)
    noise = synthetic.make_noise(size, 31)
    fp = np.array([float(sd["blocks.0.conv1.weight"].double().sum()), float(styles.double().sum()), float(noise.double().sum())])
    assert np.allclose(fp, z["fp"], rtol=1e-9, atol=1e-6), "torch CPU RNG drifted: seeded inputs differ"
    rgb, grads, gst = O.generator_grads(sd, styles, noise, torch.from_numpy(z["go"]))
    assert np.abs(rgb.float().numpy() - z["rgb"]).max() <= 5e-6
    names = [k[2:] for k in z.files if k.startswith("g.")]
    assert sorted(names) == sorted(grads)
    for n in names:
        ref = z["g." + n]
        assert np.abs(grads[n].float().numpy() - ref).max() <= 5e-5 * max(1.0, np.abs(ref).max()), n
    assert np.abs(gst.float().numpy() - z["g_styles"]).max() <= 5e-5 * max(1.0, np.abs(z["g_styles"]).max())


@pytest.mark.parametrize("tag,enc,seed", [("dis", False, 22), ("enc", True, 21)])
def test_discriminator_gradients_match_reference_autograd(golden, tag, enc, seed):
    z = golden("discriminator_grad.npz")
    size, cap = int(z["image_size"]), int(z["network_capacity"])
    sd = synthetic.make_discriminator_state(size, seed=seed, network_capacity=cap, encoder=enc)
    with torch.enable_grad():
        ps = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        x = torch.from_numpy(z["images"]).requires_grad_(True)
        out = O.discriminator_forward(ps, x)
        keys = [k[len(tag) + 3:] for k in z.files if k.startswith(tag + ".g.")]
        grads = torch.autograd.grad(out, [x] + [ps[k] for k in keys], torch.from_numpy(z[f"{tag}.go"]))
    assert np.abs(out.detach().numpy() - z[f"{tag}.out"]).max() <= 5e-6
    assert np.abs(grads[0].numpy() - z[f"{tag}.g_images"]).max() <= 2e-5 * max(1.0, np.abs(z[f"{tag}.g_images"]).max())
    assert len(keys) >= 10
    for k, g in zip(keys, grads[1:]):
        ref = z[f"{tag}.g.{k}"]
        assert np.abs(g.numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), k
