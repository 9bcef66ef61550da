"""-m gpu: the CUDA path (through the C ABI) against the oracle and the committed golden vectors.

Tolerances (BASELINE north_star): images max-abs <= 1e-4 in fp32, <= 2e-2 in bf16; selected top-k
(direction, sindex) sets exact.  Nothing here reads /root/reference.
"""
import copy
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import stylex_b200 as sx
from stylex_b200 import _native, synthetic
from oracle import stylex_oracle as O
from helpers import grad_digest, selection_margin_report, state_from_npz, tiny_cnn_from

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FP32_TOL = 1e-4
BF16_TOL = 2e-2


@pytest.fixture(scope="module", autouse=True)
def _exact_torch():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.set_grad_enabled(False)
    yield


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="session")
def tc_ok():
    """Run the tcgen05 kernel once in a SUBPROCESS with a timeout: a trap or hang there must not take the
    whole test process (or the box) with it."""
    code = ("import sys; sys.path.insert(0, %r); import ctypes, stylex_b200; from stylex_b200 import _native as N; "
            "e = ctypes.c_float(-1); rc = N.lib().sx_tc_selftest(5e-2, ctypes.byref(e)); "
            "print('rc', rc, 'err', e.value, N.lib().sx_last_error().decode()); sys.exit(0 if rc == 0 else 3)" % ROOT)
    try:
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=180)
    except subprocess.TimeoutExpired:
        return False, "tcgen05 selftest timed out"
    return r.returncode == 0, (r.stdout + r.stderr)[-2000:]


def _need_tc(tc_ok):
    ok, msg = tc_ok
    if not ok:
        pytest.fail("tcgen05 selftest failed, not launching bf16 kernels in-process: " + msg)


def g_module(sd, size, cap, dev):
    G = sx.Generator(size, 514, network_capacity=cap).to(dev)
    missing, unexpected = G.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith(".f") for k in missing)
    return G


# ------------------------------------------------------------------------------------------------
# L1 ops (module level, NCHW fp32 boundary) vs the oracle
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,Ci,Co,H,k,demod", [
    (2, 16, 8, 8, 3, True), (3, 64, 64, 16, 3, True), (1, 32, 16, 32, 3, True), (2, 12, 20, 4, 3, True),
    (2, 64, 3, 16, 1, False), (1, 8, 3, 64, 1, False), (5, 128, 64, 8, 3, False), (1, 512, 512, 4, 3, True)])
def test_conv2dmod_fp32(dev, B, Ci, Co, H, k, demod):
    g = torch.Generator().manual_seed(B * 1000 + Ci + Co + H)
    x = torch.randn(B, Ci, H, H, generator=g)
    y = torch.randn(B, Ci, generator=g)
    m = sx.Conv2DMod(Ci, Co, k, demod=demod).to(dev)
    ref = O.modconv(x, m.weight.detach().cpu(), y, demod=demod)
    out = m(x.to(dev), y.to(dev)).cpu()
    assert out.shape == ref.shape
    assert (out - ref).abs().max().item() <= FP32_TOL * max(1.0, ref.abs().max().item())


def test_tcgen05_selftest(tc_ok):
    ok, msg = tc_ok
    assert ok, msg


@pytest.mark.parametrize("B,Ci,Co,H", [(2, 64, 64, 16), (3, 32, 32, 8), (1, 64, 32, 64), (2, 512, 512, 4), (9, 256, 128, 4),
                                        (1, 128, 64, 128), (2, 64, 32, 256), (2, 32, 32, 64), (1, 256, 128, 32),
                                        (3, 128, 128, 16), (1, 512, 256, 32), (5, 32, 32, 16)])
def test_conv2dmod_bf16(dev, tc_ok, B, Ci, Co, H):
    _need_tc(tc_ok)
    g = torch.Generator().manual_seed(B * 1000 + Ci + Co + H)
    x = torch.randn(B, Ci, H, H, generator=g)
    y = torch.randn(B, Ci, generator=g) * 0.5
    m = sx.Conv2DMod(Ci, Co, 3, precision="bf16").to(dev)
    ref = O.modconv(x, m.weight.detach().cpu(), y, demod=True)
    out = m(x.to(dev), y.to(dev)).cpu()
    # demodulated outputs are O(1); bf16 inputs with fp32 accumulation
    assert (out - ref).abs().max().item() <= BF16_TOL * max(1.0, ref.abs().max().item())


def test_conv2dmod_bf16_unsupported_shape_is_an_error(dev):
    m = sx.Conv2DMod(12, 20, 3, precision="bf16").to(dev)
    with pytest.raises(RuntimeError, match="tcgen05"):
        m(torch.randn(1, 12, 8, 8, device=dev), torch.randn(1, 12, device=dev))


def test_cpu_tensor_is_an_error():
    m = sx.Conv2DMod(8, 8, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 8, 4, 4), torch.randn(1, 8))


def test_upsample_blur_noise_linear(dev):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 5, 8, 8, generator=g)
    up = sx.modules.upsample2x(x.to(dev)).cpu()
    assert (up - O.upsample2x(x)).abs().max().item() <= 1e-6
    bl = sx.Blur().to(dev)(x.to(dev)).cpu()
    assert (bl - O.blur3x3_reflect(x)).abs().max().item() <= 1e-6
    lin = torch.nn.Linear(1, 5)
    inoise = torch.rand(3, 16, 16, 1, generator=g)
    for nz in (inoise, inoise[:1]):
        ref = O.leaky_relu(x + lin(nz[:, :8, :8, :]).permute((0, 3, 2, 1)))
        out = sx.modules.noise_lrelu(x.to(dev), nz.to(dev), lin.to(dev)).cpu()
        lin = lin.cpu()
        assert (out - ref).abs().max().item() <= 1e-6
    w, b = torch.randn(7, 514, generator=g), torch.randn(7, generator=g)
    xi = torch.randn(4, 514, generator=g)
    out = sx.modules.linear(xi.to(dev), w.to(dev), b.to(dev)).cpu()
    assert (out - torch.nn.functional.linear(xi, w, b)).abs().max().item() <= 2e-5


def test_rgb_block_and_generator_block(dev):
    sd = synthetic.make_generator_state(16, seed=9, network_capacity=4)   # blocks (32,32) (32,16) (16,8)
    G = g_module(sd, 16, 4, dev)
    g = torch.Generator().manual_seed(4)
    w = torch.randn(2, 514, generator=g)
    noise = torch.rand(1, 16, 16, 1, generator=g)
    x = torch.randn(2, 32, 4, 4, generator=g)
    prev = torch.randn(2, 3, 8, 8, generator=g)
    xr, rgbr, scr = O.generator_block(sd, 1, x, prev, w, noise, upsample=True, upsample_rgb=True)
    xo, rgbo, sco = G.blocks[1](x.to(dev), prev.to(dev), w.to(dev), noise.to(dev))
    assert (xo.cpu() - xr).abs().max().item() <= FP32_TOL
    assert (rgbo.cpu() - rgbr).abs().max().item() <= FP32_TOL
    assert (sco.cpu() - scr).abs().max().item() <= 2e-5
    # last block: no rgb upsample, first block: no feature upsample / no prev
    x0 = torch.randn(2, 32, 4, 4, generator=g)
    xr, rgbr, _ = O.generator_block(sd, 0, x0, None, w, noise, upsample=False, upsample_rgb=True)
    xo, rgbo, _ = G.blocks[0](x0.to(dev), None, w.to(dev), noise.to(dev))
    assert (xo.cpu() - xr).abs().max().item() <= FP32_TOL and (rgbo.cpu() - rgbr).abs().max().item() <= FP32_TOL


# ------------------------------------------------------------------------------------------------
# Generator.forward vs the golden images produced by the reference itself
# ------------------------------------------------------------------------------------------------
def test_generator_small_fp32_vs_reference_golden(dev, golden):
    z = golden("gen_small.npz")
    sd = state_from_npz(z, "G.")
    G = g_module(sd, 32, 4, dev)
    w = torch.from_numpy(z["latents"]).to(dev)
    noise = torch.from_numpy(z["noise"]).to(dev)
    img, sc = G(sx.styles_def_to_tensor([(w, G.num_layers)]), noise, get_style_coords=True)
    assert np.abs(img.cpu().numpy() - z["image"]).max() <= FP32_TOL
    assert np.abs(sc.cpu().numpy() - z["style_coords"]).max() <= 2e-5
    img2, sc2 = G(torch.from_numpy(z["styles_per_layer"]).to(dev), noise, get_style_coords=True)
    assert np.abs(img2.cpu().numpy() - z["image_per_layer"]).max() <= FP32_TOL
    assert np.abs(sc2.cpu().numpy() - z["style_coords_per_layer"]).max() <= 2e-5
    # per-sample noise maps (input_noise batch == B) behave like the broadcast one when they are equal
    img3 = G(sx.styles_def_to_tensor([(w, G.num_layers)]), noise.expand(w.shape[0], -1, -1, -1).contiguous())
    assert torch.equal(img3, img)


@pytest.mark.parametrize("size", [64, 256])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_generator_full_vs_reference_golden(dev, golden, tc_ok, size, precision):
    if precision == "bf16":
        _need_tc(tc_ok)
    z = golden("gen_full.npz")
    sd = synthetic.make_generator_state(size, seed=42)
    n = z[f"image_{size}"].shape[0]
    G = g_module(sd, size, 16, dev)
    G.precision = precision
    w = synthetic.make_latents(n, 42).to(dev)
    noise = synthetic.make_noise(size, 42).to(dev)
    img, sc = G(sx.styles_def_to_tensor([(w, G.num_layers)]), noise, get_style_coords=True)
    err = np.abs(img.cpu().numpy() - z[f"image_{size}"]).max()
    print(f"generator {size}px {precision}: max-abs image error {err:.3e} (max|img| {np.abs(z[f'image_{size}']).max():.2f})")
    assert err <= (FP32_TOL if precision == "fp32" else BF16_TOL)
    assert np.abs(sc.cpu().numpy() - z[f"style_coords_{size}"]).max() <= 5e-5
    assert sc.shape[1] == {64: 2464, 256: 4512}[size]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_suffix_forward_equals_full_forward(dev, tc_ok, precision):
    """clean-prefix reuse must not change the result: for every conv c, re-running only the suffix from c with a
    perturbed style row equals the full forward with that row (same kernels downstream => tight tolerance)."""
    if precision == "bf16":
        _need_tc(tc_ok)
    size, cap = 32, 16   # channels 128/64/32: eligible for the tensor-core kernel
    sd = synthetic.make_generator_state(size, seed=21, network_capacity=cap)
    G = g_module(sd, size, cap, dev)
    plan = G.plan()
    w = synthetic.make_latents(1, 21).to(dev)
    noise = synthetic.make_noise(size, 21).to(dev)
    base = plan.styles(sx.styles_def_to_tensor([(w, G.num_layers)]).contiguous())
    plan.reserve(8, precision)
    plan.forward(base, noise, save_cache=True, precision=precision)
    for conv, (off, width) in enumerate(plan.conv_coords):
        rows = base.repeat(4, 1)
        for j in range(4):
            rows[j, off + (j * 7) % width] += 0.5 * (j + 1)
        full = plan.forward(rows, noise, precision=precision)
        # the full forward did not touch the cache (save_cache=False)
        suffix = plan.forward(rows, noise, start_conv=conv, precision=precision)
        # both paths feed bit-identical tensors to identical kernels (activations are rounded to the storage type
        # before every modulation), so a zero shift gives an effect of exactly zero
        assert torch.equal(full, suffix), (conv, (full - suffix).abs().max().item())


# ------------------------------------------------------------------------------------------------
# the sweep and the selection
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["mobilenet", "resnet"])
def test_attfind_sweep_small_vs_verbatim_notebook_golden(dev, golden, kind):
    z = golden("attfind_small.npz")
    sd = state_from_npz(z, "G.")
    size, cap = int(z["image_size"]), int(z["network_capacity"])
    G = g_module(sd, size, cap, dev)
    clf = sx.make_classifier(kind, tiny_cnn_from(z, f"{kind}.clf.").to(dev), size)
    latents = torch.from_numpy(z[f"{kind}.latents"]).to(dev)
    noise = torch.from_numpy(z["noise"]).to(dev)
    res = sx.attfind_sweep(G, clf, latents, noise, shift_size=1.0, precision="fp32", max_batch=64)
    assert np.abs(res["style_coordinates"].cpu().numpy() - z[f"{kind}.style_coordinates"]).max() <= 2e-5
    assert np.abs(res["base_prob"].cpu().numpy() - z[f"{kind}.base_prob"]).max() <= 2e-4
    assert np.abs(res["minima"].cpu().numpy() - z[f"{kind}.minima"][0]).max() <= 2e-5
    assert np.abs(res["maxima"].cpu().numpy() - z[f"{kind}.maxima"][0]).max() <= 2e-5
    err = np.abs(res["style_change"].cpu().numpy() - z[f"{kind}.style_change"]).max()
    print(f"sweep {kind}: max|effect error| {err:.3e} vs max|effect| {np.abs(z[f'{kind}.style_change']).max():.3f}")
    assert err <= 3e-4
    picks, merged, _ = sx.attfind_select(res["style_change"], res["base_prob"], 5, 0.5)
    assert picks[0] == [tuple(p) for p in z[f"{kind}.picks0"]]
    assert picks[1] == [tuple(p) for p in z[f"{kind}.picks1"]]
    assert merged == [tuple(p) for p in z[f"{kind}.merged"]]


def test_select_kernel_cases_exact(dev, golden):
    z = golden("select_cases.npz")
    for name in z["cases"]:
        eff = z[f"{name}.effects"]
        mie, k = z[f"{name}.params"]
        for c in (0, 1):
            got = sx.find_significant_styles(eff, int(k), c, None, None, None, None, None, max_image_effect=float(mie))
            assert got == [tuple(p) for p in z[f"{name}.picks{c}"]], (name, c)
    # sindex_offset and tensor input
    eff = torch.from_numpy(z["plain.effects"]).to(dev)
    got = sx.find_significant_styles(eff, 3, 0, max_image_effect=2.5, sindex_offset=10)
    assert got == [(d, s + 10) for d, s in [tuple(p) for p in z["plain.picks0"]][:3]]


def test_select_with_class_split_matches_oracle(dev):
    rng = np.random.RandomState(5)
    for N, S in ((64, 300), (7, 33), (200, 1000)):
        eff = (rng.randn(N, 2, S, 2) * 0.6).astype(np.float32)
        base = rng.randn(N, 2).astype(np.float32)
        base[3] = base[3, 0]  # a tie in the logits: np.argmax picks class 0
        picks_ref, merged_ref, _ = O.attfind_select(eff, base, 5, 0.5)
        picks, merged, _ = sx.attfind_select(torch.from_numpy(eff).to(dev), torch.from_numpy(base).to(dev), 5, 0.5)
        assert picks == picks_ref and merged == merged_ref


def test_select_empty_class_raises(dev):
    eff = torch.zeros(4, 2, 6, 2, device=dev)
    base = torch.tensor([[1.0, 0.0]] * 4, device=dev)
    with pytest.raises(IndexError):
        sx.attfind_select(eff, base, 2, 0.5)


def _config64(dev, kind, n_lat):
    sd = synthetic.make_generator_state(64, seed=42)
    G = g_module(sd, 64, 16, dev)
    lat = synthetic.make_latents(n_lat, 42)
    noise = synthetic.make_noise(64, 42)
    model = synthetic.make_classifier_model(kind, 42)
    clf_cpu = sx.make_classifier(kind, model, 64)
    L = len(O.generator_layout(sd))
    calib = O.generator_forward(sd, O.styles_def_to_tensor([(synthetic.make_latents(16, 7), L)]), noise)
    synthetic.calibrate_classifier(model, clf_cpu.preprocess, calib, target_std=1.0, chunk=8)
    clf_gpu = sx.make_classifier(kind, copy.deepcopy(model).to(dev), 64)
    return sd, G, lat, noise, clf_cpu, clf_gpu


@pytest.mark.parametrize("kind", ["mobilenet", "resnet"])
def test_attfind_sweep_64px_subset_vs_oracle(dev, kind):
    """BASELINE config shapes (64px generator, real torchvision classifier), bounded: 2 latents x 24 coordinates spread
    over every conv x 2 directions, oracle = batch-1 full forwards on the CPU."""
    sd, G, lat, noise, clf_cpu, clf_gpu = _config64(dev, kind, 2)
    S = G.num_style_coords
    sind = sorted(set(list(range(0, S, 107)) + [S - 1, 1023, 1024, 2367, 2368]))
    ref = O.attfind_sweep(sd, clf_cpu.classify_images, lat, noise, sindices=sind)
    res = sx.attfind_sweep(G, clf_gpu, lat.to(dev), noise.to(dev), precision="fp32", sindices=sind, max_batch=32)
    err = (res["style_change"].cpu() - ref["style_change"]).abs().max().item()
    mag = ref["style_change"].abs().max().item()
    print(f"64px {kind}: max|effect err| {err:.3e}, max|effect| {mag:.3e}")
    assert mag > 1e-3, "degenerate classifier: effects are ~0"
    assert err <= 1e-3 * max(1.0, mag)
    assert (res["base_prob"].cpu() - ref["base_prob"]).abs().max().item() <= 1e-3


def test_attfind_bf16_sweep_close_to_fp32(dev, tc_ok):
    """bf16 generator vs fp32 generator under the same fp32 classifier, coordinates spread over every conv.  The bf16
    perturbation of an effect is ABSOLUTE (rounding noise of the activations seen through the classifier), not relative to
    the effect: measured max 5.6e-2 / rms 1.3e-2 here and max 7.3e-2 / rms 1.0e-2 on the 256-latent config-2 job
    (profiles/r02_topk_parity_64.json).  Bounds = those with 2x head-room.  NOT a statement about picks -- the picks
    are covered by test_topk_* below, whose verification pass is what makes them exact."""
    _need_tc(tc_ok)
    sd, G, lat, noise, clf_cpu, clf_gpu = _config64(dev, "resnet", 3)
    S = G.num_style_coords
    sind = list(range(0, S, 61))
    r32 = sx.attfind_sweep(G, clf_gpu, lat.to(dev), noise.to(dev), precision="fp32", sindices=sind, max_batch=32)
    r16 = sx.attfind_sweep(G, clf_gpu, lat.to(dev), noise.to(dev), precision="bf16", sindices=sind, max_batch=32)
    d = (r32["style_change"] - r16["style_change"])[:, :, sind]
    err, rms = d.abs().max().item(), d.pow(2).mean().sqrt().item()
    mag = r32["style_change"].abs().max().item()
    print(f"bf16 vs fp32 effects: max err {err:.3e}, rms {rms:.3e}, max|effect| {mag:.3e}")
    assert mag > 1e-2
    assert err <= 0.15
    assert rms <= 0.03


def _bench_mode_classifier(dev, model, G, noise, size=64):
    """the throughput configuration bench.py times: bf16 + channels_last, BN folded / fused cuDNN ops, s2d stem, native
    max-pool, native preprocessing -- each validated against the eager module on generated images."""
    clf = sx.make_classifier("resnet", copy.deepcopy(model).to(dev), size)
    G.precision = "fp32"
    probe = G(sx.styles_def_to_tensor([(synthetic.make_latents(8, 7).to(dev), G.num_layers)]).contiguous(), noise.to(dev))
    info = clf.configure_throughput(probe, dtype=torch.bfloat16)
    assert info["classifier_mode"].startswith("fused") and info["preprocess"].startswith("native"), info
    return clf


def test_topk_throughput_mode_plus_verification_equals_parity_mode_config2_shape(dev, tc_ok):
    """VERDICT r1 item 1, at the BASELINE config-2 shape (64px generator, ResNet-18@224, ALL 2464 coordinates; 32 latents
    so the fp32 arm stays in seconds -- the 256-latent job is profiles/r02_topk_parity_64.json):

    arm 1  parity mode: fp32 generator kernels + fp32 eager classifier, TF32 off  -> the reference's picks
    arm 2  throughput mode exactly as benchmarked: bf16 tcgen05 generator + bf16 fused classifier
    arm 2v arm 2 + attfind_verify_topk (candidates re-evaluated in the parity mode)

    The synthetic job is tie-dominated (leading column means ~0.19, 0.186, 0.182 ...; bf16 moves them by ~6e-3), so arm 2
    alone need not reproduce the picks -- the margin report says which rounds it can.  Arm 2v MUST: identical per-class
    picks and merged list, re-evaluating only a few per cent of the columns."""
    _need_tc(tc_ok)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    n = int(os.environ.get("SX_TOPK_LATENTS", "32"))
    sd, G, lat, noise, clf_cpu, clf32 = _config64(dev, "resnet", n)
    lat, noise = lat.to(dev), noise.to(dev)
    S = G.num_style_coords
    r32 = sx.attfind_sweep(G, clf32, lat, noise, precision="fp32", max_batch=256)
    picks32, merged32, _ = sx.attfind_select(r32["style_change"], r32["base_prob"], 5, 0.5)
    clf16 = _bench_mode_classifier(dev, clf32.model, G, noise)
    r16 = sx.attfind_sweep(G, clf16, lat, noise, precision="bf16", max_batch=256)
    picks16, merged16, _ = sx.attfind_select(r16["style_change"], r16["base_prob"], 5, 0.5)
    rep = selection_margin_report(r32["style_change"].cpu().numpy(), r32["base_prob"].cpu().numpy(),
                                  r16["style_change"].cpu().numpy(), r16["base_prob"].cpu().numpy())
    print(f"throughput mode alone: picks {'==' if picks16 == picks32 else '!='} parity picks; smallest gap {rep['min_gap']:.2e}, "
          f"worst 2*colmean err / gap {rep['worst_2err_over_gap']:.1f}, label flips {rep['label_flips']}")
    if rep["picks_provably_equal"]:
        assert picks16 == picks32 and merged16 == merged32
    picks, merged, scores, info = sx.attfind_verify_topk(G, clf32, lat, noise, r16, 5, 0.5, precision="fp32", max_batch=128)
    frac = info["exact_evals"] / (2 * S * n)
    print(f"throughput mode + verification: {info['candidates']} of {2 * S} columns re-evaluated in fp32 ({100 * frac:.1f} % of the "
          f"coord-evals), band {info['band']:.2e}, {info['passes']} pass(es), verified={info['verified']}")
    assert info["verified"]
    assert picks == picks32, (picks, picks32)
    assert merged == merged32, (merged, merged32)
    assert frac <= 0.15
    # the exact columns of the hybrid tensor are the parity-mode effects (same kernels, batch-size independent arithmetic)
    cols = [d * S + s for c in (0, 1) for d, s in picks[c]]
    a = info["style_change"].reshape(n, 2 * S, 2)[:, cols]
    b = r32["style_change"].reshape(n, 2 * S, 2)[:, cols]
    assert float((a - b).abs().max()) <= 2e-4 * max(1.0, float(b.abs().max()))


@pytest.mark.parametrize("size,precision", [(64, "fp32"), (256, "bf16")])
def test_sweep_properties_at_baseline_sizes(dev, tc_ok, size, precision):
    """size-independent properties at the BASELINE generator shapes (no oracle needed):
    (1) shift_size = 0 => every effect is EXACTLY 0 (suffix path == clean path bit for bit, drift-free);
    (2) sharding invariance: sweeping with (rank, world) = (0,2),(1,2) and concatenating == the single-rank sweep;
    (3) a coordinate already at its minimum / maximum gives an exactly-zero effect in that direction."""
    if precision == "bf16":
        _need_tc(tc_ok)
    sd = synthetic.make_generator_state(size, seed=42)
    G = g_module(sd, size, 16, dev)
    lat = synthetic.make_latents(2, 42).to(dev)
    noise = synthetic.make_noise(size, 42).to(dev)
    torch.manual_seed(0)
    # batch-size independent classifier (no cuDNN algorithm choice): the property is about OUR path
    coef = torch.randn(2, 16).tolist()

    class Pool:
        """logits = fixed linear readout of 16 pixels, accumulated with elementwise ops only: bit-identical for
        every batch size, so any non-zero effect can only come from the generator path under test."""

        def classify_images(self, x):
            f = x[:, :, 1::(size // 3), 2::(size // 3)].reshape(x.shape[0], -1)
            out = []
            for c in range(2):
                acc = f[:, 0] * coef[c][0]
                for j in range(1, 16):
                    acc = acc + f[:, j] * coef[c][j]
                out.append(acc)
            return torch.stack(out, dim=1)
    S = G.num_style_coords
    sind = sorted(set(range(0, S, max(1, S // 40))) | {S - 1})
    r0 = sx.attfind_sweep(G, Pool(), lat, noise, shift_size=0.0, precision=precision, sindices=sind, max_batch=16)
    assert float(r0["style_change"].abs().max()) == 0.0
    full = sx.attfind_sweep(G, Pool(), lat, noise, precision=precision, sindices=sind, max_batch=16)
    parts = [sx.attfind_sweep(G, Pool(), lat, noise, precision=precision, sindices=sind, max_batch=16, rank=r, world_size=2,
                              gather=False)["style_change"] for r in range(2)]
    assert torch.equal(torch.cat(parts), full["style_change"])
    sc, mn, mx = full["style_coordinates"], full["minima"], full["maxima"]
    eff = full["style_change"]
    for n in range(2):
        for s in sind:
            if sc[n, s] == mn[s]:
                assert float(eff[n, 0, s].abs().max()) == 0.0
            if sc[n, s] == mx[s]:
                assert float(eff[n, 1, s].abs().max()) == 0.0
    assert float(eff.abs().max()) > 0


def test_attfind_extraction_entry_point(dev, tmp_path):
    """the notebook-signature entry point end to end on a tiny model, writing the 9 datasets."""
    size, cap = 16, 4
    sd = synthetic.make_generator_state(size, seed=11, network_capacity=cap)
    G = g_module(sd, size, cap, dev)

    class Enc(torch.nn.Module):
        def forward(self, img):
            return (torch.nn.functional.adaptive_avg_pool2d(img, 16).reshape(img.shape[0], -1)[:, :512] * 0.5).squeeze()

    class Stylex:
        pass

    st = Stylex()
    st.G, st.encoder, st.D = G, Enc(), None
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(),
                                torch.nn.Linear(4, 2)).to(dev).eval()
    clf = sx.make_classifier("mobilenet", model, size)
    images = [torch.rand(1, 3, size, size) for _ in range(3)]
    noise = synthetic.make_noise(size, 1).to(dev)
    out = sx.attfind_extraction(images, 3, str(tmp_path), st, clf, None, noise, G.num_style_coords, 1, -0.5,
                                image_size=size, batch_size=1, cuda_rank=0)
    assert set(out) == set(sx.attfind.DATASET_NAMES)
    assert out["style_change"].shape == (3, 2, G.num_style_coords, 2)
    f = [p for p in os.listdir(tmp_path) if p.startswith("style_change_records")]
    assert len(f) == 1
    with pytest.raises(ValueError):
        sx.attfind_extraction(images, 3, None, st, clf, None, noise, 2464, 1, -0.5, image_size=size)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 6e-2)])
def test_fused_classifier_matches_eager(dev, dtype, tol):
    """the opt-in fused ResNet inference path (folded BN + PyTorch's fused cuDNN ops) is the same function."""
    model = synthetic.make_classifier_model("resnet", 3)
    g = torch.Generator().manual_seed(0)
    imgs = torch.rand(16, 3, 64, 64, generator=g) * 2 - 0.5
    clf = sx.make_classifier("resnet", model, 64)
    synthetic.calibrate_classifier(model, clf.preprocess, imgs, chunk=8)
    clf.to(dev).set_compute(dtype, channels_last=True)
    ref = clf.classify_images(imgs.to(dev))
    clf.fuse_for_inference()
    got = clf.classify_images(imgs.to(dev))
    assert got.shape == ref.shape and torch.isfinite(got).all()
    assert float((got - ref).abs().max()) <= tol * max(1.0, float(ref.abs().max()))
    clf.fused.enable_native_pool()
    before = _native.launch_count()
    got2 = clf.classify_images(imgs.to(dev))
    assert _native.launch_count() - before == 1
    assert torch.equal(got2, got)        # the native pool is bit-identical, so is everything after it


@pytest.mark.parametrize("size", [256, 64, 224, 300])
def test_native_preprocess_matches_torchvision(dev, size):
    """sx_resize_aa_normalize == torchvision resize(antialias bilinear) + Normalize (+ cast, channels_last)."""
    model = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(3 * 224 * 224, 2)).to(dev)
    clf = sx.make_classifier("resnet", model, size)
    g = torch.Generator().manual_seed(size)
    imgs = (torch.rand(5, 3, size, size, generator=g) * 4 - 1.5).to(dev)
    ref = clf.preprocess(imgs)
    got = clf._native_pre(imgs)
    assert got.shape == ref.shape and got.is_contiguous(memory_format=torch.channels_last)
    err = float((got - ref).abs().max())
    assert err <= 2e-5 * max(1.0, float(ref.abs().max())), err
    clf.set_compute(torch.bfloat16, channels_last=True)
    got16 = clf._native_pre(imgs)
    assert got16.dtype == torch.bfloat16
    assert float((got16.float() - ref.to(torch.bfloat16).float()).abs().max()) <= 0.07   # <= 1 bf16 ulp at |x| < 16
    clf.set_compute(torch.float32)
    clf.use_native_preprocess(True)
    a = clf.classify_images(imgs)
    clf.use_native_preprocess(False)
    b = clf.classify_images(imgs)
    assert float((a - b).abs().max()) <= 1e-3 * max(1.0, float(b.abs().max()))
    with pytest.raises(NotImplementedError):
        sx.make_classifier("mobilenet", model, size).use_native_preprocess(True)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 6e-2)])
def test_s2d_stem_matches_eager_and_native_s2d_preprocess_is_exact(dev, dtype, tol):
    """the re-expressed ResNet stem (4x4 stride-1 conv on the space-to-depth image) is the same function, and the native
    kernel that writes the space-to-depth network input equals space_to_depth(native preprocess) bit for bit."""
    from stylex_b200.classifiers import space_to_depth_input
    model = synthetic.make_classifier_model("resnet", 5)
    g = torch.Generator().manual_seed(1)
    imgs = (torch.rand(6, 3, 256, 256, generator=g) * 2 - 0.5)
    clf = sx.make_classifier("resnet", model, 256)
    synthetic.calibrate_classifier(model, clf.preprocess, imgs, chunk=6)
    clf.to(dev).set_compute(dtype, channels_last=True)
    x = imgs.to(dev)
    ref = clf.classify_images(x)
    clf.fuse_for_inference()
    clf.fused.enable_s2d_stem()
    got = clf.classify_images(x)                      # torch preprocessing + torch space-to-depth
    assert float((got - ref).abs().max()) <= tol * max(1.0, float(ref.abs().max()))
    plain = clf._native_pre(x)
    s2d = clf._native_pre(x, s2d=True)
    assert s2d.shape == (6, 16, 115, 115) and s2d.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(s2d, space_to_depth_input(plain))
    clf.use_native_preprocess(True)
    got2 = clf.classify_images(x)                     # native space-to-depth preprocessing
    assert float((got2 - ref).abs().max()) <= tol * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(3, 64, 112, 112), (2, 8, 7, 9), (1, 16, 1, 1), (0, 8, 4, 4)])
def test_native_maxpool_is_bit_identical(dev, dtype, shape):
    """sx_maxpool3x3s2_nhwc == F.max_pool2d(x, 3, 2, 1) on channels_last tensors, bit for bit (NaNs included)."""
    from stylex_b200.classifiers import FusedResNetInference
    g = torch.Generator().manual_seed(shape[2])
    x = (torch.randn(shape, generator=g) * 3).to(dev).to(dtype).contiguous(memory_format=torch.channels_last)
    if x.numel() > 64:
        x[0, 5, 0, 0] = float("nan")
        x[-1, 3, -1, -1] = float("-inf")
    f = FusedResNetInference.__new__(FusedResNetInference)
    f.native_pool = True
    before = _native.launch_count()
    got = f._pool(x)
    assert _native.launch_count() - before == (1 if shape[0] else 0)
    ref = torch.nn.functional.max_pool2d(x, 3, 2, 1)
    assert got.shape == ref.shape and got.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(torch.nan_to_num(got.float(), nan=12345.0), torch.nan_to_num(ref.float(), nan=12345.0))
    assert bool(torch.isnan(got).any()) == bool(torch.isnan(ref).any())


@pytest.mark.parametrize("kind", ["mobilenet", "resnet"])
def test_counterfactual_rendering_vs_verbatim_notebook_golden(dev, golden, kind):
    """NB cells 17-20 through the native generator plan vs the golden vectors of the executed reference."""
    z = golden("attfind_small.npz")
    c = golden("counterfactual_small.npz")
    size, cap = int(z["image_size"]), int(z["network_capacity"])
    sd = state_from_npz(z, "G.")
    G = g_module(sd, size, cap, dev)
    G.precision = "fp32"
    clf = sx.make_classifier(kind, tiny_cnn_from(z, f"{kind}.clf.").to(dev), size)
    latents = torch.from_numpy(z[f"{kind}.latents"]).to(dev)
    noise = torch.from_numpy(z["noise"]).to(dev)
    smin, smax = z[f"{kind}.minima"][0], z[f"{kind}.maxima"][0]
    for i, (n, sindex, d, cls, shift) in enumerate(c[f"{kind}.cases"]):
        n, sindex, d, cls = int(n), int(sindex), int(d), int(cls)
        img, prob = sx.generate_change_image_given_dlatent([(latents[n:n + 1], G.num_layers)], G, clf, cls, sindex,
                                                           smin[sindex], smax[sindex], d, float(shift), 2, noise, 0)
        assert float((img[0].cpu() - torch.from_numpy(c[f"{kind}.images"][i])).abs().max()) <= FP32_TOL
        assert abs(float(prob) - c[f"{kind}.change_prob"][i]) <= 1e-4
        panel, cp, bp = sx.generate_images_given_dlatent(latents[n:n + 1].cpu().numpy(), G, clf, cls, sindex, smin[sindex],
                                                         smax[sindex], d, None, noise, shift_size=float(shift),
                                                         resolution=size, gen_num_layers=G.num_layers)
        assert abs(bp - c[f"{kind}.base_prob"][i]) <= 1e-4 and abs(cp - c[f"{kind}.change_prob"][i]) <= 1e-4
        diff = np.abs(panel.astype(np.int32) - c[f"{kind}.panels"][i].astype(np.int32))
        assert diff.max() <= 1 and (diff > 0).mean() <= 0.02
    # batched primitive == one-by-one; visualize_style keeps the notebook's selection rules
    sindex, d = int(c[f"{kind}.cases"][0][1]), int(c[f"{kind}.cases"][0][2])
    imgs, probs = sx.render_counterfactuals(G, clf, latents, sindex, smin[sindex], smax[sindex], d, 2.0, noise, 1)
    one, p1 = sx.render_counterfactuals(G, clf, latents[2:3], sindex, smin[sindex], smax[sindex], d, 2.0, noise, 1)
    assert torch.equal(imgs[2:3], one) and float((probs[2:3] - p1).abs().max()) <= 1e-6
    eff = np.zeros((latents.shape[0], 2, G.num_style_coords, 2), np.float32)
    eff[:, d, sindex, 1] = 1.0
    grid = sx.visualize_style(G, clf, latents.cpu().numpy(), eff, smin, smax, sindex, d, max_images=4, shift_size=2.0,
                              noise=noise, class_index=1, effect_threshold=0.0, seed=3)
    assert grid.shape == (4 * size, 2 * size, 3) and grid.dtype == np.uint8
    none = sx.visualize_style(G, clf, latents.cpu().numpy(), eff, smin, smax, sindex, d, max_images=4, shift_size=2.0,
                              noise=noise, class_index=1, effect_threshold=5.0, seed=3)
    assert none.size == 0
    with pytest.raises(IndexError):
        sx.render_counterfactuals(G, clf, latents, G.num_style_coords, 0.0, 1.0, 0, 1.0, noise)


def test_edge_cases_empty_and_ragged(dev):
    """empty batches are no-ops, a single latent works (min == max: every shift is 0), odd batch sizes and coordinate
    subsets that straddle conv boundaries are handled."""
    m = sx.Conv2DMod(16, 8, 3).to(dev)
    out = m(torch.zeros(0, 16, 8, 8, device=dev), torch.zeros(0, 16, device=dev))
    assert out.shape == (0, 8, 8, 8)
    assert sx.modules.upsample2x(torch.zeros(0, 3, 4, 4, device=dev)).shape == (0, 3, 8, 8)
    sd = synthetic.make_generator_state(16, seed=3, network_capacity=4)
    G = g_module(sd, 16, 4, dev)
    noise = synthetic.make_noise(16, 3).to(dev)
    assert G(torch.zeros(0, G.num_layers, 514, device=dev), noise).shape == (0, 3, 16, 16)

    class Mean:
        def classify_images(self, x):
            return torch.stack([x[:, 0, 2, 3], x[:, 1, 5, 1]], 1)

    lat = synthetic.make_latents(1, 3).to(dev)
    r = sx.attfind_sweep(G, Mean(), lat, noise, max_batch=7)          # odd max_batch -> rounded down to 6
    assert r["style_change"].shape == (1, 2, G.num_style_coords, 2)
    assert float(r["style_change"].abs().max()) == 0.0               # one latent: minima == maxima == its own coords
    lat3 = synthetic.make_latents(3, 3).to(dev)
    S = G.num_style_coords
    sub = [0, 31, 32, 63, 64, S - 1]                                  # straddles conv1/conv2 and block boundaries
    full = sx.attfind_sweep(G, Mean(), lat3, noise, max_batch=16)
    part = sx.attfind_sweep(G, Mean(), lat3, noise, max_batch=4, sindices=sub)
    assert torch.equal(part["style_change"][:, :, sub], full["style_change"][:, :, sub])
    mask = torch.ones(S, dtype=torch.bool)
    mask[sub] = False
    assert float(part["style_change"][:, :, mask].abs().max()) == 0.0
    with pytest.raises(ValueError):
        sx.attfind_sweep(G, Mean(), lat3, noise, minmax=(torch.zeros(3, device=dev), torch.zeros(3, device=dev)))


def test_native_launches_counted(dev):
    before = _native.launch_count()
    sx.modules.upsample2x(torch.zeros(1, 1, 4, 4, device=dev))
    assert _native.launch_count() == before + 1


# ---- phase A front end + StylEx container -----------------------------------
def _frontend(z, dev):
    size, cap = int(z["image_size"]), int(z["network_capacity"])
    enc = sx.DiscriminatorE(size, cap, encoder=True)
    dis = sx.DiscriminatorE(size, cap)
    assert not enc.load_state_dict(synthetic.make_discriminator_state(size, seed=21, network_capacity=cap, encoder=True), strict=False).unexpected_keys
    assert not dis.load_state_dict(synthetic.make_discriminator_state(size, seed=22, network_capacity=cap), strict=False).unexpected_keys
    G = g_module(synthetic.make_generator_state(size, seed=23, network_capacity=cap), size, cap, dev)

    class Stylex:
        pass

    st = Stylex()
    st.G, st.encoder, st.D = G, enc.to(dev).eval(), dis.to(dev).eval()
    clf = sx.make_classifier("mobilenet", tiny_cnn_from(z, "clf.").to(dev), size)
    return size, st, clf


def test_encoder_discriminator_match_reference_golden(dev, golden):
    """DiscriminatorE (cuDNN convolutions + the native blur) against the executed reference, batch and single image."""
    z = golden("frontend_small.npz")
    _, st, _ = _frontend(z, dev)
    images = torch.from_numpy(z["images"]).to(dev)
    before = _native.launch_count()
    e = st.encoder(images)
    assert _native.launch_count() - before == 3                      # one native blur per down-sampling block
    assert float((e.cpu() - torch.from_numpy(z["enc_batch"])).abs().max()) <= FP32_TOL
    assert float((st.D(images).cpu() - torch.from_numpy(z["dis_batch"])).abs().max()) <= FP32_TOL
    single = st.encoder(images[:1])
    assert single.shape == (512,)
    assert float((single.cpu() - torch.from_numpy(z["enc_single"])).abs().max()) <= FP32_TOL
    assert st.D(images[:1]).shape == ()


def test_attfind_extraction_phase_a_matches_verbatim_notebook(dev, golden, tmp_path):
    """the notebook-signature entry point with the real encoder / discriminator in place: the batched phase A gives the
    datasets the VERBATIM notebook loop wrote (tests/golden/frontend_small.npz)."""
    z = golden("frontend_small.npz")
    size, st, clf = _frontend(z, dev)
    images = [torch.from_numpy(z["images"][i: i + 1]) for i in range(z["images"].shape[0])]
    noise = torch.from_numpy(z["noise"]).to(dev)
    n = len(images)
    out = sx.attfind_extraction(images + images[:1], n, str(tmp_path), st, clf, None, noise, st.G.num_style_coords, 1.0, -0.5,
                                image_size=size, batch_size=1, cuda_rank=0, front_batch=4)     # 5 images: chunks of 4 + 1
    for k in ("latents", "base_prob", "style_coordinates", "discriminator", "original_images", "minima", "maxima"):
        ref = torch.from_numpy(z["nb." + k])
        assert out[k].shape == ref.shape, k
        assert float((out[k].cpu() - ref).abs().max()) <= FP32_TOL, k
    tot = float(out["style_change"].abs().sum())
    assert abs(tot - float(z["nb.style_change_abs_sum"])) <= 1e-3 * float(z["nb.style_change_abs_sum"])


def test_attfind_extraction_discriminator_filter_matches_verbatim_notebook(dev, golden):
    """``use_discriminator=True`` against the VERBATIM notebook loop (tests/golden/frontend_filter.npz): the executed
    reference keeps the images whose discriminator output is BELOW the threshold (NB:262-266 returns the flag that NB:327
    reads as `skip`), in dataloader order; and when fewer images pass than ``num_images`` its zero rows take part in
    minima / maxima (NB:291, 340), which changes every shift."""
    z = golden("frontend_small.npz")
    f = golden("frontend_filter.npz")
    size, st, clf = _frontend(z, dev)
    images = [torch.from_numpy(z["images"][i: i + 1]) for i in range(z["images"].shape[0])]
    noise = torch.from_numpy(z["noise"]).to(dev)
    thr = float(f["threshold"])
    d = f["d_all"]
    want = [i for i in range(len(images)) if d[i] < thr]
    assert len(want) == 3
    for tag, num_images, loader in (("exact", 3, images + images[:1]), ("short", 4, images)):
        out = sx.attfind_extraction(loader, num_images, None, st, clf, None, noise, st.G.num_style_coords, 1.0, thr,
                                    image_size=size, use_discriminator=True, front_batch=2)
        for k in ("latents", "base_prob", "style_coordinates", "discriminator", "original_images", "minima", "maxima"):
            ref = torch.from_numpy(f[f"{tag}.{k}"])
            assert out[k].shape == ref.shape, (tag, k)
            assert float((out[k].cpu() - ref).abs().max()) <= FP32_TOL, (tag, k)
        assert float((out["latents"].cpu()[:3] - torch.from_numpy(z["nb.latents"][want])).abs().max()) <= FP32_TOL
        ref = torch.from_numpy(f[f"{tag}.style_change"])
        err = float((out["style_change"].cpu() - ref).abs().max())
        assert err <= 2e-4 * max(1.0, float(ref.abs().max())), (tag, err)


def test_stylex_container_checkpoint_on_device(dev, tmp_path):
    """StylEx container: reference-format checkpoint -> device model whose G / encoder / D run the native path."""
    torch.manual_seed(5)
    m = sx.StylEx(image_size=16, network_capacity=4)
    assert next(m.parameters()).is_cuda                                  # ST:965
    path = sx.save_checkpoint(m, tmp_path, "default", 7, sx.stylex_config(16, network_capacity=4))
    m2 = sx.load_stylex(path, 16, network_capacity=4)
    img = torch.rand(3, 3, 16, 16, device=dev)
    w = m2.encoder(img)
    assert w.shape == (3, 512) and torch.equal(w, m.eval().encoder(img))
    lat = torch.cat([w, torch.zeros(3, 2, device=dev)], dim=1)
    noise = synthetic.make_noise(16, 1).to(dev)
    rgb = m2.G(sx.styles_def_to_tensor([(lat, m2.G.num_layers)]), noise)
    assert rgb.shape == (3, 3, 16, 16) and torch.isfinite(rgb).all()
    assert m2.D(rgb).shape == (3,)
    assert torch.equal(m2.GE(sx.styles_def_to_tensor([(lat, m2.G.num_layers)]), noise), rgb)   # GE == G after reset_parameter_averaging


# ---- Conv2DMod backward ---------------------------------------------------------------
def _rel(got, ref):
    return float((got.double().cpu() - ref.double().cpu()).abs().max()) / max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("case", ["k3_demod", "k1_rgb", "k3_wide"])
def test_conv2dmod_backward_matches_reference_autograd(dev, golden, case):
    """sx_conv2dmod_bwd through torch.autograd against the gradients the UNMODIFIED reference module produced."""
    z = golden("conv2dmod_grad.npz")
    b, ci, co, hw, k, demod = (int(v) for v in z[f"{case}.cfg"])
    t = {key: torch.from_numpy(z[f"{case}.{key}"]).to(dev) for key in ("x", "y", "w", "go", "out", "gx", "gy", "gw")}
    conv = sx.Conv2DMod(ci, co, k, demod=bool(demod)).to(dev)
    with torch.no_grad():
        conv.weight.copy_(t["w"])
    before = _native.launch_count()
    with torch.enable_grad():
        xs, ys = t["x"].clone().requires_grad_(True), t["y"].clone().requires_grad_(True)
        out = conv(xs, ys)
        gx, gy, gw = torch.autograd.grad(out, (xs, ys, conv.weight), t["go"])
    assert _native.launch_count() - before >= 10                       # forward + backward kernels, all native
    assert _rel(out, t["out"]) <= FP32_TOL
    assert _rel(gx, t["gx"]) <= FP32_TOL
    assert _rel(gy, t["gy"]) <= FP32_TOL
    assert _rel(gw, t["gw"]) <= FP32_TOL


@pytest.mark.parametrize("b,ci,co,hw,k,demod", [(4, 64, 32, 32, 3, True), (3, 128, 3, 16, 1, False), (2, 20, 36, 12, 3, True),
                                                (1, 512, 512, 4, 3, True)])
def test_conv2dmod_backward_matches_oracle_fp64(dev, b, ci, co, hw, k, demod):
    """generator-sized and ragged shapes against torch autograd through the oracle's literal restatement in float64;
    .backward() accumulates into .grad like any autograd op; deterministic (fixed-order split-K)."""
    g = torch.Generator().manual_seed(b * 1000 + ci)
    x = torch.randn(b, ci, hw, hw, generator=g)
    y = torch.randn(b, ci, generator=g) * 0.5
    w = torch.randn(co, ci, k, k, generator=g) * (2.0 / (ci * k * k)) ** 0.5
    go = torch.randn(b, co, hw, hw, generator=g)
    out_ref, gx_ref, gy_ref, gw_ref = O.modconv_grads(x, w, y, go, demod=demod)
    conv = sx.Conv2DMod(ci, co, k, demod=demod).to(dev)
    with torch.no_grad():
        conv.weight.copy_(w)
    res = []
    for _ in range(2):
        conv.weight.grad = None
        with torch.enable_grad():
            xs, ys = x.to(dev).requires_grad_(True), y.to(dev).requires_grad_(True)
            out = conv(xs, ys)
            (out * go.to(dev)).sum().backward()
        res.append((xs.grad.clone(), ys.grad.clone(), conv.weight.grad.clone()))
    assert _rel(out.detach(), out_ref) <= FP32_TOL
    assert _rel(res[0][0], gx_ref) <= FP32_TOL
    assert _rel(res[0][1], gy_ref) <= FP32_TOL
    assert _rel(res[0][2], gw_ref) <= FP32_TOL
    for a, c in zip(res[0], res[1]):
        assert torch.equal(a, c)                                       # bit-reproducible


def test_conv2dmod_backward_properties(dev):
    """size-independent properties: linear in the upstream gradient; a zero upstream gradient gives exact zeros; the
    torch.no_grad() records no graph; an empty batch gives a zero weight gradient."""
    g = torch.Generator().manual_seed(9)
    b, ci, co, hw = 2, 32, 32, 16
    x = torch.randn(b, ci, hw, hw, generator=g).to(dev)
    y = (torch.randn(b, ci, generator=g) * 0.5).to(dev)
    conv = sx.Conv2DMod(ci, co, 3).to(dev)
    g1 = torch.randn(b, co, hw, hw, generator=g).to(dev)
    g2 = torch.randn(b, co, hw, hw, generator=g).to(dev)

    def grads(go):
        with torch.enable_grad():
            xs, ys = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
            return torch.autograd.grad(conv(xs, ys), (xs, ys, conv.weight), go)

    a, c, both = grads(g1), grads(g2), grads(2.0 * g1 - 0.5 * g2)
    for ga, gc, gb in zip(a, c, both):
        assert _rel(gb, 2.0 * ga - 0.5 * gc) <= FP32_TOL
    for gz in grads(torch.zeros_like(g1)):
        assert float(gz.abs().max()) == 0.0
    with torch.no_grad():
        assert not conv(x, y).requires_grad
    lib = _native.lib()
    gw = torch.ones(co, ci, 3, 3, device=dev)
    _native.check(lib.sx_conv2dmod_bwd(0, 0, 0, 0, 0, 0, gw.data_ptr(), 0, 0, ci, co, hw, hw, 3, 1, 1e-8, 0, 0, 0, _native.stream_ptr()),
                  "sx_conv2dmod_bwd")
    assert float(gw.abs().max()) == 0.0


def test_generator_backward_matches_reference_autograd(dev, golden):
    """Generator.forward with gradients recorded (every op a native forward + native backward) against the gradients the
    UNMODIFIED reference Generator produced under torch autograd: all 42 parameters and the styles."""
    z = golden("generator_grad.npz")
    size, cap = int(z["image_size"]), int(z["network_capacity"])
    sd = synthetic.make_generator_state(size, seed=31, network_capacity=cap)
    G = g_module(sd, size, cap, dev).train()
    b = z["go"].shape[0]
    styles = synthetic.make_latents(b * G.num_layers, 32).reshape(b, G.num_layers, -1).to(dev)
    noise = synthetic.make_noise(size, 31).to(dev)
    go = torch.from_numpy(z["go"]).to(dev)
    before = _native.launch_count()
    with torch.enable_grad():
        st = styles.clone().requires_grad_(True)
        rgb = G(st, noise)
        assert rgb.requires_grad
        (rgb * go).sum().backward()
    assert _native.launch_count() - before > 100
    assert _rel(rgb.detach(), torch.from_numpy(z["rgb"])) <= FP32_TOL
    # the same forward through the fused plan (no graph) agrees with the module-level autograd path
    assert _rel(G(styles, noise), rgb.detach()) <= FP32_TOL
    names = [k[2:] for k in z.files if k.startswith("g.")]
    params = dict(G.named_parameters())
    assert sorted(names) == sorted(params)
    for n in names:
        assert params[n].grad is not None, n
        assert _rel(params[n].grad, torch.from_numpy(z["g." + n])) <= FP32_TOL, n
    assert _rel(st.grad, torch.from_numpy(z["g_styles"])) <= FP32_TOL


def test_generator_backward_64px_vs_oracle(dev):
    """BASELINE-shaped 64px generator (capacity 16, 512-channel blocks): image-loss gradients against float64 autograd
    through the oracle for a sample of parameters of every kind."""
    size = 64
    sd = synthetic.make_generator_state(size, seed=42)
    G = g_module(sd, size, 16, dev).train()
    b = 2
    lat = synthetic.make_latents(b, 5)
    styles = O.styles_def_to_tensor([(lat, G.num_layers)])
    noise = synthetic.make_noise(size, 42)
    go = torch.randn(b, 3, size, size, generator=torch.Generator().manual_seed(1))
    _, ref, ref_st = O.generator_grads(sd, styles, noise, go, dtype=torch.float64)
    with torch.enable_grad():
        st = styles.to(dev).requires_grad_(True)
        rgb = G(st, noise.to(dev))
        (rgb * go.to(dev)).sum().backward()
    params = dict(G.named_parameters())
    for n in ("initial_block", "initial_conv.weight", "blocks.0.conv1.weight", "blocks.0.to_style1.weight", "blocks.1.to_noise1.weight",
              "blocks.2.conv2.weight", "blocks.3.to_rgb.conv.weight", "blocks.3.to_rgb.to_style.bias", "blocks.4.conv1.weight",
              "blocks.4.to_noise2.bias", "blocks.4.to_style2.weight"):
        assert _rel(params[n].grad, ref[n]) <= 2e-4, n          # fp32 sums over up to 64*64*2 pixels and 512 channels
    assert _rel(st.grad, ref_st) <= 2e-4


def test_bandwidth_op_adjoints(dev):
    """<op(x), g> == <x, op_bwd(g)> for the linear maps (upsample, blur) on ragged sizes, and the closed forms of the
    noise / linear backward against torch autograd of the same expressions."""
    g = torch.Generator().manual_seed(3)
    for (b, c, h, w) in [(2, 3, 2, 2), (1, 5, 7, 4), (3, 4, 16, 16)]:
        x = torch.randn(b, c, h, w, generator=g, dtype=torch.float32).to(dev)
        gu = torch.randn(b, c, 2 * h, 2 * w, generator=g).to(dev)
        gb = torch.randn(b, c, h, w, generator=g).to(dev)
        with torch.enable_grad():
            xs = x.clone().requires_grad_(True)
            (gxu,) = torch.autograd.grad(sx.modules.upsample2x(xs), xs, gu)
            xs2 = x.clone().requires_grad_(True)
            (gxb,) = torch.autograd.grad(sx.Blur().to(dev)(xs2), xs2, gb)
            xr = x.clone().requires_grad_(True)
            ref_u = torch.nn.functional.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=False)
            (ru,) = torch.autograd.grad(ref_u, xr, gu)
            xr2 = x.cpu().clone().requires_grad_(True)
            (rb,) = torch.autograd.grad(O.blur3x3_reflect(xr2), xr2, gb.cpu())
        assert _rel(gxu, ru) <= 1e-5 and _rel(gxb, rb) <= 1e-5
    # linear
    x = torch.randn(5, 514, generator=g).to(dev)
    lin = torch.nn.Linear(514, 48).to(dev)
    go = torch.randn(5, 48, generator=g).to(dev)
    with torch.enable_grad():
        xs = x.clone().requires_grad_(True)
        got = torch.autograd.grad(sx.modules.linear(xs, lin.weight, lin.bias), (xs, lin.weight, lin.bias), go)
        xr = x.clone().requires_grad_(True)
        ref = torch.autograd.grad(torch.nn.functional.linear(xr, lin.weight, lin.bias), (xr, lin.weight, lin.bias), go)
    for a, r in zip(got, ref):
        assert _rel(a, r) <= 1e-5
    # noise + leaky-ReLU (transposed noise, quirk Q1), per-sample and broadcast noise maps
    for nb in (1, 3):
        xin = torch.randn(3, 6, 8, 8, generator=g).to(dev)
        inoise = torch.rand(nb, 16, 16, 1, generator=g).to(dev)
        tn = torch.nn.Linear(1, 6).to(dev)
        go = torch.randn(3, 6, 8, 8, generator=g).to(dev)
        with torch.enable_grad():
            xs = xin.clone().requires_grad_(True)
            got = torch.autograd.grad(sx.modules.noise_lrelu(xs, inoise, tn), (xs, tn.weight, tn.bias), go)
            xr = xin.clone().requires_grad_(True)
            nz = tn(inoise[:, :8, :8, :]).permute((0, 3, 2, 1))
            ref = torch.autograd.grad(torch.nn.functional.leaky_relu(xr + nz, 0.2), (xr, tn.weight, tn.bias), go)
        for a, r in zip(got, ref):
            assert a.shape == r.shape and _rel(a, r) <= 1e-5


@pytest.mark.parametrize("tag,enc,seed", [("dis", False, 22), ("enc", True, 21)])
def test_discriminator_backward_matches_reference_autograd(dev, golden, tag, enc, seed):
    """DiscriminatorE under autograd (cuDNN convolutions + the native blur forward / adjoint kernels) against the executed
    reference: input-image gradient (what the gradient penalty and the encoder path differentiate) and parameter gradients."""
    z = golden("discriminator_grad.npz")
    size, cap = int(z["image_size"]), int(z["network_capacity"])
    m = sx.DiscriminatorE(size, cap, encoder=enc)
    m.load_state_dict(synthetic.make_discriminator_state(size, seed=seed, network_capacity=cap, encoder=enc), strict=False)
    m = m.to(dev).train()
    before = _native.launch_count()
    with torch.enable_grad():
        x = torch.from_numpy(z["images"]).to(dev).requires_grad_(True)
        out = m(x)
        (out * torch.from_numpy(z[f"{tag}.go"]).to(dev)).sum().backward()
    assert _native.launch_count() - before == 6                         # 3 blurs forward + 3 adjoints
    assert _rel(out.detach(), torch.from_numpy(z[f"{tag}.out"])) <= FP32_TOL
    assert _rel(x.grad, torch.from_numpy(z[f"{tag}.g_images"])) <= FP32_TOL
    params = dict(m.named_parameters())
    keys = [k[len(tag) + 3:] for k in z.files if k.startswith(tag + ".g.")]
    for k in keys:
        assert _rel(params[k].grad, torch.from_numpy(z[f"{tag}.g.{k}"])) <= FP32_TOL, k


@pytest.mark.parametrize("size", [256, 300, 64, 224])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_resize_s2d_separable_equals_direct(dev, size, dtype):
    """the separable (shared-memory, two-pass) space-to-depth preprocessing kernel performs the direct kernel's fp32
    operations in the same order: bit-identical outputs for down-scaling, up-scaling and identity resizes."""
    from stylex_b200.classifiers import space_to_depth_input
    model = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(3 * 224 * 224, 2)).to(dev)
    clf = sx.make_classifier("resnet", model, size)
    clf.to(dev).set_compute(dtype, channels_last=True)
    g = torch.Generator().manual_seed(size)
    x = (torch.rand(5, 3, size, size, generator=g) * 4 - 1.5).to(dev)
    sep = clf._native_pre(x, s2d=True)
    os.environ["SX_RESIZE_DIRECT"] = "1"
    try:
        direct = clf._native_pre(x, s2d=True)
    finally:
        del os.environ["SX_RESIZE_DIRECT"]
    assert sep.shape == (5, 16, 115, 115)
    assert torch.equal(sep, direct)
    assert torch.equal(sep, space_to_depth_input(clf._native_pre(x)))


@pytest.mark.parametrize("b,ci,co,hw,k,demod", [(2, 128, 128, 64, 3, True), (2, 32, 32, 64, 3, True), (4, 64, 32, 32, 3, True),
                                                (3, 64, 3, 16, 1, False), (2, 32, 32, 128, 3, True)])
def test_conv2dmod_backward_bf16_tensor_cores(dev, tc_ok, b, ci, co, hw, k, demod):
    """precision bf16: dgrad on the tcgen05 conv kernels, wgrad as a tcgen05 GEMM over K = pixels (FFMA where a kernel does
    not take the shape: k = 1 for the dgrad, maps narrower than 64 pixels for the wgrad), against float64 autograd through the oracle; tolerance 2e-2 of the largest
    gradient entry (bf16 operands, fp32 accumulation); bit-reproducible."""
    _need_tc(tc_ok)
    g = torch.Generator().manual_seed(b * 100 + ci + hw)
    x = torch.randn(b, ci, hw, hw, generator=g)
    y = torch.randn(b, ci, generator=g) * 0.5
    w = torch.randn(co, ci, k, k, generator=g) * (2.0 / (ci * k * k)) ** 0.5
    go = torch.randn(b, co, hw, hw, generator=g)
    _, gx_ref, gy_ref, gw_ref = O.modconv_grads(x, w, y, go, demod=demod)
    lib = _native.lib()
    xd, yd, wd, god = x.to(dev), y.to(dev), w.to(dev), go.to(dev)
    ws = torch.empty(lib.sx_conv2dmod_workspace_bytes(b, ci, co, hw, hw, k, 0) + 256, dtype=torch.uint8, device=dev)
    out = torch.empty(b, co, hw, hw, device=dev)
    _native.check(lib.sx_conv2dmod_fwd(xd.data_ptr(), wd.data_ptr(), yd.data_ptr(), out.data_ptr(), b, ci, co, hw, hw, k, int(demod), 1e-8,
                                       0, ws.data_ptr(), ws.numel(), _native.stream_ptr()), "fwd")
    res = []
    for _ in range(2):
        gx, gy, gw = torch.empty_like(xd), torch.empty_like(yd), torch.empty_like(wd)
        wb = torch.empty(lib.sx_conv2dmod_bwd_workspace_bytes(b, ci, co, hw, hw, k, 1), dtype=torch.uint8, device=dev)
        _native.check(lib.sx_conv2dmod_bwd(xd.data_ptr(), wd.data_ptr(), yd.data_ptr(), out.data_ptr(), god.data_ptr(), gx.data_ptr(),
                                           gw.data_ptr(), gy.data_ptr(), b, ci, co, hw, hw, k, int(demod), 1e-8, 1, wb.data_ptr(),
                                           wb.numel(), _native.stream_ptr()), "sx_conv2dmod_bwd bf16")
        torch.cuda.synchronize()
        res.append((gx, gy, gw))
    assert _rel(res[0][0], gx_ref) <= BF16_TOL
    assert _rel(res[0][1], gy_ref) <= BF16_TOL
    assert _rel(res[0][2], gw_ref) <= BF16_TOL
    for a, c in zip(res[0], res[1]):
        assert torch.equal(a, c)


def test_generator_backward_bf16(dev, tc_ok):
    """Generator.forward with precision "bf16" and gradients recorded: the 3x3 modulated convs run forward / dgrad / wgrad on
    the tensor cores; parameter and style gradients within the bf16 tolerance of float64 autograd through the oracle."""
    _need_tc(tc_ok)
    size = 64
    sd = synthetic.make_generator_state(size, seed=42)
    G = g_module(sd, size, 16, dev).train()
    G.precision = "bf16"
    lat = synthetic.make_latents(2, 5)
    styles = O.styles_def_to_tensor([(lat, G.num_layers)])
    noise = synthetic.make_noise(size, 42)
    go = torch.randn(2, 3, size, size, generator=torch.Generator().manual_seed(1))
    rgb_ref, ref, ref_st = O.generator_grads(sd, styles, noise, go, dtype=torch.float64)
    with torch.enable_grad():
        st = styles.to(dev).requires_grad_(True)
        rgb = G(st, noise.to(dev))
        (rgb * go.to(dev)).sum().backward()
    assert _rel(rgb.detach(), rgb_ref) <= BF16_TOL

    def cos(a, r):
        a, r = a.double().cpu().flatten(), r.double().flatten()
        return float(torch.dot(a, r) / (a.norm() * r.norm()))

    # bf16 rounding compounds through the chain (the gradient of initial_conv has crossed all ten 3x3 convs backwards: measured
    # 4.8e-2 of its largest entry, against <= 2e-2 for a single conv): per tensor, direction within cos >= 0.99 of the float64
    # gradient and every entry within 1e-1 of the largest one -- a wrong kernel gives cos ~ 0
    params = dict(G.named_parameters())
    for n in ("initial_conv.weight", "blocks.0.conv1.weight", "blocks.1.conv2.weight", "blocks.2.conv1.weight", "blocks.3.conv2.weight",
              "blocks.4.conv1.weight", "blocks.4.to_style2.weight", "blocks.3.to_rgb.conv.weight", "blocks.2.to_noise1.weight"):
        assert _rel(params[n].grad, ref[n]) <= 1e-1, n
        assert cos(params[n].grad, ref[n]) >= 0.99, n
    assert _rel(st.grad, ref_st) <= 1e-1 and cos(st.grad, ref_st) >= 0.99


@pytest.mark.parametrize("h,batch,bcast", [(2, 3, False), (4, 2, True), (8, 5, False), (16, 3, True), (64, 2, False), (128, 2, True)])
def test_rgb_prefill_upsample_blur_vs_oracle(dev, h, batch, bcast):
    """RGBBlock.upsample = Sequential(Upsample(bilinear x2), Blur) ST:613-616 as the closed-form pre-fill kernels of the
    fused ToRGB (one 2x2 quad per thread for small planes, one 4x4 block per thread from 16 px up; borders included),
    with and without the batch broadcast of the AttFind suffix forwards."""
    torch.manual_seed(h)
    prev = torch.randn(1 if bcast else batch, 3, h, h)
    ref = O.blur3x3_reflect(O.upsample2x(prev)).expand(batch, -1, -1, -1)
    out = torch.full((batch, 3, 2 * h, 2 * h), float("nan"), device=dev)
    pd = prev.to(dev).contiguous()
    _native.check(_native.lib().sx_rgb_prefill_upsample_blur(pd.data_ptr(), pd.shape[0], out.data_ptr(), batch, h, h,
                                                             _native.stream_ptr()), "sx_rgb_prefill_upsample_blur")
    assert float((out.cpu() - ref).abs().max()) <= 2e-6


def _digest_close(got, ref, rel, what):
    """compare a gradient with its golden digest (strided sample + three sums, tests/helpers.grad_digest)"""
    d = grad_digest(got)
    assert d.shape == ref.shape, (what, d.shape, ref.shape)
    scale = max(1e-6, float(np.abs(ref[3:]).max()))
    err = float(np.abs(d[3:] - ref[3:]).max())
    assert err <= rel * scale, (what, err, scale)
    assert abs(d[1] - ref[1]) <= 4 * rel * max(1e-6, ref[1]), (what, "sum|g|", d[1], ref[1])


def test_training_losses_and_double_backward_match_reference_autograd(dev, golden):
    with torch.enable_grad():
        _training_losses_and_double_backward(dev, golden)


def _training_losses_and_double_backward(dev, golden):
    """SURVEY.md section 8f row 1 / VERDICT r1 item 7: the loss helpers and both phases of ``Trainer.train`` on this
    package's modules against the reference's own classes under torch autograd (tests/golden/training_small.npz):

    * ``calc_pl_lengths`` (ST:306-316) + the path-length loss: DOUBLE backward through the generator
      (``Generator.double_backward``: ConvSharedFunction / ConvWgradFunction, the upsample / blur Function pairs);
    * the discriminator phase with ``gradient_penalty`` (ST:296-303): double backward through DiscriminatorE (cuDNN
      convolutions + the native Blur pair);
    * the generator phase, encoder branch: gen hinge + reconstruction (L1 image + L1 w) + classifier KL, first-order native
      backward of every generator op, gradients flowing through the classifier and the encoder."""
    from stylex_b200 import training as T
    z = golden("training_small.npz")
    size, cap = int(z["image_size"]), int(z["network_capacity"])
    b = z["real"].shape[0]
    G = g_module(synthetic.make_generator_state(size, seed=41, network_capacity=cap), size, cap, dev).train()
    enc = sx.DiscriminatorE(size, cap, encoder=True)
    dis = sx.DiscriminatorE(size, cap)
    enc.load_state_dict(synthetic.make_discriminator_state(size, seed=42, network_capacity=cap, encoder=True), strict=False)
    dis.load_state_dict(synthetic.make_discriminator_state(size, seed=43, network_capacity=cap), strict=False)
    enc, dis = enc.to(dev).train(), dis.to(dev).train()
    clf = sx.make_classifier("mobilenet", tiny_cnn_from(z, "clf.").to(dev), size)
    noise = torch.from_numpy(z["noise"]).to(dev)
    real = torch.from_numpy(z["real"]).to(dev)
    enc_batch = torch.from_numpy(z["enc_batch"]).to(dev)
    gnames = [n for n, _ in G.named_parameters()]
    # ---- path-length penalty: double backward through the generator
    G.double_backward = True
    styles = torch.from_numpy(z["pl.styles"]).to(dev).requires_grad_(True)
    images = G(styles, noise)
    torch.manual_seed(47)
    pl_noise = torch.randn(images.shape).to(dev)                         # the reference's CPU draw under the same seed
    pl_lengths = T.calc_pl_lengths(styles, images, pl_noise)
    assert float((pl_lengths.detach().cpu() - torch.from_numpy(z["pl.lengths"])).abs().max()) <= 2e-4
    pl_loss = ((pl_lengths - 0.05) ** 2).mean()
    assert abs(float(pl_loss) - float(z["pl.loss"])) <= 2e-4 * max(1.0, float(z["pl.loss"]))
    grads = torch.autograd.grad(pl_loss, list(G.parameters()), allow_unused=True)
    for n, gr in zip(gnames, grads):
        ref = z["pl.g." + n]
        if gr is None:
            assert ref.shape == (4,) and float(np.abs(ref).max()) == 0.0, n    # the reference has no gradient there either
            continue
        _digest_close(gr, ref, 2e-3, "pl." + n)
    G.double_backward = False
    # ---- discriminator phase with the gradient penalty: double backward through the discriminator
    w = torch.cat((synthetic.make_latents(b, 48)[:, :512].to(dev), clf.classify_images(enc_batch).detach()), dim=1)
    gen = G(sx.styles_def_to_tensor([(w, G.num_layers)]), noise)
    rb = real.clone().requires_grad_(True)
    fake_output = dis(gen.clone().detach())
    real_output = dis(rb)
    divergence = T.hinge_loss(real_output, fake_output)
    gp = T.gradient_penalty(rb, real_output)
    assert abs(float(divergence) - float(z["d.divergence"])) <= 2e-4 * max(1.0, float(z["d.divergence"]))
    assert abs(float(gp) - float(z["d.gp"])) <= 1e-3 * max(1.0, float(z["d.gp"]))
    grads = torch.autograd.grad(divergence + gp, list(dis.parameters()))
    for (n, _), gr in zip(dis.named_parameters(), grads):
        _digest_close(gr, z["d.g." + n], 2e-3, "d." + n)
    # ---- generator phase, encoder branch
    eb = enc_batch.clone().requires_grad_(True)
    encoder_output = enc(eb)
    real_logits = clf.classify_images(eb)
    w_styles = sx.styles_def_to_tensor([(torch.cat((encoder_output, real_logits), dim=1), G.num_layers)])
    gen = G(w_styles, noise)
    assert float((gen.detach().cpu() - torch.from_numpy(z["g.image"])).abs().max()) <= 2e-4
    gen_logits = clf.classify_images(gen)
    fake_output = dis(gen)
    rec = 2 * 10 * T.reconstruction_loss(eb, gen, enc(gen), encoder_output)
    kl = 2 * 1 * T.classifier_kl_loss(real_logits, gen_logits)
    gen_loss = T.gen_hinge_loss(fake_output, None)
    for got, key in ((gen_loss, "g.gen_loss"), (rec, "g.rec"), (kl, "g.kl")):
        assert abs(float(got) - float(z[key])) <= 2e-4 * max(1.0, abs(float(z[key]))), key
    grads = torch.autograd.grad(gen_loss + rec + kl, list(G.parameters()) + list(enc.parameters()))
    names = gnames + ["enc." + n for n, _ in enc.named_parameters()]
    for n, gr in zip(names, grads):
        _digest_close(gr, z["g.g." + n], 1e-3, "g." + n)


def test_train_step_runs_and_learns(dev):
    with torch.enable_grad():
        _train_step_runs_and_learns(dev)


def _train_step_runs_and_learns(dev):
    """``training.TrainStep`` (ST:1249-1506): a few optimisation steps at 16 px with gradient accumulation (both the
    noise and the encoder branch of the alternating schedule), the gradient penalty (step 0) and the path-length penalty
    (forced early): finite losses, every trainable parameter of G / S / encoder / D receives updates."""
    from stylex_b200 import training as T
    torch.manual_seed(0)
    size, cap, bs = 16, 4, 4
    st = sx.StylEx(size, network_capacity=cap, rank=dev.index or 0)
    with torch.no_grad():                         # the reference zero-initialises to_noise: give the noise path something to learn
        for blk in st.G.blocks:
            blk.to_noise1.weight.normal_(0, 0.1)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, stride=2, padding=1), torch.nn.ReLU(), torch.nn.AdaptiveAvgPool2d(1),
                                torch.nn.Flatten(), torch.nn.Linear(8, 2)).to(dev)
    clf = sx.make_classifier("mobilenet", model, size)
    g = torch.Generator().manual_seed(1)
    data = torch.rand(64, 3, size, size, generator=g).to(dev)

    def loader():
        i = 0
        while True:
            yield data[(i * bs) % 60: (i * bs) % 60 + bs].clone()
            i += 1
    it = loader()
    ts = T.TrainStep(st, clf, batch_size=bs, gradient_accumulate_every=2, rank=dev.index or 0, pl_after=0)
    before = {n: p.detach().clone() for n, p in st.named_parameters() if p.requires_grad}
    logs = []
    for step in range(34):                        # step 32 applies the path-length penalty (steps > pl_after and % 32 == 0)
        if step in (3, 4, 5) or 8 <= step < 31:
            ts.steps += 1                         # skip ahead: only the interesting steps are run
            continue
        logs.append(ts.train_step(it))
    assert all(np.isfinite([v for v in lg.values() if v is not None]).all() for lg in logs), logs
    assert logs[0]["GP"] is not None and ts.pl_mean is not None
    moved = {n: float((p.detach() - before[n]).abs().max()) for n, p in st.named_parameters() if p.requires_grad}
    frozen = [n for n, v in moved.items() if v == 0.0 and not n.startswith(("SE.", "GE."))]
    assert not frozen, frozen


def _small_tc_setup(dev, n, seed=11):
    """32 px / capacity 16 generator (every conv tensor-core eligible) + a tiny conv classifier with O(1) logits"""
    size, cap = 32, 16
    sd = synthetic.make_generator_state(size, seed=seed, network_capacity=cap)
    G = g_module(sd, size, cap, dev)
    lat = synthetic.make_latents(n, seed).to(dev)
    noise = synthetic.make_noise(size, seed).to(dev)
    torch.manual_seed(1)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, stride=2, padding=1), torch.nn.ReLU(), torch.nn.AdaptiveAvgPool2d(1),
                                torch.nn.Flatten(), torch.nn.Linear(4, 2)).eval()
    with torch.no_grad():
        model[-1].weight.mul_(20)
    clf = sx.make_classifier("mobilenet", model.to(dev), size)
    G.precision = "fp32"
    lg = clf.classify_images(G(sx.styles_def_to_tensor([(lat, G.num_layers)]).contiguous(), noise))
    with torch.no_grad():                          # decision boundary between the base images: both classes populated
        d = torch.sort(lg[:, 1] - lg[:, 0]).values
        mid = float(d[(len(d) - 1) // 2] + d[(len(d) - 1) // 2 + 1]) / 2
        model[-1].bias[1] -= mid / 2
        model[-1].bias[0] += mid / 2
    return G, clf, lat, noise


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_sweep_edge_cases(dev, tc_ok, precision):
    """Edge cases of the sweep (the reference has no tests; these follow from NB:340-387): ONE latent -- minima == maxima ==
    its own coordinates, so every shift and therefore every effect is exactly 0, and the selection fails like the notebook
    (one class is empty, quirk Q7); an EMPTY coordinate subset; an odd max_batch; a batch larger than the coordinate count."""
    if precision == "bf16":
        _need_tc(tc_ok)
    G, clf, lat, noise = _small_tc_setup(dev, 3)
    S = G.num_style_coords
    one = sx.attfind_sweep(G, clf, lat[:1], noise, precision=precision, max_batch=64)
    # the generator output is bit-identical to the base image (zero shift); the tiny classifier is one cuDNN convolution whose
    # algorithm may depend on the batch size, hence a rounding-level bound instead of == 0 (the exact-zero property itself is
    # test_sweep_properties_at_baseline_sizes, with a batch-size independent classifier)
    assert one["style_change"].shape == (1, 2, S, 2) and float(one["style_change"].abs().max()) <= 1e-4
    assert torch.equal(one["minima"], one["maxima"])
    with pytest.raises(IndexError):
        sx.attfind_select(one["style_change"], one["base_prob"], 5, 0.5)
    none = sx.attfind_sweep(G, clf, lat, noise, precision=precision, sindices=[], max_batch=64)
    assert float(none["style_change"].abs().max()) == 0.0
    sind = list(range(100, 131)) + [S - 1]
    a = sx.attfind_sweep(G, clf, lat, noise, precision=precision, sindices=sind, max_batch=7)      # odd: rounded down to 6
    b = sx.attfind_sweep(G, clf, lat, noise, precision=precision, sindices=sind, max_batch=4096)   # one batch per conv
    # our kernels are batch-size independent; the tiny classifier is one cuDNN conv whose algorithm may not be
    assert float((a["style_change"] - b["style_change"]).abs().max()) <= 1e-4
    # images per classifier call: one generator launch each, or several launches gathered (flush in the middle of a latent)
    c = sx.attfind_sweep(G, clf, lat, noise, precision=precision, sindices=sind, max_batch=8, classify_batch=8)
    d = sx.attfind_sweep(G, clf, lat, noise, precision=precision, sindices=sind, max_batch=8, classify_batch=20)
    assert float((c["style_change"] - a["style_change"]).abs().max()) <= 1e-4
    assert float((d["style_change"] - a["style_change"]).abs().max()) <= 1e-4
    mask = torch.ones(S, dtype=torch.bool)
    mask[sind] = False
    assert float(a["style_change"][:, :, mask].abs().max()) == 0.0


def test_verify_topk_small_generator_equals_full_fp32_sweep(dev, tc_ok):
    """attfind_verify_topk end to end on a small tensor-core generator, every coordinate: the picks and the merged list of
    (bf16 sweep + verification) are those of the full fp32 sweep, and the oracle's selection on the hybrid effects agrees."""
    _need_tc(tc_ok)
    G, clf, lat, noise = _small_tc_setup(dev, 12, seed=13)
    r32 = sx.attfind_sweep(G, clf, lat, noise, precision="fp32", max_batch=256)
    want = sx.attfind_select(r32["style_change"], r32["base_prob"], 5, 0.5)
    r16 = sx.attfind_sweep(G, clf, lat, noise, precision="bf16", max_batch=256)
    picks, merged, scores, info = sx.attfind_verify_topk(G, clf, lat, noise, r16, 5, 0.5, precision="fp32", max_batch=64,
                                                         min_candidates=8)
    assert info["verified"]
    assert picks == want[0] and merged == want[1]
    ref = O.attfind_select(info["style_change"].cpu().numpy(), info["base_prob"].cpu().numpy(), 5, 0.5)
    assert picks == ref[0] and merged == ref[1]
    print(f"verification re-evaluated {info['exact_evals']} of {lat.shape[0] * 2 * G.num_style_coords} entries, {info['passes']} pass(es)")


def test_native_call_on_a_tensor_of_another_device_is_an_error(dev):
    """launches go to the CURRENT device's stream; a tensor that lives elsewhere must raise, not fault."""
    if torch.cuda.device_count() < 2:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            sx.modules.upsample2x(torch.zeros(1, 1, 2, 2))
        return
    other = torch.device("cuda", 1 if (dev.index or 0) == 0 else 0)
    with pytest.raises(RuntimeError, match="current CUDA device"):
        sx.modules.upsample2x(torch.zeros(1, 1, 2, 2, device=other))
    with torch.cuda.device(other):
        assert sx.modules.upsample2x(torch.ones(1, 1, 2, 2, device=other)).shape == (1, 1, 4, 4)


def test_attfind_extraction_with_verification_writes_records_that_select_like_fp32(dev, tc_ok, tmp_path):
    """the notebook-signature entry point in the throughput mode + ``verify_classifier``: the records it writes carry the hybrid
    effects, and the notebook's own selection on the LOADED records (cells 12, 14-16: ``load_records`` -> class split ->
    ``find_significant_styles`` per class) returns the picks of the same extraction run entirely in fp32."""
    _need_tc(tc_ok)
    G, clf, _, noise = _small_tc_setup(dev, 4, seed=17)
    size = G.image_size

    class Enc(torch.nn.Module):
        def forward(self, img):
            f = torch.nn.functional.adaptive_avg_pool2d(img, 16).reshape(img.shape[0], -1)[:, :512]
            return ((f - 0.5) * 6).squeeze()

    class Stylex:
        pass

    st = Stylex()
    st.G, st.encoder, st.D = G, Enc(), None
    g = torch.Generator().manual_seed(5)
    images = [torch.rand(1, 3, size, size, generator=g) for _ in range(12)]
    S = G.num_style_coords
    G.precision = "fp32"
    ref = sx.attfind_extraction(images, 12, None, st, clf, None, noise, S, 1.0, -0.5, image_size=size, precision="fp32", max_batch=128)
    labels = torch.argmax(ref["base_prob"], dim=1)
    if int((labels == 0).sum()) in (0, 12):
        pytest.skip("degenerate class split for this seed")
    want = sx.attfind_select(ref["style_change"], ref["base_prob"], 5, 0.5)
    out = sx.attfind_extraction(images, 12, str(tmp_path), st, clf, None, noise, S, 1.0, -0.5, image_size=size, precision="bf16",
                                max_batch=128, verify_classifier=clf)
    assert out["verify"]["verified"] and out["picks"] == want[0] and out["merged"] == want[1]
    rec = sx.load_records(os.path.join(str(tmp_path), "style_change_records.hdf5"))
    eff, base = rec["style_change"], rec["base_prob"]
    lab = np.argmax(base, axis=1)
    for c in (0, 1):                                                     # cells 14-16 on the loaded arrays
        per_class = eff[lab == c].astype(np.float64)
        got = sx.find_significant_styles(per_class, 5, c, max_image_effect=2.5, device=dev)
        assert got == want[0][c], (c, got, want[0][c])


@pytest.mark.parametrize("fuse_pool", [False, True])
@pytest.mark.parametrize("shape", [(3, 115, 115), (2, 23, 30), (1, 4, 4), (5, 35, 19), (0, 8, 8)])
def test_native_stem_kernel_matches_conv_relu_pool(dev, tc_ok, shape, fuse_pool):
    """sx_stem_s2d_conv_relu (tcgen05, 4x4 taps on the 16-channel space-to-depth input, SWIZZLE_32B halo boxes) against
    relu(conv2d + bias) in fp32 on the same bf16 operands, rounded to bf16 once, then max_pool2d: equal up to the order of
    the fp32 accumulation (<= 1 bf16 ulp on a few elements), ragged sizes and the pool's borders included."""
    _need_tc(tc_ok)
    from stylex_b200.classifiers import FusedResNetInference
    b, h, w = shape
    g = torch.Generator().manual_seed(h * 131 + w)
    x = torch.randn(b, 16, h, w, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(64, 16, 4, 4, generator=g) * 0.1).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(64, generator=g).to(dev).to(torch.bfloat16)
    f = FusedResNetInference.__new__(FusedResNetInference)
    f.dtype, f.stem_s2d, f.native_stem = torch.bfloat16, (wt, bias), None
    f.enable_native_stem(fuse_pool=fuse_pool)
    got = f._stem_native(x)
    ref = torch.relu(torch.nn.functional.conv2d(x.float(), wt.float(), bias.float())).to(torch.bfloat16)
    if fuse_pool:
        ref = torch.nn.functional.max_pool2d(ref, 3, 2, 1)
    assert got.shape == ref.shape and got.dtype == torch.bfloat16
    if b == 0:
        return
    assert got.is_contiguous(memory_format=torch.channels_last)
    diff = (got.float() - ref.float()).abs()
    assert float(diff.max()) <= 2.0 ** -7 * max(1.0, float(ref.float().abs().max())), float(diff.max())
    assert float((diff > 0).float().mean()) < 0.02


def test_native_stem_in_the_classifier(dev, tc_ok):
    """configure_throughput's default (bf16, native stem with the pool fused) stays within the reduced-precision tolerance
    of the fp32 module, reports the native stem as active, and matches the cuDNN s2d stem it replaces far more closely."""
    _need_tc(tc_ok)
    model = synthetic.make_classifier_model("resnet", 5)
    g = torch.Generator().manual_seed(1)
    imgs = (torch.rand(6, 3, 256, 256, generator=g) * 2 - 0.5)
    clf = sx.make_classifier("resnet", model, 256)
    synthetic.calibrate_classifier(model, clf.preprocess, imgs, chunk=6)
    clf.to(dev)
    x = imgs.to(dev)
    ref32 = clf.classify_images(x).float()
    info = clf.configure_throughput(x)
    assert "sx_stem_s2d_conv_relu" in info["classifier_mode"] and "fused into the stem" in info["classifier_mode"], info
    assert clf.fused.native_stem is not None
    got = clf.classify_images(x).float()
    # bf16 network vs the fp32 module: the bound configure_throughput itself accepts a reduced-precision network under
    assert float((got - ref32).abs().max()) <= 3 * (0.05 * float(ref32.abs().max()) + 0.05)
    stem = clf.fused.native_stem
    clf.fused.native_stem = None
    cudnn = clf.classify_images(x).float()
    clf.fused.native_stem = stem
    assert float((got - cudnn).abs().max()) <= 2e-2 * max(1.0, float(ref32.abs().max()))
