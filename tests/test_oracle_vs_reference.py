"""Runs only where the reference checkout exists (the build container): the oracle and the product's classifier
wrappers against the UNMODIFIED reference code, beyond what the committed fixtures cover."""
import numpy as np
import pytest
import torch

import stylex_b200 as sx
from stylex_b200 import synthetic
from oracle import ref_loader as RL, stylex_oracle as O

pytestmark = pytest.mark.skipif(not RL.available(), reason="reference checkout not present (GPU box)")
torch.set_grad_enabled(False)


def test_oracle_ops_bit_match_reference_modules():
    R = RL.load_reference()
    g = torch.Generator().manual_seed(0)
    for ci, co, k, demod in ((8, 12, 3, True), (16, 3, 1, False)):
        m = R.Conv2DMod(ci, co, k, demod=demod)
        x, y = torch.randn(2, ci, 8, 8, generator=g), torch.randn(2, ci, generator=g)
        assert torch.equal(m(x, y), O.modconv(x, m.weight.detach(), y, demod=demod))
    x = torch.randn(2, 3, 8, 8, generator=g)
    assert torch.equal(R.Blur()(x), O.blur3x3_reflect(x))


def test_oracle_generator_matches_reference_with_shift():
    """the oracle's functional coord_shift == the notebook's `bias += shift` patch (NB:381) up to fp32 rounding."""
    sd = synthetic.make_generator_state(16, seed=2, network_capacity=4)
    G = RL.reference_generator(sd, 16, network_capacity=4)
    w = synthetic.make_latents(1, 2)
    noise = synthetic.make_noise(16, 2)
    styles = O.styles_def_to_tensor([(w, G.num_layers)])
    S = O.num_style_coords(sd)
    for sindex in (0, 31, 32, 70, S - 1):
        shift = torch.zeros(1, S)
        shift[0, sindex] = 0.7
        mine = O.generator_forward(sd, styles, noise, coord_shift=shift)
        blk, idx = O.sindex_to_block_idx_and_index(O.generator_layout(sd), sindex)
        b = G.blocks[blk]
        layer, j = (b.to_style1, idx) if idx < b.input_channels else (b.to_style2, idx - b.input_channels)
        layer.bias[j] += 0.7
        ref = G(styles, noise)
        layer.bias[j] -= 0.7
        assert (ref - mine).abs().max().item() <= 2e-6


def test_product_classifier_wrappers_match_reference_wrappers():
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(), torch.nn.Linear(4, 2)).eval()
    x = torch.rand(2, 3, 64, 64) * 2 - 0.5
    for kind in ("resnet", "mobilenet"):
        ref = RL.reference_classifier(kind, net, 64).classify_images(x)
        mine = sx.make_classifier(kind, net, 64).classify_images(x)
        assert torch.equal(ref, mine), kind


def test_oracle_discriminator_matches_reference_module():
    """DiscriminatorE as encoder and as discriminator at 32px / capacity 8, fresh seeds (beyond frontend_small.npz)."""
    R = RL.load_reference()
    g = torch.Generator().manual_seed(5)
    images = torch.rand(3, 3, 32, 32, generator=g)
    for enc in (True, False):
        m = R.DiscriminatorE(32, 8, encoder=enc).eval()
        sd = synthetic.make_discriminator_state(32, seed=77, network_capacity=8, encoder=enc)
        r = m.load_state_dict(sd, strict=False)
        assert not r.unexpected_keys and all(k.endswith(".f") for k in r.missing_keys)
        assert (m(images) - O.discriminator_forward(sd, images)).abs().max().item() <= 2e-6
        assert m(images[:1]).shape == O.discriminator_forward(sd, images[:1]).shape


def test_product_stylex_loads_a_reference_state_dict_strictly():
    """a state dict produced by the reference StylEx class loads into the drop-in container with strict=True, and the
    reverse (what Trainer.load does with a checkpoint written by save_checkpoint)."""
    R = RL.load_reference()
    real = torch.cuda.is_available
    torch.cuda.is_available = lambda: False
    try:
        with RL.cpu_cuda_identity():
            ref = R.StylEx(image_size=32, network_capacity=4)
        mine = sx.StylEx(image_size=32, network_capacity=4)
    finally:
        torch.cuda.is_available = real
    mine.load_state_dict(ref.state_dict())                   # strict
    ref.load_state_dict(mine.state_dict())
    for (k, a), (k2, b) in zip(ref.state_dict().items(), mine.state_dict().items()):
        assert k == k2 and torch.equal(a, b), k


@pytest.mark.parametrize("ci,co,k,demod,hw", [(6, 10, 3, True, 5), (12, 3, 1, False, 7), (32, 32, 3, True, 8)])
def test_oracle_conv_gradients_match_reference_autograd(ci, co, k, demod, hw):
    R = RL.load_reference()
    g = torch.Generator().manual_seed(ci * 10 + hw)
    m = R.Conv2DMod(ci, co, k, demod=demod)
    x, y = torch.randn(2, ci, hw, hw, generator=g), torch.randn(2, ci, generator=g) * 0.6
    go = torch.randn(2, co, hw, hw, generator=g)
    with torch.enable_grad():
        xs, ys = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
        ref = torch.autograd.grad(m(xs, ys), (xs, ys, m.weight), go)
    _, gx, gy, gw = O.modconv_grads(x, m.weight.detach(), y, go, demod=demod)
    for got, r in zip((gx, gy, gw), ref):
        assert (got.float() - r).abs().max().item() <= 2e-5 * max(1.0, r.abs().max().item())


def test_filter_unstable_images_matches_notebook_cell():
    ns = RL.notebook_namespace(True)
    rng = np.random.RandomState(3)
    eff = rng.randn(8, 2, 60, 2) * 0.25
    eff[5] *= 6
    a = ns["filter_unstable_images"](eff.copy(), 0.3, 100)
    b = sx.filter_unstable_images(eff.copy(), 0.3, 100)
    assert np.array_equal(a, b) and np.all(b[5] == 0) and np.any(b[0] != 0)
