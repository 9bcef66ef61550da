"""Seeded synthetic inputs for the AttFind hot path.

There is no network for datasets or checkpoints, so every BASELINE config runs on random-init
weights of the reference's architecture and synthetic latents.  Everything is drawn from a
``torch.Generator`` on the CPU so the same seed gives the same tensors on every box.

Shapes follow ``Generator.__init__`` (reference ``stylex/stylex_train.py:748-792``); the init
follows ``StylEx._init_weights`` (``stylex_train.py:974-983``: kaiming-normal, fan_in, gain
sqrt(2)) except ``to_noise*.weight ~ N(0, 0.1)`` instead of zeros so the noise path is exercised.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Tuple

import torch


def generator_pairs(image_size: int, network_capacity: int = 16, fmap_max: int = 512) -> List[Tuple[int, int]]:
    """[(Ci, Co)] per GeneratorBlock -- reference stylex_train.py:753-763."""
    num_layers = int(math.log2(image_size) - 1)
    filters = [network_capacity * (2 ** (i + 1)) for i in range(num_layers)][::-1]
    filters = [min(fmap_max, f) for f in filters]
    filters = [filters[0], *filters]
    return list(zip(filters[:-1], filters[1:]))


def num_style_coords(image_size: int, network_capacity: int = 16, fmap_max: int = 512) -> int:
    return sum(ci + co for ci, co in generator_pairs(image_size, network_capacity, fmap_max))


def _kaiming(gen: torch.Generator, *shape: int) -> torch.Tensor:
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    return torch.randn(*shape, generator=gen) * math.sqrt(2.0 / fan_in)


def _bias(gen: torch.Generator, n: int, fan_in: int) -> torch.Tensor:
    bound = 1.0 / math.sqrt(fan_in)
    return (torch.rand(n, generator=gen) * 2 - 1) * bound


def make_generator_state(image_size: int, seed: int = 42, latent_dim: int = 514, network_capacity: int = 16,
                         fmap_max: int = 512, noise_std: float = 0.1) -> "OrderedDict[str, torch.Tensor]":
    """State dict with the reference Generator's key names, fp32, CPU."""
    g = torch.Generator().manual_seed(seed)
    pairs = generator_pairs(image_size, network_capacity, fmap_max)
    c0 = pairs[0][0]
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["initial_block"] = torch.randn(1, c0, 4, 4, generator=g)
    sd["initial_conv.weight"] = _kaiming(g, c0, c0, 3, 3)
    sd["initial_conv.bias"] = _bias(g, c0, c0 * 9)
    for i, (ci, co) in enumerate(pairs):
        p = f"blocks.{i}."
        sd[p + "to_style1.weight"] = _kaiming(g, ci, latent_dim)
        sd[p + "to_style1.bias"] = _bias(g, ci, latent_dim)
        sd[p + "to_noise1.weight"] = torch.randn(co, 1, generator=g) * noise_std
        sd[p + "to_noise1.bias"] = torch.randn(co, generator=g) * noise_std
        sd[p + "conv1.weight"] = _kaiming(g, co, ci, 3, 3)
        sd[p + "to_style2.weight"] = _kaiming(g, co, latent_dim)
        sd[p + "to_style2.bias"] = _bias(g, co, latent_dim)
        sd[p + "to_noise2.weight"] = torch.randn(co, 1, generator=g) * noise_std
        sd[p + "to_noise2.bias"] = torch.randn(co, generator=g) * noise_std
        sd[p + "conv2.weight"] = _kaiming(g, co, co, 3, 3)
        sd[p + "to_rgb.to_style.weight"] = _kaiming(g, co, latent_dim)
        sd[p + "to_rgb.to_style.bias"] = _bias(g, co, latent_dim)
        sd[p + "to_rgb.conv.weight"] = _kaiming(g, 3, co, 1, 1)
    return sd


def discriminator_filters(image_size: int, network_capacity: int = 16, fmap_max: int = 512) -> List[Tuple[int, int]]:
    """[(Ci, Co)] per DiscriminatorBlock -- reference stylex_train.py:846-853."""
    num_layers = int(math.log2(image_size) - 1)
    filters = [3] + [min(fmap_max, (network_capacity * 4) * (2 ** i)) for i in range(num_layers + 1)]
    return list(zip(filters[:-1], filters[1:]))


def make_discriminator_state(image_size: int, seed: int = 42, network_capacity: int = 16, fmap_max: int = 512,
                             encoder: bool = False, encoder_dim: int = 512) -> "OrderedDict[str, torch.Tensor]":
    """State dict with the reference DiscriminatorE's key names (stylex_train.py:721-744, 842-887; SURVEY.md
    Appendix B), kaiming weights + uniform biases like nn.Conv2d / nn.Linear under StylEx._init_weights."""
    g = torch.Generator().manual_seed(seed + 3000017)
    pairs = discriminator_filters(image_size, network_capacity, fmap_max)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()

    def conv(key, co, ci, k):
        sd[key + ".weight"] = _kaiming(g, co, ci, k, k)
        sd[key + ".bias"] = _bias(g, co, ci * k * k)

    for i, (ci, co) in enumerate(pairs):
        p = f"blocks.{i}."
        conv(p + "conv_res", co, ci, 1)
        conv(p + "net.0", co, ci, 3)
        conv(p + "net.2", co, co, 3)
        if i != len(pairs) - 1:
            conv(p + "downsample.1", co, co, 3)
    c = pairs[-1][1]
    conv("final_conv", c, c, 3)
    out_dim = encoder_dim if encoder else 1
    sd["fc.weight"] = _kaiming(g, out_dim, 4 * c)
    sd["fc.bias"] = _bias(g, out_dim, 4 * c)
    return sd


def make_latents(n: int, seed: int = 42, latent_dim: int = 514) -> torch.Tensor:
    """w[N,512] ~ N(0,1) (+) 2 'classifier logits' ~ N(0,1): the concat_w_tensor of the notebook."""
    g = torch.Generator().manual_seed(seed + 1000003)
    return torch.randn(n, latent_dim, generator=g)


def make_noise(image_size: int, seed: int = 42) -> torch.Tensor:
    """image_noise(1, S): U[0,1) [1,S,S,1] (reference stylex_train.py:336-337)."""
    g = torch.Generator().manual_seed(seed + 2000003)
    return torch.rand(1, image_size, image_size, 1, generator=g)


def make_classifier_model(kind: str, seed: int = 42) -> torch.nn.Module:
    """torchvision resnet18 / mobilenet_v2, weights=None, 2-way head, eval mode (seeded init)."""
    import torchvision

    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed + 3000003)
        if kind == "resnet":
            m = torchvision.models.resnet18(weights=None)
            m.fc = torch.nn.Linear(512, 2)
        elif kind == "mobilenet":
            m = torchvision.models.mobilenet_v2(weights=None)
            m.classifier[1] = torch.nn.Linear(1280, 2)
        else:
            raise ValueError(f"unknown classifier kind {kind!r}")
    for p in m.parameters():
        p.requires_grad = False
    return m.eval()


def _head(model: torch.nn.Module) -> torch.nn.Linear:
    return model.fc if hasattr(model, "fc") else model.classifier[1]


@torch.no_grad()
def calibrate_classifier(model: torch.nn.Module, preprocess, images: torch.Tensor, target_std: float = 1.0,
                         chunk: int = 32) -> torch.nn.Module:
    """Make a random-init classifier non-degenerate (hard part 5).

    One pass over ``images`` in train mode sets the BatchNorm running statistics (momentum=None:
    cumulative average, so the result does not depend on chunking order beyond fp rounding); the
    2-way head is then rescaled / recentred so the logit *difference* over ``images`` has median 0
    (both classes populated) and standard deviation ``target_std``.
    ``preprocess`` is the wrapper's resize+normalise (everything of classify_images but the model).
    """
    bns = [m for m in model.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
    saved = [b.momentum for b in bns]
    for b in bns:
        b.reset_running_stats()
        b.momentum = None
    model.train()
    for i in range(0, images.shape[0], chunk):
        model(preprocess(images[i: i + chunk]))
    model.eval()
    for b, m in zip(bns, saved):
        b.momentum = m
    logits = torch.cat([model(preprocess(images[i: i + chunk])) for i in range(0, images.shape[0], chunk)])
    diff = (logits[:, 1] - logits[:, 0]).double()
    std = float(diff.std()) or 1.0
    head = _head(model)
    scale = target_std / std
    head.weight.mul_(scale)
    head.bias.mul_(scale)
    med = float(diff.median()) * scale
    head.bias[1] -= med / 2
    head.bias[0] += med / 2
    return model
