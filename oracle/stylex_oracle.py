"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's AttFind hot path.

This file is the *oracle*: a plain torch-CPU / numpy restatement of the algorithm in
NoahVl/Explaining-In-Style-Reproducibility-Study (``R/`` = the reference checkout).  It is
NOT part of the product: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker.  The
product path (``explaining-in-style-reproducibility-study_b200``) never imports this module and
fails loudly when its CUDA library is missing.

Parity pinning: every function here is checked against the *imported, unmodified* reference
classes / notebook cells in this container by ``tests/golden/make_golden.py`` (which also
writes the committed fixtures under ``tests/golden/``) and by ``tests/test_oracle_vs_reference.py``
(runs only where ``/root/reference`` exists).  One boundary stays **parity unpinned**: the blur
(``kornia.filters.filter2d``, kornia==0.6.2 per ``R/environment.yml:211``) is third-party code
that is absent from ``R/``; ``blur3x3_reflect`` restates its published algorithm (normalised
kernel, 'reflect' border, depthwise correlation) -- call site ``R/stylex/stylex_train.py:153``.

Citations: ST = ``R/stylex/stylex_train.py``; NB = ``R/stylex/run_attfind_combined.ipynb``
(raw JSON line numbers).

The oracle works on a *state dict* (``{key: tensor}`` with the reference's key names, see
SURVEY.md Appendix B) so it is independent of every module class in the product.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------
# L1 ops
# --------------------------------------------------------------------------------------
def modconv(x: Tensor, weight: Tensor, style: Tensor, demod: bool = True, eps: float = 1e-8) -> Tensor:
    """Conv2DMod.forward, ST:647-667 -- literal restatement (per-sample weights, grouped conv).

    x [B,Ci,H,W], weight [Co,Ci,k,k], style [B,Ci] -> [B,Co,H,W].
    """
    b, c, h, w = x.shape
    co, ci, k, _ = weight.shape
    w1 = style[:, None, :, None, None]                     # ST:650
    w2 = weight[None, :, :, :, :]                          # ST:651
    weights = w2 * (w1 + 1)                                # ST:652
    if demod:
        d = torch.rsqrt((weights ** 2).sum(dim=(2, 3, 4), keepdim=True) + eps)   # ST:655
        weights = weights * d                              # ST:656
    x = x.reshape(1, -1, h, w)                             # ST:658
    weights = weights.reshape(b * co, ci, k, k)            # ST:660-661
    padding = (k - 1) // 2                                 # ST:644-645,663 (stride=1, dilation=1)
    x = F.conv2d(x, weights, padding=padding, groups=b)    # ST:664
    return x.reshape(-1, co, h, w)                         # ST:666


def upsample2x(x: Tensor) -> Tensor:
    """nn.Upsample(scale_factor=2, mode='bilinear', align_corners=False), ST:614,679."""
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)


def blur3x3_reflect(x: Tensor) -> Tensor:
    """Blur.forward ST:144-153 -> kornia.filters.filter2d(x, [1,2,1] (x) [1,2,1], normalized=True).

    kornia 0.6.2 algorithm (restated, third-party, parity unpinned): kernel / sum(|kernel|),
    pad 1 px with border_type='reflect', depthwise cross-correlation, same output size.
    """
    f = torch.tensor([1.0, 2.0, 1.0], dtype=x.dtype, device=x.device)
    k = f[None, :] * f[:, None]
    k = k / k.abs().sum()
    c = x.shape[1]
    xp = F.pad(x, (1, 1, 1, 1), mode="reflect")
    return F.conv2d(xp, k[None, None].expand(c, 1, 3, 3).contiguous(), groups=c)


def leaky_relu(x: Tensor) -> Tensor:
    """leaky_relu(0.2), ST:340-341."""
    return F.leaky_relu(x, 0.2)


# --------------------------------------------------------------------------------------
# L2 networks, driven by a state dict with the reference's key names
# --------------------------------------------------------------------------------------
def generator_layout(params: Dict[str, Tensor]) -> List[Tuple[int, int]]:
    """[(Ci, Co)] per GeneratorBlock, read from the state dict (Generator.__init__ ST:748-792)."""
    pairs = []
    i = 0
    while f"blocks.{i}.conv1.weight" in params:
        co, ci = params[f"blocks.{i}.conv1.weight"].shape[:2]
        pairs.append((int(ci), int(co)))
        i += 1
    return pairs


def num_style_coords(params: Dict[str, Tensor]) -> int:
    """sum over blocks of (input_channels + filters), ST:677."""
    return sum(ci + co for ci, co in generator_layout(params))


def generator_block(params, i: int, x: Tensor, prev_rgb: Optional[Tensor], istyle: Tensor, inoise: Tensor,
                    upsample: bool, upsample_rgb: bool, coord_shift: Optional[Tensor] = None):
    """GeneratorBlock.forward ST:692-718 (+ RGBBlock.forward ST:618-629).

    coord_shift: optional [B, Ci+Co] additive shift of this block's style coordinates -- the
    functional form of the notebook's ``to_style{1,2}.bias += shift`` (NB:381).
    """
    p = lambda k: params[f"blocks.{i}.{k}"]
    if upsample:
        x = upsample2x(x)                                                        # ST:693-694
    inoise = inoise[:, : x.shape[2], : x.shape[3], :]                            # ST:696
    noise1 = F.linear(inoise, p("to_noise1.weight"), p("to_noise1.bias")).permute((0, 3, 2, 1))   # ST:697
    noise2 = F.linear(inoise, p("to_noise2.weight"), p("to_noise2.bias")).permute((0, 3, 2, 1))   # ST:698
    ci = p("conv1.weight").shape[1]
    style1 = F.linear(istyle, p("to_style1.weight"), p("to_style1.bias"))       # ST:700
    if coord_shift is not None:
        style1 = style1 + coord_shift[:, :ci]
    x = modconv(x, p("conv1.weight"), style1, demod=True)                        # ST:704
    x = leaky_relu(x + noise1)                                                   # ST:705
    style2 = F.linear(istyle, p("to_style2.weight"), p("to_style2.bias"))       # ST:707
    if coord_shift is not None:
        style2 = style2 + coord_shift[:, ci:]
    style_coords = torch.cat([style1, style2], dim=-1)                           # ST:709
    x = modconv(x, p("conv2.weight"), style2, demod=True)                        # ST:713
    x = leaky_relu(x + noise2)                                                   # ST:714
    # RGBBlock ST:618-629
    rstyle = F.linear(istyle, p("to_rgb.to_style.weight"), p("to_rgb.to_style.bias"))   # ST:620
    rgb = modconv(x, p("to_rgb.conv.weight"), rstyle, demod=False)               # ST:621
    if prev_rgb is not None:
        rgb = rgb + prev_rgb                                                     # ST:623-624
    if upsample_rgb:
        rgb = blur3x3_reflect(upsample2x(rgb))                                   # ST:626-627
    return x, rgb, style_coords


def generator_forward(params: Dict[str, Tensor], styles: Tensor, input_noise: Tensor,
                      get_style_coords: bool = False, coord_shift: Optional[Tensor] = None):
    """Generator.forward ST:794-825.  styles [B,L,latent], input_noise [B|1,S,S,1].

    coord_shift: optional [B, S_total] additive shift in StyleSpace (see generator_block).
    """
    pairs = generator_layout(params)
    L = len(pairs)
    b = styles.shape[0]
    x = params["initial_block"].expand(b, -1, -1, -1)                            # ST:802
    x = F.conv2d(x, params["initial_conv.weight"], params["initial_conv.bias"], padding=1)   # ST:806
    rgb = None
    coords = []
    off = 0
    for i, (ci, co) in enumerate(pairs):                                         # ST:811-818
        sh = None if coord_shift is None else coord_shift[:, off: off + ci + co]
        x, rgb, sc = generator_block(params, i, x, rgb, styles[:, i], input_noise,
                                     upsample=(i != 0), upsample_rgb=(i != L - 1), coord_shift=sh)
        coords.append(sc)
        off += ci + co
    if get_style_coords:
        return rgb, torch.cat(coords, dim=1)                                     # ST:820-822
    return rgb


def styles_def_to_tensor(styles_def):
    """ST:352-353."""
    return torch.cat([t[:, None, :].expand(-1, n, -1) for t, n in styles_def], dim=1)


# --------------------------------------------------------------------------------------
# L4 AttFind sweep (NB cell 5) -- functional, drift-free form of the bias-patching loop
# --------------------------------------------------------------------------------------
def sindex_to_block_idx_and_index(pairs: Sequence[Tuple[int, int]], sindex: int) -> Tuple[int, int]:
    """NB:220-234."""
    tmp = sindex
    for idx, (ci, co) in enumerate(pairs):
        if tmp < ci + co:
            return idx, tmp
        tmp -= ci + co
    raise IndexError(sindex)


def get_min_max_style_vectors(style_coordinates: Tensor) -> Tuple[Tensor, Tensor]:
    """NB:237-252 (elementwise min / max over the image axis)."""
    return style_coordinates.min(dim=0).values, style_coordinates.max(dim=0).values


@torch.no_grad()
def attfind_sweep(params: Dict[str, Tensor], classify: Callable[[Tensor], Tensor], latents: Tensor,
                  noise: Tensor, shift_size: float = 1.0, sindices: Optional[Sequence[int]] = None,
                  image_indices: Optional[Sequence[int]] = None, minmax_from: Optional[Tensor] = None):
    """attfind_extraction phases A(second half)-C, NB:316-389, batch 1, one G forward per coord-eval.

    latents [N,514] play the role of ``concat_w_tensor`` (NB:311-314).  Returns a dict with the
    notebook's dataset names (NB:395-403): 'style_change' [N,2,S,2], 'base_prob' [N,2] (raw
    logits, quirk Q3), 'style_coordinates' [N,S], 'minima'/'maxima' [S], 'latents'.
    ``sindices`` / ``image_indices`` restrict the loop (bounded CPU samples); untouched entries
    of 'style_change' stay 0.  Drift-free: the shift is applied functionally, never by mutating
    weights (quirk Q2).
    """
    pairs = generator_layout(params)
    L = len(pairs)
    S = sum(ci + co for ci, co in pairs)
    N = latents.shape[0]
    style_coordinates = torch.zeros(N, S)
    base_logits = torch.zeros(N, 2)
    for n in range(N):                                                           # NB:300-336
        w = styles_def_to_tensor([(latents[n: n + 1], L)])
        img, sc = generator_forward(params, w, noise, get_style_coords=True)     # NB:318
        style_coordinates[n] = sc[0]
        base_logits[n] = classify(img)[0]                                        # NB:334
    src = style_coordinates if minmax_from is None else minmax_from
    minima, maxima = get_min_max_style_vectors(src)                              # NB:340
    effects = torch.zeros(N, 2, S, 2)
    s_list = range(S) if sindices is None else sindices
    n_list = range(N) if image_indices is None else image_indices
    for n in n_list:                                                             # NB:346
        w = styles_def_to_tensor([(latents[n: n + 1], L)])
        for s in s_list:                                                         # NB:356
            for d, target in enumerate((minima, maxima)):                        # NB:374-377
                shift = torch.zeros(1, S)
                shift[0, s] = (target[s] - style_coordinates[n, s]) * shift_size
                img = generator_forward(params, w, noise, coord_shift=shift)     # NB:381-382
                effects[n, d, s] = classify(img)[0] - base_logits[n]             # NB:384-385
    return {
        "style_change": effects, "latents": latents.clone(), "base_prob": base_logits,
        "minima": minima, "maxima": maxima, "style_coordinates": style_coordinates,
    }


# --------------------------------------------------------------------------------------
# L4 AttFind selection (NB cells 14, 15, 16) -- numpy, float64 exactly like the notebook
# --------------------------------------------------------------------------------------
def split_by_class(style_change_effect: np.ndarray, base_probs: np.ndarray) -> Dict[int, np.ndarray]:
    """NB:695-714: argmax label per image; per-class float64 copy of the effects."""
    all_labels = np.argmax(base_probs, axis=1)
    out = {}
    for c in range(2):
        idx = np.array([i for i in range(all_labels.shape[0]) if all_labels[i] == c], dtype=np.int64)
        cur = np.zeros((len(idx),) + style_change_effect.shape[1:])          # float64 buffer NB:703
        for k, i in enumerate(idx):
            cur[k] = style_change_effect[i]                                  # NB:707
        out[c] = cur
    return out


def find_significant_styles(style_change_effect: np.ndarray, num_indices: int, class_index: int,
                            max_image_effect: float = 0.2, sindex_offset: int = 0) -> List[Tuple[int, int]]:
    """NB:731-758 (unused generator/classifier/dlatent/min/max arguments dropped)."""
    num_images = style_change_effect.shape[0]
    S = style_change_effect.shape[2]
    eff = np.maximum(0, style_change_effect[:, :, :, class_index].reshape((num_images, -1)))   # NB:745
    images_effect = np.zeros(num_images)
    picks: List[int] = []
    while len(picks) < num_indices:                                          # NB:751
        with np.errstate(invalid="ignore"), __import__("warnings").catch_warnings():
            __import__("warnings").simplefilter("ignore")
            col_mean = np.mean(eff[images_effect < max_image_effect], axis=0)   # empty mask -> NaN (quirk Q7)
        next_s = int(np.argmax(col_mean))                                    # NB:752 first-max tie-break
        picks.append(next_s)
        images_effect += eff[:, next_s]                                      # NB:755
        eff[:, next_s] = 0                                                   # NB:756
    return [(x // S, (x % S) + sindex_offset) for x in picks]                # NB:758


def attfind_select(style_change_effect: np.ndarray, base_probs: np.ndarray, num_indices: int = 5,
                   effect_threshold: float = 0.5):
    """NB:775-814: per-class greedy picks + merged ranking.

    Returns (picks_per_class {0: [(dir, s)], 1: [...]}, merged [(dir, s)] sorted by score desc,
    scores in the merged order).
    """
    classes = split_by_class(style_change_effect, base_probs)
    picks = {}
    for c in (0, 1):
        picks[c] = find_significant_styles(classes[c], num_indices, c, max_image_effect=effect_threshold * 5)
    s0 = [s for _, s in picks[0]]
    joined = [(1 - d, s) for d, s in picks[1] if s not in s0]                # NB:802
    joined += picks[0]                                                       # NB:803
    scores = []
    for d, s in joined:                                                      # NB:806-809
        od = 1 if d == 0 else 0
        scores.append(np.mean(style_change_effect[:, d, s, 0]) + np.mean(style_change_effect[:, od, s, 1]))
    order = np.argsort(scores)[::-1]                                         # NB:811
    merged = [joined[i] for i in order]
    return picks, merged, [float(scores[i]) for i in order]


# --------------------------------------------------------------------------------------
# L5 counterfactual rendering (NB cells 17-19) -- the step after the selection
# --------------------------------------------------------------------------------------
@torch.no_grad()
def generate_change_image_given_dlatent(params: Dict[str, Tensor], classify: Callable[[Tensor], Tensor], dlatent: Tensor,
                                        class_index: int, sindex: int, s_style_min: float, s_style_max: float,
                                        style_direction_index: int, shift_size: float, noise: Tensor):
    """NB cell 17 (``generate_change_image_given_dlatent``), functional form of its bias patch.

    dlatent [B,514].  Style coordinate ``sindex`` moves by ``(target - coord) * shift_size`` with target = the
    coordinate's minimum (direction 0) or maximum (direction 1); returns (images [B,3,S,S], softmax probability of
    ``class_index`` per image [B]).  The notebook only ever passes B = 1 and returns element 0 of the probabilities.
    """
    pairs = generator_layout(params)
    S = sum(ci + co for ci, co in pairs)
    w = styles_def_to_tensor([(dlatent, len(pairs))])
    _, coords = generator_forward(params, w, noise, get_style_coords=True)
    target = s_style_min if style_direction_index == 0 else s_style_max
    shift = torch.zeros(dlatent.shape[0], S)
    shift[:, sindex] = (float(target) - coords[:, sindex]) * shift_size
    images = generator_forward(params, w, noise, coord_shift=shift)
    probs = torch.softmax(classify(images), dim=1)[:, class_index]
    return images, probs


def draw_on_image(image: np.ndarray) -> np.ndarray:
    """NB cell 18 with its text drawing commented out, as in the reference: CHW float -> HWC uint8 of clip(x, 0, 1) * 255."""
    image = np.clip(np.transpose(image, (1, 2, 0)), 0, 1)
    return (image * 255).astype(np.uint8)


@torch.no_grad()
def generate_images_given_dlatent(params, classify, dlatent: Tensor, class_index: int, sindex: int, s_style_min: float,
                                  s_style_max: float, style_direction_index: int, noise: Tensor, shift_size: float = 2):
    """NB cell 19 (``draw_results_on_image=True`` branch): (panel uint8 [res, 2*res, 3], change_prob, base_prob)."""
    pairs = generator_layout(params)
    w = styles_def_to_tensor([(dlatent, len(pairs))])
    base = generator_forward(params, w, noise)
    base_prob = float(torch.softmax(classify(base), dim=1)[0, class_index])
    change, prob = generate_change_image_given_dlatent(params, classify, dlatent, class_index, sindex, s_style_min,
                                                       s_style_max, style_direction_index, shift_size, noise)
    res = base.shape[-1]
    panel = np.zeros((res, 2 * res, 3), np.uint8)
    panel[:, :res] = draw_on_image(base[0].numpy())
    panel[:, res:] = draw_on_image(change[0].numpy())
    return panel, float(prob[0]), base_prob


# --------------------------------------------------------------------------------------
# phase A front end: encoder / discriminator
# --------------------------------------------------------------------------------------
def discriminator_forward(params: Dict[str, Tensor], x: Tensor) -> Tensor:
    """DiscriminatorE.forward ST:889-909 over DiscriminatorBlock.forward ST:738-744, on a state dict with the
    reference's keys (``blocks.i.conv_res / net.0 / net.2 / downsample.1``, ``final_conv``, ``fc``); encoder and
    discriminator differ only in the width of ``fc`` (ST:884-887).  fq / attention blocks are default-off."""
    n_blocks = len({k.split(".")[1] for k in params if k.startswith("blocks.")})
    for i in range(n_blocks):
        pre = f"blocks.{i}."
        down = (pre + "downsample.1.weight") in params                                   # ST:733-736: all but the last
        res = F.conv2d(x, params[pre + "conv_res.weight"], params[pre + "conv_res.bias"], stride=2 if down else 1)
        x = leaky_relu(F.conv2d(x, params[pre + "net.0.weight"], params[pre + "net.0.bias"], padding=1))
        x = leaky_relu(F.conv2d(x, params[pre + "net.2.weight"], params[pre + "net.2.bias"], padding=1))
        if down:
            x = F.conv2d(blur3x3_reflect(x), params[pre + "downsample.1.weight"], params[pre + "downsample.1.bias"],
                         padding=1, stride=2)
        x = (x + res) * (1 / math.sqrt(2))                                               # ST:743
    x = F.conv2d(x, params["final_conv.weight"], params["final_conv.bias"], padding=1)  # ST:904
    x = x.reshape(x.shape[0], -1)                                                        # Flatten ST:905
    return F.linear(x, params["fc.weight"], params["fc.bias"]).squeeze()                 # ST:907-909


def encode_images(enc_params: Dict[str, Tensor], classify: Callable[[Tensor], Tensor], images: Tensor,
                  use_old_architecture: bool = True) -> Tuple[Tensor, Tensor]:
    """NB:300-314, image by image like the notebook: w = encoder(image); logits = classify(image);
    concat_w = cat(w, logits) (old architecture) or cat(w, softmax(logits)).  Returns (latents [N,514], logits)."""
    lat, lgs = [], []
    for i in range(images.shape[0]):
        w = discriminator_forward(enc_params, images[i: i + 1]).unsqueeze(0)             # NB:306
        lg = classify(images[i: i + 1])                                                  # NB:307
        lat.append(torch.cat((w, lg if use_old_architecture else torch.softmax(lg, dim=1)), dim=1))
        lgs.append(lg)
    return torch.cat(lat), torch.cat(lgs)


# --------------------------------------------------------------------------------------
# training-step slice: gradients of Conv2DMod
# --------------------------------------------------------------------------------------
def modconv_grads(x: Tensor, weight: Tensor, style: Tensor, grad_out: Tensor, demod: bool = True, eps: float = 1e-8,
                  dtype: torch.dtype = torch.float64) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """What autograd derives for Conv2DMod.forward ST:647-667: (out, dL/dx, dL/dstyle, dL/dweight) for an upstream
    gradient ``grad_out`` -- torch autograd through the literal restatement ``modconv`` (float64 by default, so the
    oracle's own rounding is far below the 1e-4 parity tolerance)."""
    with torch.enable_grad():
        xs = x.detach().to(dtype).requires_grad_(True)
        ws = weight.detach().to(dtype).requires_grad_(True)
        ys = style.detach().to(dtype).requires_grad_(True)
        out = modconv(xs, ws, ys, demod=demod, eps=eps)
        gx, gy, gw = torch.autograd.grad(out, (xs, ys, ws), grad_out.to(dtype))
    return out.detach(), gx, gy, gw


def generator_grads(params: Dict[str, Tensor], styles: Tensor, input_noise: Tensor, grad_out: Tensor,
                    dtype: torch.dtype = torch.float64) -> Tuple[Tensor, Dict[str, Tensor], Tensor]:
    """What autograd derives for Generator.forward ST:794-825: (rgb, {parameter key: dL/dparam}, dL/dstyles) for an
    upstream gradient ``grad_out`` on the image -- torch autograd through ``generator_forward`` above."""
    with torch.enable_grad():
        ps = {k: v.detach().to(dtype).requires_grad_(True) for k, v in params.items() if not k.endswith(".f")}
        st = styles.detach().to(dtype).requires_grad_(True)
        rgb = generator_forward(ps, st, input_noise.to(dtype))
        keys = list(ps)
        grads = torch.autograd.grad(rgb, [ps[k] for k in keys] + [st], grad_out.to(dtype), allow_unused=True)
    return rgb.detach(), {k: g for k, g in zip(keys, grads[:-1])}, grads[-1]

