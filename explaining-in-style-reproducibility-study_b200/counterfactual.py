"""Counterfactual rendering -- the step after the AttFind selection (reference ``run_attfind_combined.ipynb`` cells 17-20;
SURVEY.md section 8(f), rank 3).

Same function names and arguments as the notebook; the generator forwards run through the native plan
(``GeneratorPlan.forward``: the same sm_100a kernels as the sweep), the classifier stays PyTorch.  The single-coordinate
perturbation is data (one entry of the per-sample style row), never a patched ``nn.Linear`` bias (quirk Q2 of the
reference is not reproduced), so a whole batch of latents renders in ONE generator forward:
``render_counterfactuals`` is the batched primitive, the notebook-shaped functions are thin wrappers over it.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

from . import _native as N
from .modules import Generator, styles_def_to_tensor


@torch.no_grad()
def render_counterfactuals(generator: Generator, classifier, latents: torch.Tensor, sindex: int, s_style_min, s_style_max,
                           style_direction_index: int, shift_size: float, noise: torch.Tensor, class_index: int = 0,
                           precision: Optional[str] = None, with_base: bool = False):
    """latents [B, latent_dim] -> (perturbed images [B,3,S,S], P(class_index | perturbed) [B]) and, with ``with_base``,
    additionally (base images, base probabilities).  Style coordinate ``sindex`` of every latent moves by
    ``(target - coord) * shift_size`` (NB cell 17: ``shift = one_hot * ((s_style_m - style_coords[:, sindex]) * shift_size)``)."""
    N.require_cuda(latents, noise)
    plan = generator.plan()
    if not 0 <= sindex < plan.S:
        raise IndexError(f"sindex {sindex} out of range [0, {plan.S})")
    if style_direction_index not in (0, 1):
        raise ValueError("style_direction_index must be 0 (towards the minimum) or 1 (towards the maximum)")
    prec = precision or generator.precision
    w = styles_def_to_tensor([(latents.float(), generator.num_layers)]).contiguous()
    styles = plan.styles(w)                                             # [B, style_row]; columns [0, S) are StyleSpace
    target = float(s_style_min if style_direction_index == 0 else s_style_max)
    moved = styles.clone()
    moved[:, sindex] = styles[:, sindex] + (target - styles[:, sindex]) * float(shift_size)
    images = plan.forward(moved, noise, precision=prec)
    probs = torch.softmax(classifier.classify_images(images).float(), dim=1)[:, class_index]
    if not with_base:
        return images, probs
    base = plan.forward(styles, noise, precision=prec)
    base_probs = torch.softmax(classifier.classify_images(base).float(), dim=1)[:, class_index]
    return images, probs, base, base_probs


def generate_change_image_given_dlatent(dlatent, generator, classifier, class_index, sindex, s_style_min, s_style_max,
                                        style_direction_index, shift_size, label_size=2, noise=None, cuda_rank=0):
    """NB cell 17, same arguments: ``dlatent`` is the notebook's ``[(w[1,latent_dim], num_layers)]`` styles definition.
    Returns (perturbed_generated_images, change_prob) with change_prob the numpy scalar of image 0."""
    w = dlatent[0][0] if isinstance(dlatent, (list, tuple)) else dlatent
    images, probs = render_counterfactuals(generator, classifier, w, int(sindex), s_style_min, s_style_max,
                                           int(style_direction_index), shift_size, noise, class_index)
    return images, probs.cpu().numpy()[0]


def draw_on_image(image, number=None, font_file=None, font_fill=(0, 0, 255)):
    """NB cell 18 (the reference commented the text drawing out): CHW float image -> HWC uint8 of clip(x, 0, 1) * 255."""
    image = np.clip(np.transpose(image, (1, 2, 0)), 0, 1)
    return (image * 255).astype(np.uint8)


def generate_images_given_dlatent(dlatent, generator, classifier, class_index, sindex, s_style_min, s_style_max,
                                  style_direction_index, font_file=None, noise=None, shift_size=2, label_size=2,
                                  draw_results_on_image=True, resolution=64, cuda_rank=0, gen_num_layers=5):
    """NB cell 19: (side-by-side uint8 panel [resolution, 2*resolution, 3], change_prob, base_prob).  The reference's
    ``draw_results_on_image=False`` branch calls ``np.maxiumum`` (a typo that raises); here it does what it meant to:
    clamp to [-1, 1] and map to 0..255."""
    dev = noise.device
    w = torch.as_tensor(np.asarray(dlatent), dtype=torch.float32, device=dev).reshape(1, -1)
    change, cprob, base, bprob = render_counterfactuals(generator, classifier, w, int(sindex), s_style_min, s_style_max,
                                                        int(style_direction_index), shift_size, noise, class_index,
                                                        with_base=True)
    panel = np.zeros((resolution, 2 * resolution, 3), np.uint8)
    b, c = base[0].cpu().numpy(), change[0].cpu().numpy()
    if draw_results_on_image:
        panel[:, :resolution] = draw_on_image(b)
        panel[:, resolution:] = draw_on_image(c)
    else:
        panel[:, :resolution] = (np.transpose(b, (1, 2, 0)) * 127.5 + 127.5).astype(np.uint8)
        panel[:, resolution:] = (np.transpose(np.clip(c, -1, 1), (1, 2, 0)) * 127.5 + 127.5).astype(np.uint8)
    return panel, float(cprob[0]), float(bprob[0])


def visualize_style(generator, classifier, all_dlatents, style_change_effect, style_min, style_max, sindex,
                    style_direction_index, max_images, shift_size, font_file=None, noise=None, label_size=2, class_index=0,
                    effect_threshold=0.3, seed=None, allow_both_directions_change=False, draw_results_on_image=True,
                    render_batch: int = 64):
    """NB cell 20: stack up to ``max_images`` base | counterfactual panels of the images whose recorded effect for
    (direction, sindex, class) exceeds ``effect_threshold`` and whose rendered probability change does too.  Selection,
    shuffling (``np.random.seed(seed)``) and the keep / stop rules are the notebook's; the candidates are rendered
    ``render_batch`` at a time instead of one by one."""
    eff = np.asarray(style_change_effect)[:, style_direction_index, sindex, class_index]
    images_idx = (np.abs(eff) > effect_threshold).nonzero()[0] if allow_both_directions_change else (eff > effect_threshold).nonzero()[0]
    if images_idx.size == 0:
        return np.array([])
    if seed is not None:
        np.random.seed(seed)
    np.random.shuffle(images_idx)
    images_idx = images_idx[: min(max_images * 10, len(images_idx))]
    dev = noise.device
    res = generator.image_size
    result_images = []
    for i in range(0, len(images_idx), render_batch):
        w = torch.as_tensor(np.asarray(all_dlatents)[images_idx[i: i + render_batch]], dtype=torch.float32, device=dev)
        change, cprob, base, bprob = render_counterfactuals(generator, classifier, w, int(sindex), style_min[sindex],
                                                            style_max[sindex], int(style_direction_index), shift_size,
                                                            noise, class_index, with_base=True)
        change, base, cprob, bprob = change.cpu().numpy(), base.cpu().numpy(), cprob.cpu().numpy(), bprob.cpu().numpy()
        for j in range(w.shape[0]):
            if abs(float(cprob[j]) - float(bprob[j])) < effect_threshold:
                continue
            panel = np.zeros((res, 2 * res, 3), np.uint8)
            if draw_results_on_image:
                panel[:, :res], panel[:, res:] = draw_on_image(base[j]), draw_on_image(change[j])
            else:
                panel[:, :res] = (np.transpose(base[j], (1, 2, 0)) * 127.5 + 127.5).astype(np.uint8)
                panel[:, res:] = (np.transpose(np.clip(change[j], -1, 1), (1, 2, 0)) * 127.5 + 127.5).astype(np.uint8)
            result_images.append(panel)
            if len(result_images) == max_images:
                break
        if len(result_images) == max_images:
            break
    if len(result_images) < 3:
        return np.array([])           # "No point in returning results with very little images" (NB cell 20)
    return np.concatenate(result_images[:max_images], axis=0)


def visualize_style_by_distance_in_s(generator, classifier, all_dlatents, all_style_vectors_distances, style_min, style_max, sindex,
                                     style_sign_index, max_images, shift_size, font_file=None, noise=None, label_size=2,
                                     class_index=0, draw_results_on_image=True, effect_threshold=0.1, cuda_rank=0):
    """NB cell 21: the images whose style coordinate ``sindex`` lies FARTHEST from the extreme it is pushed to
    (``all_style_vectors_distances[:, sindex, style_sign_index]`` descending, cell 12), rendered base | counterfactual and
    stacked; at most ``max_images`` panels out of the first ``10 * max_images`` candidates, nothing below 3 panels.  (Like the
    notebook, ``effect_threshold`` is accepted and unused.)"""
    images_idx = np.argsort(np.asarray(all_style_vectors_distances)[:, sindex, style_sign_index])[::-1]
    if images_idx.size == 0:
        return np.array([])
    images_idx = images_idx[: min(max_images * 10, len(images_idx))]
    dlatents = np.asarray(all_dlatents)[images_idx]
    result_images = []
    for i in range(min(len(images_idx), max_images)):      # the notebook renders all candidates and keeps the first max_images
        panel, _, _ = generate_images_given_dlatent(dlatent=dlatents[i: i + 1], generator=generator, classifier=classifier,
                                                    class_index=class_index, sindex=sindex, noise=noise,
                                                    s_style_min=style_min[sindex], s_style_max=style_max[sindex],
                                                    style_direction_index=style_sign_index, font_file=font_file,
                                                    shift_size=shift_size, label_size=label_size,
                                                    draw_results_on_image=draw_results_on_image,
                                                    resolution=generator.image_size, cuda_rank=cuda_rank,
                                                    gen_num_layers=generator.num_layers)
        result_images.append(panel)
    if len(images_idx) < 3:
        return np.array([])
    return np.concatenate(result_images[:max_images], axis=0)

