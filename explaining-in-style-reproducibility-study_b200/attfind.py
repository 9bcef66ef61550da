"""AttFind: the StyleSpace perturbation sweep and the greedy top-k selection.

Host-side mirror of the reference notebook ``stylex/run_attfind_combined.ipynb`` (NB; raw JSON line
numbers): ``attfind_extraction`` NB:269-417, ``sindex_to_block_idx_and_index`` NB:220-234,
``get_min_max_style_vectors`` NB:237-252, ``find_significant_styles`` NB:731-758, the class split and
merged ranking of cells 14/16 (NB:695-714, NB:775-814).  Same names, argument meaning and outputs.

What changed underneath:

* no per-shift Python loop and no ``bias += shift`` patching of the model (quirk Q2): for every conv of
  the generator, all (coordinate, direction) pairs that perturb that conv's style vector are one batch;
  ``sx_attfind_make_styles`` writes the shifted style rows, ``sx_generator_forward(start_conv=c)`` re-runs
  only the suffix of the network from that conv on, reusing the latent's clean prefix;
* the classifier stays PyTorch on the same stream (BASELINE north_star); ``sx_attfind_scatter_effects``
  stores the logit deltas; ``sx_attfind_select`` does the class split + greedy selection on device with
  numpy-exact float64 column means;
* latents are sharded across ranks; one all-gather of the effects at the end (``dist.py``).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as N
from .modules import Generator, styles_def_to_tensor

_sel_ws = N.Workspace()


# ---------------------------------------------------------------------------------------------
# helpers with the notebook's names
# ---------------------------------------------------------------------------------------------
def sindex_to_block_idx_and_index(generator, sindex):
    """NB:220-234."""
    tmp_idx = sindex
    for idx, block in enumerate(generator.blocks):
        if tmp_idx < block.num_style_coords:
            return idx, tmp_idx
        tmp_idx = tmp_idx - block.num_style_coords
    return None, None


def get_min_max_style_vectors(style_coordinates: torch.Tensor):
    """NB:237-252: elementwise min / max over the image axis (native kernel)."""
    if style_coordinates is None or style_coordinates.shape[0] == 0:
        raise ValueError('No images pass the threshold check')
    N.require_cuda(style_coordinates)
    N.device_check()
    sc = N.f32c(style_coordinates)
    n, s = sc.shape
    mn = torch.empty(s, device=sc.device, dtype=torch.float32)
    mx = torch.empty_like(mn)
    N.check(N.lib().sx_attfind_minmax(sc.data_ptr(), n, s, s, mn.data_ptr(), mx.data_ptr(), N.stream_ptr()), "sx_attfind_minmax")
    return mn, mx


def discriminator_filter(discriminator, generated_image, threshold, probabilities=None):
    """NB:255-266 (PyTorch discriminator; not on the hot path)."""
    if probabilities is not None:
        output_generated = discriminator(generated_image, probabilities=probabilities)
    else:
        output_generated = discriminator(generated_image)
    if threshold is None:
        return output_generated
    if output_generated < threshold:
        return (False, output_generated)
    return (True, output_generated)


# ---------------------------------------------------------------------------------------------
# the sweep
# ---------------------------------------------------------------------------------------------
def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous latent shard of `rank`; remainder spread over the first ranks."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def default_eval_batch(image_size: int) -> int:
    """Perturbed images per generator / classifier launch when the caller does not say.  The notebook evaluates ONE per
    forward (NB:383); the batch only changes how many loop iterations share a launch, never a value.  256 at 256 px (the
    32-bit element indices of the narrow-layer kernels end at batch 511 there); small maps want more rows per launch --
    measured at 64 px with ResNet-18@224: 140.1 k coord-evals/s at 256, 146.6 k at 512, 155.9 k at 1024."""
    return 256 if image_size >= 256 else 512 if image_size >= 128 else 1024


def default_classify_batch(image_size: int = 256) -> int:
    """Images per classifier call of the sweep (the outputs of several generator launches): 1024 while the fp32 image
    buffer stays within 1 GiB (up to 256 px); measured at 256 px: 256 -> 1024 +2.3 %, 2048 nothing more."""
    env = os.environ.get("SX_CLASSIFY_BATCH")
    if env:
        return int(env)
    return max(1, min(1024, (1 << 30) // (12 * image_size * image_size)))


@torch.no_grad()
def attfind_sweep(G: Generator, classifier, latents: torch.Tensor, noise: torch.Tensor, shift_size: float = 1.0,
                  precision: Optional[str] = None, max_batch: Optional[int] = None, rank: int = 0, world_size: int = 1,
                  sindices: Optional[Sequence[int]] = None, image_indices: Optional[Sequence[int]] = None,
                  gather: bool = True, stats: Optional[dict] = None,
                  minmax: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                  zero_row: bool = False, classify_batch: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Phases A(second half)-C of ``attfind_extraction`` (NB:316-389) for latents ``[N, latent]``.

    Every rank computes the style coordinates, base images and base logits of ALL N latents (cheap, and it
    makes minima/maxima identical everywhere without a collective), then sweeps its own contiguous shard of
    latents over all S coordinates x 2 directions.  Returns the notebook's datasets as device tensors:
    'style_change' [N,2,S,2] (after the all-gather when ``gather``; the local shard otherwise), 'base_prob'
    [N,2] (raw logits, quirk Q3), 'style_coordinates' [N,S], 'minima' / 'maxima' [S], 'latents'.

    ``sindices`` / ``image_indices`` restrict the sweep to a subset (bounded benchmark / test samples);
    untouched entries of 'style_change' stay 0.  ``minmax`` = (minima, maxima) [S] supplies the global
    per-coordinate extrema when ``latents`` is only a slice of the job (they must come from ALL latents,
    NB:340); by default they are computed from ``latents``.  ``zero_row``: an all-zero style row takes part in the
    extrema -- what the notebook computes when fewer images pass the discriminator filter than ``num_images`` (its
    ``style_coordinates`` buffer keeps zero rows for the images it never found, and NB:340 reduces over the whole buffer).
    ``classify_batch``: images per classifier call (>= ``max_batch``; default ``default_classify_batch()``).
    """
    precision = precision or G.precision
    N.require_cuda(latents, noise)
    N.device_check()
    lib = N.lib()
    plan = G.plan()
    dev = latents.device
    L = G.num_layers
    S, row = plan.S, plan.row
    n_all = latents.shape[0]
    max_batch = default_eval_batch(G.image_size) if max_batch is None else max_batch
    max_batch = max(2, max_batch - (max_batch % 2))
    half = max_batch // 2
    plan.reserve(max_batch, precision)
    stream = N.stream_ptr()

    lat = N.f32c(latents)
    styles_all = plan.styles(styles_def_to_tensor([(lat, L)]).contiguous())           # [N, row]
    base_logits = torch.empty(n_all, 2, device=dev, dtype=torch.float32)
    for i in range(0, n_all, max_batch):                                                # NB:316-334, batched
        rgb = plan.forward(styles_all[i: i + max_batch].contiguous(), noise, precision=precision)
        base_logits[i: i + max_batch] = classifier.classify_images(rgb).float()
    if minmax is None:
        minima = torch.empty(S, device=dev, dtype=torch.float32)
        maxima = torch.empty_like(minima)
        N.check(lib.sx_attfind_minmax(styles_all.data_ptr(), n_all, S, row, minima.data_ptr(), maxima.data_ptr(), stream),
                "sx_attfind_minmax")                                                    # NB:340
        if zero_row:
            minima.clamp_(max=0.0)
            maxima.clamp_(min=0.0)
    else:
        minima, maxima = (N.f32c(t) for t in minmax)
        N.require_cuda(minima, maxima)
        if minima.shape != (S,) or maxima.shape != (S,):
            raise ValueError(f"minmax must be two [{S}] tensors")

    lo, hi = shard_range(n_all, rank, world_size)
    mine = list(range(lo, hi))
    if image_indices is not None:
        keep = set(int(i) for i in image_indices)
        mine = [i for i in mine if i in keep]
    effects = torch.zeros(hi - lo, 2, S, 2, device=dev, dtype=torch.float32)
    # coordinate runs per conv: [(conv index, first sindex, count)] -- contiguous runs of the requested sindices
    runs = _coord_runs(plan.conv_coords, sindices, half)
    # per-plan scratch, kept across calls (bench.py sweeps one latent per call: 201 MB of rgb at batch 256 / 256 px)
    styles_b = plan.scratch("sweep_styles", (max_batch, row), dev)
    # the classifier takes the images of several generator launches at once (its small-map layers want more rows than the
    # generator's index range allows per launch at 256 px): generator batches fill slices of one image buffer
    cb = max(max_batch, default_classify_batch(G.image_size) if classify_batch is None else int(classify_batch))
    rgb_b = plan.scratch("sweep_rgb", (cb, 3, G.image_size, G.image_size), dev)
    evals = 0
    for n in mine:                                                                      # NB:346
        base_row = styles_all[n]
        plan.forward(styles_all[n: n + 1], noise, save_cache=True, precision=precision, out=rgb_b[:1])
        pending, fill = [], 0

        def flush():
            logits = classifier.classify_images(rgb_b[:fill]).float().contiguous()      # NB:384 (raw rgb, quirk Q4)
            for off, first, cnt in pending:
                N.check(lib.sx_attfind_scatter_effects(logits.data_ptr() + off * logits.stride(0) * 4, base_logits[n].data_ptr(),
                                                       effects[n - lo].data_ptr(), 0, S, first, cnt, stream),
                        "sx_attfind_scatter_effects")

        for conv, first, cnt in runs:                                                   # NB:356-387, batched per conv
            b = 2 * cnt
            if fill + b > cb:
                flush()
                pending, fill = [], 0
            N.check(lib.sx_attfind_make_styles(base_row.data_ptr(), minima.data_ptr(), maxima.data_ptr(), styles_b.data_ptr(),
                                               row, first, cnt, float(shift_size), stream), "sx_attfind_make_styles")
            plan.forward(styles_b[:b], noise, start_conv=conv, precision=precision, out=rgb_b[fill: fill + b])
            pending.append((fill, first, cnt))
            fill += b
            evals += b
        if fill:
            flush()
    if stats is not None:
        stats["coord_evals"] = stats.get("coord_evals", 0) + evals
    if gather and world_size > 1:
        from .dist import gather_effects
        effects = gather_effects(effects, n_all, world_size)
    return {"style_change": effects, "latents": lat, "base_prob": base_logits, "minima": minima, "maxima": maxima,
            "style_coordinates": styles_all[:, :S].contiguous()}


def _coord_runs(conv_coords, sindices, half) -> List[Tuple[int, int, int]]:
    runs = []
    if sindices is None:
        for conv, (off, width) in enumerate(conv_coords):
            for s0 in range(0, width, half):
                runs.append((conv, off + s0, min(half, width - s0)))
        return runs
    want = sorted(set(int(s) for s in sindices))
    for conv, (off, width) in enumerate(conv_coords):
        sel = [s for s in want if off <= s < off + width]
        i = 0
        while i < len(sel):
            j = i
            while j + 1 < len(sel) and sel[j + 1] == sel[j] + 1 and (j + 1 - i) < half:
                j += 1
            runs.append((conv, sel[i], j - i + 1))
            i = j + 1
    return runs


# ---------------------------------------------------------------------------------------------
# selection
# ---------------------------------------------------------------------------------------------
def _native_select(effects: torch.Tensor, base_logits: Optional[torch.Tensor], k: int, max_image_effect: float,
                   class_index: int) -> List[Tuple[int, int]]:
    N.require_cuda(effects, base_logits)
    N.device_check()
    lib = N.lib()
    if effects.dtype not in (torch.float32, torch.float64):
        effects = effects.float()
    effects = effects.contiguous()
    n, two, S, two2 = effects.shape
    assert two == 2 and two2 == 2, "effects must be [N, 2, S, 2]"
    if n == 0:
        # the notebook dies here too (quirk Q7: "arrays used as indices must be of integer (or boolean) type")
        raise IndexError(f"AttFind selection: class {class_index} has no images")
    bl = None if base_logits is None else N.f32c(base_logits)
    ws = _sel_ws.get(lib.sx_attfind_select_workspace_bytes(n, S), effects.device)
    picks = torch.empty(2 * k, device=effects.device, dtype=torch.int32)
    N.check(lib.sx_attfind_select(effects.data_ptr(), 1 if effects.dtype == torch.float64 else 0, N.ptr(bl), n, S, k,
                                  float(max_image_effect), class_index, picks.data_ptr(), ws.data_ptr(), ws.numel(),
                                  N.stream_ptr()), "sx_attfind_select")
    p = picks.cpu().tolist()
    return [(p[2 * i], p[2 * i + 1]) for i in range(k)]


def find_significant_styles(style_change_effect, num_indices, class_index, generator=None, classifier=None,
                            all_dlatents=None, style_min=None, style_max=None, max_image_effect=0.2, label_size=2,
                            sindex_offset=0, device=None):
    """NB:731-758, same signature.  ``style_change_effect`` is the per-class array [N_c, 2, S, 2] (numpy or
    tensor; the notebook passes a float64 copy).  Unlike the notebook the input is NOT modified in place."""
    if isinstance(style_change_effect, np.ndarray):
        dev = device or torch.device("cuda", torch.cuda.current_device())
        eff = torch.from_numpy(np.ascontiguousarray(style_change_effect)).to(dev)
    else:
        eff = style_change_effect
    picks = _native_select(eff, None, int(num_indices), float(max_image_effect), int(class_index))
    return [(d, s + sindex_offset) for d, s in picks]


def attfind_select(style_change_effect: torch.Tensor, base_probs: torch.Tensor, num_indices: int = 5,
                   effect_threshold: float = 0.5):
    """Cells 14 + 16 (NB:695-714, NB:775-814): per-class greedy picks on device, merged ranking on the host.

    Returns (picks {0: [(direction, sindex)], 1: [...]}, merged [(direction, sindex)], scores).  The merged
    ranking evaluates the notebook's own numpy expression on the <= 2k picked columns only.
    """
    labels = torch.argmax(base_probs, dim=1)
    for c in (0, 1):
        if int((labels == c).sum()) == 0:
            raise IndexError(f"AttFind selection: class {c} has no images (the notebook fails here too, quirk Q7)")
    picks = {c: _native_select(style_change_effect, base_probs, num_indices, effect_threshold * 5, c) for c in (0, 1)}
    s0 = [s for _, s in picks[0]]
    joined = [(1 - d, s) for d, s in picks[1] if s not in s0]                            # NB:802
    joined += picks[0]                                                                   # NB:803
    scores = []
    for d, s in joined:                                                                  # NB:806-809
        od = 1 if d == 0 else 0
        col_a = style_change_effect[:, d, s, 0].float().cpu().numpy()
        col_b = style_change_effect[:, od, s, 1].float().cpu().numpy()
        scores.append(np.mean(col_a) + np.mean(col_b))
    order = np.argsort(scores)[::-1]                                                     # NB:811
    return picks, [joined[i] for i in order], [float(scores[i]) for i in order]


# ---------------------------------------------------------------------------------------------
# exact top-k from a throughput-mode sweep: screen in bf16, verify the candidates in fp32
# ---------------------------------------------------------------------------------------------
@torch.no_grad()
def _exact_entries(G: Generator, classifier, noise: torch.Tensor, styles_all: torch.Tensor, base_exact: torch.Tensor,
                   minima: torch.Tensor, maxima: torch.Tensor, latent_idx: torch.Tensor, columns: torch.Tensor,
                   shift_size: float, precision: str, max_batch: int) -> torch.Tensor:
    """Parity-mode effects of the (latent n_j, flat column x_j = d*S + s) pairs: [len, 2].

    Full generator forwards (start_conv = 0 needs no per-latent prefix cache), so one batch mixes latents freely:
    ``sx_attfind_make_styles_pairs`` writes row j = style row of latent n_j with coordinate s_j moved to its minimum /
    maximum (NB:374-381 as data), then generator + classifier + logit delta (NB:382-385)."""
    plan = G.plan()
    lib = N.lib()
    S, row = plan.S, plan.row
    dev = styles_all.device
    n = int(latent_idx.numel())
    out = torch.empty(n, 2, device=dev, dtype=torch.float32)
    if n == 0:
        return out
    li = latent_idx.to(device=dev, dtype=torch.int32).contiguous()
    ci = columns.to(device=dev, dtype=torch.int32).contiguous()
    plan.reserve(max_batch, precision)
    stream = N.stream_ptr()
    styles_b = torch.empty(max_batch, row, device=dev, dtype=torch.float32)
    # every batch has exactly max_batch rows (the tail repeats the last pair): one shape for the classifier's cuDNN plans
    # (with cudnn.benchmark each new batch size re-tunes ~20 fp32 convolutions: seconds per odd-sized tail, measured)
    if n % max_batch:
        pad = max_batch - n % max_batch
        li = torch.cat([li, li[-1:].expand(pad)]).contiguous()
        ci = torch.cat([ci, ci[-1:].expand(pad)]).contiguous()
    for j0 in range(0, n, max_batch):
        cnt = min(max_batch, n - j0)
        N.check(lib.sx_attfind_make_styles_pairs(styles_all.data_ptr(), styles_all.stride(0), minima.data_ptr(), maxima.data_ptr(),
                                                 styles_b.data_ptr(), row, S, li[j0:].data_ptr(), ci[j0:].data_ptr(), max_batch,
                                                 float(shift_size), stream), "sx_attfind_make_styles_pairs")
        rgb = plan.forward(styles_b, noise, precision=precision)
        logits = classifier.classify_images(rgb).float()
        out[j0: j0 + cnt] = (logits - base_exact[li[j0: j0 + max_batch].long()])[:cnt]
    return out


def sharded_entries(evaluate, latent_idx: torch.Tensor, columns: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    """Split the (latent, column) pair list evenly over the ranks (contiguous shares, ``shard_range``), let this rank
    ``evaluate(latent_idx_share, columns_share) -> [share, 2]`` its part and all-gather the parts back in order: every rank
    returns the same [P, 2] tensor, each entry computed by exactly one rank."""
    total = int(latent_idx.numel())
    if world_size == 1:
        return evaluate(latent_idx, columns)
    from .dist import gather_effects
    plo, phi = shard_range(total, rank, world_size)
    return gather_effects(evaluate(latent_idx[plo:phi], columns[plo:phi]).contiguous(), total, world_size)


def screen_and_verify(approx: torch.Tensor, base_exact: torch.Tensor, exact_entries, select, num_indices: int = 5,
                      effect_threshold: float = 0.5, min_candidates: int = 32, band_sigmas: float = 8.0,
                      band_floor: float = 2.0, max_passes: int = 6, audit_columns: int = 8):
    """The selection logic of ``attfind_verify_topk`` (device-agnostic torch; see there).

    approx [N,2,S,2]: effects of the throughput-mode sweep; base_exact [N,2]: parity-mode base logits;
    ``exact_entries(latent_idx [P], columns [P]) -> [P, 2]``: parity-mode effects of the (latent, flat column x = d*S + s)
    pairs; ``select(effects, base_logits, k, effect_threshold) -> (picks, merged, scores)``: the exact selection
    (``attfind_select``).  Candidate columns of class c are re-evaluated for the images of class c only (the greedy
    selection of class c reads nothing else); the picked columns and their opposite directions, which the merged ranking
    reads over ALL images (NB:806-809), are completed at the end.  ``audit_columns`` (per class): once every round is
    verified, that many RANDOM columns that were not candidates are re-evaluated too and their approximate-vs-exact
    column-mean discrepancy is compared with the band -- the band was measured on the leaders only, this checks it on the
    population it is applied to; a discrepancy above band / 2 doubles the band and the judgement is repeated
    (``info['audit_max']``).  Returns (picks, merged, scores, info)."""
    n_all, _, S, _ = approx.shape
    dev = approx.device
    k = int(num_indices)
    T = effect_threshold * 5                                                      # NB:776-777,795
    labels = torch.argmax(base_exact, dim=1)
    for c in (0, 1):
        if int((labels == c).sum()) == 0:
            raise IndexError(f"AttFind selection: class {c} has no images (the notebook fails here too, quirk Q7)")
    hybrid = approx.float().clone()
    flat = hybrid.view(n_all, 2 * S, 2)
    rows = {c: (labels == c).nonzero().flatten() for c in (0, 1)}
    is_exact = {c: torch.zeros(2 * S, dtype=torch.bool, device=dev) for c in (0, 1)}
    E_apx = {c: approx[rows[c]][:, :, :, c].reshape(rows[c].numel(), 2 * S).double().clamp_(min=0) for c in (0, 1)}
    info = {"passes": 0, "exact_evals": 0, "band": None, "candidates": [0, 0], "verified": False}
    neg_inf = torch.full((2 * S,), float("-inf"), device=dev, dtype=torch.float64)

    def make_exact(per_class_columns):
        """{class: columns}: one call of ``exact_entries`` for the (row of class c, new column of class c) pairs"""
        li, ci, todo = [], [], {}
        for c, columns in per_class_columns.items():
            cols = sorted(set(int(x) for x in columns) - set(is_exact[c].nonzero().flatten().tolist()))
            if not cols:
                continue
            cols_t = torch.tensor(cols, device=dev, dtype=torch.long)
            todo[c] = cols_t
            li.append(rows[c].repeat_interleave(len(cols)))
            ci.append(cols_t.repeat(rows[c].numel()))
        if not todo:
            return
        vals = exact_entries(torch.cat(li), torch.cat(ci))                         # [P, 2]
        off = 0
        for c, cols_t in todo.items():
            cnt = rows[c].numel() * cols_t.numel()
            flat[rows[c][:, None], cols_t[None, :]] = vals[off: off + cnt].reshape(rows[c].numel(), cols_t.numel(), 2).to(flat.dtype)
            off += cnt
            is_exact[c][cols_t] = True
        info["exact_evals"] += int(vals.shape[0])
        info["candidates"] = [int(is_exact[0].sum()), int(is_exact[1].sum())]

    def replay(band=None, collect_top=0):
        """the greedy selection (NB:751-756) on the hybrid effects, winners restricted to exact columns.
        -> (every round verified, {class: columns to add}, flat picks per class, approx-minus-exact column-mean discrepancies)"""
        ok, need, picks, disc = True, {0: set(), 1: set()}, {}, []
        for c in (0, 1):
            ex = is_exact[c]
            have_exact = bool(ex.any())
            E = hybrid[rows[c]][:, :, :, c].reshape(rows[c].numel(), 2 * S).double().clamp_(min=0)
            Ea = E_apx[c].clone()
            img = torch.zeros(E.shape[0], device=dev, dtype=torch.float64)
            picks[c] = []
            for _ in range(k):
                mask = img < T
                if not bool(mask.any()):                                           # quirk Q7: NaN means, argmax = 0
                    picks[c].append(0)
                    continue
                cm = E[mask].mean(dim=0)
                if collect_top:
                    need[c].update(torch.topk(cm, min(collect_top, cm.numel())).indices.tolist())
                x = int(torch.argmax(torch.where(ex, cm, neg_inf))) if have_exact else int(torch.argmax(cm))
                if have_exact:
                    disc.append((Ea[mask].mean(dim=0) - cm)[ex])
                if band is not None:
                    viol = ((~ex) & (cm + band >= cm[x])).nonzero().flatten()
                    if viol.numel():
                        ok = False
                        need[c].update(viol.tolist())
                picks[c].append(x)
                img += E[:, x]
                E[:, x] = 0
                Ea[:, x] = 0
        return ok, need, picks, disc

    # first candidates: the leaders of every round of the approximate selection
    _, need, _, _ = replay(collect_top=max(2, min_candidates))
    make_exact(need)
    picks_r = None
    audited = audit_columns <= 0
    band_min = 0.0
    gen = torch.Generator().manual_seed(20211)            # the same audit columns on every rank
    for p in range(max_passes):
        info["passes"] = p + 1
        # the band from the measured discrepancy of the candidates' masked column means
        _, _, _, disc = replay()
        d = torch.cat(disc)
        band = max(band_sigmas * float(d.std()), band_floor * float(d.abs().max()), band_min)
        info.update(band=band, colmean_err_std=float(d.std()), colmean_err_max=float(d.abs().max()))
        ok, need, picks_r, _ = replay(band=band)
        if not ok:
            make_exact(need)
            continue
        if audited:
            info["verified"] = True
            break
        # audit: the band on a random sample of the columns it was applied to without being measured on them
        audited = True
        worst = 0.0
        extra = {}
        for c in (0, 1):
            free = (~is_exact[c]).nonzero().flatten().cpu()
            if free.numel() == 0:
                continue
            pick = free[torch.randperm(free.numel(), generator=gen)[:audit_columns]]
            extra[c] = pick.tolist()
        make_exact(extra)
        for c, cols in extra.items():
            idx = torch.tensor(cols, device=dev, dtype=torch.long)
            exact_cm = hybrid[rows[c]][:, :, :, c].reshape(rows[c].numel(), 2 * S)[:, idx].double().clamp_(min=0).mean(dim=0)
            worst = max(worst, float((E_apx[c][:, idx].mean(dim=0) - exact_cm).abs().max()))
        info["audit_max"] = worst
        if worst <= band / 2:
            info["verified"] = True
            break
        band_min = 2 * worst                              # the leaders under-estimated the error: widen and judge again
    # the merged ranking (NB:806-809) reads both directions of every picked coordinate over ALL images
    both = [x for c in (0, 1) for x in picks_r[c]]
    both += [(1 - x // S) * S + x % S for x in both]
    make_exact({0: both, 1: both})
    picks, merged, scores = select(hybrid, base_exact, k, effect_threshold)
    want = {c: [(x // S, x % S) for x in picks_r[c]] for c in (0, 1)}
    if picks != want:                 # can only happen on an exact float64 tie between two exact columns
        info["verified"] = False
        info["replay_mismatch"] = {"native": picks, "replay": want}
    info["style_change"] = hybrid
    info["base_prob"] = base_exact
    return picks, merged, scores, info


@torch.no_grad()
def attfind_verify_topk(G: Generator, classifier, latents: torch.Tensor, noise: torch.Tensor, sweep: Dict[str, torch.Tensor],
                        num_indices: int = 5, effect_threshold: float = 0.5, shift_size: float = 1.0, precision: str = "fp32",
                        max_batch: int = 128, rank: int = 0, world_size: int = 1, min_candidates: int = 32,
                        band_sigmas: float = 8.0, band_floor: float = 2.0, max_passes: int = 6, audit_columns: int = 8):
    """The reference's EXACT top-k (NB:731-814) out of a throughput-mode sweep: screen in bf16, verify in fp32.

    A bf16 generator + bf16 classifier moves every effect by ~1e-2 of its size; the greedy selection of NB:751-756 is an
    argmax over 2S masked column means whose leaders can be closer than that, so the picks of a pure bf16 sweep are not the
    reference's (measured: profiles/README.md, top-k parity).  But the bf16 sweep is accurate enough to tell which FEW
    columns can possibly win.  This function

    1. recomputes the base logits of every latent in the parity mode (``classifier`` = the fp32 wrapper, ``precision`` =
       "fp32"), each latent by the rank that owns it: the class split (NB:695-714) is exact;
    2. replays the greedy selection on the approximate effects ``sweep['style_change']`` and takes, per class and round,
       the ``min_candidates`` leading columns as candidates;
    3. re-evaluates (image of class c, candidate column of class c) in the parity mode -- full forwards batched across
       latents (``sx_attfind_make_styles_pairs``), the pair list split evenly over the ranks, one small all-gather -- and
       overwrites those entries of the effects;
    4. measures the approximate-vs-exact discrepancy d of the candidates' masked column means and sets
       ``band = max(band_sigmas * std(d), band_floor * max|d|)``;
    5. replays the selection on the hybrid effects: a round is VERIFIED when its winner is an exact column and no
       non-candidate column comes within ``band`` of it (its true mean then cannot exceed the winner's).  Columns that do
       are added to the candidates and steps 3-5 repeat (``max_passes``); once every round passes, ``audit_columns`` random
       non-candidate columns per class are re-evaluated as well and the band is checked (and widened if need be) on them;
    6. the picked columns and their opposite directions (the merged ranking reads them over ALL images, NB:806-809) are
       completed, and the final picks / merged list come from ``attfind_select`` (numpy-exact float64 kernels) on the
       hybrid effects.

    The picks equal those of a full parity-mode sweep whenever the band holds for the columns that were NOT re-evaluated
    (8 sigma of the measured error by default).  Returns (picks, merged, scores, info): ``info['verified']`` says whether
    every round passed, ``info['style_change']`` is the hybrid effects tensor, ``info['base_prob']`` the exact logits."""
    N.require_cuda(latents, noise)
    plan = G.plan()
    S, L = plan.S, G.num_layers
    n_all = latents.shape[0]
    dev = latents.device
    lat = N.f32c(latents)
    minima, maxima = N.f32c(sweep["minima"]).reshape(-1), N.f32c(sweep["maxima"]).reshape(-1)
    styles_all = plan.styles(styles_def_to_tensor([(lat, L)]).contiguous())
    approx = sweep["style_change"]
    if tuple(approx.shape) != (n_all, 2, S, 2):
        raise ValueError(f"sweep['style_change'] must be the gathered [N,2,S,2] tensor, got {tuple(approx.shape)}")

    def gather_rows(t, total):                                                     # contiguous shares [cnt_r, 2] -> [total, 2]
        if world_size == 1:
            return t
        from .dist import gather_effects
        return gather_effects(t.contiguous(), total, world_size)

    # exact base logits: latent n by rank owner(n), gathered (bit-identical everywhere)
    lo, hi = shard_range(n_all, rank, world_size)
    plan.reserve(max_batch, precision)
    base_loc = torch.empty(hi - lo, 2, device=dev, dtype=torch.float32)
    for i in range(lo, hi, max_batch):
        j = min(hi, i + max_batch)
        rows_b = styles_all[i:j]
        if j - i < max_batch:                           # fixed batch shape (see _exact_entries)
            rows_b = torch.cat([rows_b, rows_b[-1:].expand(max_batch - (j - i), -1)])
        rgb = plan.forward(rows_b.contiguous(), noise, precision=precision)
        base_loc[i - lo: j - lo] = classifier.classify_images(rgb).float()[: j - i]
    base_exact = gather_rows(base_loc, n_all)

    def exact_entries(latent_idx, columns):
        return sharded_entries(
            lambda li, ci: _exact_entries(G, classifier, noise, styles_all, base_exact, minima, maxima, li, ci, shift_size,
                                          precision, max_batch),
            latent_idx, columns, rank, world_size)

    return screen_and_verify(approx, base_exact, exact_entries, attfind_select, num_indices, effect_threshold,
                             min_candidates, band_sigmas, band_floor, max_passes, audit_columns)


# ---------------------------------------------------------------------------------------------
# the notebook's entry point
# ---------------------------------------------------------------------------------------------
DATASET_NAMES = ("style_change", "latents", "base_prob", "minima", "maxima", "style_coordinates", "original_images",
                 "noise", "discriminator")


@torch.no_grad()
def attfind_extraction(dataloader, num_images, results_folder, stylex, classifier, dataset_name, noise, num_style_coords,
                       shift_size, discriminator_threshold, image_size=64, batch_size=1, cuda_rank=0,
                       use_discriminator=False, use_old_architecture=True, precision=None, max_batch=None,
                       rank=0, world_size=1, front_batch=256, verify_classifier=None, num_indices=5, effect_threshold=0.5):
    """``attfind_extraction`` of NB:269-417 with the same arguments (extra keyword arguments have defaults).

    Phase A (encode each image, classify it, build ``concat_w``, discriminator output and optional filter;
    NB:300-336) runs batched through ``stylex.encode_images`` (``front_batch`` images per launch; the dataloader
    still yields one image per item, NB:284-285); phases B-C are ``attfind_sweep``.  Writes ``style_change_records.hdf5`` with the 9 datasets of NB:395-403 (h5py when importable, else
    the built-in minimal writer); also returns them as a dict.

    ``verify_classifier`` (a parity-mode wrapper: fp32, eager): when the sweep ran in a reduced precision, the top-k
    candidates are re-evaluated in fp32 (``attfind_verify_topk``) and the records carry the HYBRID effects + exact base
    logits, so that the notebook's selection cells 14-16 (``find_significant_styles`` on the loaded records) return the
    picks of a full fp32 sweep; the returned dict then also has 'picks', 'merged' and 'verify'.
    """
    if batch_size != 1:
        raise ValueError('Please use a batch_size equal to 1')                          # NB:284-285
    dev = torch.device("cuda", cuda_rank)
    with torch.cuda.device(dev):      # native launches go to the CURRENT device's stream: make cuda_rank current
        return _attfind_extraction(dataloader, num_images, results_folder, stylex, classifier, noise, num_style_coords,
                                   shift_size, discriminator_threshold, image_size, dev, use_discriminator,
                                   use_old_architecture, precision, max_batch, rank, world_size, front_batch,
                                   verify_classifier, num_indices, effect_threshold)


def _attfind_extraction(dataloader, num_images, results_folder, stylex, classifier, noise, num_style_coords, shift_size,
                        discriminator_threshold, image_size, dev, use_discriminator, use_old_architecture, precision,
                        max_batch, rank, world_size, front_batch, verify_classifier=None, num_indices=5, effect_threshold=0.5):
    G = stylex.G
    if num_style_coords != G.num_style_coords:
        raise ValueError(f"num_style_coords={num_style_coords} but the generator has {G.num_style_coords} (quirk Q5)")
    latent_dim = G.latent_dim
    image_latents = torch.zeros((num_images, latent_dim), device=dev)
    original_images = torch.zeros((num_images, 3, image_size, image_size), device=dev)
    discriminator_results = torch.zeros((num_images, 1), device=dev)
    images_found = 0
    from .stylex import encode_images
    have_d = use_discriminator or getattr(stylex, "D", None) is not None
    it = iter(dataloader)                                                                # NB:298
    exhausted = False
    while images_found < num_images and not exhausted:
        # the notebook handles one image per iteration (NB:300-336); here up to `front_batch` images go through the
        # encoder, the classifier, the generator and the discriminator together -- same arithmetic per image, and the
        # images that pass the filter are kept in dataloader order, exactly like the sequential loop
        chunk = []
        for batch in it:
            chunk.append(batch.to(dev).reshape(-1, 3, image_size, image_size))
            if len(chunk) >= min(front_batch, num_images - images_found):
                break
        else:
            exhausted = True
        if not chunk:
            break
        x = torch.cat(chunk)
        fe = encode_images(stylex, classifier, x, noise, batch=front_batch, use_old_architecture=use_old_architecture,
                           discriminator=have_d)
        keep = torch.ones(x.shape[0], dtype=torch.bool, device=dev)
        if use_discriminator and discriminator_threshold is not None:
            # NB:262-266 returns (False, out) when out < threshold and (True, out) otherwise; NB:323,327 bind that flag to
            # `skip` and `continue` when it is set: the executed reference KEEPS the images with out < threshold (pinned by
            # tests/golden/frontend_filter.npz, the verbatim loop run with use_discriminator=True)
            keep = fe["discriminator"][:, 0] < discriminator_threshold
        idx = keep.nonzero().flatten()[: num_images - images_found]
        k = int(idx.numel())
        original_images[images_found: images_found + k] = x[idx]
        image_latents[images_found: images_found + k] = fe["latents"][idx]
        if have_d:
            discriminator_results[images_found: images_found + k] = fe["discriminator"][idx]
        images_found += k
    if images_found == 0:
        raise ValueError('No images pass the threshold check')
    # fewer images than asked for: the notebook's unfilled (zero) rows take part in its min / max (NB:291, 340)
    res = attfind_sweep(G, classifier, image_latents[:images_found], noise, shift_size=shift_size, precision=precision,
                        max_batch=max_batch, rank=rank, world_size=world_size, zero_row=images_found < num_images)
    extra = {}
    if verify_classifier is not None:
        picks, merged, scores, info = attfind_verify_topk(G, verify_classifier, image_latents[:images_found], noise, res,
                                                          num_indices, effect_threshold, shift_size=shift_size, precision="fp32",
                                                          max_batch=128 if max_batch is None else min(max_batch, 256),
                                                          rank=rank, world_size=world_size)
        res = dict(res, style_change=info["style_change"], base_prob=info["base_prob"])
        extra = {"picks": picks, "merged": merged, "verify": {k: v for k, v in info.items() if k not in ("style_change", "base_prob")}}
    out = {
        "style_change": _pad(res["style_change"], num_images), "latents": image_latents,
        "base_prob": _pad(res["base_prob"], num_images), "minima": res["minima"][None], "maxima": res["maxima"][None],
        "style_coordinates": _pad(res["style_coordinates"], num_images), "original_images": original_images,
        "noise": noise.reshape(1, image_size, image_size, 1), "discriminator": discriminator_results,
    }
    if rank == 0 and results_folder is not None:
        save_records(results_folder, out)
    out.update(extra)
    return out


def _pad(t: torch.Tensor, n: int) -> torch.Tensor:
    if t.shape[0] == n:
        return t
    out = torch.zeros((n,) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
    out[: t.shape[0]] = t
    return out


def save_records(results_folder: str, datasets: Dict[str, torch.Tensor]) -> str:
    """NB:394-417: the 9 float32 datasets of ``style_change_records.hdf5`` -- through h5py when it is importable, else through
    the built-in minimal HDF5 writer (``hdf5_lite``: superblock v0, contiguous datasets in the root group)."""
    arrays = {k: datasets[k].detach().float().cpu().numpy() for k in DATASET_NAMES}
    os.makedirs(results_folder, exist_ok=True)
    path = os.path.join(results_folder, "style_change_records.hdf5")
    try:
        import h5py
    except ImportError:
        from . import hdf5_lite
        hdf5_lite.write_hdf5(path, arrays)
        return path
    with h5py.File(path, "w") as f:
        for k, v in arrays.items():
            f.create_dataset(k, v.shape, dtype="f")[:] = v
    return path


def load_records(path: str, threshold_index: Optional[int] = None) -> Dict[str, np.ndarray]:
    """NB cell 12: read ``style_change_records`` back (the ``.hdf5`` of ``save_records`` / of the reference, or an ``.npz``;
    through h5py or the built-in reader) and derive what the selection / visualisation cells use.  ``threshold_index`` = the notebook's
    ``load_hdf5_results(..., threshold)`` row cap (501 there).  Returns the nine datasets (``noise`` / ``minima`` / ``maxima``
    unsliced, like cell 12) plus ``style_min`` / ``style_max`` [S] and ``all_style_vectors_distances`` [N, S, 2]."""
    if path.endswith((".hdf5", ".h5")):
        try:
            import h5py
        except ImportError:
            from . import hdf5_lite
            skipped: Dict[str, str] = {}
            allv = hdf5_lite.read_hdf5(path, skipped)
            missing = [k for k in DATASET_NAMES if k not in allv]
            if missing:
                why = "; ".join(f"'{k}': {skipped.get(k, 'not in the file')}" for k in missing)
                raise ValueError(f"{path}: the built-in HDF5 reader (h5py is not installed) cannot provide {why}. It parses "
                                 "contiguous / compact datasets only -- install h5py to read chunked or compressed records.")
            raw = {k: allv[k] for k in DATASET_NAMES}
        else:
            with h5py.File(path, "r") as f:
                raw = {k: np.array(f[k]) for k in DATASET_NAMES}
    else:
        with np.load(path) as z:
            raw = {k: np.array(z[k]) for k in DATASET_NAMES}
    out = {k: (v if k in ("noise", "minima", "maxima") else v[:threshold_index]) for k, v in raw.items()}
    out["style_min"] = np.squeeze(out["minima"])
    out["style_max"] = np.squeeze(out["maxima"])
    sc = out["style_coordinates"]
    dist = np.zeros((sc.shape[0], sc.shape[1], 2))                                       # float64, like cell 12
    dist[:, :, 0] = sc - np.tile(out["style_min"], (sc.shape[0], 1))
    dist[:, :, 1] = np.tile(out["style_max"], (sc.shape[0], 1)) - sc
    out["all_style_vectors_distances"] = dist
    return out


def filter_unstable_images(style_change_effect, effect_threshold=0.3, num_indices_threshold=150):
    """NB cell 11 (defined there, its call is commented out at NB:664): zero the rows of images for which more than
    ``num_indices_threshold`` (direction, coordinate, class) entries move the logits by more than ``effect_threshold``.
    In place, like the notebook; returns the array."""
    unstable_images = (np.sum(np.abs(style_change_effect) > effect_threshold, axis=(1, 2, 3)) > num_indices_threshold)
    style_change_effect[unstable_images] = 0
    return style_change_effect

