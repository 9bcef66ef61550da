"""Shared test helpers (test infrastructure; may import the oracle)."""
import numpy as np
import torch


def state_from_npz(z, prefix):
    return {k[len(prefix):]: torch.from_numpy(np.array(z[k])) for k in z.files if k.startswith(prefix)}


class TinyCNN(torch.nn.Module):
    """same architecture as tests/golden/make_golden.py::TinyCNN (weights come from the fixture)."""

    def __init__(self):
        super().__init__()
        self.net = torch.nn.Sequential(
            torch.nn.Conv2d(3, 8, 3, stride=2, padding=1), torch.nn.BatchNorm2d(8), torch.nn.ReLU(),
            torch.nn.Conv2d(8, 16, 3, stride=2, padding=1), torch.nn.BatchNorm2d(16), torch.nn.ReLU(),
            torch.nn.AdaptiveAvgPool2d(2), torch.nn.Flatten())
        self.fc = torch.nn.Linear(64, 2)

    def forward(self, x):
        return self.fc(self.net(x))


def tiny_cnn_from(z, prefix):
    m = TinyCNN()
    m.load_state_dict(state_from_npz(z, prefix))
    return m.eval()


# ---------------------------------------------------------------------------------------------------------------------
# top-k selection margins: how far the greedy picks of NB:731-758 are from flipping, and how large the perturbation
# (another precision mode of the same sweep) actually is in the quantity the selection looks at
# ---------------------------------------------------------------------------------------------------------------------
def selection_margin_report(eff_ref, base_ref, eff_other=None, base_other=None, k=5, max_image_effect=2.5):
    """Replays the notebook's greedy selection (class split NB:695-714, find_significant_styles NB:731-758) on
    ``eff_ref`` [N,2,S,2] / ``base_ref`` [N,2] in float64 and reports, per class and round:

    * ``top``/``second``: the largest and second-largest masked column mean, ``gap`` = their difference -- the pick of
      this round flips only if the column means move by more than gap/2;
    * ``mask_margin``: min over images of |images_effect - max_image_effect| -- the saturation mask (which rows enter
      the mean) flips only if an image's accumulated effect moves by more than that;
    * with ``eff_other``: ``colmean_err`` = max over ALL 2S columns of |masked column mean(other) - masked column
      mean(ref)| under the reference's masks and picks, ``image_err`` = max |accumulated image effect(other) - (ref)|.

    A mode ``other`` provably reproduces the reference's picks when, in every round, 2*colmean_err < gap and
    image_err < mask_margin, and no image changes class (``label_margin`` = min |logit1 - logit0| vs ``base_err``).
    Returns a dict (JSON-serialisable)."""
    eff_ref = np.asarray(eff_ref, dtype=np.float64)
    base_ref = np.asarray(base_ref, dtype=np.float64)
    labels = np.argmax(base_ref, axis=1)
    out = {"classes": {}, "label_margin": float(np.abs(base_ref[:, 1] - base_ref[:, 0]).min()),
           "class_sizes": [int((labels == 0).sum()), int((labels == 1).sum())],
           "max_abs_effect": float(np.abs(eff_ref).max())}
    if eff_other is not None:
        eff_other = np.asarray(eff_other, dtype=np.float64)
        out["max_abs_effect_err"] = float(np.abs(eff_other - eff_ref).max())
        out["rms_effect_err"] = float(np.sqrt(np.mean((eff_other - eff_ref) ** 2)))
    if base_other is not None:
        base_other = np.asarray(base_other, dtype=np.float64)
        out["base_err"] = float(np.abs(base_other - base_ref).max())
        out["label_flips"] = int((np.argmax(base_other, axis=1) != labels).sum())
    S = eff_ref.shape[2]
    worst = {"gap": np.inf, "ratio": 0.0, "mask_ratio": 0.0}
    for c in (0, 1):
        rows = labels == c
        E = np.maximum(0, eff_ref[rows][:, :, :, c].reshape(int(rows.sum()), -1))
        Eo = None if eff_other is None else np.maximum(0, eff_other[rows][:, :, :, c].reshape(int(rows.sum()), -1))
        img = np.zeros(E.shape[0])
        imgo = np.zeros(E.shape[0])
        rounds = []
        for _ in range(k):
            mask = img < max_image_effect
            if not mask.any():
                rounds.append({"empty_mask": True})
                break
            cm = E[mask].mean(axis=0)
            order = np.argsort(-cm, kind="stable")
            x, x2 = int(order[0]), int(order[1])
            r = {"pick": [x // S, x % S], "top": float(cm[x]), "second": float(cm[x2]), "gap": float(cm[x] - cm[x2]),
                 "rows_in_mean": int(mask.sum()), "mask_margin": float(np.abs(img - max_image_effect).min())}
            if Eo is not None:
                cmo = Eo[mask].mean(axis=0)
                r["colmean_err"] = float(np.abs(cmo - cm).max())
                r["image_err"] = float(np.abs(imgo - img).max())
                r["gap_over_2err"] = float(r["gap"] / (2 * r["colmean_err"])) if r["colmean_err"] > 0 else float("inf")
                worst["ratio"] = max(worst["ratio"], 2 * r["colmean_err"] / r["gap"] if r["gap"] > 0 else np.inf)
                worst["mask_ratio"] = max(worst["mask_ratio"], r["image_err"] / r["mask_margin"] if r["mask_margin"] > 0 else np.inf)
                imgo = imgo + Eo[:, x]
                Eo[:, x] = 0
            worst["gap"] = min(worst["gap"], r["gap"])
            rounds.append(r)
            img = img + E[:, x]
            E[:, x] = 0
        out["classes"][str(c)] = rounds
    out["min_gap"] = float(worst["gap"])
    if eff_other is not None:
        out["worst_2err_over_gap"] = float(worst["ratio"])
        out["worst_image_err_over_mask_margin"] = float(worst["mask_ratio"])
        out["picks_provably_equal"] = bool(worst["ratio"] < 1 and worst["mask_ratio"] < 1 and out.get("label_flips", 0) == 0)
    return out


def grad_digest(t, cap=4096):
    """same as tests/golden/make_golden.py::grad_digest: strided sample of at most ``cap`` entries + (sum, sum |.|, sum .^2)."""
    f = t.detach().reshape(-1).double().cpu()
    step = max(1, (f.numel() + cap - 1) // cap)
    return np.concatenate([[float(f.sum()), float(f.abs().sum()), float((f * f).sum())], f[::step].numpy()])
