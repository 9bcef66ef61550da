"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference in the build container.

Used by ``tests/golden/make_golden.py`` (fixture generation) and
``tests/test_oracle_vs_reference.py`` (skipped where ``/root/reference`` is absent, e.g. on the
GPU box).  No reference source is copied: ``stylex_train.py`` is imported from where it lies and
notebook cells are ``exec``'d from the ``.ipynb`` JSON.  What blocks a plain import, and the shim
for each:

* missing third-party modules ``kornia`` / ``lpips`` / ``vector_quantize_pytorch`` / ``aim`` /
  ``fire`` (ST:4,7,30,37,49)  -> inert stand-ins in ``sys.modules``; ``kornia.filters.filter2d``
  is restated per kornia 0.6.2 (parity unpinned boundary, see stylex_oracle.blur3x3_reflect).
* ``assert torch.cuda.is_available()`` ST:51 and ``lpips.LPIPS(net="alex").cuda(0)`` ST:404
  -> ``torch.cuda.is_available`` patched to True during the import only.
* classifier constructors call ``torch.hub.load`` (network)  -> wrappers are built with
  ``__new__`` and a local torchvision model, ``classify_images`` then runs unmodified.
* notebook cell 5 calls ``.cuda(rank)`` and ``h5py``  -> identity ``.cuda`` on CPU and an
  in-memory h5py stand-in.
"""
from __future__ import annotations

import contextlib
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")   # oracle/make_ref.sh (git-ignored copy)


def _find_root() -> str:
    """The reference checkout: $STYLEX_REFERENCE_ROOT, else /root/reference (build container), else the files
    ``oracle/make_ref.sh`` staged under ``oracle/_ref`` (the GPU box, where /root/reference does not exist)."""
    env = os.environ.get("STYLEX_REFERENCE_ROOT")
    if env:
        return env
    for root in ("/root/reference", _STAGED):
        if os.path.isfile(os.path.join(root, "stylex", "stylex_train.py")):
            return root
    return "/root/reference"


REFERENCE_ROOT = _find_root()
_R = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "stylex", "stylex_train.py"))


def _filter2d(x, kernel, border_type="reflect", normalized=False, padding="same"):
    # kornia 0.6.2 filter2d restated: kernel [1,kh,kw]; normalise by sum(|k|); reflect pad; depthwise corr.
    k = kernel.to(x)
    if normalized:
        k = k / k.abs().sum(dim=(-2, -1), keepdim=True)
    kh, kw = k.shape[-2:]
    c = x.shape[1]
    xp = F.pad(x, (kw // 2, kw // 2, kh // 2, kh // 2), mode=border_type)
    return F.conv2d(xp, k[:, None].expand(c, 1, kh, kw).contiguous(), groups=c)


def load_reference():
    """import R/stylex/stylex_train.py unmodified; returns the module."""
    global _R
    if _R is not None:
        return _R
    if not available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")

    class _LPIPS:
        def __init__(self, *a, **k):
            pass

        def cuda(self, *a, **k):
            return self

        def __call__(self, *a, **k):
            raise RuntimeError("lpips stub")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules.setdefault(name, m)
        return sys.modules[name]

    mod("lpips", LPIPS=_LPIPS)
    k = mod("kornia")
    kf = mod("kornia.filters", filter2d=_filter2d)
    k.filters = kf
    mod("vector_quantize_pytorch", VectorQuantize=object)
    mod("aim", Session=object)
    mod("fire", Fire=lambda *a, **k: None)
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, "stylex"))
    real = torch.cuda.is_available
    torch.cuda.is_available = lambda: True
    try:
        import stylex_train as R  # noqa: the reference module, imported where it lies
    finally:
        torch.cuda.is_available = real
    _R = R
    return R


def reference_generator(state_dict, image_size, latent_dim=514, network_capacity=16, fmap_max=512):
    """The reference's own Generator (ST:747-840) carrying ``state_dict``."""
    R = load_reference()
    G = R.Generator(image_size, latent_dim, network_capacity=network_capacity, fmap_max=fmap_max)
    missing, unexpected = G.load_state_dict(state_dict, strict=False)
    # the only keys our synthetic state dicts leave out are the constant Blur buffers 'f'
    assert all(k.endswith(".f") for k in missing), missing
    assert not unexpected, unexpected
    return G.eval()


def reference_classifier(kind, model, image_size):
    """The reference's ResNet / MobileNet wrapper (classify_images unmodified) around ``model``."""
    load_reference()
    from torchvision.transforms import transforms
    if kind == "resnet":
        import resnet_classifier as rc
        c = rc.ResNet.__new__(rc.ResNet)
        c.resnet_dim = 224
    elif kind == "mobilenet":
        import mobilenet_classifier as mc
        c = mc.MobileNet.__new__(mc.MobileNet)
        c.mobilenet_dim = 224
    else:
        raise ValueError(kind)
    c.model = model.eval()
    c.image_size = image_size
    c.normalize = True
    c.tensor_transform = transforms.Compose(
        [transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    return c


class _MemDataset:
    def __init__(self, shape):
        self.arr = np.zeros(shape, dtype=np.float32)

    def __setitem__(self, k, v):
        self.arr[k] = v.numpy() if isinstance(v, torch.Tensor) else v

    def __getitem__(self, k):
        return self.arr[k]

    def __array__(self, dtype=None, copy=None):
        return self.arr if dtype is None else self.arr.astype(dtype)


class _MemFile:
    store = {}

    def __init__(self, path, mode="r"):
        self.path = path
        if mode == "w":
            _MemFile.store[path] = {}
        self.d = _MemFile.store[path]

    def create_dataset(self, name, shape, dtype="f"):
        self.d[name] = _MemDataset(shape)
        return self.d[name]

    def __getitem__(self, k):
        return self.d[k]

    def close(self):
        pass


def notebook_namespace(use_old_architecture=True, sindex_subset=None):
    """exec NB cells 5, 11 and 15 verbatim (attfind_extraction, find_significant_styles, helpers).

    ``sindex_subset``: bounded samples only -- the notebook wraps its style-coordinate loop in ``tqdm.tqdm`` (NB:356);
    the progress-bar stand-in injected here then yields just these coordinates.  The cell itself stays verbatim."""
    R = load_reference()
    nb = json.load(open(os.path.join(REFERENCE_ROOT, "stylex", "run_attfind_combined.ipynb")))
    h5 = types.ModuleType("h5py")
    h5.File = _MemFile

    class _tqdm_mod:
        @staticmethod
        def tqdm(x, *a, **k):
            if sindex_subset is not None and isinstance(x, range):
                keep = set(int(s) for s in sindex_subset)
                return [s for s in x if s in keep]
            return x

    ns = {
        "torch": torch, "np": np, "F": F, "os": os, "math": __import__("math"),
        "multiprocessing": __import__("multiprocessing"), "h5py": h5, "tqdm": _tqdm_mod,
        "USE_OLD_ARCHITECTURE": use_old_architecture,
        "styles_def_to_tensor": R.styles_def_to_tensor, "image_noise": R.image_noise,
        "Dataset": R.Dataset, "DistributedSampler": R.DistributedSampler, "MNIST_1vA": getattr(R, "MNIST_1vA", None),
        "cycle": R.cycle, "default": R.default, "DataLoader": torch.utils.data.DataLoader,
        "make_grid": None, "Image": None, "display": None, "print": lambda *a, **k: None,
    }
    for cell in (5, 11, 15):
        exec("".join(nb["cells"][cell]["source"]), ns)
    ns["_memfile"] = _MemFile
    return ns


@contextlib.contextmanager
def cpu_cuda_identity():
    """make ``.cuda(rank)`` a no-op so the notebook's loop runs on CPU."""
    t_old, m_old = torch.Tensor.cuda, torch.nn.Module.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = t_old, m_old


class FakeStylEx:
    """Just the attributes NB cell 5 touches: .G, .encoder, .D (NB:306,318,322)."""

    def __init__(self, G, encoder, D):
        self.G, self.encoder, self.D = G, encoder, D


def run_reference_attfind(G, classifier, images, encoder, noise, num_style_coords, shift_size=1.0,
                          discriminator=None, results_folder="mem://attfind", sindex_subset=None,
                          use_discriminator=False, discriminator_threshold=-0.5, num_images=None, extra_tail=True):
    """Run the VERBATIM notebook ``attfind_extraction`` (NB:269-417) on CPU; returns the 9 datasets.
    ``sindex_subset`` bounds the coordinate loop (see ``notebook_namespace``); untouched effect entries are whatever
    ``torch.Tensor(1, 2, S, 2)`` held (uninitialised, NB:354) -- callers of a bounded run must only read the subset.
    ``num_images`` < len(images) together with ``use_discriminator`` exercises the filter (NB:322-327)."""
    ns = notebook_namespace(True, sindex_subset=sindex_subset)
    stylex = FakeStylEx(G, encoder, discriminator if discriminator is not None else (lambda img: torch.zeros(1)))
    loader = [images[i: i + 1] for i in range(images.shape[0])]
    if extra_tail:
        loader = loader + [images[:1]]                                           # NB:300 reads one past the end
    with cpu_cuda_identity(), torch.no_grad():
        ns["attfind_extraction"](loader, images.shape[0] if num_images is None else num_images, results_folder, stylex,
                                 classifier, None, noise, num_style_coords, shift_size, discriminator_threshold,
                                 image_size=images.shape[-1], batch_size=1, cuda_rank=0,
                                 use_discriminator=use_discriminator)
    d = _MemFile.store[os.path.join(results_folder, "style_change_records.hdf5")]
    return {k: np.array(v.arr) for k, v in d.items()}
