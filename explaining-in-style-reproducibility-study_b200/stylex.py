"""The callers either side of the sweep: the StylEx container with its encoder /
discriminator, the reference's checkpoint format, and the batched phase-A front end of ``attfind_extraction``.

Mirrors ``stylex/stylex_train.py`` (ST): ``DiscriminatorBlock`` ST:721-744, ``DiscriminatorE`` ST:842-909,
``EqualLinear`` / ``StyleVectorizer`` ST:576-601, ``StylEx`` ST:912-1000, ``Trainer.save/load/config`` ST:1198-1218,
1736-1774, and the notebook's ``model_loader`` (NB cell 6) -- same class names, constructor arguments, attribute
names and state-dict keys, so a ``model_<n>.pt`` written by the reference loads with ``strict=True``.

The encoder / discriminator are plain convolutions: they run through PyTorch (cuDNN) exactly like the classifier
("plain convs -> cuDNN is fine, just batch it"); the depthwise ``Blur`` in front of every
down-sampling conv is the native kernel (``sx_blur3x3_reflect``).  What changes is the batching: the notebook
encodes, classifies and generates image by image (NB:300-336, batch 1); ``encode_images`` does the same arithmetic
on ``batch`` images per launch.

Inference only: no optimisers, no augmentation wrapper state (``D_aug`` shares ``D``'s tensors), no apex.
"""
from __future__ import annotations

import json
import math
import os
from functools import partial
from math import log2
from pathlib import Path
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from .modules import Blur, Generator, exists, leaky_relu, styles_def_to_tensor

__reference_version__ = "1.8.7"     # the 'version' entry the reference's Trainer.save writes (ST:1738-1741)


class Flatten(nn.Module):
    def forward(self, x):
        return x.reshape(x.shape[0], -1)


class EqualLinear(nn.Module):
    """ST:576-586."""

    def __init__(self, in_dim, out_dim, lr_mul=1, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_dim))
        self.lr_mul = lr_mul

    def forward(self, input):
        return F.linear(input, self.weight * self.lr_mul, bias=self.bias * self.lr_mul)


class StyleVectorizer(nn.Module):
    """ST:589-601 (the mapping network; not used by AttFind, kept for checkpoint compatibility)."""

    def __init__(self, emb, depth, lr_mul=0.1):
        super().__init__()
        layers = []
        for _ in range(depth):
            layers.extend([EqualLinear(emb, emb, lr_mul), leaky_relu()])
        self.net = nn.Sequential(*layers)

    def forward(self, x):
        x = F.normalize(x, dim=1)
        return self.net(x)


class DiscriminatorBlock(nn.Module):
    """ST:721-744."""

    def __init__(self, input_channels, filters, downsample=True):
        super().__init__()
        self.conv_res = nn.Conv2d(input_channels, filters, 1, stride=(2 if downsample else 1))
        self.net = nn.Sequential(
            nn.Conv2d(input_channels, filters, 3, padding=1),
            leaky_relu(),
            nn.Conv2d(filters, filters, 3, padding=1),
            leaky_relu()
        )
        self.downsample = nn.Sequential(
            Blur(),
            nn.Conv2d(filters, filters, 3, padding=1, stride=2)
        ) if downsample else None

    def forward(self, x):
        res = self.conv_res(x)
        x = self.net(x)
        if exists(self.downsample):
            x = self.downsample(x)
        x = (x + res) * (1 / math.sqrt(2))
        return x


class DiscriminatorE(nn.Module):
    """ST:842-909: the discriminator, and with ``encoder=True`` the image encoder (fc -> encoder_dim)."""

    def __init__(self, image_size, network_capacity=16, fq_layers=[], fq_dict_size=256, attn_layers=[],
                 transparent=False, encoder=False, encoder_dim=512, fmap_max=512):
        super().__init__()
        if fq_layers or attn_layers or transparent:
            raise NotImplementedError("stylex_b200 DiscriminatorE: fq_layers / attn_layers / transparent are default-off "
                                      "extras of the reference outside the AttFind path (SURVEY.md section 2)")
        num_layers = int(log2(image_size) - 1)
        num_init_filters = 3
        filters = [num_init_filters] + [(network_capacity * 4) * (2 ** i) for i in range(num_layers + 1)]
        set_fmap_max = partial(min, fmap_max)
        filters = list(map(set_fmap_max, filters))
        chan_in_out = list(zip(filters[:-1], filters[1:]))

        blocks = []
        for ind, (in_chan, out_chan) in enumerate(chan_in_out):
            is_not_last = ind != (len(chan_in_out) - 1)
            blocks.append(DiscriminatorBlock(in_chan, out_chan, downsample=is_not_last))
        self.blocks = nn.ModuleList(blocks)
        self.attn_blocks = nn.ModuleList([None] * len(blocks))        # key-less, like the reference's lists of None
        self.quantize_blocks = nn.ModuleList([None] * len(blocks))

        chan_last = filters[-1]
        latent_dim = 2 * 2 * chan_last
        self.final_conv = nn.Conv2d(chan_last, chan_last, 3, padding=1)
        self.flatten = Flatten()
        self.encoder_dim = encoder_dim
        self.fc = nn.Linear(latent_dim, 1 if not encoder else self.encoder_dim)

    def forward(self, x):
        for block in self.blocks:
            x = block(x)
        x = self.final_conv(x)
        x = self.flatten(x)
        x = self.fc(x)
        return x.squeeze()          # ST:909: [B, D] for B > 1, [D] for a single image (the notebook unsqueezes it back)


class AugWrapper(nn.Module):
    """ST:558-571 without the augmentation (training only): holds ``D`` so that the ``D_aug.D.*`` keys exist."""

    def __init__(self, D, image_size):
        super().__init__()
        self.D = D

    def forward(self, images, prob=0., types=[], detach=False):
        if prob > 0:
            raise NotImplementedError("DiffAugment is part of the training step (SURVEY.md section 8f row 1)")
        if detach:
            images = images.detach()
        return self.D(images)


class StylEx(nn.Module):
    """ST:912-1000, inference side: ``.encoder .S .G .D .SE .GE .D_aug`` with the reference's state-dict keys."""

    def __init__(self, image_size, latent_dim=514, fmap_max=512, style_depth=8, network_capacity=16, transparent=False,
                 fp16=False, cl_reg=False, steps=1, lr=1e-4, ttur_mult=2, fq_layers=[], fq_dict_size=256,
                 attn_layers=[], no_const=False, lr_mlp=0.1, rank=0, classifier_labels=2, encoder_class=None,
                 kl_rec_during_disc=False):
        super().__init__()
        if fp16 or cl_reg or encoder_class is not None:
            raise NotImplementedError("stylex_b200 StylEx: fp16 (apex) / cl_reg / debug encoders are outside the AttFind path")
        self.lr = lr
        self.steps = steps
        self.image_size = image_size
        self.latent_dim = latent_dim
        self.encoder = DiscriminatorE(image_size, network_capacity, encoder=True, fq_layers=fq_layers, fq_dict_size=fq_dict_size,
                                      attn_layers=attn_layers, transparent=transparent, fmap_max=fmap_max)
        self.S = StyleVectorizer(latent_dim, style_depth, lr_mul=lr_mlp)
        self.G = Generator(image_size, latent_dim, network_capacity, transparent=transparent, attn_layers=attn_layers,
                           no_const=no_const, fmap_max=fmap_max)
        self.D = DiscriminatorE(image_size, network_capacity, fq_layers=fq_layers, fq_dict_size=fq_dict_size,
                                attn_layers=attn_layers, transparent=transparent, fmap_max=fmap_max)
        self.SE = StyleVectorizer(latent_dim, style_depth, lr_mul=lr_mlp)
        self.GE = Generator(image_size, latent_dim, network_capacity, transparent=transparent, attn_layers=attn_layers,
                            no_const=no_const)
        self.D_cl = None
        self.D_aug = AugWrapper(self.D, image_size)
        for p in list(self.SE.parameters()) + list(self.GE.parameters()):
            p.requires_grad_(False)
        self._init_weights()
        self.reset_parameter_averaging()
        self.fp16 = False
        if torch.cuda.is_available():
            self.cuda(rank)             # ST:965

    def _init_weights(self):
        """ST:974-983."""
        for m in self.modules():
            if type(m) in {nn.Conv2d, nn.Linear}:
                nn.init.kaiming_normal_(m.weight, a=0, mode='fan_in', nonlinearity='leaky_relu')
        for block in self.G.blocks:
            nn.init.zeros_(block.to_noise1.weight)
            nn.init.zeros_(block.to_noise2.weight)
            nn.init.zeros_(block.to_noise1.bias)
            nn.init.zeros_(block.to_noise2.bias)

    def reset_parameter_averaging(self):
        self.SE.load_state_dict(self.S.state_dict())
        self.GE.load_state_dict(self.G.state_dict())

    def forward(self, x):
        return x


# ---------------------------------------------------------------------------------------------
# on-disk formats: model_<n>.pt + .config.json (ST:1198-1218, 1736-1774)
# ---------------------------------------------------------------------------------------------
CONFIG_KEYS = ("image_size", "network_capacity", "lr_mlp", "transparent", "fq_layers", "fq_dict_size", "attn_layers", "no_const")


def stylex_config(image_size, network_capacity=16, lr_mlp=0.1, transparent=False, fq_layers=(), fq_dict_size=256,
                  attn_layers=(), no_const=False) -> Dict:
    """``Trainer.config()`` ST:1213-1216."""
    return {'image_size': image_size, 'network_capacity': network_capacity, 'lr_mlp': lr_mlp, 'transparent': transparent,
            'fq_layers': list(fq_layers), 'fq_dict_size': fq_dict_size, 'attn_layers': list(attn_layers), 'no_const': no_const}


def model_name(models_dir, name, num) -> str:
    """``Trainer.model_name`` : <models_dir>/<name>/model_<num>.pt"""
    return str(Path(models_dir) / name / f'model_{num}.pt')


def config_path(models_dir, name) -> str:
    return str(Path(models_dir) / name / '.config.json')


def save_checkpoint(stylex: StylEx, models_dir, name, num, config: Optional[Dict] = None) -> str:
    """``Trainer.save`` ST:1736-1746: ``{'StylEx': state_dict, 'version': ...}`` + ``.config.json``."""
    path = model_name(models_dir, name, num)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save({'StylEx': stylex.state_dict(), 'version': __reference_version__}, path)
    cfg = config or stylex_config(stylex.image_size)
    Path(config_path(models_dir, name)).write_text(json.dumps(cfg))
    return path


def latest_checkpoint(models_dir, name) -> Optional[int]:
    """``Trainer.load(num=-1)`` ST:1752-1758: the highest ``model_<n>.pt``."""
    nums = sorted(int(p.stem.split('_')[1]) for p in (Path(models_dir) / name).glob('model_*.pt'))
    return nums[-1] if nums else None


def load_checkpoint(models_dir, name='default', num=-1, rank=0, **stylex_kwargs) -> Tuple[StylEx, Dict]:
    """``Trainer.load`` ST:1748-1774: read ``.config.json`` (defaults where absent, ST:1201-1211), build the StylEx it
    describes and load ``model_<num>.pt`` strictly.  Returns (model in eval mode, config)."""
    cp = Path(config_path(models_dir, name))
    cfg = json.loads(cp.read_text()) if cp.exists() else None
    if num == -1:
        num = latest_checkpoint(models_dir, name)
        if num is None:
            raise FileNotFoundError(f"no model_*.pt under {Path(models_dir) / name}")
    data = torch.load(model_name(models_dir, name, num), map_location="cpu")
    sd = data['StylEx']
    if cfg is None:     # the reference falls back to the Trainer's constructor arguments; here: read them off the tensors
        cfg = config_from_state_dict(sd)
    model = StylEx(cfg['image_size'], network_capacity=cfg['network_capacity'], transparent=cfg.get('transparent', False),
                   fq_layers=cfg.get('fq_layers', []), fq_dict_size=cfg.get('fq_dict_size', 256),
                   attn_layers=cfg.get('attn_layers', []), no_const=cfg.get('no_const', False),
                   lr_mlp=cfg.get('lr_mlp', 0.1), fmap_max=cfg.get('fmap_max', 512), rank=rank, **stylex_kwargs)
    model.load_state_dict(sd)           # strict, like ST:1768
    return model.eval(), cfg


def config_from_state_dict(sd: Dict[str, torch.Tensor]) -> Dict:
    """image_size / network_capacity / fmap_max recovered from the generator tensors of a StylEx state dict."""
    n_blocks = len({k.split('.')[2] for k in sd if k.startswith('G.blocks.')})
    image_size = 4 << (n_blocks - 1)
    last_co = sd[f'G.blocks.{n_blocks - 1}.conv2.weight'].shape[0]          # = network_capacity * 2 (ST:755)
    fmap_max = sd['G.initial_block'].shape[1]
    c0 = fmap_max
    uncapped = last_co * (2 ** (n_blocks - 1))                                # first filter count before min(fmap_max, .)
    cfg = stylex_config(image_size, network_capacity=last_co // 2)
    cfg['fmap_max'] = c0 if c0 < uncapped else max(512, c0)
    return cfg


def load_stylex(stylex_path, image_size, rank=0, **kwargs) -> StylEx:
    """The notebook's way (NB cell 6): ``StylEx(image_size=...)`` + ``load_state_dict(torch.load(path)["StylEx"])``."""
    model = StylEx(image_size=image_size, rank=rank, **kwargs)
    model.load_state_dict(torch.load(stylex_path, map_location="cpu")["StylEx"])
    return model.eval()


def model_loader(stylex_path, classifier_name, image_size, cuda_rank):
    """NB cell 6, same signature: (StylEx, classifier wrapper)."""
    from .classifiers import MobileNet, ResNet
    init_stylex = load_stylex(stylex_path, image_size, rank=cuda_rank)
    if "mobilenet" in classifier_name.lower():
        init_classifier = MobileNet(classifier_name, cuda_rank=cuda_rank, output_size=2, image_size=image_size)
    elif "resnet" in classifier_name.lower():
        init_classifier = ResNet(classifier_name, cuda_rank=cuda_rank, output_size=2, image_size=image_size)
    else:
        raise NotImplementedError("This classifier is not supported yet, please add support or change the filename to "
                                  "contain MobileNet or ResNet.")
    return init_stylex, init_classifier


# ---------------------------------------------------------------------------------------------
# phase A of attfind_extraction, batched (NB:300-336)
# ---------------------------------------------------------------------------------------------
@torch.no_grad()
def encode_images(stylex, classifier, images: torch.Tensor, noise: Optional[torch.Tensor] = None, batch: int = 256,
                  use_old_architecture: bool = True, discriminator: bool = False) -> Dict[str, torch.Tensor]:
    """NB:300-336 for ``images`` [N,3,S,S] on the device, ``batch`` images per launch instead of one:

        w          = encoder(image)                                  NB:306
        logits     = classifier.classify_images(image)               NB:307
        concat_w   = cat(w, logits | softmax(logits))                NB:310-314
        generated  = G(styles_def_to_tensor([(concat_w, L)]), noise) NB:318-320 (only with ``discriminator``)
        d_out      = D(generated)                                    NB:322-325

    Returns 'latents' [N, latent_dim], 'logits' [N,2] and, with ``discriminator``, 'discriminator' [N,1] (+ 'generated').
    """
    n = images.shape[0]
    G = stylex.G
    lat = torch.empty(n, G.latent_dim, device=images.device, dtype=torch.float32)
    logits = torch.empty(n, 2, device=images.device, dtype=torch.float32)
    out = {"latents": lat, "logits": logits}
    if discriminator:
        if noise is None:
            raise ValueError("the discriminator filter needs the generator noise")
        out["discriminator"] = torch.zeros(n, 1, device=images.device, dtype=torch.float32)
    for i in range(0, n, batch):
        x = images[i: i + batch].float()
        w = stylex.encoder(x).reshape(x.shape[0], -1)
        lg = classifier.classify_images(x).float()
        logits[i: i + batch] = lg
        lat[i: i + batch] = torch.cat((w, lg if use_old_architecture else torch.softmax(lg, dim=1)), dim=1)
        if discriminator:
            gen = G(styles_def_to_tensor([(lat[i: i + batch], G.num_layers)]), noise)
            if use_old_architecture:
                d = stylex.D(gen)
            else:
                # the new architecture's discriminator (stylex_train_new.py) is conditioned on the class probabilities; a
                # discriminator with that forward signature works here, this package's DiscriminatorE is the old one
                d = stylex.D(gen, probabilities=torch.softmax(classifier.classify_images(gen), dim=1))
            out["discriminator"][i: i + batch] = d.reshape(-1, 1)
    return out


@torch.no_grad()
def find_discriminator_threshold(stylex, classifier, dataloader, num_images, threshold_folder, dataset_name=None, image_size=64,
                                 batch_size=1, cuda_rank=0, noise=None, use_old_architecture: bool = True, front_batch: int = 256):
    """NB cell 5 ``find_discriminator_threshold``: the discriminator's output on the reconstruction of the first ``num_images``
    images of ``dataloader`` (encode -> classify -> generate -> discriminate), written to ``discriminator_threshold.hdf5``
    (datasets ``discriminator_outputs`` [N,1] and ``generated_images`` [N,3,S,S]) and returned.  ``front_batch`` images per
    launch instead of the notebook's one; the notebook reads ``noise`` from its global scope, here it is an argument."""
    if noise is None:
        raise ValueError("find_discriminator_threshold needs the generator noise (a notebook global in the reference)")
    dev = noise.device
    G = stylex.G
    outputs = torch.zeros(num_images, 1, device=dev)
    generated = torch.zeros(num_images, 3, image_size, image_size, device=dev)
    it = iter(dataloader)
    done = 0
    while done < num_images:
        chunk = []
        for image in it:
            chunk.append(image.to(dev).reshape(-1, 3, image_size, image_size))
            if len(chunk) >= min(front_batch, num_images - done):
                break
        if not chunk:
            raise StopIteration(f"the dataloader ran out after {done} of {num_images} images")      # next(dataloader) in the notebook
        x = torch.cat(chunk)[: num_images - done]
        logits = classifier.classify_images(x).float()
        w = stylex.encoder(x).reshape(x.shape[0], -1)
        lat = torch.cat((w, logits if use_old_architecture else torch.softmax(logits, dim=1)), dim=1)
        gen = G(styles_def_to_tensor([(lat, G.num_layers)]), noise)
        if use_old_architecture:
            d = stylex.D(gen)
        else:
            d = stylex.D(gen, probabilities=torch.softmax(classifier.classify_images(gen), dim=1))
        k = x.shape[0]
        outputs[done: done + k] = d.reshape(-1, 1)
        generated[done: done + k] = gen
        done += k
    result = {"discriminator_outputs": outputs, "generated_images": generated}
    if threshold_folder is not None:
        os.makedirs(threshold_folder, exist_ok=True)
        path = os.path.join(threshold_folder, "discriminator_threshold.hdf5")
        arrays = {k: v.float().cpu().numpy() for k, v in result.items()}
        try:
            import h5py
        except ImportError:
            from . import hdf5_lite
            hdf5_lite.write_hdf5(path, arrays)
        else:
            with h5py.File(path, "w") as f:
                for k, v in arrays.items():
                    f.create_dataset(k, v.shape, dtype="f")[:] = v
    return result

