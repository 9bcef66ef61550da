"""The StylEx training step (BASELINE config 5): host-side mirror of ``Trainer.train`` of the
reference (``stylex/stylex_train.py`` ST:1249-1506) and its loss helpers, over this package's modules.

What runs where: the generator's modulated convolutions, upsample, blur, noise / leaky-ReLU and style affines run on the
native kernels behind ``torch.autograd.Function``s (``modules.py``; first-order backward native too), the encoder /
discriminator convolutions and the classifier go through PyTorch / cuDNN (plain convolutions), the
optimiser is ``torch.optim.Adam`` like the reference's (ST:957-959).  The two penalties that differentiate twice --
``gradient_penalty`` through the discriminator (ST:296-303, every 4th step ST:1273,1345-1349) and ``calc_pl_lengths``
through the generator (ST:306-316, ST:1274,1422-1429) -- use the twice-differentiable compositions of ``modules.py``
(``Generator.double_backward``; the Blur / upsample Function pairs).

Not reproduced (default-off or impossible here): apex fp16, DiffAugment (``aug_prob = 0``), top-k training, the dual
contrastive loss, and the LPIPS term of ``reconstruction_loss`` -- its AlexNet weights are not in this image; pass
``lpips_fn`` to include it.
"""
from __future__ import annotations

import math
from contextlib import ExitStack, contextmanager
from random import random
from typing import Callable, Iterator, Optional

import torch
import torch.nn.functional as F
from torch.autograd import grad as torch_grad

from .modules import image_noise, styles_def_to_tensor


class NanException(Exception):
    """ST:68"""


def raise_if_nan(t):
    """ST:269-271"""
    if torch.isnan(t):
        raise NanException


def gen_hinge_loss(fake, real):
    """ST:382-383"""
    return fake.mean()


def hinge_loss(real, fake):
    """ST:386-387"""
    return (F.relu(1 + real) + F.relu(1 - fake)).mean()


def gradient_penalty(images, output, weight=10):
    """ST:296-303: the double backward runs through the discriminator (cuDNN convolutions + the native Blur pair)."""
    batch_size = images.shape[0]
    gradients = torch_grad(outputs=output, inputs=images, grad_outputs=torch.ones(output.size(), device=images.device),
                           create_graph=True, retain_graph=True, only_inputs=True)[0]
    gradients = gradients.reshape(batch_size, -1)
    return weight * ((gradients.norm(2, dim=1) - 1) ** 2).mean()


def calc_pl_lengths(styles, images, pl_noise: Optional[torch.Tensor] = None):
    """ST:306-316.  ``pl_noise`` (default: ``randn`` like the reference) can be handed in so that a test reproduces the
    reference's draw.  Needs a generator forward recorded with ``Generator.double_backward = True``."""
    device = images.device
    num_pixels = images.shape[2] * images.shape[3]
    if pl_noise is None:
        pl_noise = torch.randn(images.shape, device=device)
    pl_noise = pl_noise / math.sqrt(num_pixels)
    outputs = (images * pl_noise).sum()
    pl_grads = torch_grad(outputs=outputs, inputs=styles, grad_outputs=torch.ones(outputs.shape, device=device),
                          create_graph=True, retain_graph=True, only_inputs=True)[0]
    return (pl_grads ** 2).sum(dim=2).mean(dim=1).sqrt()


def lpips_normalize(images):
    """ST:370-377: per-image min / max rescaling to [-1, 1] (what the LPIPS AlexNet expects)."""
    flat = images.reshape(images.shape[0], -1)
    _max = flat.max(dim=1)[0].view(-1, 1, 1, 1)
    _min = flat.min(dim=1)[0].view(-1, 1, 1, 1)
    return (images - _min) / (_max - _min) * 2 - 1


def reconstruction_loss(encoder_batch, generated_images, generated_images_w, encoder_w, lpips_fn: Optional[Callable] = None):
    """ST:409-418: 0.1 LPIPS + 0.1 L1(w) + 1 L1(image).  The LPIPS term needs the AlexNet-LPIPS weights (absent offline):
    it is included only when ``lpips_fn(a, b) -> [B,...]`` is given."""
    loss = 0.1 * F.l1_loss(encoder_w, generated_images_w) + 1 * F.l1_loss(encoder_batch, generated_images)
    if lpips_fn is not None:
        loss = loss + 0.1 * lpips_fn(lpips_normalize(encoder_batch), lpips_normalize(generated_images)).mean()
    return loss


def classifier_kl_loss(real_classifier_logits, fake_classifier_logits):
    """ST:421-438: KLDivLoss(batchmean, log_target) between the log-softmaxes."""
    real = F.log_softmax(real_classifier_logits, dim=1)
    fake = F.log_softmax(fake_classifier_logits, dim=1)
    return F.kl_div(fake, real, reduction="batchmean", log_target=True)


class EMA:
    """ST:72-80"""

    def __init__(self, beta):
        self.beta = beta

    def update_average(self, old, new):
        if old is None:
            return new
        return old * self.beta + (1 - self.beta) * new


def noise_list(n, layers, latent_dim, device):
    """ST:319-324"""
    return [(torch.randn(n, latent_dim, device=device), layers)]


def mixed_list(n, layers, latent_dim, device):
    """ST:327-329"""
    tt = int(torch.rand(()).numpy() * layers)
    return noise_list(n, tt, latent_dim, device) + noise_list(n, layers - tt, latent_dim, device)


def latent_to_w(style_vectorizer, latent_descr):
    """ST:332-333"""
    return [(style_vectorizer(z), num_layers) for z, num_layers in latent_descr]


@contextmanager
def _no_sync(modules):
    with ExitStack() as stack:
        for m in modules:
            if hasattr(m, "no_sync"):
                stack.enter_context(m.no_sync())
        yield


class TrainStep:
    """One optimisation step of ``Trainer.train`` (ST:1249-1506): discriminator phase, generator phase, EMA bookkeeping.

    ``stylex``: the ``StylEx`` container (``.encoder .S .G .D .D_aug .SE .GE``); ``classifier``: a wrapper with
    ``classify_images`` (its parameters are frozen, gradients flow through it to the generator, ST:1390,1415-1416);
    ``loader``: an iterator of image batches [B,3,S,S] in [0,1] on the device.  With ``ddp=True`` (torch.distributed
    initialised) S / G / D and -- unlike the reference, which forgot it (ST:1190-1193) -- the encoder are wrapped in
    DistributedDataParallel; gradient accumulation uses ``no_sync`` like ``gradient_accumulate_contexts`` (ST:274-285).
    """

    def __init__(self, stylex, classifier, batch_size=4, lr=2e-4, ttur_mult=2, mixed_prob=0.9, gradient_accumulate_every=1,
                 rel_disc_loss=False, no_pl_reg=False, kl_scaling=1, rec_scaling=10, alternating_training=True,
                 lpips_fn: Optional[Callable] = None, ddp=False, rank=0, steps=0, pl_after=5000):
        self.StylEx, self.classifier = stylex, classifier
        self.batch_size, self.mixed_prob = batch_size, mixed_prob
        self.gradient_accumulate_every = gradient_accumulate_every
        self.rel_disc_loss, self.no_pl_reg = rel_disc_loss, no_pl_reg
        self.kl_scaling, self.rec_scaling = kl_scaling, rec_scaling
        self.alternating_training = alternating_training
        self.lpips_fn = lpips_fn
        self.rank, self.steps, self.pl_after = rank, steps, pl_after
        self.pl_mean, self.pl_length_ma = None, EMA(0.99)
        self.d_loss = self.g_loss = self.total_rec_loss = self.total_kl_loss = 0.0
        self.last_gp_loss = None
        generator_params = list(stylex.G.parameters()) + list(stylex.S.parameters()) + list(stylex.encoder.parameters())
        self.G_opt = torch.optim.Adam(generator_params, lr=lr, betas=(0.5, 0.9))                       # ST:957-958
        self.D_opt = torch.optim.Adam(stylex.D.parameters(), lr=lr * ttur_mult, betas=(0.5, 0.9))      # ST:959
        self.is_ddp = ddp
        self.S, self.G, self.D, self.E = stylex.S, stylex.G, stylex.D, stylex.encoder
        if ddp:
            from torch.nn.parallel import DistributedDataParallel as DDP
            kw = {"device_ids": [rank]}
            self.S, self.G, self.D = DDP(stylex.S, **kw), DDP(stylex.G, **kw), DDP(stylex.D, **kw)
            self.E = DDP(stylex.encoder, **kw)

    # ------------------------------------------------------------------------------------------------------------
    def _latents(self, loader, use_encoder, batch_size):
        """-> (w_styles [B,L,latent], noise, encoder_batch | None, encoder_output | None, real_logits | None)"""
        G = self.StylEx.G
        if use_encoder:
            batch = next(loader).requires_grad_()
            encoder_output = self.E(batch)                                             # ST:1310 / 1378
            real_logits = self.classifier.classify_images(batch)                       # ST:1311 / 1379
            style = [(torch.cat((encoder_output, real_logits), dim=1), G.num_layers)]  # ST:1312-1313
            return styles_def_to_tensor(style), image_noise(batch_size, G.image_size, self.rank), batch, encoder_output, real_logits
        fn = mixed_list if random() < self.mixed_prob else noise_list                 # ST:1320
        style = fn(batch_size, G.num_layers, G.latent_dim, device=torch.device("cuda", self.rank))
        w_space = latent_to_w(self.S, style)
        return styles_def_to_tensor(w_space), image_noise(batch_size, G.image_size, self.rank), None, None, None

    @torch.enable_grad()
    def train_step(self, loader: Iterator[torch.Tensor]) -> dict:
        st = self.StylEx
        st.train()
        bs = self.batch_size
        every = self.gradient_accumulate_every
        apply_gradient_penalty = self.steps % 4 == 0                                                   # ST:1273
        apply_path_penalty = (not self.no_pl_reg) and self.steps > self.pl_after and self.steps % 32 == 0   # ST:1274
        total_disc = total_gen = total_rec = total_kl = 0.0

        # ---------------- discriminator phase, ST:1293-1360 ----------------
        self.D_opt.zero_grad()
        encoder_input = False
        st.G.double_backward = False
        for i in range(every):
            with ExitStack() as stack:
                if self.is_ddp and i < every - 1:
                    stack.enter_context(_no_sync([self.D, self.S, self.G, self.E]))
                discriminator_batch = next(loader).requires_grad_()
                use_enc = (not self.alternating_training) or encoder_input
                w_styles, noise, _, _, _ = self._latents(loader, use_enc, bs)
                encoder_input = False if use_enc else (True if self.alternating_training else encoder_input)
                # ST:1330-1331 detaches the fake images before the discriminator sees them: no gradient reaches G / S / the
                # encoder in this phase, so the generator runs without a graph (the fused plan; same values)
                with torch.no_grad():
                    generated_images = self.G(w_styles.detach().contiguous(), noise)
                fake_output = self.D(generated_images.clone().detach())
                real_output = self.D(discriminator_batch)
                real_l, fake_l = real_output, fake_output
                if self.rel_disc_loss:
                    real_l = real_l - fake_output.mean()
                    fake_l = fake_l - real_output.mean()
                divergence = hinge_loss(real_l, fake_l)
                disc_loss = divergence
                if apply_gradient_penalty:
                    gp = gradient_penalty(discriminator_batch, real_output)
                    self.last_gp_loss = gp.detach().item()
                    disc_loss = disc_loss + gp
                disc_loss = disc_loss / every
                disc_loss.register_hook(raise_if_nan)
                disc_loss.backward()
                total_disc += divergence.detach().item() / every
        self.d_loss = float(total_disc)
        self.D_opt.step()

        # ---------------- generator phase, ST:1362-1467 ----------------
        encoder_input = False
        self.G_opt.zero_grad()
        st.G.double_backward = bool(apply_path_penalty)
        avg_pl_length = self.pl_mean
        for i in range(every):
            with ExitStack() as stack:
                if self.is_ddp and i < every - 1:
                    stack.enter_context(_no_sync([self.S, self.G, self.D, self.E]))
                use_enc = (not self.alternating_training) or encoder_input
                w_styles, noise, image_batch, encoder_output, real_logits = self._latents(loader, use_enc, bs)
                generated_images = self.G(w_styles, noise)
                gen_logits = self.classifier.classify_images(generated_images)                         # ST:1390
                fake_output = self.D(generated_images)
                if use_enc:   # ST:1412-1416 (x2: these losses exist every other iteration under alternating training)
                    rec_loss = 2 * self.rec_scaling * reconstruction_loss(image_batch, generated_images, self.E(generated_images),
                                                                          encoder_output, self.lpips_fn) / every
                    kl_loss = 2 * self.kl_scaling * classifier_kl_loss(real_logits, gen_logits) / every
                loss = gen_hinge_loss(fake_output, None)
                gen_loss = loss
                if apply_path_penalty:
                    pl_lengths = calc_pl_lengths(w_styles, generated_images)
                    avg_pl_length = float(pl_lengths.detach().mean().item())
                    if self.pl_mean is not None:
                        pl_loss = ((pl_lengths - self.pl_mean) ** 2).mean()
                        if not torch.isnan(pl_loss):
                            gen_loss = gen_loss + pl_loss
                gen_loss = gen_loss / every
                gen_loss.register_hook(raise_if_nan)
                if use_enc:
                    gen_loss.backward(retain_graph=True)                                               # ST:1436-1438
                    rec_loss.backward(retain_graph=True)
                    kl_loss.backward()
                    total_rec += rec_loss.detach().item()
                    total_kl += kl_loss.detach().item()
                else:
                    gen_loss.backward()
                total_gen += loss.detach().item() / every
                encoder_input = not encoder_input
        self.g_loss, self.total_rec_loss, self.total_kl_loss = float(total_gen), float(total_rec), float(total_kl)
        self.G_opt.step()
        st.G.double_backward = False

        if apply_path_penalty and avg_pl_length is not None and not math.isnan(avg_pl_length):       # ST:1471-1473
            self.pl_mean = self.pl_length_ma.update_average(self.pl_mean, avg_pl_length)
        if self.steps % 10 == 0 and self.steps > 20000:                                                 # ST:1475-1476
            self._ema()
        if self.steps <= 25000 and self.steps % 1000 == 2:                                              # ST:1478-1479
            st.reset_parameter_averaging()
        for v in (total_gen, total_disc):                                                               # ST:1483-1486
            if math.isnan(v):
                raise NanException
        self.steps += 1
        return {"D": self.d_loss, "G": self.g_loss, "Rec": self.total_rec_loss, "KL": self.total_kl_loss,
                "GP": self.last_gp_loss if apply_gradient_penalty else None, "PL": self.pl_mean}

    def _ema(self, beta=0.995):
        """StylEx.EMA ST:985-993"""
        st = self.StylEx
        with torch.no_grad():
            for ma, cur in ((st.SE, st.S), (st.GE, st.G)):
                for cp, mp in zip(cur.parameters(), ma.parameters()):
                    mp.data = mp.data * beta + (1 - beta) * cp.data
