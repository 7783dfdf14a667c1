"""ORACLE (test infrastructure only) -- stock ``torch.nn`` restatement of the reference hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; the product package never does (no CPU fallback).

What is restated, and from where (``/root/reference`` is cited as ``ref:``):

* ``DownSampleConv``            ref:src/model.py:42-65
* ``Generator``                 ref:src/model.py:15-39
* ``Discriminator``             ref:src/model.py:68-92
* ``BasicUNet`` and its blocks  third-party: ``monai==1.3.0`` (ref:requirements.txt:3, call site
  ref:src/model.py:22-28).  MONAI is absent from the container (no network), so its published
  structure is restated: ``monai/networks/nets/basic_unet.py`` (TwoConv / Down / UpCat /
  BasicUNet), ``monai/networks/blocks/convolutions.py`` (Convolution = conv -> ADN, ordering "NDA"),
  ``monai/networks/blocks/acti_norm.py`` (ADN submodule names N, D, A) and
  ``monai/networks/blocks/upsample.py`` (UpSample mode "deconv": ConvTranspose3d(k=2, s=2) named
  ``deconv``).  Defaults used by the reference call: act LeakyReLU(0.1, inplace), norm
  InstanceNorm3d(affine=True), bias=True, dropout=0.05 (element-wise nn.Dropout), upsample="deconv".
* ``gan_step``                  ref:src/model.py:170-193 and 259-281 (training_step semantics)

PARITY PIN: the reference ships no tests / golden vectors of its own for this path.
The restatement is pinned by OUTPUTS OF THE REFERENCE ITSELF run in the build container:
``tests/golden/make_golden_ref.py`` imports /root/reference/src/model.py unmodified (stand-ins for the absent
monai / lightning / torchio / nibabel packages; ``monai.networks.nets.BasicUNet`` is the restatement below) and
records module trees, initialisation checksums, forward values, ``_gen_step`` / ``_discr_step`` losses, phase
gradients under ``toggle_optimizer`` and three whole ``training_step`` calls into
``tests/golden/golden_ref_v1.npz``; ``tests/test_reference_golden_cpu.py`` checks this module against them.
Status: pinned by the reference for everything in model.py; the BasicUNet internals (third-party monai 1.3.0,
absent here) remain a restatement pinned by the structural known answers of SURVEY.md Appendix A
(22 645 318 U-Net parameters; key names and shapes of the published monai 1.3.0 source).
"""
from __future__ import annotations

import torch
from torch import nn
import torch.nn.functional as F

MODALITIES = ("dwi-tensor", "pc-bssfp", "bssfp", "t1w")
UNET_FEATURES = (32, 64, 128, 256, 512, 32)


# --------------------------------------------------------------------------------------------
# model.py restatement
# --------------------------------------------------------------------------------------------
class DownSampleConv(nn.Module):
    """ref:src/model.py:42-65 -- Conv3d -> optional BatchNorm3d -> optional LeakyReLU(0.2)."""

    def __init__(self, in_channels, out_channels, kernel=4, strides=2, padding=1, activation=True,
                 batchnorm=True):
        super().__init__()
        self.activation = activation
        self.batchnorm = batchnorm
        self.conv = nn.Conv3d(in_channels, out_channels, kernel, strides, padding)
        if batchnorm:
            self.bn = nn.BatchNorm3d(out_channels)
        if activation:
            self.act = nn.LeakyReLU(0.2)

    def forward(self, x):
        x = self.conv(x)
        if self.batchnorm:
            x = self.bn(x)
        if self.activation:
            x = self.act(x)
        return x


class _ADN(nn.Sequential):
    """monai ADN with ordering "NDA": InstanceNorm3d(affine) -> Dropout(p) -> LeakyReLU(0.1)."""

    def __init__(self, channels, dropout):
        super().__init__()
        self.add_module("N", nn.InstanceNorm3d(channels, affine=True))
        if dropout is not None and dropout > 0:
            self.add_module("D", nn.Dropout(dropout))
        self.add_module("A", nn.LeakyReLU(negative_slope=0.1, inplace=True))


class _Convolution(nn.Sequential):
    """monai Convolution (3-D, kernel 3, stride 1, padding 1, bias) followed by ADN."""

    def __init__(self, in_chns, out_chns, dropout):
        super().__init__()
        self.add_module("conv", nn.Conv3d(in_chns, out_chns, kernel_size=3, stride=1, padding=1, bias=True))
        self.add_module("adn", _ADN(out_chns, dropout))


class _TwoConv(nn.Sequential):
    def __init__(self, in_chns, out_chns, dropout):
        super().__init__()
        self.add_module("conv_0", _Convolution(in_chns, out_chns, dropout))
        self.add_module("conv_1", _Convolution(out_chns, out_chns, dropout))


class _Down(nn.Sequential):
    def __init__(self, in_chns, out_chns, dropout):
        super().__init__()
        self.add_module("max_pooling", nn.MaxPool3d(kernel_size=2))
        self.add_module("convs", _TwoConv(in_chns, out_chns, dropout))


class _UpSample(nn.Sequential):
    def __init__(self, in_chns, out_chns):
        super().__init__()
        self.add_module("deconv", nn.ConvTranspose3d(in_chns, out_chns, kernel_size=2, stride=2, bias=True))


class _UpCat(nn.Module):
    def __init__(self, in_chns, cat_chns, out_chns, dropout, halves=True):
        super().__init__()
        up_chns = in_chns // 2 if halves else in_chns
        self.upsample = _UpSample(in_chns, up_chns)
        self.convs = _TwoConv(cat_chns + up_chns, out_chns, dropout)

    def forward(self, x, x_e):
        x_0 = self.upsample(x)
        # replicate-pad one voxel at the end of any axis whose size differs (odd edge lengths)
        dims = x.dim() - 2
        sp = [0] * (dims * 2)
        for i in range(dims):
            if x_e.shape[-i - 1] != x_0.shape[-i - 1]:
                sp[i * 2 + 1] = 1
        if any(sp):
            x_0 = F.pad(x_0, sp, "replicate")
        return self.convs(torch.cat([x_e, x_0], dim=1))


class BasicUNet(nn.Module):
    """monai==1.3.0 BasicUNet(spatial_dims=3, in_channels, out_channels, features, dropout)."""

    def __init__(self, in_channels=24, out_channels=6, features=UNET_FEATURES, dropout=0.05):
        super().__init__()
        f = tuple(features)
        self.conv_0 = _TwoConv(in_channels, f[0], dropout)
        self.down_1 = _Down(f[0], f[1], dropout)
        self.down_2 = _Down(f[1], f[2], dropout)
        self.down_3 = _Down(f[2], f[3], dropout)
        self.down_4 = _Down(f[3], f[4], dropout)
        self.upcat_4 = _UpCat(f[4], f[3], f[3], dropout)
        self.upcat_3 = _UpCat(f[3], f[2], f[2], dropout)
        self.upcat_2 = _UpCat(f[2], f[1], f[1], dropout)
        self.upcat_1 = _UpCat(f[1], f[0], f[5], dropout, halves=False)
        self.final_conv = nn.Conv3d(f[5], out_channels, kernel_size=1)

    def forward(self, x):
        x0 = self.conv_0(x)
        x1 = self.down_1(x0)
        x2 = self.down_2(x1)
        x3 = self.down_3(x2)
        x4 = self.down_4(x3)
        u4 = self.upcat_4(x4, x3)
        u3 = self.upcat_3(u4, x2)
        u2 = self.upcat_2(u3, x1)
        u1 = self.upcat_1(u2, x0)
        return self.final_conv(u1)


class Generator(nn.Module):
    """ref:src/model.py:15-39 -- 1x1x1 input head (shared objects under two keys each) + BasicUNet."""

    def __init__(self, input_modality):
        super().__init__()
        self.input_modality = input_modality
        dwi_tensor_input = DownSampleConv(6, 24, kernel=1, strides=1, padding=0)
        bssfp_input = DownSampleConv(24, 24, kernel=1, strides=1, padding=0)
        unet = BasicUNet(in_channels=24, out_channels=6, features=UNET_FEATURES, dropout=0.05)
        self.blocks = nn.ModuleDict({
            "dwi-tensor": dwi_tensor_input,
            "pc-bssfp": bssfp_input,
            "bssfp": bssfp_input,
            "t1w": dwi_tensor_input,
            "unet": unet,
        })

    def forward(self, x):
        x = self.blocks[self.input_modality](x)
        return self.blocks["unet"](x)


class Discriminator(nn.Module):
    """ref:src/model.py:68-92 -- PatchGAN on cat[x, y]; ``d1`` and ``blocks`` alias one ModuleDict."""

    def __init__(self, modality):
        super().__init__()
        self.modality = modality
        d1_bssfp = DownSampleConv(30, 32, batchnorm=False)
        d1_dwi = DownSampleConv(12, 32, batchnorm=False)
        self.d1 = self.blocks = nn.ModuleDict({
            "dwi-tensor": d1_dwi,
            "pc-bssfp": d1_bssfp,
            "bssfp": d1_bssfp,
            "t1w": d1_dwi,
        })
        self.d2 = DownSampleConv(32, 64)
        self.d3 = DownSampleConv(64, 128)
        self.d4 = DownSampleConv(128, 256)
        self.d5 = DownSampleConv(256, 512)
        self.final = nn.Conv3d(512, 1, kernel_size=1)

    def forward(self, x, y):
        x = torch.cat([x, y], dim=1)
        x = self.d1[self.modality](x)
        x = self.d2(x)
        x = self.d3(x)
        x = self.d4(x)
        x = self.d5(x)
        return self.final(x)


def in_channels_of(modality: str) -> int:
    return 24 if modality in ("pc-bssfp", "bssfp") else 6


# --------------------------------------------------------------------------------------------
# training_step restatement (ref:src/model.py:259-281), perceptual term == 0
# --------------------------------------------------------------------------------------------
RECON_FACTOR = 1e2       # ref:src/model.py:147
N_RECON_TERMS = 2        # L1 and Perceptual (ref:src/model.py:138,209); Perceptual is disabled (=0)


def recon_loss(y_hat, y):
    """ref:src/model.py:201-213 with the perceptual term fixed to 0 (weights need the network,
    SURVEY.md section 2 row 8): loss_tot = (L1 + 0) / 2 * recon_factor."""
    return F.l1_loss(y_hat, y) / N_RECON_TERMS * RECON_FACTOR


def gen_loss(gen, discr, x, y):
    """ref:src/model.py:170-181."""
    y_hat = gen(x)
    logits = discr(x, y_hat)
    adv = F.binary_cross_entropy_with_logits(logits, torch.ones_like(logits))
    return adv + recon_loss(y_hat, y), y_hat


def discr_loss(gen, discr, x, y):
    """ref:src/model.py:183-193."""
    with torch.no_grad():
        y_hat = gen(x)
    y_hat = y_hat.detach()
    logits_hat = discr(x, y_hat)
    logits = discr(x, y)
    loss_hat = F.binary_cross_entropy_with_logits(logits_hat, torch.zeros_like(logits_hat))
    loss = F.binary_cross_entropy_with_logits(logits, torch.ones_like(logits))
    return (loss + loss_hat) / 2


def _set_requires_grad(module, flag):
    for p in module.parameters():
        p.requires_grad_(flag)


def gan_step(gen, discr, opt_g, opt_d, x, y):
    """One optimisation step with the ordering of ref:src/model.py:259-281: G phase (D frozen by
    toggle_optimizer), AdamW step, then D phase on a fresh G forward (G frozen), AdamW step."""
    _set_requires_grad(discr, False)
    g_loss, _ = gen_loss(gen, discr, x, y)
    g_loss.backward()
    opt_g.step()
    opt_g.zero_grad()
    _set_requires_grad(discr, True)

    _set_requires_grad(gen, False)
    d_loss = discr_loss(gen, discr, x, y)
    d_loss.backward()
    opt_d.step()
    opt_d.zero_grad()
    _set_requires_grad(gen, True)
    return g_loss.detach(), d_loss.detach()


def make_optimizers(gen, discr, lr=1e-3):
    """ref:src/model.py:359-361 (AdamW, lr 1e-3, torch defaults)."""
    return (torch.optim.AdamW(gen.parameters(), lr=lr), torch.optim.AdamW(discr.parameters(), lr=lr))
