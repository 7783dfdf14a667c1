"""ORACLE (test infrastructure only) -- restatement of the reference's LightningModule shell with the network and
loss classes INJECTED, so that the same ``training_step`` statements can drive either the oracle's stock
``torch.nn`` classes (CPU, pinned against the reference's own run) or the sm_100a drop-in classes (GPU box).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.

Why a restatement: ``/root/reference`` does not exist on the GPU box and the reference's ``model.py`` cannot be
imported without lightning / monai / torchio / nibabel (absent from the image). Where the reference IS mounted
(the build container), ``tests/golden/make_golden_ref.py`` executes the real ``bSSFPToDWITensorModel`` and the
goldens it wrote (``bssfp_train3_*``) pin this shell: ``tests/test_reference_golden_cpu.py`` runs the shell with
the oracle classes and reproduces them; ``tests/test_reference_golden_gpu.py`` then runs the SAME shell with
``unet_bssfp_b200.Generator / Discriminator / L1Loss / BCEWithLogitsLoss`` injected.

Restated (statement by statement):
* ``bSSFPToDWITensorModel.__init__``           ref:src/model.py:141-165 (metrics / FID / perceptual network omitted)
* ``PerceptualL1Loss.forward``                 ref:src/model.py:135-138 (perceptual term == 0: its MedicalNet weights
                                               need the network; the product fixes it to 0 the same way)
* ``forward / _gen_step / _discr_step``        ref:src/model.py:167-193
* ``unpack_batch / compute_recon_loss``        ref:src/model.py:195-213
* ``training_step``                            ref:src/model.py:259-281
* ``configure_optimizers``                     ref:src/model.py:359-361
and the five trainer services ``training_step`` calls, with Lightning 2.2's documented semantics (the same stand-in
``make_golden_ref.py`` gives the real module): ``optimizers()`` returns what ``configure_optimizers`` built,
``toggle_optimizer`` sets ``requires_grad=False`` on every parameter the other optimizers own, ``untoggle_optimizer``
restores it, ``manual_backward(loss)`` is ``loss.backward()``, ``log`` records the value.
"""
import torch

DATA = "data"   # torchio.DATA


class _LightningServices(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.logged = {}
        self._optimizers = None

    def log(self, name, value, **k):
        self.logged[name] = float(value.detach()) if torch.is_tensor(value) else float(value)

    def optimizers(self):
        if self._optimizers is None:
            self._optimizers = self.configure_optimizers()
        return self._optimizers

    def toggle_optimizer(self, optimizer):
        mine = {id(p) for g in optimizer.param_groups for p in g["params"]}
        self._toggled = {}
        for opt in self.optimizers():
            for g in opt.param_groups:
                for p in g["params"]:
                    if id(p) not in mine and id(p) not in self._toggled:
                        self._toggled[id(p)] = (p, p.requires_grad)
                        p.requires_grad = False

    def untoggle_optimizer(self, optimizer):
        for p, flag in self._toggled.values():
            p.requires_grad = flag
        self._toggled = {}

    def manual_backward(self, loss):
        loss.backward()


class _PerceptualL1Loss(torch.nn.Module):
    def __init__(self, perceptual_factor, l1_class):
        super().__init__()
        self.l1 = l1_class()
        self.perceptual_factor = perceptual_factor

    def forward(self, y_hat, y):
        l1 = self.l1(y_hat, y)
        perceptual = torch.zeros((), dtype=y_hat.dtype, device=y_hat.device) * self.perceptual_factor
        return {"L1": l1, "Perceptual": perceptual}


class LightningShell(_LightningServices):
    """``classes``: any namespace with Generator, Discriminator, L1Loss, BCEWithLogitsLoss (``torch.nn`` losses for the
    oracle, the drop-in package for the product). ``optimizer_class`` as the reference: ``torch.optim.AdamW``."""

    def __init__(self, input_modality, classes, lr=1e-3, batch_size=8, perceptual_factor=1e3, recon_factor=1e2,
                 optimizer_class=torch.optim.AdamW):
        super().__init__()
        self.automatic_optimization = False
        self.input_modality = input_modality
        self.gen = classes.Generator(input_modality)
        self.discr = classes.Discriminator(input_modality)
        self.recon_criterion = _PerceptualL1Loss(perceptual_factor, classes.L1Loss)
        self.adversarial_criterion = classes.BCEWithLogitsLoss()
        self.recon_factor = recon_factor
        self.lr = lr
        self.optimizer_class = optimizer_class
        self.batch_size = batch_size

    def forward(self, x):
        return self.gen(x)

    def _gen_step(self, x, y, step_name):
        y_hat = self.gen(x)
        discr_logits = self.discr(x, y_hat)
        valid = torch.ones_like(discr_logits)
        valid = valid.type_as(x)
        adv_loss = self.adversarial_criterion(discr_logits, valid)
        recon_loss = self.compute_recon_loss(y_hat, y, step_name + "_gen")
        self.log(f"{step_name}_gen_loss_adversarial", adv_loss)
        return adv_loss + recon_loss, y_hat

    def _discr_step(self, x, y):
        y_hat = self.gen(x).detach()
        logits_hat = self.discr(x, y_hat)
        logits = self.discr(x, y)
        invalid = torch.zeros_like(logits_hat)
        invalid = invalid.type_as(x)
        valid = torch.ones_like(logits)
        valid = valid.type_as(x)
        loss_hat = self.adversarial_criterion(logits_hat, invalid)
        loss = self.adversarial_criterion(logits, valid)
        return (loss + loss_hat) / 2

    def unpack_batch(self, batch, test=False):
        x = batch[self.input_modality][DATA]
        y = batch["dwi-tensor"][DATA] if test else batch["dwi-tensor_orig"][DATA]
        return x, y

    def compute_recon_loss(self, y_hat, y, step_name):
        losses = self.recon_criterion(y_hat, y)
        loss_tot = 0
        for name, loss in losses.items():
            self.log(f"{step_name}_loss_recon_{name}", loss)
            loss_tot += loss
        loss_tot = loss_tot / len(losses.keys()) * self.recon_factor
        self.log(f"{step_name}_loss_recon", loss_tot)
        return loss_tot

    def training_step(self, batch, batch_idx):
        x, y = self.unpack_batch(batch)
        gen_optimizer, discr_optimizer = self.optimizers()

        # Train Generator
        self.toggle_optimizer(gen_optimizer)
        loss, _ = self._gen_step(x, y, "train")
        self.log("train_gen_loss", loss)
        self.manual_backward(loss)
        gen_optimizer.step()
        gen_optimizer.zero_grad()
        self.untoggle_optimizer(gen_optimizer)

        # Train Discriminator
        self.toggle_optimizer(discr_optimizer)
        loss = self._discr_step(x, y)
        self.log("train_discr_loss", loss)
        self.manual_backward(loss)
        discr_optimizer.step()
        discr_optimizer.zero_grad()
        self.untoggle_optimizer(discr_optimizer)

    def configure_optimizers(self):
        return (self.optimizer_class(self.gen.parameters(), lr=self.lr),
                self.optimizer_class(self.discr.parameters(), lr=self.lr))


class OracleClasses:
    """The stock ``torch.nn`` restatement (``oracle.model_oracle``) + torch's own losses, as the reference uses."""
    from oracle import model_oracle as _O
    Generator = _O.Generator
    Discriminator = _O.Discriminator
    L1Loss = torch.nn.L1Loss
    BCEWithLogitsLoss = torch.nn.BCEWithLogitsLoss
