"""One training iteration of the reference's GAN trainer, for the "ResNet-GAN iters/s" half of
BASELINE.json's metric.

Reference                                   here
---------------------------------------------------------------------------------------------
GAN.get_*_adversarial_loss ('hinge' ...)    generator_loss / discriminator_loss      gan/gan.py:69-128
ProjectiveGAN (D(x, cls), D(G(z, cls), cls)) the calls in GanTrainer                 gan/projective_gan.py:12-26
Trainer.train_one_step                      GanTrainer.train_one_step                gan/train.py:78-99
Adam(lr, beta_1=0, beta_2=0.9), linear decay GanTrainer.__init__                      run.py:58-59,99-144

An iteration is ``training_ratio`` discriminator steps at batch B (generator forward only, in training phase, so its
moving statistics are updated as in the reference's graph: gan/gan.py:254-255) followed by ONE generator step at
``generator_batch_multiple * B``.  Data is synthetic (no datasets in this environment); the convolutions are cuDNN.
With a process group the generator's WC layers run synchronised-batch whitening and the parameter gradients are
averaged with one all-reduce per step.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F


def generator_loss(fake_logits: torch.Tensor, kind: str = 'hinge') -> torch.Tensor:
    if kind in ('hinge', 'wgan'):
        return -fake_logits.mean()
    if kind == 'ns-gan':
        return F.binary_cross_entropy_with_logits(fake_logits, torch.ones_like(fake_logits))
    if kind == 'lsgan':
        return ((fake_logits - 1) ** 2).mean()
    raise ValueError(kind)


def discriminator_loss(real_logits: torch.Tensor, fake_logits: torch.Tensor, kind: str = 'hinge') -> torch.Tensor:
    if kind == 'hinge':
        return torch.relu(1.0 - real_logits).mean() + torch.relu(1.0 + fake_logits).mean()
    if kind == 'wgan':
        return -real_logits.mean() + fake_logits.mean()
    if kind == 'ns-gan':
        return (F.binary_cross_entropy_with_logits(real_logits, torch.ones_like(real_logits)) +
                F.binary_cross_entropy_with_logits(fake_logits, torch.zeros_like(fake_logits)))
    if kind == 'lsgan':
        return ((real_logits - 1) ** 2).mean() + (fake_logits ** 2).mean()
    raise ValueError(kind)


class GanTrainer:
    def __init__(self, generator, discriminator, *, batch_size: int, noise_size: int = 128, number_of_classes: int = 10,
                 image_shape=(32, 32, 3), training_ratio: int = 5, generator_batch_multiple: int = 2, lr: float = 2e-4,
                 beta1: float = 0.0, beta2: float = 0.9, objective: str = 'hinge', total_generator_iters: Optional[float] = None,
                 process_group=None, device='cuda'):
        self.G, self.D = generator, discriminator
        self.B, self.Z, self.K, self.image_shape = batch_size, noise_size, number_of_classes, tuple(image_shape)
        self.training_ratio, self.gen_batch_mul, self.objective = training_ratio, generator_batch_multiple, objective
        self.group, self.device = process_group, torch.device(device)
        self.opt_g = torch.optim.Adam(self.G.parameters(), lr=lr, betas=(beta1, beta2), eps=1e-7)      # keras epsilon
        self.opt_d = torch.optim.Adam(self.D.parameters(), lr=lr, betas=(beta1, beta2), eps=1e-7)
        self.sched = None
        if total_generator_iters:                                          # --lr_decay_schedule linear (run.py:107-109)
            ng, nd = float(total_generator_iters), float(total_generator_iters) * training_ratio
            self.sched = (torch.optim.lr_scheduler.LambdaLR(self.opt_g, lambda it: max(0.0, 1.0 - it / ng)),
                          torch.optim.lr_scheduler.LambdaLR(self.opt_d, lambda it: max(0.0, 1.0 - it / nd)))
        self.gen = torch.Generator(device=self.device).manual_seed(1234)

    # ---- synthetic samples (gan/dataset.py next_generator_sample / next_discriminator_sample)
    def _noise(self, n):
        z = torch.randn(n, self.Z, device=self.device, generator=self.gen)
        cls = torch.randint(0, self.K, (n, 1), device=self.device, generator=self.gen, dtype=torch.int32)
        return z, cls

    def _real(self, n):
        x = torch.rand(n, *self.image_shape, device=self.device, generator=self.gen) * 2 - 1
        cls = torch.randint(0, self.K, (n, 1), device=self.device, generator=self.gen, dtype=torch.int32)
        return x, cls

    def _sync_grads(self, params):
        if self.group is None or dist.get_world_size(self.group) == 1:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in params if p.grad is not None])
        dist.all_reduce(flat, group=self.group)
        flat /= dist.get_world_size(self.group)
        off = 0
        for p in params:
            if p.grad is not None:
                n = p.grad.numel()
                p.grad.copy_(flat[off:off + n].reshape(p.grad.shape))
                off += n

    def discriminator_step(self) -> torch.Tensor:
        x, xc = self._real(self.B)
        z, zc = self._noise(self.B)
        with torch.no_grad():
            fake = self.G(z, zc)                           # training phase: the WC layers' moving statistics are updated
        loss = discriminator_loss(self.D(x, xc), self.D(fake, zc), self.objective)
        self.opt_d.zero_grad(set_to_none=True)
        loss.backward()
        self._sync_grads(list(self.D.parameters()))
        self.opt_d.step()
        if self.sched:
            self.sched[1].step()
        return loss.detach()

    def generator_step(self) -> torch.Tensor:
        z, zc = self._noise(self.B * self.gen_batch_mul)   # concatenate_generator_batches (gan/train.py:85-91)
        loss = generator_loss(self.D(self.G(z, zc), zc), self.objective)
        self.opt_g.zero_grad(set_to_none=True)
        loss.backward()
        self._sync_grads(list(self.G.parameters()))
        self.opt_g.step()
        if self.sched:
            self.sched[0].step()
        return loss.detach()

    def train_one_step(self):
        """gan/train.py:78-99."""
        self.G.train(); self.D.train()
        d = [self.discriminator_step() for _ in range(self.training_ratio)]
        g = self.generator_step()
        return torch.stack(d).mean(), g
