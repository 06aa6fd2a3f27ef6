"""Training-loop pieces of the reference driver that touch the hot path (ref GNS/main.py:243-309),
without wandb: Adam on the flat parameter buffer (one fused kernel instead of ~150 small ones),
batch-mean loss, early stop after three non-improving epochs, reference checkpoint naming."""
from __future__ import annotations

import os

import torch

from . import _lib
from .parallel import allreduce_gradients, flat_gradient


class FlatAdam:
    """torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8) semantics (the reference's optimizer,
    ref GNS/main.py:243) applied by ``gns_adam_step`` to the model's flat parameter buffer."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        self.model, self.lr, self.betas, self.eps, self.step_count = model, lr, betas, eps, 0
        self.exp_avg = self.exp_avg_sq = None

    def zero_grad(self):
        self.model.zero_grad(set_to_none=True)

    @torch.no_grad()
    def step(self):
        lib = _lib.load_library()
        flat = self.model.flat_parameters()
        params = list(self.model.parameters())
        if self.exp_avg is None or self.exp_avg.device != flat.device:
            self.exp_avg, self.exp_avg_sq = torch.zeros_like(flat), torch.zeros_like(flat)
        # The alias of all parameter gradients is only usable when it covers EVERY parameter in order: with frozen
        # parameters (grad None) it is shorter or starts at an offset, and the kernel would pair gradients with
        # the wrong parameters.
        grad = flat_gradient(params)
        frozen = None
        if grad is None or grad.numel() != flat.numel():
            grad = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(torch.float32)
                              for p in params])
            # torch.optim.Adam skips parameters without a gradient (no moment update, no step): restore them after
            off, frozen = 0, []
            for p in params:
                if p.grad is None:
                    sl = slice(off, off + p.numel())
                    frozen.append((sl, flat[sl].clone(), self.exp_avg[sl].clone(), self.exp_avg_sq[sl].clone()))
                off += p.numel()
        self.step_count += 1
        rc = lib.gns_adam_step(flat.data_ptr(), grad.contiguous().data_ptr(), self.exp_avg.data_ptr(),
                               self.exp_avg_sq.data_ptr(), flat.numel(), self.lr, self.betas[0], self.betas[1],
                               self.eps, self.step_count, torch.cuda.current_stream(flat.device).cuda_stream)
        _lib.check(rc, "gns_adam_step")
        for sl, pv, m1, m2 in frozen or ():
            flat[sl].copy_(pv); self.exp_avg[sl].copy_(m1); self.exp_avg_sq[sl].copy_(m2)


def checkpoint_name(case_nr, model, optimizer_name="Adam"):
    """File name used by the reference (ref GNS/main.py:308-309)."""
    return (f"best_model_c{case_nr}_K{model.K}_L{model.latent_dim}_H{model.hidden_dim}_"
            f"{model.multiple_phis}_optim{optimizer_name}.pth")


def fit(model, buses, lines, generators, epochs=101, batch_size=128, lr=1e-3, case_nr=None, save_dir=None,
        log=print):
    """Epoch loop of ref GNS/main.py:274-309 with one batched call per batch instead of the
    per-sample Python loop.  Returns the list of epoch final losses."""
    opt = FlatAdam(model, lr=lr)
    n = buses.shape[0] - buses.shape[0] % batch_size
    best, worse, history = float("inf"), 0, []
    for epoch in range(epochs):
        finals = []
        for a in range(0, n, batch_size):
            opt.zero_grad()
            _, _, total, last = model(buses[a:a + batch_size], lines[a:a + batch_size], generators[a:a + batch_size])
            total.mean().backward()                               # ref :284-288
            allreduce_gradients(model.parameters(), average=True)  # no-op without a process group
            opt.step()
            finals.append(last.detach().mean())
        final = float(torch.stack(finals).mean())
        history.append(final)
        if final >= best:                                         # early stop, ref :296-300
            worse += 1
            if worse > 2:
                log("Loss is increasing")
                break
        else:
            best, worse = final, 0
        log(f"Epoch: {epoch}, Final Loss: {final}, best loss: {best}")
        if save_dir is not None and case_nr is not None:
            torch.save(model.state_dict(), os.path.join(save_dir, checkpoint_name(case_nr, model)))
    return history
