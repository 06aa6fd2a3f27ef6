"""Generate golden vectors from the LIVE reference implementation.

Run in the build container only (needs /root/reference; the GPU box has none):

    python tests/golden/make_golden.py

It imports /root/reference/GNS/main.py *unmodified* (cwd = its directory, because the
reference opens ``../data/...`` relative paths, GNS/utils.py:18) behind a one-function
``torch_scatter`` stand-in (the package is not installed; the reference only uses
``scatter_add(src, index, out=..., dim=0)``, GNS/main.py:2) and writes
``tests/golden/*.npz``.  Nothing of the reference's source is copied.

Each file holds: the packed inputs of case14 samples 1..32 produced by the
reference's own ``prepare_grid``; the reference module's seed-0 ``state_dict``;
its per-sample outputs (v, theta, total_loss, last_loss); and the gradients of
``mean(total_loss)`` (the training reduction, GNS/main.py:284-288).
"""
import os
import sys
import tempfile
import types
import warnings

import numpy as np
import torch

REF = os.environ.get("GNS_REF_DIR", "/root/reference/GNS")
OUT = os.path.dirname(os.path.abspath(__file__))
N_SAMPLES = 32


def _install_scatter_shim():
    mod = types.ModuleType("torch_scatter")

    def scatter_add(src, index, dim=0, out=None):
        idx = index
        if src.dim() > 1:
            idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
        return out.scatter_add_(dim, idx, src)

    mod.scatter_add = scatter_add
    sys.modules["torch_scatter"] = mod


def main():
    warnings.filterwarnings("ignore")
    os.environ["WANDB_MODE"] = "disabled"
    _install_scatter_shim()
    os.chdir(REF)
    sys.path.insert(0, REF)
    import main as ref  # noqa: E402  (the reference)
    from utils import get_BLG, prepare_grid  # noqa: E402

    B, L, G = get_BLG()
    grids = [prepare_grid(14, i) for i in range(1, N_SAMPLES + 1)]
    buses = torch.stack([g[0] for g in grids])
    lines = torch.stack([g[1] for g in grids])
    gens = torch.stack([g[2] for g in grids])

    # base (un-perturbed) IEEE-14 table, used by the synthetic input generator
    import pickle
    base = pickle.load(open(os.path.join(REF, "../data/case14/augmented_case14_0.pkl"), "rb"))
    np.savez_compressed(os.path.join(OUT, "case14_base.npz"),
                        baseMVA=np.float64(base["baseMVA"]),
                        bus=np.asarray(base["bus"], dtype=np.float64),
                        branch=np.asarray(base["branch"], dtype=np.float64),
                        gen=np.asarray(base["gen"], dtype=np.float64))

    # raw float64 tables of samples 1..4 + their packed form: pins the packing transform
    raw = [pickle.load(open(os.path.join(REF, f"../data/case14/augmented_case14_{i}.pkl"), "rb"))
           for i in range(1, 5)]
    np.savez_compressed(os.path.join(OUT, "case14_raw_1_4.npz"),
                        baseMVA=np.float64(raw[0]["baseMVA"]),
                        bus=np.stack([np.asarray(r["bus"], dtype=np.float64) for r in raw]),
                        branch=np.stack([np.asarray(r["branch"], dtype=np.float64) for r in raw]),
                        gen=np.stack([np.asarray(r["gen"], dtype=np.float64) for r in raw]),
                        buses=buses[:4].numpy(), lines=lines[:4].numpy(), gens=gens[:4].numpy())

    configs = {
        # name: (latent, hidden, K, gamma, multiple_phi, Pd scale)
        "k4_l20_multi": (20, 10, 4, 0.9, True, 1.0),
        "k4_l20_single": (20, 10, 4, 0.9, False, 1.0),
        "k8_l64_multi": (64, 10, 8, 0.9, True, 1.0),
        "k4_l20_multi_lowload": (20, 10, 4, 0.9, True, 0.3),   # forces the `if` arms of both lambda branches
        "k3_l10_single_g08": (10, 10, 3, 0.8, False, 1.0),
    }
    for name, (lat, hid, K, gamma, multi, pd_scale) in configs.items():
        torch.manual_seed(0)
        model = ref.GNS(latent_dim=lat, hidden_dim=hid, K=K, gamma=gamma, multiple_phi=multi)
        b = buses.clone()
        b[:, :, 2] *= pd_scale
        vs, ths, tots, lasts = [], [], [], []
        for i in range(N_SAMPLES):
            v, th, tot, last = model(b[i], lines[i], gens[i], B, L, G)
            vs.append(v.detach()), ths.append(th.detach()), tots.append(tot), lasts.append(last.detach())
        torch.stack(tots).mean().backward()
        out = {
            "buses": b.numpy(), "lines": lines.numpy(), "gens": gens.numpy(),
            "v": torch.stack(vs).numpy(), "theta": torch.stack(ths).numpy(),
            "total_loss": torch.stack(tots).detach().numpy(), "last_loss": torch.stack(lasts).numpy(),
            "hyper": np.array([lat, hid, K, int(multi)], dtype=np.int64), "gamma": np.float64(gamma),
        }
        none_grads = []
        for n, p in model.named_parameters():
            out["param/" + n] = p.detach().numpy()
            if p.grad is None:
                none_grads.append(n)
                out["grad/" + n] = np.zeros(tuple(p.shape), dtype=np.float32)
            else:
                out["grad/" + n] = p.grad.numpy()
        out["none_grads"] = np.array(none_grads)
        np.savez_compressed(os.path.join(OUT, f"ref_case14_{name}.npz"), **out)
        print(name, "total[0]=%.6f last[0]=%.6f none_grads=%d" %
              (out["total_loss"][0], out["last_loss"][0], len(none_grads)))


if __name__ == "__main__":
    main()
