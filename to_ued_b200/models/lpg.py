"""LPG network (reference models/lpg.py:39-96) as a flat parameter vector + CUDA forward/backward.

Layout of the flat vector (to_ued_b200/csrc/lpg_common.cuh):
    Wh[H,3H] Wi[X,3H] bi[3H] bhn[H] w_pi[H] W_y[H,Y] b_y[Y] | e_w0[Y,E] e_b0[E] e_w1[E] e_b1[1] | b_pi[1]
``named_params`` exposes it under the reference's flax names (MLP_0/Dense_0, LPGGRU_0/GRUCell_0/ir ...).

Deviation Q5: hidden size = lpg_gru_width (the reference passes ``features=len(gru_state)``, i.e. the
batch size, to GRUCell — evidently unintended).  Q6 is reproduced (raw step / lifetime inputs)."""
from __future__ import annotations

import numpy as np
import torch

from ..util import prng


class LPG:
    def __init__(self, embedding_net_width: int = 16, gru_width: int = 256, target_width: int = 8,
                 lifetime_conditioning: bool = False):
        if (embedding_net_width, gru_width, target_width) != (16, 256, 8):
            raise NotImplementedError(
                "the sm_100a kernels are specialised for embedding 16 / GRU 256 / target 8 (reference defaults)")
        self.embedding_net_width, self.gru_width, self.target_width = embedding_net_width, gru_width, target_width
        self.lifetime_conditioning = lifetime_conditioning
        E, H, Y = embedding_net_width, gru_width, target_width
        X = 7 if lifetime_conditioning else 5
        self.X = X
        self.shapes = [("Wh", (H, 3 * H)), ("Wi", (X, 3 * H)), ("bi", (3 * H,)), ("bhn", (H,)), ("w_pi", (H,)),
                       ("W_y", (H, Y)), ("b_y", (Y,)), ("e_w0", (Y, E)), ("e_b0", (E,)), ("e_w1", (E,)),
                       ("e_b1", (1,)), ("b_pi", (1,))]
        self.offsets, off = {}, 0
        for name, shp in self.shapes:
            n = int(np.prod(shp))
            self.offsets[name] = (off, n, shp)
            off += n
        self.size = off

    def views(self, flat):
        return {name: flat[o:o + n].view(shp) for name, (o, n, shp) in self.offsets.items()}

    def named_params(self, flat):
        """flax-style nested dict (models/lpg.py): kernels are [in, out]."""
        v = self.views(flat)
        H = self.gru_width
        cell = {}
        for gi, g in enumerate("rzn"):
            cell["i" + g] = {"kernel": v["Wi"][:, gi * H:(gi + 1) * H], "bias": v["bi"][gi * H:(gi + 1) * H]}
            cell["h" + g] = {"kernel": v["Wh"][:, gi * H:(gi + 1) * H]}
        cell["hn"]["bias"] = v["bhn"]
        return {"MLP_0": {"Dense_0": {"kernel": v["e_w0"], "bias": v["e_b0"]},
                          "Dense_1": {"kernel": v["e_w1"].view(-1, 1), "bias": v["e_b1"]}},
                "LPGGRU_0": {"GRUCell_0": cell},
                "Dense_0": {"kernel": v["w_pi"].view(-1, 1), "bias": v["b_pi"]},
                "Dense_1": {"kernel": v["W_y"], "bias": v["b_y"]}}

    def get_init_vector(self):
        """models/lpg.py:87-96 (shapes only; kept for API parity)."""
        Y = self.target_width
        return (np.ones([1, 1]), np.ones([1, 1]), np.ones([1, 1]), np.ones([1, 1, Y]), np.ones([1, 1, Y]), 1.0, 1.0)

    def init(self, rng, device="cuda") -> torch.Tensor:
        """flax default initialisers: lecun-normal Dense / input kernels, orthogonal recurrent
        kernels, zero biases.  Drawn on the host from the threefry key (init is an input of the hot
        path, not part of it; the exact jax draw is not reproduced)."""
        from .agent import lecun_normal
        keys = prng.split(np.asarray(rng, np.uint32), 8)
        H, Y, E, X = self.gru_width, self.target_width, self.embedding_net_width, self.X
        parts = {
            "e_w0": lecun_normal(keys[0], (Y, E), Y), "e_b0": np.zeros(E, np.float32),
            "e_w1": lecun_normal(keys[1], (E,), E), "e_b1": np.zeros(1, np.float32),
            "Wi": lecun_normal(keys[2], (X, 3 * H), X), "bi": np.zeros(3 * H, np.float32),
            "bhn": np.zeros(H, np.float32),
            "w_pi": lecun_normal(keys[3], (H,), H), "b_pi": np.zeros(1, np.float32),
            "W_y": lecun_normal(keys[4], (H, Y), H), "b_y": np.zeros(Y, np.float32),
        }
        orth = []
        for g in range(3):
            a = _normal(keys[5 + g], (H, H)).astype(np.float64)
            q, r = np.linalg.qr(a)
            orth.append((q * np.sign(np.diag(r))[None, :]).astype(np.float32))
        parts["Wh"] = np.concatenate(orth, axis=1)
        flat = np.concatenate([parts[name].reshape(-1) for name, _ in self.shapes]).astype(np.float32)
        t = torch.from_numpy(flat)
        return t.to(device) if device != "cpu" else t


def _normal(key, shape):
    from scipy.special import erfinv
    u = prng.uniform(key, shape, -1.0 + 2.0 ** -23, 1.0)
    return (np.sqrt(2.0) * erfinv(u.astype(np.float64))).astype(np.float32)
