"""
Familiarity evaluation of rendered views: views -> PN responses -> PCA whitening ->
perfect-memory familiarity -> the reference's `familiarity_par[m, k, ori]` / heat-map layout.

What is pinned to the reference and what is not:

* `pn_responses` is the reference's own code: `r = np.clip(omm_responses.mean(axis=1), 0, 1)`
  (agent/agent.py:742-744, `get_pn_responses`; the same line in `get_omm_responses`, :788).
* `familiarity_indices` is the reference's index arithmetic (sim/simulation.py:3085-3090).
* The whitening and the memory are InvertPy classes (`invertpy.brain.preprocessing.Whitening` with
  `w_method=pca`, `invertpy.brain.memory.PerfectMemory`), which the reference imports (agent/agent.py:30-40,
  :613-616) but does not vendor: **parity unpinned** there.  What is implemented is stated here in full:
    - PCA whitening fitted on S calibration samples x (S, N): m = mean(x, 0); C = (x - m)^T (x - m) / S;
      C = E diag(d) E^T with d descending; W = E[:, :K] / sqrt(d[:K] + eps); y = (x - m) W.
    - Perfect memory over the M stored route vectors: familiarity(q) = 1 - min_k RMSE(q, m_k),
      RMSE(q, m) = sqrt(mean_i (q_i - m_i)^2).

Everything is torch on whatever device the tensors live on.  The M x Q distance matrix is a plain GEMM
(|q - m|^2 = |q|^2 + |m|^2 - 2 q.m, cuBLAS through torch.matmul), evaluated in blocks of queries so that it
never has to exist as a whole, and used only to pick the nearest stored vector; on several GPUs every rank evaluates the views it rendered and only the
familiarity rows travel (`distributed.gather_views`, NCCL over NVLink) -- the one collective of the design.
"""
import numpy as np
import torch


def pn_responses(omm_responses):
    """(…, N, C) ommatidia responses -> (…, N) PN input: clip(mean over channels, 0, 1)  (agent.py:742-744)."""
    r = torch.as_tensor(omm_responses)
    return r.mean(dim=-1).clamp(0.0, 1.0)


def familiarity_indices(nb_views, route_length, nb_orientations):
    """
    The reference's placement of the j-th view of the second phase in `familiarity_par[m, k, ori]`
    (simulation.py:3085-3090): m = (j // nb_orientations) % route_length, k = j // (route_length *
    nb_orientations), ori = j % nb_orientations.
    """
    j = np.arange(nb_views)
    return (j // nb_orientations) % route_length, j // (route_length * nb_orientations), j % nb_orientations


class Whitening:
    """PCA whitening (see the module docstring for the definition); `fit` = the reference's calibration step
    (`agent.calibrate` -> `Whitening.reset(samples)`, agent/agent.py:706-712)."""

    def __init__(self, nb_output=None, eps=1e-5, dtype=torch.float32):
        self.nb_output, self.eps, self.dtype = nb_output, eps, dtype
        self.mean = None
        self.w = None

    def fit(self, samples):
        x = torch.as_tensor(samples).to(torch.float64)            # the decomposition in fp64, the map in `dtype`
        m = x.mean(dim=0)
        xc = x - m
        cov = xc.T @ xc / x.shape[0]
        d, e = torch.linalg.eigh(cov)
        order = torch.argsort(d, descending=True)
        k = self.nb_output or x.shape[1]
        d, e = d[order][:k], e[:, order][:, :k]
        self.mean = m.to(self.dtype)
        self.w = (e / torch.sqrt(d.clamp_min(0) + self.eps)).to(self.dtype)
        return self

    def __call__(self, x):
        assert self.w is not None, "fit() the whitening on calibration samples first"
        x = torch.as_tensor(x).to(self.dtype)
        return (x - self.mean.to(x.device)) @ self.w.to(x.device)


class PerfectMemory:
    """Stores every training vector; familiarity(q) = 1 - min over the stored vectors of RMSE(q, m)."""

    def __init__(self, block=65536):
        self.block = block
        self.database = None

    def store(self, vectors):
        v = torch.as_tensor(vectors)
        v = v.reshape(-1, v.shape[-1])
        self.database = v if self.database is None else torch.cat([self.database, v.to(self.database.device)], dim=0)
        return self

    def familiarity(self, queries):
        assert self.database is not None and self.database.shape[0] > 0, "nothing stored"
        q = torch.as_tensor(queries)
        q = q.reshape(-1, q.shape[-1])
        db = self.database.to(device=q.device, dtype=q.dtype)
        m2 = (db * db).sum(dim=1)
        out = torch.empty(q.shape[0], dtype=q.dtype, device=q.device)
        for a in range(0, q.shape[0], self.block):
            qb = q[a:a + self.block]
            d2 = (qb * qb).sum(dim=1, keepdim=True) + m2[None, :] - 2.0 * (qb @ db.T)    # (block, M) GEMM
            # the GEMM form only picks the nearest stored vector; its distance is then formed from the difference
            # itself (no cancellation: a stored view is exactly familiar with itself)
            diff = qb - db[d2.argmin(dim=1)]
            out[a:a + self.block] = 1.0 - torch.sqrt((diff * diff).sum(dim=1) / q.shape[1])
        return out


def evaluate(views, route_views, whitening=None, memory=None):
    """
    views (V, N[, C]) and route_views (L, N[, C]): ommatidia responses (or PN vectors when 2-D) of the second phase and
    of the training route, e.g. `stats["ommatidia"]` / `stats["ommatidia_out"]` of `datasets.collect_familiarity_dataset`.
    Trains a perfect memory on the (whitened) route views and returns the familiarity of every view, shape (V,).
    """
    pv = pn_responses(views) if torch.as_tensor(views).ndim == 3 else torch.as_tensor(views)
    pr = pn_responses(route_views) if torch.as_tensor(route_views).ndim == 3 else torch.as_tensor(route_views)
    if whitening is not None:
        pv, pr = whitening(pv), whitening(pr)
    memory = PerfectMemory() if memory is None else memory
    memory.store(pr)
    return memory.familiarity(pv)


def familiarity_par(fam, route_length, nb_parallel, nb_orientations):
    """(V,) familiarities of a parallel-lines dataset -> the reference's `familiarity_par[m, k, ori]` array."""
    fam = torch.as_tensor(fam).detach().cpu().numpy()
    m, k, ori = familiarity_indices(fam.shape[0], route_length, nb_orientations)
    out = np.zeros((route_length, nb_parallel, nb_orientations), dtype=fam.dtype)
    out[m, k, ori] = fam
    return out


def familiarity_map(fam, nb_rows, nb_cols, nb_orientations):
    """(V,) familiarities of a grid dataset (row-major row, col, ori: simulation.py:2675-2681) -> (rows, cols, oris)."""
    return torch.as_tensor(fam).detach().cpu().numpy().reshape(nb_rows, nb_cols, nb_orientations)


def evaluate_sharded(local_views, route_views, nb_positions, views_per_position, whitening=None, dst=None, group=None):
    """
    Multi-GPU: every rank passes the views of ITS block of positions (`distributed.shard`) and the whole training
    route; the familiarities are gathered in position order (all ranks, or rank `dst` only).
    """
    from . import distributed
    fam = evaluate(local_views, route_views, whitening=whitening)
    return distributed.gather_views({"familiarity": fam}, nb_positions, views_per_position, dst=dst, group=group)["familiarity"]
