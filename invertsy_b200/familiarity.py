"""
Familiarity evaluation of rendered views: views -> PN responses -> (PCA whitening) ->
perfect-memory familiarity -> the reference's `familiarity_par[m, k, ori]` / heat-map layout.

What is pinned to the reference and what is not:

* `pn_responses` is the reference's own code: `r = np.clip(omm_responses.mean(axis=1), 0, 1)`
  (agent/agent.py:742-744, `get_pn_responses`; the same line in `get_omm_responses`, :788).  The render kernel
  writes exactly this when asked for one-channel responses (`resp`, include/isy_b200.h), so batches never make
  the three-channel detour.
* `familiarity_indices` is the reference's index arithmetic (sim/simulation.py:3085-3090).
* The whitening and the memory are InvertPy classes (`invertpy.brain.preprocessing.Whitening` with
  `w_method=pca`, `invertpy.brain.memory.PerfectMemory`), which the reference imports (agent/agent.py:30-40,
  :613-616) but does not vendor: **parity unpinned** there.  What is implemented is stated here in full:
    - PCA whitening fitted on S calibration samples x (S, N): m = mean(x, 0); C = (x - m)^T (x - m) / S;
      C = E diag(d) E^T with d descending; W = E[:, :K] / sqrt(d[:K] + eps); y = (x - m) W.
    - Perfect memory over the M stored route vectors: familiarity(q) = 1 - min_k RMSE(q, m_k),
      RMSE(q, m) = sqrt(mean_i (q_i - m_i)^2).  `reference_familiarity` states it in NumPy float64 (tests).

The memory runs in the CUDA library (`isy_memory_create` / `isy_familiarity`, csrc/isy_fam.cuh): a hand-written
TF32 tcgen05 kernel (TMA-staged operands, accumulators in TMEM, running top-4 per query in the epilogue) filters
the M stored vectors down to four candidates per query, an fp32 kernel re-checks those exactly.  The Q x M
distance matrix never exists in HBM.  On several GPUs every rank evaluates the views it rendered and only the
familiarity rows travel (`distributed.gather_views`, NCCL over NVLink) -- the one collective of the design.
"""
import ctypes

import numpy as np
import torch

from . import _lib


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


def reference_familiarity(queries, stored):
    """NumPy float64 statement of the perfect memory: (familiarity (Q,), nearest index (Q,)); blocks of queries."""
    q = np.asarray(queries, dtype=np.float64)
    m = np.asarray(stored, dtype=np.float64)
    fam, idx = np.empty(q.shape[0]), np.empty(q.shape[0], dtype=np.int64)
    for a in range(0, q.shape[0], 256):
        d2 = ((q[a:a + 256, None, :] - m[None, :, :]) ** 2).mean(axis=2)
        idx[a:a + 256] = d2.argmin(axis=1)
        fam[a:a + 256] = 1.0 - np.sqrt(d2.min(axis=1))
    return fam, idx


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
    """
    Stores every training vector; familiarity(q) = 1 - min over the stored vectors of RMSE(q, m).
    The stored vectors live in an `isy_memory` handle on `device`; queries that are CUDA tensors are evaluated in
    place (no copy when they are contiguous float32 with a row length that is a multiple of 4), anything else goes
    through the host-buffer entry point.  No CPU fallback.
    """

    def __init__(self, device=0):
        self.device = int(device)
        self.database = None          # (M, K) float32 NumPy copy of what is stored
        self._h = None
        self._ws = None

    def store(self, vectors):
        v = torch.as_tensor(vectors).detach().to("cpu", torch.float32).numpy()
        v = np.ascontiguousarray(v.reshape(-1, v.shape[-1]))
        self.database = v if self.database is None else np.concatenate([self.database, v], axis=0)
        self._release()
        return self

    def _handle(self):
        assert self.database is not None and self.database.shape[0] > 0, "nothing stored"
        if self._h is None:
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().isy_memory_create(_lib.ptr(self.database), self.database.shape[0],
                                                    self.database.shape[1], self.device, ctypes.byref(h)))
            self._h = h
        return self._h

    def _release(self):
        if self._h is not None:
            try:
                _lib.lib().isy_memory_destroy(self._h)
            except Exception:
                pass
            self._h = None

    def __del__(self):
        self._release()

    def unresolved(self):
        """queries of the last device call whose four candidates all lay inside the filter's error margin (csrc/isy_fam.cuh)"""
        q, Q = self._inflight
        n = ctypes.c_int64(0)
        stream = torch.cuda.current_stream(q.device).cuda_stream
        _lib.check(_lib.lib().isy_familiarity_unresolved(self._handle(), Q, self._ws.data_ptr(), ctypes.c_void_p(stream),
                                                         ctypes.byref(n)))
        return int(n.value)

    def unique_rows(self):
        """indices of the stored vectors the device keeps (bit-identical copies are stored once, lowest index first)"""
        _, first = np.unique(self.database, axis=0, return_index=True)
        return np.sort(first)

    def candidates(self):
        """diagnostics: (candidates (Q, 4) int32 -- indices into the UNIQUE stored vectors, `unique_rows()` --, filter
        scores (Q, 4) float32) of the last device call"""
        q, Q = self._inflight
        cand, score = np.empty((Q, 4), dtype=np.int32), np.empty((Q, 4), dtype=np.float32)
        stream = torch.cuda.current_stream(q.device).cuda_stream
        _lib.check(_lib.lib().isy_familiarity_candidates(self._handle(), Q, self._ws.data_ptr(), ctypes.c_void_p(stream),
                                                         _lib.ptr(cand), _lib.ptr(score)))
        return cand, score

    def familiarity(self, queries, nearest=False, out=None):
        """queries (Q, K) -> familiarity (Q,) float32 [, nearest stored index (Q,) int32]; CUDA in, CUDA out."""
        h = self._handle()
        L = _lib.lib()
        K = self.database.shape[1]
        if torch.is_tensor(queries) and queries.is_cuda:
            q = queries.reshape(-1, queries.shape[-1])
            if q.shape[1] != K:
                raise ValueError("queries have %d values, the stored vectors %d" % (q.shape[1], K))
            if q.dtype != torch.float32 or q.stride(1) != 1 or q.stride(0) % 4 or q.data_ptr() % 16 or \
                    q.device.index != self.device:
                ld = (K + 3) // 4 * 4
                buf = torch.zeros((q.shape[0], ld), dtype=torch.float32, device="cuda:%d" % self.device)
                buf[:, :K] = q
                q = buf[:, :K]
            Q = q.shape[0]
            dev = q.device
            fam = out if out is not None else torch.empty(Q, dtype=torch.float32, device=dev)
            idx = torch.empty(Q, dtype=torch.int32, device=dev) if nearest else None
            need = int(L.isy_familiarity_workspace_bytes(h, Q))
            if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
                self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(L.isy_familiarity(h, q.data_ptr(), Q, q.stride(0), fam.data_ptr(),
                                         None if idx is None else idx.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                         ctypes.c_void_p(stream)))
            self._inflight = (q, Q)
            return (fam, idx) if nearest else fam
        q = np.ascontiguousarray(torch.as_tensor(queries).detach().cpu().numpy(), dtype=np.float32).reshape(-1, K)
        fam = np.empty(q.shape[0], dtype=np.float32)
        idx = np.empty(q.shape[0], dtype=np.int32)
        _lib.check(L.isy_familiarity_host(h, _lib.ptr(q), q.shape[0], _lib.ptr(fam), _lib.ptr(idx)))
        fam, idx = torch.from_numpy(fam), torch.from_numpy(idx)
        return (fam, idx) if nearest else fam


def _hash32(x):
    """the "lowbias32" integer hash of csrc/isy_willshaw.cuh, vectorised (uint32 arithmetic)"""
    x = np.asarray(x, dtype=np.uint32).copy()
    x ^= x >> np.uint32(16)
    x *= np.uint32(0x7feb352d)
    x ^= x >> np.uint32(15)
    x *= np.uint32(0x846ca68b)
    x ^= x >> np.uint32(16)
    return x


def willshaw_wiring(nb_pn, nb_kc, fan_in, seed):
    """(nb_kc, fan_in) PN index of every KC input: umulhi(hash32(j * fan_in + i + seed), nb_pn)"""
    j = np.arange(nb_kc, dtype=np.uint64)[:, None] * np.uint64(fan_in) + np.arange(fan_in, dtype=np.uint64)[None, :] + np.uint64(seed)
    h = _hash32((j & np.uint64(0xffffffff)).astype(np.uint32)).astype(np.uint64)
    return ((h * np.uint64(nb_pn)) >> np.uint64(32)).astype(np.int64)


def reference_willshaw(stored, queries, nb_kc, fan_in=8, sparseness=0.01, seed=2021):
    """
    NumPy statement of the Willshaw memory (see `WillshawMemory`): returns the familiarity of every query after all
    `stored` vectors have been learnt.  KC inputs are summed in float32 in input order; the k = round(sparseness * nb_kc)
    largest fire, ties going to the lower cell index.
    """
    stored, queries = np.asarray(stored, dtype=np.float32), np.asarray(queries, dtype=np.float32)
    wiring = willshaw_wiring(stored.shape[1], nb_kc, fan_in, seed)
    k = int(max(1, min(nb_kc, round(sparseness * nb_kc))))

    def firing(x):
        s = np.zeros((x.shape[0], nb_kc), dtype=np.float32)
        for i in range(fan_in):
            s = s + x[:, wiring[:, i]]
        order = np.argsort(-s, axis=1, kind="stable")[:, :k]          # stable: ties keep the lower index first
        act = np.zeros((x.shape[0], nb_kc), dtype=bool)
        np.put_along_axis(act, order, True, axis=1)
        return act

    synapse = ~firing(stored).any(axis=0)                             # cleared by every stored view's firing cells
    out = np.empty(queries.shape[0])
    for a in range(0, queries.shape[0], 64):
        out[a:a + 64] = 1.0 - (firing(queries[a:a + 64]) & synapse[None, :]).sum(axis=1) / k
    return out


class WillshawMemory:
    """
    Willshaw-type mushroom body (the default memory of the reference's `VisualNavigationAgent`, agent.py:560-575:
    "#KC equal to 40 x #PN, sparseness is 1 %"; InvertPy's `WillshawNetwork` is not vendored: **parity unpinned**).
    Definition (csrc/isy_willshaw.cuh, `reference_willshaw`):
      * every Kenyon cell sums `fan_in` projection neurons chosen by a fixed integer hash of (cell, input, seed);
      * the k = round(sparseness * nb_kc) cells with the largest input fire (ties: the lower cell index);
      * KC -> MBON synapses start at 1; `store` sets the synapses of the firing cells of every stored view to 0;
      * familiarity(view) = 1 - (firing cells whose synapse is still 1) / k  -- 1 for a stored view, ~0 for a novel one.
    Runs on the device behind the C ABI (`isy_willshaw_*`); vectors are float32 CUDA tensors.  No CPU fallback.
    """

    def __init__(self, nb_pn, nb_kc=None, fan_in=8, sparseness=0.01, seed=2021, device=0):
        self.nb_pn, self.nb_kc = int(nb_pn), int(40 * nb_pn if nb_kc is None else nb_kc)
        self.fan_in, self.sparseness, self.seed, self.device = int(fan_in), float(sparseness), int(seed), int(device)
        self._h = None

    def _handle(self):
        if self._h is None:
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().isy_willshaw_create(self.nb_pn, self.nb_kc, self.fan_in, self.sparseness, self.seed,
                                                      self.device, ctypes.byref(h)))
            self._h = h
        return self._h

    def __del__(self):
        if self._h is not None:
            try:
                _lib.lib().isy_willshaw_destroy(self._h)
            except Exception:
                pass
            self._h = None

    @property
    def nb_active(self):
        return int(_lib.lib().isy_willshaw_active(self._handle()))

    def reset(self):
        _lib.check(_lib.lib().isy_willshaw_reset(self._handle()))
        return self

    def _dev(self, v):
        v = torch.as_tensor(v)
        v = v.reshape(-1, v.shape[-1]).to(device="cuda:%d" % self.device, dtype=torch.float32).contiguous()
        if v.shape[1] != self.nb_pn:
            raise ValueError("vectors have %d values, the memory %d projection neurons" % (v.shape[1], self.nb_pn))
        return v

    def store(self, vectors):
        v = self._dev(vectors)
        stream = torch.cuda.current_stream(v.device).cuda_stream
        _lib.check(_lib.lib().isy_willshaw_train(self._handle(), v.data_ptr(), v.shape[0], v.stride(0), ctypes.c_void_p(stream)))
        self._inflight = v
        return self

    def familiarity(self, queries, out=None):
        q = self._dev(queries)
        fam = out if out is not None else torch.empty(q.shape[0], dtype=torch.float32, device=q.device)
        stream = torch.cuda.current_stream(q.device).cuda_stream
        _lib.check(_lib.lib().isy_willshaw_familiarity(self._handle(), q.data_ptr(), q.shape[0], q.stride(0), fam.data_ptr(),
                                                       ctypes.c_void_p(stream)))
        self._inflight = q
        return fam


def evaluate(views, route_views, whitening=None, memory=None, device=0):
    """
    views (V, N[, C]) and route_views (L, N[, C]): ommatidia responses (or PN vectors when 2-D) of the second phase and
    of the training route, e.g. `stats["ommatidia"]` / `stats["ommatidia_out"]` of `datasets.collect_familiarity_dataset`.
    Trains a perfect memory on the (whitened) route views and returns the familiarity of every view, shape (V,).
    """
    pv = pn_responses(views) if torch.as_tensor(views).ndim == 3 else torch.as_tensor(views)
    pr = pn_responses(route_views) if torch.as_tensor(route_views).ndim == 3 else torch.as_tensor(route_views)
    if whitening is not None:
        pv, pr = whitening(pv), whitening(pr)
    memory = PerfectMemory(device=pv.device.index if pv.is_cuda else device) if memory is None else memory
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


def evaluate_sharded(local_views, route_views, nb_positions, views_per_position, whitening=None, dst=None, group=None,
                     memory=None):
    """
    Multi-GPU: every rank passes the views of ITS block of positions (`distributed.shard`) and the whole training
    route; the familiarities are gathered in position order (all ranks, or rank `dst` only).
    """
    from . import distributed
    fam = evaluate(local_views, route_views, whitening=whitening, memory=memory)
    return distributed.gather_views({"familiarity": fam}, nb_positions, views_per_position, dst=dst, group=group)["familiarity"]


class HeatMap(object):
    """
    The whole job of the reference's familiarity heat-map / parallel-lines scripts for one set of views, on however many
    GPUs the process group has (sim/simulation.py:2675-2722 renders the views, :3061-3092 evaluates them):

        hm = HeatMap(renderer, memory, sensor, nb_positions=P, nb_headings=H)
        fam = hm(pos, yaw)            # host poses of THIS rank's positions in, the whole map out (pinned host tensor)

    Positions are dealt round-robin over the ranks (`views.shard_positions_interleaved`); each rank renders its share,
    turns the PN responses the render kernel writes into familiarities with the perfect memory, one NCCL all_gather +
    an on-device transpose give every rank the map in view order p * H + h.  All buffers are allocated once.
    """

    def __init__(self, renderer, memory, sensor, nb_positions, nb_headings, outputs=("resp",), group=None):
        from . import distributed
        self.r, self.memory, self.sensor, self.group = renderer, memory, sensor, group
        self.P, self.H = int(nb_positions), int(nb_headings)
        self.rank, self.world = distributed.rank_and_world(group)
        self.per = (self.P + self.world - 1) // self.world          # positions of the largest shard
        self.outputs = tuple(outputs) if "resp" in outputs else tuple(outputs) + ("resp",)
        dev = "cuda:%d" % renderer.device
        self.views = {}                                             # the rank's rendered outputs (device, re-used)
        self.local = torch.zeros(self.per * self.H, dtype=torch.float32, device=dev)
        self.scratch = torch.empty(self.world * self.per * self.H, dtype=torch.float32, device=dev) if self.world > 1 else None
        self.full = torch.empty(self.world * self.per * self.H, dtype=torch.float32, device=dev) if self.world > 1 else None
        self.host = torch.empty(self.P * self.H, dtype=torch.float32, pin_memory=True)

    def positions(self):
        """indices of the positions this rank takes"""
        from .views import shard_positions_interleaved
        return shard_positions_interleaved(self.P, self.rank, self.world)

    def on_device(self, pos_dev, yaw_dev):
        """this rank's poses (device tensors) -> the whole map (device), asynchronous on the current stream"""
        from . import distributed
        n = pos_dev.shape[0] * self.H
        self.r.render(pos_dev, yaw_dev, outputs=self.outputs, out=self.views, sensor=self.sensor)
        self.memory.familiarity(self.views["resp"], out=self.local[:n])
        return distributed.all_gather_interleaved(self.local, self.P, self.H, out=self.full, scratch=self.scratch,
                                                  group=self.group)

    def __call__(self, pos_host, yaw_host):
        dev = self.local.device
        pos_d = torch.as_tensor(pos_host).to(dev, non_blocking=True)
        yaw_d = torch.as_tensor(yaw_host).to(dev, non_blocking=True)
        fam = self.on_device(pos_d, yaw_d)
        self.host.copy_(fam, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return self.host
