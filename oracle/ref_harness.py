"""oracle/ref_harness.py -- TEST INFRASTRUCTURE (build container only).

Imports the UNMODIFIED reference from /root/reference and drives it with an injected,
counter-based uniform stream so that its outputs can be committed as golden fixtures
(tests/golden/make_golden.py).  /root/reference does not exist on the GPU box: nothing
that runs there imports this file.

Injection recipe (SURVEY.md Appendix C probe8):
  * ``np.random.choice`` is replaced, for the duration of a sampling call, by numpy's own
    documented algorithm for ``choice(a, p=p)`` -- ``cdf = p.cumsum(); cdf /= cdf[-1];
    a[cdf.searchsorted(u, side='right')]`` -- with ``u`` taken from the shared Philox
    stream instead of the global MT19937.  ``check_choice_rule`` verifies on the spot that
    this is what the installed numpy computes (same u => same pick).
  * ``_single_walk`` is wrapped (not modified) to tell the injector which walk is running.
Everything else -- adjacency order, dead ends, Counter, sorted, weights -- is the
reference's own code.
"""
from __future__ import annotations

import sys
import types

import numpy as np

REF = "/root/reference"


def install_torch_geometric_stub():
    """model/pinsage.py:5-6 only needs these names to import."""
    if "torch_geometric" in sys.modules:
        return
    import torch.nn as nn
    tg = types.ModuleType("torch_geometric")
    tg_nn = types.ModuleType("torch_geometric.nn")
    tg_utils = types.ModuleType("torch_geometric.utils")

    class MessagePassing(nn.Module):
        def __init__(self, aggr="add"):
            super().__init__()

    tg_nn.MessagePassing = MessagePassing
    tg_utils.to_dense_batch = lambda *a, **k: None
    tg.nn, tg.utils = tg_nn, tg_utils
    sys.modules.update({"torch_geometric": tg, "torch_geometric.nn": tg_nn,
                        "torch_geometric.utils": tg_utils})


def import_reference():
    install_torch_geometric_stub()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import utils.random_walk as rw          # noqa: E402
    import model.pinsage as ps              # noqa: E402
    import model.layers as ly               # noqa: E402
    import model.aggregators as ag          # noqa: E402
    import utils.evaluation as ev           # noqa: E402
    return rw, ps, ly, ag, ev


def check_choice_rule(n_trials=2000, seed=7):
    """np.random.choice(a, p=p) == a[searchsorted(cumsum(p)/last, u, 'right')] with ONE
    random_sample() per call, for the installed numpy."""
    rs = np.random.RandomState(seed)
    gen = np.random.RandomState(seed + 1)
    for _ in range(n_trials):
        n = int(gen.randint(1, 400))
        w = gen.randint(1, 11, size=n) * 0.5
        p = w / w.sum()
        a = gen.randint(0, 10**6, size=n).tolist()
        state = rs.get_state()
        picked = rs.choice(a, p=p)
        rs.set_state(state)
        u = rs.random_sample()
        cdf = p.cumsum()
        cdf /= cdf[-1]
        if a[int(cdf.searchsorted(u, side="right"))] != picked:
            return False
    return True


class UniformInjector:
    """Feeds u[start, walk, step] (a callable) to the reference sampler."""

    def __init__(self, sampler, uniform_fn):
        self.sampler = sampler
        self.uniform_fn = uniform_fn      # (start, walk, step) -> float in [0,1)
        self.start = self.walk = self.step = None

    def _choice(self, a, p=None, **_kw):
        u = self.uniform_fn(self.start, self.walk, self.step)
        self.step += 1
        cdf = np.asarray(p, dtype=np.float64).cumsum()
        cdf /= cdf[-1]
        return np.asarray(a)[int(cdf.searchsorted(u, side="right"))]

    def batch_sample(self, nodes, num_neighbors):
        orig_choice = np.random.choice
        orig_walk = self.sampler._single_walk
        inj = self

        def wrapped_walk(start_node):
            inj.walk += 1
            inj.step = 0
            return orig_walk(start_node)

        np.random.choice = self._choice
        self.sampler._single_walk = wrapped_walk
        try:
            all_n, all_w = [], []
            for node in nodes:
                self.start, self.walk = int(node), -1
                n, w = self.sampler.sample_neighbors(int(node), num_neighbors)
                all_n.append(n)
                all_w.append(w)
            return all_n, all_w
        finally:
            np.random.choice = orig_choice
            del self.sampler._single_walk
