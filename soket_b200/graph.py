"""soket_b200.graph -- a whole static-shape training step as ONE CUDA-graph launch.

SURVEY.md section 8f-1: at config 1 (batch 100, hidden 100) the step is ~70 tiny kernels and
its cost is the per-op host work of the eager engine (the reference spends ~10 us per array
call in tensor.pyx:_make_from_op / autodiff.pyx, this engine ~5 us), not the GPU.  Capturing
the step removes the host from the loop:

    x_buf, y_buf = soket.Tensor(X0), soket.Tensor(y0)          # static input buffers
    def step():
        loss = crit(model(x_buf), y_buf); loss.backward(); opt.step(); return loss
    g = StaticStep(step)            # 2 eager dry runs + 1 captured run (3 real steps)
    for xb, yb in batches:
        x_buf._data[...] = xb; y_buf._data[...] = yb           # device-side copies
        g.launch()
        print(g.loss.item())

What makes a replay equal to an eager step:
  * every buffer the captured kernels touch comes from a private allocator arena that is
    reserved for the graph (soket_b200.arena_*); parameters are updated in place, so the
    Python-visible tensors (params, their .grad, the loss) are the graph's own buffers;
  * dropout draws a fresh mask per replay: the kernels mix a device-side epoch counter into
    their seeds at run time and the graph's first node advances it;
  * the optimiser's kernel arguments must not change from step to step: SGD qualifies as is;
    Adam's bias corrections are host floats in the reference (optim.pyx:191-195) and would be
    frozen, so the default Adam is refused and ``Adam(..., capturable=True)`` keeps beta^t in
    device memory instead (same update, bit for bit; tests/test_graph_gpu.py).
"""
from __future__ import annotations

import soket_b200 as sk


class StaticStep:
    def __init__(self, fn, dry_runs: int = 2):
        if dry_runs < 2:
            raise ValueError("StaticStep: at least 2 dry runs are needed for the allocator to reach its steady state")
        self._fn = fn
        self._arena = sk.arena_create()
        self._graph = sk.Graph()
        self.loss = None
        sk.arena_begin(self._arena)
        try:
            for _ in range(dry_runs):
                self.loss = fn()
            sk.synchronize()
            self._graph.begin()
            try:
                sk.rng_epoch_advance()
                self.loss = fn()
            finally:
                self._graph.end()
        finally:
            sk.arena_end()
        self._graph.keep(self.loss)
        # a capture records, it does not execute: run the captured step once so that the third
        # step really happens and `self.loss` (the graph's own buffer) holds its value
        self._graph.launch()

    def launch(self):
        """Replay the captured step (asynchronous; read `self.loss.item()` to synchronise)."""
        self._graph.launch()

    def close(self):
        self._graph = None
        self.loss = None
        if self._arena is not None:
            sk.arena_destroy(self._arena)
            self._arena = None
