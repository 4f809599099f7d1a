"""Host side of the fused training step: pinned host batch -> device -> one CUDA-graph replay per step.

The reference's loop (training/change_detection_trainer.py:113-177, segmentation_trainer.py:70-130) copies the batch with
`.to(device)` on the compute stream and synchronises twice per step through `loss.item()`.  Here the copy of batch i+1 runs on a
second stream while step i computes (`prefetch`), the step itself is a replayed CUDA graph over static input buffers (captured after
two eager steps, so no training step is ever spent on warm-up data), and the optimizer launch stays outside the graph so the
per-epoch learning-rate schedule keeps working without re-capturing.
"""
from __future__ import annotations

import torch

EAGER_STEPS = 2


class HostPipelineMixin:
    """Needs: self.configs, self._engine(first_input) and self._to_device(batch) -> [inputs..., mask] (device tensors)."""

    def _pipeline_state(self):
        st = self.__dict__.get("_pl")
        if st is None:
            st = self.__dict__["_pl"] = dict(copy_stream=None, staged=[], static=None, replay=None, engine=None)
        return st

    def _stage(self, batch):
        st, dev = self._pipeline_state(), torch.device(self.configs["device"])
        if st["copy_stream"] is None:
            st["copy_stream"] = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(st["copy_stream"]):
            tensors = self._to_device(batch)
            ev = torch.cuda.Event()
            ev.record()
        return tensors, ev

    def prefetch(self, batch):
        """Start the host->device copy of a batch that a later `step_host(batch)` (same object) will consume."""
        if torch.device(self.configs["device"]).type != "cuda":
            return
        st = self._pipeline_state()
        st["staged"] = st["staged"][-1:] + [(batch,) + self._stage(batch)]       # at most two batches in flight

    def step_host(self, batch):
        """One training step from a HOST batch (pinned tensors). Returns (device loss[3], device mask)."""
        st = self._pipeline_state()
        if torch.device(self.configs["device"]).type != "cuda":            # shadow-ops schedule tests only: no streams, no graphs
            tensors = self._to_device(batch)
            return self._engine(tensors[0]).train_step(*tensors), tensors[-1]
        hit = [e for e in st["staged"] if e[0] is batch]
        if hit:
            tensors, ev = hit[0][1:]
            st["staged"] = [e for e in st["staged"] if e[0] is not batch]
        else:
            tensors, ev = self._stage(batch)
        main = torch.cuda.current_stream()
        main.wait_event(ev)
        for t in tensors:
            t.record_stream(main)
        eng = self._engine(tensors[0])
        # the warm-up count lives ON the engine object: engines are rebuilt per batch geometry and freed, and CPython reuses ids
        n_eager = getattr(eng, "_eager_steps", 0)
        if not self.configs.get("cuda_graph", True) or n_eager < EAGER_STEPS:
            eng._eager_steps = n_eager + 1
            return eng.train_step(*tensors), tensors[-1]
        if st["engine"] is not eng:
            # first capture, or another batch geometry took over (e.g. after a ragged last batch the model builds a new engine for the
            # full-size batches): the old graph and its static inputs are dropped and this engine is captured instead
            st["replay"], st["static"] = None, None
            st["static"] = [torch.empty_like(t) for t in tensors]
            st["replay"] = eng.capture(*st["static"], warmup=0, optimizer_in_graph=False)
            st["engine"] = eng
        for s, t in zip(st["static"], tensors):
            s.copy_(t, non_blocking=True)
        st["replay"]()
        return eng.loss3, st["static"][-1]


def lookahead(loader):
    """Yields (batch, next_batch or None)."""
    it = iter(loader)
    try:
        cur = next(it)
    except StopIteration:
        return
    for nxt in it:
        yield cur, nxt
        cur = nxt
    yield cur, None
