"""Pieces shared by the model engines: the flat parameter/gradient store and the fused training step
(forward -> CE+Dice(+argmax) -> backward -> gradient all-reduce -> Adam) with CUDA-graph capture.

Reference path: training/change_detection_trainer.py:136-177 (forward, loss, backward, optimizer step).
An engine provides `ops, params, device, N, H, W, logits` (the [N,K,H,W] fp32 tensor the criterion reads),
`forward(xA, xB, training)` and `backward(dlogits)`.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch


def _align(n: int, a: int = 64) -> int:
    return (n + a - 1) // a * a


class FlatParams:
    """All trainable parameters of a module as views into one flat fp32 buffer (+ flat grad)."""

    def __init__(self, module: torch.nn.Module):
        self.module = module
        self.names: List[str] = []
        self.offsets: Dict[str, Tuple[int, torch.Size]] = {}
        off = 0
        for name, p in module.named_parameters():
            self.names.append(name)
            self.offsets[name] = (off, p.shape)
            off += _align(p.numel(), 4)
        self.numel = _align(off, 4)
        self.flat: Optional[torch.Tensor] = None
        self.grad: Optional[torch.Tensor] = None

    def ensure(self, device) -> bool:
        """(Re)flatten if the module's parameters are not views of the flat buffer (e.g. after .to())."""
        params = dict(self.module.named_parameters())
        ok = self.flat is not None and self.flat.device == torch.device(device)
        if ok:
            base = self.flat.data_ptr()
            for name in self.names:
                if params[name].data_ptr() != base + 4 * self.offsets[name][0]:
                    ok = False
                    break
        if ok:
            return False
        if self.flat is not None:
            import warnings
            warnings.warn("module parameters no longer alias the flat parameter buffer (e.g. after .to() or load of new Parameter objects): "
                          "re-flattening; the flat layout is deterministic, so existing optimizer moments stay aligned", RuntimeWarning, stacklevel=2)
        flat = torch.zeros(self.numel, dtype=torch.float32, device=device)
        grad = torch.zeros(self.numel, dtype=torch.float32, device=device)
        for name in self.names:
            off, shape = self.offsets[name]
            p = params[name]
            flat[off:off + p.numel()].copy_(p.data.reshape(-1).to(device=device, dtype=torch.float32))
            p.data = flat[off:off + p.numel()].view(shape)
        self.flat, self.grad = flat, grad
        return True

    def p(self, name: str) -> torch.Tensor:
        off, shape = self.offsets[name]
        return self.flat[off:off + shape.numel()]

    def g(self, name: str) -> torch.Tensor:
        off, shape = self.offsets[name]
        return self.grad[off:off + shape.numel()]

    def grad_views(self):
        return [self.g(n).view(self.offsets[n][1]) for n in self.names]


class TrainStepMixin:
    # ------------------------------------------------------------------------------------------
    # fused training step: forward -> CE+Dice (+argmax) -> backward -> (all-reduce) -> Adam
    # (training/change_detection_trainer.py:136-177 without the two loss.item() host syncs)
    # ------------------------------------------------------------------------------------------
    def init_training(self, class_weights=(1.0, 1.0, 1.0), ignore_index: int = 3, lr: float = 1e-3,
                      betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, process_group=None,
                      dice_weight: float = 1.0, optimizer: str = "adam", momentum: float = 0.0,
                      bucket_mb: float = 32.0, overlap_comm: bool = True):
        dev = self.device
        self.params.ensure(dev)
        self.cw = torch.tensor(class_weights, dtype=torch.float32, device=dev)
        self.ignore_index = ignore_index
        self.dice_weight = float(dice_weight)     # 1: CE+Dice (utilities/bce_and_dice.py), 0: plain cross-entropy (the reference default)
        self.hp = dict(lr=lr, b1=betas[0], b2=betas[1], eps=eps, wd=weight_decay, momentum=momentum)
        if optimizer not in ("adam", "adamw", "sgd"):
            raise NotImplementedError(f"fused optimizer '{optimizer}' (adam: change_detection_trainer.py:52-54, adamw: :55-60, sgd: :61-66)")
        self.optimizer = optimizer
        if getattr(self, "adam_m", None) is not None and int(self.adam_step.item()) > 0:
            import warnings
            warnings.warn("init_training() called on an engine that has already stepped: the optimizer moments and the step counter are "
                          "reset (use adopt_training_state(old_engine) to carry them over)", RuntimeWarning, stacklevel=2)
        self.adam_m = torch.zeros_like(self.params.flat)
        self.adam_v = torch.zeros_like(self.params.flat)
        self.adam_step = torch.zeros(1, dtype=torch.int32, device=dev)
        self.loss3 = torch.zeros(3, dtype=torch.float32, device=dev)
        self.dlogits = torch.zeros_like(self.logits)
        self.pred = torch.zeros(self.N, self.H, self.W, dtype=torch.uint8, device=dev)
        self.loss_ws = self.ops.ce_dice_workspace(self.N, dev)
        self.pg = process_group
        self.world = 1
        if process_group is not None:
            import torch.distributed as dist
            self.world = dist.get_world_size(process_group)
            dist.broadcast(self.params.flat, src=dist.get_global_rank(process_group, 0) if hasattr(dist, "get_global_rank") else 0,
                           group=process_group)
            # data-parallel ranks must draw DIFFERENT dropout / drop-path masks (torch DDP: every process has its own RNG stream)
            rank = dist.get_rank(process_group)
            for attr in ("dropout_seed", "seed"):
                if hasattr(self, attr) and not getattr(self, "_seed_rank_mixed", False):
                    setattr(self, attr, (int(getattr(self, attr)) + 0x9E3779B1 * rank) & 0x7FFFFFFF)
            self._seed_rank_mixed = True
        # bucketed gradient all-reduce overlapped with the rest of the backward (see _grads_ready)
        self.bucket_bytes = int(bucket_mb * 2 ** 20)
        self.overlap_comm = bool(overlap_comm) and self.world > 1
        self._ready, self._ready_bytes, self._reduced, self._comm_done, self._cap = [], 0, [], [], None
        self.comm_stream = torch.cuda.Stream(device=dev) if (self.world > 1 and torch.device(dev).type == "cuda") else None
        self.comm_stats = {"messages": 0, "bytes": 0}
        self.graph = None

    def adopt_training_state(self, old) -> bool:
        """Take over the optimizer state and the stochastic-layer step counter of the engine this one replaces (another batch
        geometry of the same model: the ragged last batch of an epoch, a validation-sized batch).  Without the counter the same
        dropout mask sequence would restart with every rebuild."""
        if old is None or old is self or not hasattr(old, "adam_m") or old.adam_m.numel() != self.adam_m.numel():
            return False
        self.adam_m.copy_(old.adam_m)
        self.adam_v.copy_(old.adam_v)
        self.adam_step.copy_(old.adam_step)
        if hasattr(self, "do_step") and hasattr(old, "do_step"):
            self.do_step.copy_(old.do_step)
        return True

    def _fwd_loss_bwd(self, *args):
        """args = the model inputs (two dates for change detection, one stacked image for segmentation) + the mask."""
        inputs, mask = args[:-1], args[-1]
        logits = self.forward(*inputs, training=True)
        self.ops.ce_dice(logits, mask, self.cw, self.ignore_index, 1.0, self.loss3, self.dlogits, self.pred, self.loss_ws, self.dice_weight)
        self.backward(self.dlogits)

    # ------------------------------------------------------------------------------------------
    # Data-parallel gradient exchange: NCCL all-reduce (sum; the optimizer applies 1/world) of the flat fp32
    # gradient buffer, in BUCKETS issued on a side stream as soon as a range of the buffer is final, so that the transfer overlaps
    # the rest of the backward.  Engines call _grads_ready(lo, hi) when flat-gradient elements [lo, hi) will not be written again
    # in this step (ViT: after each transformer block's backward; the head first); whatever was never announced goes out in one last
    # message after the backward.  Under CUDA-graph capture every bucket boundary ends one graph segment and starts the next: a
    # replayed step is graph | all-reduce (side stream) | graph | ... and the GPU runs bucket k's all-reduce under segment k+1.
    # ------------------------------------------------------------------------------------------
    def _grads_ready(self, lo: int, hi: int):
        if getattr(self, "world", 1) == 1 or not self.overlap_comm or hi <= lo or getattr(self, "_skip_comm", False):   # world: set by init_training
            return
        self._ready.append((int(lo), int(hi)))
        self._ready_bytes += 4 * (hi - lo)
        if self._ready_bytes >= self.bucket_bytes:
            self._flush_ready()

    @staticmethod
    def _merge(ranges):
        out = []
        for lo, hi in sorted(ranges):
            if out and lo <= out[-1][1]:
                out[-1] = (out[-1][0], max(out[-1][1], hi))
            else:
                out.append((lo, hi))
        return out

    def _flush_ready(self):
        if not self._ready:
            return
        ranges = self._merge(self._ready)
        self._ready, self._ready_bytes = [], 0
        self._reduced += ranges
        if self._cap is not None:
            self._segment_break(ranges)
        else:
            self._launch_allreduce(ranges)

    def _launch_allreduce(self, ranges):
        import torch.distributed as dist
        g = self.params.grad
        self.comm_stats["messages"] += len(ranges)
        self.comm_stats["bytes"] += sum(4 * (hi - lo) for lo, hi in ranges)
        if self.comm_stream is None:                       # CPU tensors (gloo, host-logic tests): synchronous
            for lo, hi in ranges:
                dist.all_reduce(g[lo:hi], group=self.pg)
            return
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(ev)
            for lo, hi in ranges:
                dist.all_reduce(g[lo:hi], group=self.pg)   # NCCL over NVLink / NVSwitch
            done = torch.cuda.Event()
            done.record(self.comm_stream)
        self._comm_done.append(done)

    def _wait_comm(self):
        main = torch.cuda.current_stream() if self.comm_stream is not None else None
        for ev in self._comm_done:
            main.wait_event(ev)
        self._comm_done = []

    def _allreduce(self):
        """After the backward: everything not yet announced goes out, then the compute stream waits for all buckets."""
        if self.world == 1:
            return
        if getattr(self, "_skip_comm", False):             # bench.py: the step without its exchange (exposed-time measurement)
            self._ready, self._ready_bytes, self._reduced = [], 0, []
            return
        done = self._merge(self._reduced + self._ready)
        rest, pos = [], 0
        for lo, hi in done:
            if lo > pos:
                rest.append((pos, lo))
            pos = max(pos, hi)
        if pos < self.params.numel:
            rest.append((pos, self.params.numel))
        self._ready += rest
        self._flush_ready()
        self._reduced = []
        if self._cap is None:
            self._wait_comm()

    # -- segmented CUDA-graph capture (data parallel) -------------------------------------------------
    def _segment_begin(self):
        g = torch.cuda.CUDAGraph()
        g.capture_begin(pool=self._cap["pool"], capture_error_mode="thread_local")
        self._cap["g"] = g

    def _segment_break(self, ranges):
        self._cap["g"].capture_end()
        self._cap["segs"].append(("graph", self._cap["g"]))
        self._cap["segs"].append(("comm", ranges))
        self._segment_begin()

    def _capture_segments(self, fn):
        """Capture fn() (forward + loss + backward + _allreduce) as graph segments separated by the bucket all-reduces."""
        self._cap = {"pool": torch.cuda.graph_pool_handle(), "segs": [], "g": None}
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream())
        try:
            with torch.cuda.stream(s):
                self._segment_begin()
                fn()
                self._cap["g"].capture_end()
                self._cap["segs"].append(("graph", self._cap["g"]))
        finally:
            segs, self._cap = self._cap["segs"], None
        torch.cuda.current_stream().wait_stream(s)
        return segs

    def _replay_segments(self, segs):
        for kind, x in segs:
            if kind == "graph":
                x.replay()
            else:
                self._launch_allreduce(x)
        self._wait_comm()

    def _optimizer(self):
        hp = self.hp
        if self.optimizer == "sgd":
            self.ops.sgd_step(self.params.flat, self.params.grad, self.adam_m, hp["lr"], hp["momentum"], hp["wd"], 1.0 / self.world)
            return
        step = self.ops.adamw_step if self.optimizer == "adamw" else self.ops.adam_step
        step(self.params.flat, self.params.grad, self.adam_m, self.adam_v, hp["lr"], hp["b1"], hp["b2"], hp["eps"],
             hp["wd"], 1.0 / self.world, self.adam_step)

    def train_step(self, *args) -> torch.Tensor:
        """One optimizer step on (inputs..., mask); returns the device tensor [total, dice, ce] (no host sync)."""
        self._fwd_loss_bwd(*args)
        self._allreduce()
        self._optimizer()
        return self.loss3

    def capture(self, *args, warmup: int = 2, optimizer_in_graph: bool = True):
        """Capture the step over STATIC input tensors: one CUDA graph on a single GPU; with data parallelism two graphs
        (forward+loss+backward | optimizer) around the eager NCCL all-reduce.  Returns a callable that replays one step.
        warmup: real training steps run on `args` before capturing (0 when the caller has already stepped this engine eagerly);
        optimizer_in_graph=False keeps the optimizer launch eager, so a learning-rate change (hp["lr"], a by-value kernel argument)
        takes effect without re-capturing."""
        self.params.ensure(self.device)
        if warmup > 0:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(warmup):
                    self.train_step(*args)
            torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if self.world == 1 and optimizer_in_graph:
            self.graph = torch.cuda.CUDAGraph()
            # thread_local everywhere: a DataLoader pin-memory thread (or NCCL's watchdog) may issue CUDA calls during the capture
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.train_step(*args)
            self.replay = self.graph.replay
        elif self.world == 1:
            ga = torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga, capture_error_mode="thread_local"):
                self._fwd_loss_bwd(*args)
            self.graph = ga

            def replay():
                ga.replay()
                self._optimizer()
            self.replay = replay
        else:
            # data parallel: graph segments split at the bucket boundaries, NCCL launched between them on the side stream
            def body():
                self._fwd_loss_bwd(*args)
                self._allreduce()
            segs = self._capture_segments(body)
            self.graph = segs
            gb = None
            if optimizer_in_graph:
                gb = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gb, capture_error_mode="thread_local"):
                    self._optimizer()

            def replay():
                self._replay_segments(segs)
                if gb is not None:
                    gb.replay()
                else:
                    self._optimizer()
            self.replay = replay
        return self.replay
