"""The fused training step and the on-device epoch stream (sampler + shuffle + batching).

`FusedTrainStep` is what the four entry points (scripts/train_neumf.py, pretrain.py,
train_teacher.py, train_student.py) run instead of the reference inner loop
    optimizer.zero_grad(); prediction = model(user, item); loss = criterion(prediction, label)
    loss.backward(); optimizer.step(); total_loss += loss.item()
(reference scripts/train_neumf.py:106-118): per step it launches
    [teacher forward]  ->  fused forward+loss+backward  ->  sparse-row Adam / SGD
with no host synchronisation; the loss is accumulated on the device and read once per epoch.
`capture()` records a window of steps into a CUDA graph for the launch-bound small configs.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib, ops
from .models import NCF


class FusedTrainStep:
    def __init__(self, model: NCF, optimizer: str = "adam", lr: float = 1e-3, betas=(0.9, 0.999),
                 eps: float = 1e-8, max_batch: int = 256, teacher: Optional[NCF] = None,
                 alpha: float = 0.5, distillation=None):
        """`teacher` + `alpha`: response (teacher-score) KD, evaluated in the fused kernel's epilogue.
        `distillation`: any ncf_b200.distillation object (Response / SoftTarget / Feature / Attention /
        Unified); its teacher, weights and adapters are taken from it.  Objectives other than the
        response one run as forward -> ncf_loss_grad_kd -> ncf_backward (+ ncf_feature_kd)."""
        self.kd = None
        if distillation is not None:
            if distillation.student_model is not model:
                raise ValueError("distillation.student_model must be the model being trained")
            spec = distillation.fused_spec()
            teacher = distillation.teacher_model
            if spec["kd_mode"] == 0 and not spec["features"]:
                alpha = spec["w_task"]                    # plain response KD: the fused epilogue does it
            else:
                self.kd = spec
        p0 = next(model.parameters())
        if not p0.is_cuda:
            raise _lib.NcfError("FusedTrainStep needs the model on a CUDA device (no CPU path)")
        if optimizer not in ("adam", "sgd"):
            raise ValueError("optimizer must be 'adam' or 'sgd'")
        if model.dropout and model.dropout > 0:
            raise NotImplementedError("the fused step implements dropout=0.0 (reference default)")
        self.model, self.teacher = model, teacher
        self.optimizer, self.lr, self.betas, self.eps, self.alpha = optimizer, float(lr), betas, eps, float(alpha)
        self.max_batch = int(max_batch)
        dev = p0.device
        self.device = dev
        mt = model.abi_type()
        self.grads = ops.GradBuffers.allocate(mt, model.factor_num, model.num_layers, model.user_num,
                                              model.item_num, self.max_batch, dev)
        self.state = (ops.AdamBuffers.allocate(mt, model.factor_num, model.num_layers, model.user_num,
                                               model.item_num, dev) if optimizer == "adam" else None)
        self.loss_accum = torch.zeros(1, dtype=torch.float64, device=dev)
        self.num_steps = 0
        self._refresh()
        # one scratch for the fused step and for plain student forwards (the distillation objectives outside the
        # fused epilogue run forward -> loss -> backward)
        self.workspace = torch.empty(max(ops.train_workspace_bytes(self._m, self.max_batch),
                                         ops.forward_workspace_bytes(self._m, self.max_batch)),
                                     dtype=torch.uint8, device=dev)
        self.teacher_logits = (torch.empty(self.max_batch, dtype=torch.float32, device=dev)
                               if teacher is not None else None)
        if self.kd is not None:
            self.student_logits = torch.empty(self.max_batch, dtype=torch.float32, device=dev)
            self.dlogit = torch.empty(self.max_batch, dtype=torch.float32, device=dev)
        self.teacher_workspace = None
        if teacher is not None:
            teacher.eval()
            for p in teacher.parameters():  # reference src/distillation/base.py:16-18
                p.requires_grad = False
            # the teacher forward is captured into CUDA graphs with the rest of the step, so its scratch
            # must be a buffer this object owns (the shared inference workspace can be replaced later)
            nb = ops.forward_workspace_bytes(teacher.abi_struct(), self.max_batch)
            self.teacher_workspace = torch.empty(max(nb, 256), dtype=torch.uint8, device=dev)
        self._dirty = False  # True while some rows lag behind the dense-Adam state
        self.fused_prepare = True  # False: ncf_mark_rows + ncf_adam_catchup instead of ncf_adam_prepare (tests)
        # Adam over every row instead of the touched rows once a step touches this share of the tables
        # (expected distinct rows of a uniform batch; NCF_ADAM_DENSE=0/1 forces a mode)
        self.dense_share = 0.42

    def _refresh(self):
        """Re-reads the parameter pointers (they change if the module is moved or reloaded)."""
        self._m = self.model.abi_struct()
        self._g = self.grads.struct()
        self._s = self.state.struct() if self.state is not None else None
        self._tm = self.teacher.abi_struct() if self.teacher is not None else None

    def dense_adam(self, global_batch: int) -> bool:
        """Whether a step of `global_batch` samples should run the optimiser over all rows: the
        reference's dense Adam literally (ncf_adam_step_dense) instead of list + catch-up + row gather."""
        import math
        import os
        forced = os.environ.get("NCF_ADAM_DENSE")
        if forced in ("0", "1"):
            return forced == "1"
        U, I = self.model.user_num, self.model.item_num
        touched = U * -math.expm1(-global_batch / U) + I * -math.expm1(-global_batch / I)
        return touched >= self.dense_share * (U + I)

    # -- one optimisation step on a device batch ------------------------------------------------
    def step(self, user: torch.Tensor, item: torch.Tensor, label: torch.Tensor,
             logits_out: Optional[torch.Tensor] = None) -> None:
        B = user.numel()
        if B > self.max_batch:
            raise _lib.NcfError(f"batch {B} exceeds max_batch {self.max_batch}")
        t_logits = None
        if self.teacher is not None:
            t_logits = self.teacher_logits[:B]
            ops.forward(self._tm, user, item, out=t_logits, workspace=self.teacher_workspace)
        dense = self.optimizer == "adam" and self.dense_adam(B)
        if self.optimizer == "adam":
            # rows this batch reads must first catch up with the dense-Adam trajectory — unless every
            # row is current already (the previous steps ran the optimiser over all rows)
            if self._dirty or not dense:
                if self.fused_prepare:   # registration + catch-up as one launch
                    ops.adam_prepare(self._m, self._g, self._s, user, item, self.lr, self.betas[0],
                                     self.betas[1], self.eps)
                else:                    # the same as two: list first, then a walk over the list
                    ops.mark_rows(self._m, self._g, user, item)
                    ops.adam_catchup(self._m, self._g, self._s, self.lr, self.betas[0], self.betas[1], self.eps)
        else:
            ops.mark_rows(self._m, self._g, user, item)
        if self.kd is None:
            ops.train_step_grads(self._m, self._g, user, item, label, t_logits, self.alpha,
                                 self.loss_accum, self.workspace, logits_out)
        else:
            self._kd_grads(user, item, label, t_logits, logits_out)
        if dense:
            ops.adam_step_dense(self._m, self._g, self._s, self.lr, self.betas[0], self.betas[1], self.eps)
            self._dirty = False
        elif self.optimizer == "adam":
            ops.adam_step(self._m, self._g, self._s, self.lr, self.betas[0], self.betas[1], self.eps)
            self._dirty = True
        else:
            ops.sgd_step(self._m, self._g, self.lr)
        self.num_steps += 1

    def _kd_grads(self, user, item, label, t_logits, logits_out=None) -> None:
        """Objectives the fused epilogue does not evaluate: student forward -> loss + dloss/dlogit
        (ncf_loss_grad_kd) -> fused backward from that dlogit (ncf_backward) -> feature-matching terms
        (ncf_feature_kd: loss value + row gradients on the embedding-level features)."""
        kd, B = self.kd, user.numel()
        x = logits_out if logits_out is not None else self.student_logits[:B]
        ops.forward(self._m, user, item, out=x, workspace=self.workspace)
        dl = self.dlogit[:B]
        ops.loss_grad_kd(x, label, t_logits, kd["w_task"], kd["w_kd"], kd["temperature"], kd["kd_mode"],
                         self.loss_accum, dl)
        ops.backward(self._m, self._g, user, item, dl, self.workspace)
        for ft in kd["features"]:
            ops.feature_kd(self._m, self._tm, self._g, user, item, ft["kind"], ft["w"], ft["b"], ft["weight"],
                           self.loss_accum)

    def flush(self) -> None:
        """Brings every embedding row to the dense-Adam state of the current step.  Must run
        before the weights are read (evaluation, checkpoint, state_dict)."""
        if self.optimizer == "adam" and self._dirty:
            ops.adam_flush(self._m, self._s, self.lr, self.betas[0], self.betas[1], self.eps)
            self._dirty = False

    def pop_loss(self) -> float:
        """Sum of the per-batch mean losses since the last call (one device->host read)."""
        v = float(self.loss_accum.item())
        self.loss_accum.zero_()
        return v

    # -- CUDA-graph window ----------------------------------------------------------------------
    def capture(self, users: torch.Tensor, items: torch.Tensor, labels: torch.Tensor,
                batch: int, step_fn=None) -> "StepGraph":
        """Captures `len(users) // batch` consecutive steps over static window buffers.  `step_fn`: the
        step to record instead of self.step — e.g. a data-parallel wrapper's step, whose NCCL
        collectives are captured with the kernels."""
        return StepGraph(self, users, items, labels, batch, step_fn)


class StepGraph:
    """A CUDA graph of consecutive FusedTrainStep.step() calls over fixed window buffers
    (refilled in place by ncf_shuffle_epoch between replays)."""

    def __init__(self, ts: FusedTrainStep, users, items, labels, batch: int, step_fn=None):
        step_fn = step_fn or ts.step
        n = users.numel() // batch
        if n < 1:
            raise _lib.NcfError("window smaller than one batch")
        self.ts, self.n_steps, self.batch = ts, n, batch
        # a window captured while every row was current contains no catch-up for its first step
        self.assumes_current = ts.optimizer == "adam" and not ts._dirty
        self.graph = torch.cuda.CUDAGraph()
        # capture only records the launches: parameters and optimiser state are not advanced by it
        torch.cuda.synchronize()
        stream = torch.cuda.Stream()
        stream.wait_stream(torch.cuda.current_stream())
        # thread-local capture mode: NCCL's watchdog thread may query events while this thread captures
        mode = "thread_local" if (torch.distributed.is_available() and torch.distributed.is_initialized()) else "global"
        with torch.cuda.stream(stream):
            self.graph.capture_begin(capture_error_mode=mode)
            for i in range(n):
                sl = slice(i * batch, (i + 1) * batch)
                step_fn(users[sl], items[sl], labels[sl])
            self.graph.capture_end()
        torch.cuda.current_stream().wait_stream(stream)
        ts.num_steps -= n  # capture records, it does not execute
        self.dirty_after = ts._dirty
        ts._dirty = not self.assumes_current if ts.optimizer == "adam" else False

    def replay(self) -> None:
        if self.assumes_current and self.ts._dirty:
            self.ts.flush()
        self.graph.replay()
        self.ts.num_steps += self.n_steps
        if self.ts.optimizer == "adam":
            self.ts._dirty = self.dirty_after


class HostFedTrainer:
    """Training from HOST batches (pinned memory) with the copies off the critical path: two device
    buffer sets, one captured step graph per set, and a copy stream that stages batch k+1 while
    step k computes.  `prefetch(user, item, label)` stages a pinned host batch (asynchronous);
    `step()` runs the oldest staged batch and returns its loss — the same per-step `loss.item()`
    the reference loop reads (scripts/train_neumf.py:112-121).  Usage:

        hf = HostFedTrainer(ts, batch); hf.prefetch(*b[0])
        for k in range(n): hf.prefetch(*b[k + 1]); loss = hf.step()

    `launch()` / `wait()` split the step so that the GPU never waits for the host: up to `depth` steps may be
    staged or in flight, and `wait()` returns the loss of the oldest one (the reference only sums the losses for its log
    line, so reading step k's loss while step k+1 runs changes nothing):

        hf.prefetch(*b[0]); hf.launch(); hf.prefetch(*b[1])
        for k in range(1, n): hf.launch(); hf.prefetch(*b[k + 1]); loss = hf.wait()    # loss of step k-1
        loss = hf.wait()
    """

    def __init__(self, ts: FusedTrainStep, batch: int, step_fn=None, depth: int = 2):
        """`step_fn`: the step each graph records (default ts.step); a ReplicatedDataParallel.step makes this
        the host-fed trainer of a data-parallel rank (every rank must then call launch() in lock step).
        `depth`: buffer sets = steps that may be staged / in flight at once (2 hides the copies; more also rides
        out host hiccups longer than a step — a 65 536-sample step is 0.4 ms)."""
        if depth < 2:
            raise _lib.NcfError("HostFedTrainer: depth must be at least 2")
        dev = ts.device
        self.ts, self.batch, self.depth = ts, batch, depth
        self.bufs = [(torch.empty(batch, dtype=torch.int64, device=dev),
                      torch.empty(batch, dtype=torch.int64, device=dev),
                      torch.empty(batch, dtype=torch.float32, device=dev)) for _ in range(depth)]
        for b in self.bufs:  # valid indices for the capture pass
            b[0].zero_(); b[1].zero_(); b[2].zero_()
        # each graph = clear the loss accumulator, the step, park the loss in the set's own device word: the
        # read-back then runs on its own stream and the next step does not queue behind a PCIe round trip
        base_step = step_fn or ts.step
        self.loss_dev = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(depth)]

        def recorded(j):
            def fn(user, item, label):
                ts.loss_accum.zero_()
                base_step(user, item, label)
                self.loss_dev[j].copy_(ts.loss_accum)
            return fn
        self.graphs = [ts.capture(*b, batch, recorded(j)) for j, b in enumerate(self.bufs)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.read_stream = torch.cuda.Stream(device=dev)
        self.copied = [torch.cuda.Event() for _ in range(depth)]
        self.consumed = [torch.cuda.Event() for _ in range(depth)]
        for e in self.consumed:
            e.record()
        self.host_loss = torch.zeros(depth, dtype=torch.float64).pin_memory()
        self.done = [torch.cuda.Event() for _ in range(depth)]   # loss of the step on buffer set j is in host_loss[j]
        self.n_staged = 0   # batches staged so far
        self.n_run = 0      # batches trained (launched) so far
        self.n_waited = 0   # losses handed back so far

    def prefetch(self, user: torch.Tensor, item: torch.Tensor, label: torch.Tensor) -> None:
        if self.n_staged - self.n_run >= self.depth:
            raise _lib.NcfError(f"HostFedTrainer: all {self.depth} buffer sets hold batches that have not run yet")
        if user.numel() != self.batch:
            raise _lib.NcfError("HostFedTrainer: batches must have the captured size")
        j = self.n_staged % self.depth
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[j])
            self.bufs[j][0].copy_(user, non_blocking=True)
            self.bufs[j][1].copy_(item, non_blocking=True)
            self.bufs[j][2].copy_(label, non_blocking=True)
            self.copied[j].record(self.copy_stream)
        self.n_staged += 1

    def launch(self) -> None:
        """Enqueues the step on the oldest staged batch and the read-back of its loss."""
        if self.n_run >= self.n_staged:
            raise _lib.NcfError("HostFedTrainer.step: no staged batch")
        if self.n_run - self.n_waited >= self.depth:
            raise _lib.NcfError(f"HostFedTrainer.launch: {self.depth} steps are in flight, wait() for the oldest first")
        j = self.n_run % self.depth
        cur = torch.cuda.current_stream()
        cur.wait_event(self.copied[j])
        self.graphs[j].replay()
        self.consumed[j].record()
        with torch.cuda.stream(self.read_stream):
            self.read_stream.wait_event(self.consumed[j])
            self.host_loss[j:j + 1].copy_(self.loss_dev[j], non_blocking=True)
            self.done[j].record(self.read_stream)
        self.n_run += 1

    def wait(self) -> float:
        """Loss of the oldest step whose loss has not been read yet (blocks until that step is complete)."""
        if self.n_waited >= self.n_run:
            raise _lib.NcfError("HostFedTrainer.wait: no step in flight")
        j = self.n_waited % self.depth
        self.done[j].synchronize()
        self.n_waited += 1
        return float(self.host_loss[j])

    @property
    def in_flight(self) -> int:
        return self.n_run - self.n_waited

    def step(self) -> float:
        self.launch()
        return self.wait()


class EpochStream:
    """On-device replacement of `NCFData.ng_sample()` + `DataLoader(shuffle=True)` (reference
    src/data/datasets.py:53-83, scripts/train_neumf.py:55,102): a CSR of the observed pairs, a
    Philox negative sampler, and a keyed permutation that lays the epoch's S = P*(1+num_ng)
    samples out in shuffled order, window by window."""

    def __init__(self, pos_user: torch.Tensor, pos_item: torch.Tensor, user_num: int, item_num: int,
                 num_ng: int, seed: int = 0, p_offset: int = 0, observed=None):
        self.pos_user = pos_user.to(torch.int64).contiguous()
        self.pos_item = pos_item.to(torch.int64).contiguous()
        self.user_num, self.item_num, self.num_ng, self.seed = int(user_num), int(item_num), int(num_ng), int(seed)
        self.p_offset = int(p_offset)
        self.P = self.pos_user.numel()
        self.S = self.P * (1 + self.num_ng)
        # pairs to reject against (train_mat of the reference); defaults to the positives themselves
        ou, oi = observed if observed is not None else (self.pos_user, self.pos_item)
        self.rowptr, self.col = ops.csr_build(ou.to(torch.int64).contiguous(),
                                              oi.to(torch.int64).contiguous(), self.user_num)
        self.neg_item = torch.empty(self.P * self.num_ng, dtype=torch.int64, device=self.pos_user.device)
        self.epoch = -1

    def begin_epoch(self, epoch: int) -> None:
        """ng_sample(): draws this epoch's negatives."""
        self.epoch = int(epoch)
        ops.sample_neg(self.rowptr, self.col, self.pos_user, self.num_ng, self.item_num, self.seed,
                       self.epoch, self.p_offset, out=self.neg_item)

    def fill(self, q_begin: int, count: int, out_user, out_item, out_label) -> None:
        """Writes stream positions [q_begin, q_begin+count) of the current epoch."""
        ops.shuffle_epoch(self.pos_user, self.pos_item, self.neg_item, self.num_ng, self.seed,
                          self.epoch, q_begin, count, out_user, out_item, out_label)

    def num_batches(self, batch: int) -> int:
        return (self.S + batch - 1) // batch  # drop_last=False like the reference DataLoader


def train_epoch(ts: FusedTrainStep, stream: EpochStream, epoch: int, batch: int,
                window_steps: int = 64, use_graph: bool = True, cache: Optional[dict] = None):
    """One epoch of the reference loop (scripts/train_neumf.py:98-120) on the fused path.
    Returns (avg_loss, num_batches)."""
    dev = ts.device
    stream.begin_epoch(epoch)
    S, nb = stream.S, stream.num_batches(batch)
    cache = cache if cache is not None else {}
    W = max(1, min(window_steps, S // batch)) if S >= batch else 1
    key = ("win", batch, W)
    if key not in cache:
        cache[key] = (torch.empty(W * batch, dtype=torch.int64, device=dev),
                      torch.empty(W * batch, dtype=torch.int64, device=dev),
                      torch.empty(W * batch, dtype=torch.float32, device=dev))
    wu, wi, wl = cache[key]
    graph = None
    q = 0
    full_windows = (S // batch) // W if S >= batch else 0
    for w in range(full_windows):
        stream.fill(q, W * batch, wu, wi, wl)
        # the very first window runs eagerly: it loads the kernels before any capture
        if use_graph and (w > 0 or ("graph", batch, W) in cache):
            gkey = ("graph", batch, W)
            if gkey not in cache:
                cache[gkey] = ts.capture(wu, wi, wl, batch)
            graph = cache[gkey]
            graph.replay()
        else:
            for i in range(W):
                sl = slice(i * batch, (i + 1) * batch)
                ts.step(wu[sl], wi[sl], wl[sl])
        q += W * batch
    # tail: remaining full batches, then the short last batch
    while q < S:
        cnt = min(batch, S - q)
        stream.fill(q, cnt, wu, wi, wl)
        ts.step(wu[:cnt], wi[:cnt], wl[:cnt])
        q += cnt
    total = ts.pop_loss()
    return total / nb, nb
