"""Multi-GPU training (one process per GPU, torch.distributed over NCCL / NVLink).

The reference is single-process; both modes below are new design that must
reproduce the single-process semantics at the GLOBAL batch.

ReplicatedDataParallel — BASELINE config 4 (MovieLens-sized tables).  Two layouts:

* **user-partitioned (default).**  Sharding follows the data: rank r OWNS the contiguous user range
  `user_range(U, world, r)` and trains on the samples of those users only, so the user tables (84 % of
  the parameters at the ml20m shape), their gradients and their Adam state never leave the rank.  The
  item tables and the tower are replicated.  Per step:
    1. [lazy-Adam mode only] all_gather of the item indices (8 B/sample): every rank registers the
       item rows the global batch touches plus its own user rows and replays their pending
       zero-gradient Adam steps;
    2. fused forward+loss+backward on the local samples with the loss mean over the GLOBAL batch
       (ncf_train_step_grads_norm);
    3. ONE all-reduce (sum) of the contiguous tail [item GMF | item MLP | tower] of the flat gradient
       buffer — 17 MB at the ml20m shape instead of the 106 MB of the fully replicated layout.
       north_star names the tower all-reduce; the item-row gradients have to travel too or the
       replicas diverge.  NCCL leaves bit-identical sums on every rank, so the
       replicated parts stay bit-identical;
    4. Adam: all rows of the items + the rank's own user range (ncf_adam_step_dense_range) when the
       global batch touches a large share of the rows, else the touched rows (ncf_adam_step).
  Rows of other ranks' users are stale on this rank until `sync_user_tables()` (checkpoint / full
  evaluation); evaluation of the rank's own users needs no exchange at all.

* **fully replicated** (`partition_users=False`, the round-1 layout): every rank holds everything and
  takes any slice of the global batch; ONE all-reduce (average) of the whole flat gradient buffer.
  From 4 GPUs the optimiser is sharded (`_setup_sharded`): reduce-scatter of the gradients ->
  elementwise Adam on the rank's 1/N slice (ncf_adam_range) -> all-gather of the parameters;
  NCF_DP_P2P=1 does that exchange inside the optimiser kernel over CUDA-IPC peer buffers
  (ncf_adam_p2p: 0.21 ms vs 0.35 ms for the three NCCL-bracketed steps at N=2, profiles/r02).

`partition` / `user_range` are pure functions so that the plan is testable on CPU (gloo).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import ops


def partition(n: int, world: int, rank: int):
    """Contiguous slice [lo, hi) of n samples owned by `rank` (remainder to the first ranks)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def user_range(user_num: int, world: int, rank: int):
    """Contiguous range [lo, hi) of user rows owned by `rank` in the user-partitioned layout."""
    return partition(user_num, world, rank)


def gather_indices(user: torch.Tensor, item: torch.Tensor, world: int):
    """all_gather of the per-rank (user, item) index slices -> global [world * B] tensors."""
    B = user.numel()
    # two collectives straight into their rank-major outputs: no stack / slice copies around them
    gu = torch.empty(world * B, dtype=user.dtype, device=user.device)
    gi = torch.empty(world * B, dtype=item.dtype, device=item.device)
    dist.all_gather_into_tensor(gu, user.contiguous())
    dist.all_gather_into_tensor(gi, item.contiguous())
    return gu, gi


def make_rank_barrier(dev):
    """A stream-ordered rank barrier as a callable: the peer-memory barrier kernel (default on one node), or a
    one-element NCCL all-reduce (NCF_PEER_BARRIER=0)."""
    if os.environ.get("NCF_PEER_BARRIER", "1") != "0" and dist.get_world_size() <= 8 and dev.type == "cuda":
        return ops.PeerBarrier(dev).wait
    flag = torch.zeros(1, dtype=torch.float32, device=dev)
    return lambda: dist.all_reduce(flag)


def average_(t: torch.Tensor, world: int):
    """In-place rank average (NCCL has ReduceOp.AVG; gloo does not)."""
    if dist.get_backend() == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t.div_(world)
    return t


class ReplicatedDataParallel:
    """Data-parallel wrapper of a FusedTrainStep (see the module docstring for the two layouts)."""

    def __init__(self, ts, check_replicas: bool = True, partition_users=None, global_batch=None):
        """`global_batch`: samples per step over all ranks (default ts.max_batch * world); decides at
        construction whether steps run the all-rows optimiser, which the peer-memory tail requires."""
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.ts = ts
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if ts.grads.flat is None:
            raise RuntimeError("gradient buffers must be one flat allocation")
        if ts.optimizer != "adam":
            raise NotImplementedError("data-parallel training is implemented for Adam")
        if partition_users is None:
            partition_users = os.environ.get("NCF_DP_PARTITION", "1") != "0"
        if os.environ.get("NCF_DP_P2P") == "1" or os.environ.get("NCF_DP_SHARD_ADAM") == "1":
            partition_users = False    # those exchanges exist in the fully replicated layout only
        self.partition_users = bool(partition_users)
        self.user_lo, self.user_hi = (user_range(ts.model.user_num, self.world, self.rank)
                                      if self.partition_users else (0, ts.model.user_num))
        # the global batch may touch up to world * B distinct rows
        dev = ts.device
        cap = ts.max_batch * self.world
        ts.grads.user_list = torch.zeros(min(cap, ts.model.user_num), dtype=torch.int64, device=dev)
        ts.grads.item_list = torch.zeros(min(cap, ts.model.item_num), dtype=torch.int64, device=dev)
        ts._refresh()
        if check_replicas:
            for p in ts.model.parameters():  # start from rank 0's weights
                dist.broadcast(p.data, src=0)
        # offset (floats) of the replicated tail [item GMF | item MLP | tower] / of the tower in the flat buffer
        g = ts.grads
        first_item = g.g_item_gmf if g.g_item_gmf is not None else g.g_item_mlp
        self.n_user_flat = ((first_item if first_item is not None else g.g_tower).data_ptr() - g.flat.data_ptr()) // 4
        self.n_rows_flat = (g.g_tower.data_ptr() - g.flat.data_ptr()) // 4
        self.comm_stream, self.sharded, self.tail = None, None, None
        if self.partition_users:
            # The replicated tail [item GMF | item MLP | tower] over peer memory (default on one NVLink node
            # when steps run the all-rows optimiser; NCF_DP_P2P_TAIL=0 keeps the NCCL all-reduce): one kernel
            # does reduce -> Adam -> broadcast of that tail (ncf_adam_p2p) instead of all-reduce + Adam.
            want = os.environ.get("NCF_DP_P2P_TAIL", "1") != "0"
            if want and dev.type == "cuda" and dist.get_backend() == "nccl" and 2 <= self.world <= 8 \
                    and ts.dense_adam(int(global_batch) if global_batch else ts.max_batch * self.world) \
                    and os.environ.get("NCF_ADAM_DENSE") != "0" \
                    and g.g_item_gmf is not None and g.g_item_mlp is not None and ts.model.factor_num % 4 == 0:
                self._setup_p2p_tail()
            # opt-in (NCF_DP_OVERLAP=1): the item-row gradients are complete when the tower kernel is (before the
            # weight-gradient kernel on the tcgen05 path), so their all-reduce can run on a side stream under that
            # kernel.  Measured at N=2 on B200: the NCCL CTAs take SMs from the one-CTA-per-SM weight-gradient
            # kernel - 0.433 vs 0.437 ms for a synchronised step, 0.477 vs 0.428 ms in a run-ahead loop - so it
            # stays off by default.
            if dev.type == "cuda" and dist.get_backend() == "nccl" and os.environ.get("NCF_DP_OVERLAP", "0") == "1" \
                    and self.n_rows_flat > self.n_user_flat:
                self.comm_stream = torch.cuda.Stream(device=dev)
            return
        # ---- fully replicated layout ----
        # the row gradients (all of the flat buffer but its tower tail) are reduced on a side stream as
        # soon as they are complete, while the weight-gradient kernel is still running
        # (opt-in, NCF_DP_OVERLAP=1: measured on B200 it gains 7 % at 4 GPUs but loses 2-6 % at 2 and 8,
        # where the all-reduce's CTAs mostly take issue slots from the weight-gradient kernel)
        overlap = os.environ.get("NCF_DP_OVERLAP") == "1"
        self.comm_stream = torch.cuda.Stream(device=dev) if (dev.type == "cuda" and overlap) else None
        # optimiser sharding: reduce-scatter the gradients, Adam on the own slice of one flat parameter
        # buffer, all-gather the parameters - see _setup_sharded / _sharded_step.  On by default from 4
        # GPUs when the global batch runs the optimiser in its all-rows mode anyway (measured on B200 at
        # the bench workload: 394 vs 349 M samples/s at N=4, 191 vs 200 at N=2); NCF_DP_SHARD_ADAM=0/1
        # overrides.
        want_shard = os.environ.get("NCF_DP_SHARD_ADAM")
        auto = self.world >= 4 and ts.dense_adam(ts.max_batch * self.world)
        if os.environ.get("NCF_DP_P2P") == "1" and self.world >= 2:
            auto = True   # the peer-memory exchange exists in the sharded step only
        if dev.type == "cuda" and (want_shard == "1" or (want_shard is None and auto)):
            self._setup_sharded()

    # -- the replicated tail over peer memory -------------------------------------------------------------------------
    def _setup_p2p_tail(self):
        """Re-homes the gradients and parameters of the item tables and the tower into two buffers the other
        ranks map through CUDA IPC, laid out [item GMF | item MLP | tower] and padded so that they split evenly:
        rank r owns elements [r * per, (r + 1) * per) and keeps the Adam moments of that slice only."""
        ts, W, dev = self.ts, self.world, self.ts.device
        ts.flush()
        model, g, st = ts.model, ts.grads, ts.state
        tower_params = [q for lin in model.linears() for q in (lin.weight, lin.bias)]
        tower_params += [model.predict_layer.weight, model.predict_layer.bias]
        n_ig, n_im, nt = g.g_item_gmf.numel(), g.g_item_mlp.numel(), g.g_tower.numel()
        pad4 = lambda n: (n + 3) // 4 * 4
        o_im, o_t = pad4(n_ig), pad4(n_ig) + pad4(n_im)
        n = o_t + pad4(nt)
        per = -(-n // (4 * W)) * 4
        n_pad = per * W
        gbuf, pbuf = ops.PeerBuffer(n_pad, dev), ops.PeerBuffer(n_pad, dev)
        gt, pt = gbuf.tensor, pbuf.tensor
        mflat = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        vflat = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        for off, cnt, gname, param, mname, vname in ((0, n_ig, "g_item_gmf", model.embed_item_GMF.weight, "m_item_gmf", "v_item_gmf"),
                                                      (o_im, n_im, "g_item_mlp", model.embed_item_MLP.weight, "m_item_mlp", "v_item_mlp")):
            old = getattr(g, gname)
            gt[off:off + cnt].copy_(old.reshape(-1))
            pt[off:off + cnt].copy_(param.data.reshape(-1))
            mflat[off:off + cnt].copy_(getattr(st, mname).reshape(-1))
            vflat[off:off + cnt].copy_(getattr(st, vname).reshape(-1))
            setattr(g, gname, gt[off:off + cnt].view_as(old))
            param.data = pt[off:off + cnt].view_as(param.data)
        gt[o_t:o_t + nt].copy_(g.g_tower)
        mflat[o_t:o_t + nt].copy_(st.m_tower)
        vflat[o_t:o_t + nt].copy_(st.v_tower)
        g.g_tower = gt[o_t:o_t + nt]
        o = o_t
        for q in tower_params:              # the flat tower order of the C ABI (ncf_tower_param_count)
            cnt = q.numel()
            pt[o:o + cnt].copy_(q.data.reshape(-1))
            q.data = pt[o:o + cnt].view_as(q.data)
            o += cnt
        assert o == o_t + nt, "tower layout mismatch"
        g.flat = None                        # the gradients are no longer one allocation
        lo = self.rank * per
        handles = [None] * W
        dist.all_gather_object(handles, (gbuf.handle(), pbuf.handle()))
        self.tail = {"per": per, "lo": lo, "g": gt, "p": pt, "bufs": (gbuf, pbuf),
                     "m": mflat[lo:lo + per].clone(), "v": vflat[lo:lo + per].clone(),
                     "barrier": self._make_barrier(dev),
                     "gptrs": [gbuf.address if r == self.rank else gbuf.open_peer(handles[r][0]) for r in range(W)],
                     "pptrs": [pbuf.address if r == self.rank else pbuf.open_peer(handles[r][1]) for r in range(W)]}
        ts._refresh()
        dist.barrier()                       # nobody steps before every rank has mapped every buffer

    def _make_barrier(self, dev):
        return make_rank_barrier(dev)

    def _teacher_logits(self, user, item):
        """Teacher forward for response KD (reference src/distillation/response.py:15-19), or None."""
        ts = self.ts
        if ts.teacher is None:
            return None, 1.0
        t = ts.teacher_logits[:user.numel()]
        ops.forward(ts._tm, user, item, out=t, workspace=ts.teacher_workspace)
        return t, ts.alpha

    # -- user-partitioned layout -------------------------------------------------------------------------------------
    def _partitioned_step(self, user, item, label, global_batch=None):
        """`user` must lie in this rank's range [user_lo, user_hi); `global_batch` = number of samples of
        the step over all ranks (default: every rank passes the same count)."""
        ts, W = self.ts, self.world
        b = user.numel()
        if b > ts.max_batch:
            raise ops._lib.NcfError(f"batch {b} exceeds max_batch {ts.max_batch}")
        B = int(global_batch) if global_batch is not None else b * W
        if B < b or B <= 0:
            raise ValueError("global_batch must be >= the local batch and positive")
        dense = ts.dense_adam(B)
        if ts._dirty or not dense:
            # lazy mode: the item rows of the GLOBAL batch and the own user rows catch up first
            if global_batch is None:
                gi = torch.empty(W * b, dtype=item.dtype, device=item.device)
                dist.all_gather_into_tensor(gi, item.contiguous())
            else:   # uneven local batches: pad to max_batch with -1 (ignored by the marker)
                pad = torch.full((ts.max_batch,), -1, dtype=item.dtype, device=item.device)
                pad[:b] = item
                gi = torch.empty(W * ts.max_batch, dtype=item.dtype, device=item.device)
                dist.all_gather_into_tensor(gi, pad)
            if b > 0:
                ops.mark_rows_side(ts._m, ts._g, user, 0)
            ops.mark_rows_side(ts._m, ts._g, gi, 1)
            ops.adam_catchup(ts._m, ts._g, ts._s, ts.lr, ts.betas[0], ts.betas[1], ts.eps)
        if b > 0:
            t_logits, alpha = self._teacher_logits(user, item)
            ops.train_step_grads_norm(ts._m, ts._g, user, item, label, B, ts.loss_accum, ts.workspace,
                                      teacher_logits=t_logits, alpha=alpha)
        if self.tail is not None:
            # reduce -> Adam -> broadcast of the replicated tail in ONE kernel over peer memory: the rank reads its
            # slice of every rank's gradients (NVLink loads), sums them (every rank divided by the global batch
            # already), steps its slice and stores the new parameters into every rank's buffer.  The two
            # one-element all-reduces are the rank barriers around it (stream-ordered on every rank).
            if not dense:
                raise ops._lib.NcfError("the peer-memory tail runs with the all-rows optimiser only "
                                        "(set NCF_DP_P2P_TAIL=0 for batches that touch few rows)")
            tl = self.tail
            tl["barrier"]()                      # every rank's gradients are complete
            ops.adam_p2p(tl["gptrs"], tl["pptrs"], tl["m"], tl["v"], tl["lo"], self.rank, ts.state.step, ts.lr,
                         ts.betas[0], ts.betas[1], ts.eps, grad_scale=1.0)
            tl["barrier"]()                      # every rank has read these gradients and written its parameters
            tl["g"].zero_()
            ops.adam_step_dense_range(ts._m, ts._g, ts._s, self.user_lo, self.user_hi, ts.lr, ts.betas[0],
                                      ts.betas[1], ts.eps, parts=ops.PART_USERS)
            ts._dirty = False
            ts.num_steps += 1
            return
        # every rank holds (sum over its samples) / B: the global-mean gradient is the plain sum
        flat = ts.grads.flat
        if self.comm_stream is not None and b > 0:
            ops.wait_embedding_grads(self.comm_stream)           # comm stream: after this rank's tower kernel
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(flat[self.n_user_flat:self.n_rows_flat], op=dist.ReduceOp.SUM)   # item rows
            dist.all_reduce(flat[self.n_rows_flat:], op=dist.ReduceOp.SUM)                       # tower, after wgrad
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        elif self.comm_stream is not None:    # a rank without samples this step: same two collectives, in the same order
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(flat[self.n_user_flat:self.n_rows_flat], op=dist.ReduceOp.SUM)
            dist.all_reduce(flat[self.n_rows_flat:], op=dist.ReduceOp.SUM)
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        else:
            dist.all_reduce(flat[self.n_user_flat:], op=dist.ReduceOp.SUM)
        if dense:
            ops.adam_step_dense_range(ts._m, ts._g, ts._s, self.user_lo, self.user_hi, ts.lr, ts.betas[0],
                                      ts.betas[1], ts.eps)
            ts._dirty = False
        else:
            ops.adam_step(ts._m, ts._g, ts._s, ts.lr, ts.betas[0], ts.betas[1], ts.eps)
            ts._dirty = True
        ts.num_steps += 1

    def global_loss(self) -> float:
        """Sum over steps of the global-batch mean loss since the last call (one all-reduce + one read)."""
        t = self.ts.loss_accum.clone()
        if self.partition_users:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)     # local parts are already divided by the global batch
        else:
            average_(t, self.world)
        self.ts.loss_accum.zero_()
        return float(t.item())

    def sync_user_tables(self) -> None:
        """Makes every rank hold every user row (checkpoint / full evaluation): each owner broadcasts its
        range.  A no-op in the fully replicated layout."""
        if not self.partition_users:
            return
        self.ts.flush()
        m = self.ts.model
        for table in (m.embed_user_GMF.weight, m.embed_user_MLP.weight):
            for r in range(self.world):
                lo, hi = user_range(m.user_num, self.world, r)
                if hi > lo:
                    dist.broadcast(table.data[lo:hi], src=r)

    # -- optimiser sharding ----------------------------------------------------------------------------------------
    def _setup_sharded(self):
        """Re-homes parameters and gradients into flat buffers with one common layout
        [user GMF | item GMF | user MLP | item MLP | tower], padded so that it splits evenly over the
        ranks; the Adam moments exist for the own slice only.  The all-reduce of the replicated step
        (reduce-scatter + all-gather inside NCCL) becomes reduce-scatter -> Adam on 1/world of the
        elements -> all-gather of the parameters: same bytes on the wire, 1/world of the optimiser."""
        ts, W = self.ts, self.world
        ts.flush()     # the flat copies below must see every row at the current step
        model, g, st = ts.model, ts.grads, ts.state
        old = g.flat
        n = old.numel()
        per = -(-n // (4 * W)) * 4
        n_pad = per * W
        dev = old.device
        pieces = [("g_user_gmf", model.embed_user_GMF.weight, "m_user_gmf", "v_user_gmf"),
                  ("g_item_gmf", model.embed_item_GMF.weight, "m_item_gmf", "v_item_gmf"),
                  ("g_user_mlp", model.embed_user_MLP.weight, "m_user_mlp", "v_user_mlp"),
                  ("g_item_mlp", model.embed_item_MLP.weight, "m_item_mlp", "v_item_mlp")]
        # NCF_DP_P2P=1 (experimental, one NVLink node): gradients and parameters live in buffers the other
        # ranks can map (CUDA IPC), and the exchange happens inside the optimiser kernel - see _sharded_step
        p2p = os.environ.get("NCF_DP_P2P") == "1" and 2 <= W <= 8
        if p2p:
            gbuf, pbuf = ops.PeerBuffer(n_pad, dev), ops.PeerBuffer(n_pad, dev)
            gflat, pflat = gbuf.tensor, pbuf.tensor
        else:
            gflat = torch.zeros(n_pad, dtype=torch.float32, device=dev)
            pflat = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        mflat = torch.zeros(n_pad, dtype=torch.float32, device=dev)   # staging only: sliced below
        vflat = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        base = old.data_ptr()
        for gname, param, mname, vname in pieces:
            gt = getattr(g, gname)
            if gt is None:
                continue
            off, cnt = (gt.data_ptr() - base) // 4, gt.numel()
            gflat[off:off + cnt].copy_(gt.reshape(-1))
            pflat[off:off + cnt].copy_(param.data.reshape(-1))
            mflat[off:off + cnt].copy_(getattr(st, mname).reshape(-1))
            vflat[off:off + cnt].copy_(getattr(st, vname).reshape(-1))
            setattr(g, gname, gflat[off:off + cnt].view_as(gt))
            param.data = pflat[off:off + cnt].view_as(param.data)
        toff = (g.g_tower.data_ptr() - base) // 4
        nt = g.g_tower.numel()
        gflat[toff:toff + nt].copy_(g.g_tower)
        mflat[toff:toff + nt].copy_(st.m_tower)
        vflat[toff:toff + nt].copy_(st.v_tower)
        g.g_tower = gflat[toff:toff + nt]
        o = toff
        tower_params = [q for lin in model.linears() for q in (lin.weight, lin.bias)]
        tower_params += [model.predict_layer.weight, model.predict_layer.bias]
        for q in tower_params:   # the flat tower order of the C ABI (ncf_tower_param_count)
            cnt = q.numel()
            pflat[o:o + cnt].copy_(q.data.reshape(-1))
            q.data = pflat[o:o + cnt].view_as(q.data)
            o += cnt
        assert o == toff + nt, "tower layout mismatch"
        g.flat = gflat
        lo = self.rank * per
        self.sharded = {"per": per, "lo": lo, "g": gflat, "p": pflat,
                        "m": mflat[lo:lo + per].clone(), "v": vflat[lo:lo + per].clone()}
        if p2p:
            handles = [None] * W
            dist.all_gather_object(handles, (gbuf.handle(), pbuf.handle()))
            self.sharded.update(
                bufs=(gbuf, pbuf), flag=torch.zeros(1, dtype=torch.float32, device=dev),
                gptrs=[gbuf.address if r == self.rank else gbuf.open_peer(handles[r][0]) for r in range(W)],
                pptrs=[pbuf.address if r == self.rank else pbuf.open_peer(handles[r][1]) for r in range(W)])
            dist.barrier()  # nobody starts stepping before every rank has mapped every buffer
        ts._refresh()

    def _sharded_step(self, user, item, label):
        ts, sh = self.ts, self.sharded
        assert not ts._dirty, "the sharded optimiser keeps every row current"
        t_logits, alpha = self._teacher_logits(user, item)
        ops.train_step_grads(ts._m, ts._g, user, item, label, t_logits, alpha, ts.loss_accum, ts.workspace)
        lo, per = sh["lo"], sh["per"]
        if "gptrs" in sh:
            # One kernel instead of reduce-scatter -> Adam -> all-gather: it reads the own slice of every
            # rank's gradients and writes the new parameters into every rank's buffer over NVLink.  The two
            # one-element all-reduces are the rank barriers around it (stream-ordered on every rank).
            dist.all_reduce(sh["flag"])      # every rank's gradients are complete
            ops.adam_p2p(sh["gptrs"], sh["pptrs"], sh["m"], sh["v"], lo, self.rank, ts.state.step, ts.lr,
                         ts.betas[0], ts.betas[1], ts.eps)
            dist.all_reduce(sh["flag"])      # every rank has read these gradients and written its parameters
            sh["g"].zero_()
            ops.adam_finish_dense(ts._m, ts._g, ts._s)
            ts._dirty = False
            ts.num_steps += 1
            return
        mine_g, mine_p = sh["g"][lo:lo + per], sh["p"][lo:lo + per]
        if dist.get_backend() == "nccl":
            dist.reduce_scatter_tensor(mine_g, sh["g"], op=dist.ReduceOp.AVG)
        else:
            dist.reduce_scatter_tensor(mine_g, sh["g"], op=dist.ReduceOp.SUM)
            mine_g.div_(self.world)
        ops.adam_range(mine_p, sh["m"], sh["v"], mine_g, ts.state.step, ts.lr, ts.betas[0], ts.betas[1], ts.eps)
        dist.all_gather_into_tensor(sh["p"], mine_p)
        sh["g"].zero_()                      # the other ranks' slices still hold this rank's local sums
        ops.adam_finish_dense(ts._m, ts._g, ts._s)
        ts._dirty = False
        ts.num_steps += 1

    def step(self, user, item, label, global_batch=None):
        ts = self.ts
        if self.partition_users:
            return self._partitioned_step(user, item, label, global_batch)
        if self.sharded is not None:   # elementwise Adam over every element: exact for any batch size
            return self._sharded_step(user, item, label)
        # a global batch that touches a large share of the tables runs the optimiser over all rows
        # (the dense Adam of the reference as it is): no index exchange, no catch-up, no row gather
        dense = ts.dense_adam(user.numel() * self.world)
        if ts._dirty or not dense:
            gu, gi = gather_indices(user, item, self.world)
            ops.adam_prepare(ts._m, ts._g, ts._s, gu, gi, ts.lr, ts.betas[0], ts.betas[1], ts.eps)
        t_logits, alpha = self._teacher_logits(user, item)
        ops.train_step_grads(ts._m, ts._g, user, item, label, t_logits, alpha, ts.loss_accum, ts.workspace)
        flat = ts.grads.flat
        if self.comm_stream is not None and dist.get_backend() == "nccl":
            ops.wait_embedding_grads(self.comm_stream)
            with torch.cuda.stream(self.comm_stream):
                average_(flat[:self.n_rows_flat], self.world)
            average_(flat[self.n_rows_flat:], self.world)      # tower gradients: after the wgrad kernel
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        else:
            average_(flat, self.world)
        if dense:
            ops.adam_step_dense(ts._m, ts._g, ts._s, ts.lr, ts.betas[0], ts.betas[1], ts.eps)
            ts._dirty = False
        else:
            ops.adam_step(ts._m, ts._g, ts._s, ts.lr, ts.betas[0], ts.betas[1], ts.eps)
            ts._dirty = True
        ts.num_steps += 1

    def replica_divergence(self) -> float:
        """max over parameters and ranks of |w_rank - w_0| (0.0 when the replicas are identical); in the
        user-partitioned layout the user tables are synchronised first (their owners' rows win)."""
        self.ts.flush()
        self.sync_user_tables()
        worst = torch.zeros(1, device=self.ts.device)
        for p in self.ts.model.parameters():
            ref = p.data.clone()
            dist.broadcast(ref, src=0)
            worst = torch.maximum(worst, (p.data - ref).abs().max().reshape(1))
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        return float(worst.item())


# ---- row-sharded tables (BASELINE config 5) ------------------------------------------------------------------
def shard_rows(n_rows: int, world: int, rank: int) -> int:
    """Number of rows r in [0, n_rows) with r % world == rank."""
    return (n_rows - rank + world - 1) // world


def shard_state_dict(full: dict, world: int, rank: int) -> dict:
    """The slice of a full-model state_dict that lives on `rank`: table row r -> rank r % world, local
    index r // world; the tower is replicated."""
    out = {}
    for k, v in full.items():
        out[k] = v[rank::world].clone() if k.startswith("embed_") else v.clone()
    return out


class RowShardedTrainer:
    """Training with the four embedding tables (and their Adam state) row-sharded over the ranks.

    Each rank trains on samples of ITS OWN users (sharding follows the data), so user rows are
    local and only item rows travel.  Two transports for the exchange:

    **peer memory (default on one node, world <= 8; NCF_SHARD_P2P=0 disables).**  The ranks map each
    other's buffers (CUDA IPC) and the kernels of csrc/shard.cu write into them over NVLink:
      1. ncf_shard_request: one request (local row, slot) per sample into the owner's inbox;
      2. barrier; owners register the requested rows + their own users (ncf_shard_mark_requests,
         ncf_mark_rows_side), replay pending Adam steps (dense-equivalence) and PUSH the rows into the
         requesters' receive buffers at the named slots (ncf_shard_push_rows);
      3. barrier; fused forward+loss+backward on [local user tables | received item rows], loss mean
         over the GLOBAL batch (ncf_train_step_grads_norm); ncf_shard_push_grads adds every per-sample
         item-row gradient straight into the owner's gradient table (red.sys over NVLink);
      4. barrier; all-reduce (sum) of the tower gradients; sparse-row Adam on every rank's shard.
    The barriers are stream-ordered one-element all-reduces; no split size ever reaches the host, so the
    step has no host synchronisation and nothing is permuted (slot = position in the batch).

    **NCCL all-to-all** (any world size): bucket by owner -> all-to-all of the counts (one host read) ->
    all-to-all #1 indices, #2 rows, #3 row gradients (GMF and MLP parts separately).

    The result equals single-process training at the global batch (tests/shard_worker.py)."""

    def __init__(self, shard_model, user_num: int, item_num: int, lr: float = 1e-3,
                 betas=(0.9, 0.999), eps: float = 1e-8, max_batch: int = 65536, p2p=None):
        from .trainer import FusedTrainStep
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.user_num, self.item_num = int(user_num), int(item_num)
        self.model = shard_model
        if shard_model.user_num != shard_rows(user_num, self.world, self.rank) or \
                shard_model.item_num != shard_rows(item_num, self.world, self.rank):
            raise ValueError("shard_model must hold exactly this rank's rows")
        # the shard's optimiser state, gradient buffers and lists (items: up to world * max_batch requests)
        self.ts = FusedTrainStep(shard_model, "adam", lr, betas, eps, max_batch=max_batch)
        dev = self.ts.device
        self.ts.grads.item_list = torch.zeros(min(max_batch * self.world, shard_model.item_num),
                                              dtype=torch.int64, device=dev)
        self.ts._refresh()
        self.max_batch = max_batch
        self.f, self.d = shard_model.factor_num, shard_model.factor_num << (shard_model.num_layers - 1)
        self.loss_accum = torch.zeros(1, dtype=torch.float64, device=dev)
        for name in ("MLP_layers", "predict_layer"):  # identical tower everywhere
            for p in getattr(shard_model, name).parameters():
                dist.broadcast(p.data, src=0)
        if p2p is None:
            p2p = os.environ.get("NCF_SHARD_P2P", "1") != "0"
        self.p2p = bool(p2p) and 2 <= self.world <= 8 and dev.type == "cuda" and dist.get_backend() == "nccl" \
            and (self.f % 4 == 0) and shard_model.model_type != "GMF" and shard_model.model_type != "MLP"
        # per-sample item-row gradients of the fused step (both transports)
        self._g_gmf = torch.zeros(max_batch, self.f, dtype=torch.float32, device=dev)
        self._g_mlp = torch.zeros(max_batch, self.d, dtype=torch.float32, device=dev)
        self._item_pos = torch.arange(max_batch, dtype=torch.int64, device=dev)
        if self.p2p:
            self._setup_p2p()

    # -- peer-memory transport ---------------------------------------------------------------------------------------
    def _setup_p2p(self):
        ts, W, dev, cap = self.ts, self.world, self.ts.device, self.max_batch
        m, g = self.model, ts.grads
        self._bufs = {
            "inbox": ops.PeerBuffer(W * cap, dev, torch.int64),
            "count": ops.PeerBuffer(max(W, 4), dev, torch.int32),
            "rows_gmf": ops.PeerBuffer(cap * self.f, dev),
            "rows_mlp": ops.PeerBuffer(cap * self.d, dev),
            # the shard's item-gradient tables must be writable by the peers: re-home them
            "g_item_gmf": ops.PeerBuffer(m.item_num * self.f, dev),
            "g_item_mlp": ops.PeerBuffer(m.item_num * self.d, dev),
        }
        g.g_item_gmf = self._bufs["g_item_gmf"].tensor.view(m.item_num, self.f)
        g.g_item_mlp = self._bufs["g_item_mlp"].tensor.view(m.item_num, self.d)
        g.flat = None                     # no longer one allocation
        ts._refresh()
        handles = [None] * W
        dist.all_gather_object(handles, {k: b.handle() for k, b in self._bufs.items()})
        self._peer = {k: [b.address if r == self.rank else b.open_peer(handles[r][k]) for r in range(W)]
                      for k, b in self._bufs.items()}
        self._rows_gmf = self._bufs["rows_gmf"].tensor.view(cap, self.f)
        self._rows_mlp = self._bufs["rows_mlp"].tensor.view(cap, self.d)
        self._cursor = torch.zeros(max(W, 4), dtype=torch.int32, device=dev)
        self._barrier = make_rank_barrier(dev)   # stream-ordered: every rank's preceding kernels are complete
        dist.barrier()                    # nobody steps before every rank has mapped every buffer

    def _step_p2p(self, user, item, label, B_global):
        ts, W, m = self.ts, self.world, self.model
        b = user.numel()
        cap = self.max_batch
        inbox, count = self._bufs["inbox"].tensor, self._bufs["count"].tensor
        ops.shard_request(item, W, self.rank, cap, self.item_num, self._peer["inbox"], self._peer["count"],
                          self._cursor)
        self._barrier()
        local_user = (user // W).contiguous()
        if b > 0:
            ops.mark_rows_side(ts._m, ts._g, local_user, 0)
        ops.shard_mark_requests(ts._m, ts._g, inbox, count, W, cap)
        ops.adam_catchup(ts._m, ts._g, ts._s, ts.lr, ts.betas[0], ts.betas[1], ts.eps)
        ops.shard_push_rows(ts._m, inbox, count, W, cap, self._peer["rows_gmf"], self._peer["rows_mlp"])
        self._barrier()
        if b > 0:
            g_gmf, g_mlp = self._g_gmf[:b], self._g_mlp[:b]
            g_gmf.zero_()
            g_mlp.zero_()
            cm = ops.model_struct(m.abi_type(), m.factor_num, m.num_layers, m.user_num, b,
                                  (m.embed_user_GMF.weight.detach(), self._rows_gmf, m.embed_user_MLP.weight.detach(),
                                   self._rows_mlp),
                                  [(l.weight.detach(), l.bias.detach()) for l in m.linears()],
                                  (m.predict_layer.weight.detach(), m.predict_layer.bias.detach()),
                                  tower_math=m.tower_math)
            cg = ts.grads.struct()
            cg.g_item_gmf, cg.g_item_mlp = ops.ptr(g_gmf), ops.ptr(g_mlp)
            ops.train_step_grads_norm(cm, cg, local_user, self._item_pos[:b], label, B_global, self.loss_accum,
                                      ts.workspace)
            ops.shard_push_grads(item, W, self.item_num, g_gmf, g_mlp, self.f, self.d, self._peer["g_item_gmf"],
                                 self._peer["g_item_mlp"])
        # the tower all-reduce doubles as the barrier "every rank has pushed its row gradients"
        dist.all_reduce(ts.grads.g_tower, op=dist.ReduceOp.SUM)
        ops.adam_step(ts._m, ts._g, ts._s, ts.lr, ts.betas[0], ts.betas[1], ts.eps)
        ts._dirty = True
        ts.num_steps += 1

    def wire_stats(self) -> dict:
        """Bytes this rank sends over NVLink per step of `max_batch` samples (uniform item owners)."""
        W, b = self.world, self.max_batch
        away = (W - 1) / W
        row = 4 * (self.f + self.d)
        return {"transport": "peer memory (CUDA IPC, stores / REDs over NVLink)" if self.p2p else "NCCL all-to-all",
                "bytes_out_per_step": int(away * b * (8 + 2 * row)),
                "what": "requests (8 B) + item rows pushed as owner + item-row gradients pushed as requester, per sample, "
                        "times the share of samples whose item lives on another rank"}

    def _a2a(self, send: torch.Tensor, send_counts, recv_counts, row: int = 1) -> torch.Tensor:
        out = torch.empty((sum(recv_counts),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(out, send, output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts))
        return out

    def step(self, user: torch.Tensor, item: torch.Tensor, label: torch.Tensor, global_batch=None) -> None:
        """user: GLOBAL ids owned by this rank; item: GLOBAL ids; label: float32.  `global_batch`: samples
        of this step over all ranks (the loss mean runs over it); default = every rank passes the same
        count.  A rank without samples still has to call step (it owns rows the others need)."""
        b = user.numel()
        if b > self.max_batch:      # before any communication: every rank checks its own input
            raise ops._lib.NcfError(f"batch {b} exceeds max_batch {self.max_batch}")
        if self.p2p:
            B_global = int(global_batch) if global_batch is not None else b * self.world
            if B_global < max(b, 1):
                raise ValueError("global_batch must be >= the local batch and positive")
            return self._step_p2p(user, item, label, B_global)
        return self._step_nccl(user, item, label, global_batch)

    def _step_nccl(self, user, item, label, global_batch=None) -> None:
        ts, W, dev = self.ts, self.world, self.ts.device
        b = user.numel()
        # global batch size (for the loss mean) and the item-owner buckets
        perm, local_item, counts = ops.bucket_by_owner(item, W)
        meta = torch.cat([counts.to(torch.int64), torch.tensor([b], dtype=torch.int64, device=dev)])
        all_meta = torch.empty(W, W + 1, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_meta.view(-1), meta)
        all_meta = all_meta.cpu()                                # the one host sync of the step
        send_counts = all_meta[self.rank, :W].tolist()            # my samples per owner
        recv_counts = all_meta[:, self.rank].tolist()             # requests I receive per rank
        B_global = int(all_meta[:, W].sum()) if global_batch is None else int(global_batch)
        # 1-2. ask the owners
        req = self._a2a(local_item, send_counts, recv_counts)     # local item indices requested from me
        local_user = (user // W).contiguous()
        if b > 0:
            ops.mark_rows_side(ts._m, ts._g, local_user, 0)
        ops.mark_rows_side(ts._m, ts._g, req, 1)
        ops.adam_catchup(ts._m, ts._g, ts._s, ts.lr, ts.betas[0], ts.betas[1], ts.eps)
        m = self.model
        rows_gmf = ops.gather_rows(m.embed_item_GMF.weight.detach(), req)
        rows_mlp = ops.gather_rows(m.embed_item_MLP.weight.detach(), req)
        # 3. rows come back in my bucket order
        got_gmf = self._a2a(rows_gmf, recv_counts, send_counts)
        got_mlp = self._a2a(rows_mlp, recv_counts, send_counts)
        # 4. fused step on [my user shard | received item rows]; a rank without samples only serves rows
        g_gmf, g_mlp = self._g_gmf[:b], self._g_mlp[:b]
        if b > 0:
            u_sorted = ops.permute(local_user, perm)
            y_sorted = ops.permute(label, perm)
            g_gmf.zero_()
            g_mlp.zero_()
            cm = ops.model_struct(m.abi_type(), m.factor_num, m.num_layers, m.user_num, b,
                                  (m.embed_user_GMF.weight.detach(), got_gmf, m.embed_user_MLP.weight.detach(), got_mlp),
                                  [(l.weight.detach(), l.bias.detach()) for l in m.linears()],
                                  (m.predict_layer.weight.detach(), m.predict_layer.bias.detach()),
                                  tower_math=m.tower_math)
            cg = ts.grads.struct()
            cg.g_item_gmf, cg.g_item_mlp = ops.ptr(g_gmf), ops.ptr(g_mlp)
            ops.train_step_grads_norm(cm, cg, u_sorted, self._item_pos[:b], y_sorted, B_global, self.loss_accum,
                                      ts.workspace)
        # 5. item-row gradients go home
        back_gmf = self._a2a(g_gmf, send_counts, recv_counts)
        back_mlp = self._a2a(g_mlp, send_counts, recv_counts)
        ops.scatter_add_rows(ts.grads.g_item_gmf, req, back_gmf)
        ops.scatter_add_rows(ts.grads.g_item_mlp, req, back_mlp)
        # 6. tower gradients: every rank holds (local sum) / B_global -> sum over ranks
        dist.all_reduce(ts.grads.g_tower, op=dist.ReduceOp.SUM)
        ops.adam_step(ts._m, ts._g, ts._s, ts.lr, ts.betas[0], ts.betas[1], ts.eps)
        ts._dirty = True
        ts.num_steps += 1

    def flush(self):
        self.ts.flush()

    def gather_full_state(self) -> dict:
        """Reassembles the full-model state_dict on every rank (tests / checkpoints)."""
        self.flush()
        out = {}
        for k, v in self.model.state_dict().items():
            if not k.startswith("embed_"):
                out[k] = v.clone()
                continue
            n_rows = self.user_num if "user" in k else self.item_num
            per = shard_rows(n_rows, self.world, 0)
            pad = torch.zeros(per, v.shape[1], dtype=v.dtype, device=v.device)
            pad[: v.shape[0]] = v
            parts = [torch.empty_like(pad) for _ in range(self.world)]
            dist.all_gather(parts, pad)
            full = torch.empty(n_rows, v.shape[1], dtype=v.dtype, device=v.device)
            for r in range(self.world):
                full[r::self.world] = parts[r][: shard_rows(n_rows, self.world, r)]
            out[k] = full
        return out
