"""Whole-step CUDA-graph execution of the cluster-batch training step.

A Reddit cluster batch is ~2 k nodes, so the reference's step (and this repo's
eager step) is bound by kernel launches and Python, not by the GPU.  Here ONE
training step — device batch build (K3) -> L+1 SAGE layers forward (K1 + GEMM)
-> cross entropy -> backward (K2 + GEMMs) -> Adam — is captured once into a CUDA
graph and replayed per step.  Batches have different node / edge counts, so the
captured shapes are the maxima over any grouping of parts and every batch is
padded with the builder's `-1` sentinel: pad rows are isolated, have zero
features and a False train mask, hence contribute exactly 0 to the loss and to
every gradient.  Results equal the eager step's (tests/test_gpu_graphed.py).

Pipelining (``pipeline=True``, the default).  Three things overlap the training kernels of
step k instead of sitting between steps:
  * the batch build of step k+1 (mark / count / scan / fill / ndata gathers — latency-bound
    walks over hub rows that leave most SMs idle) is captured on a second stream as a parallel
    branch of the graph that trains step k.  Two graphs alternate: graph j trains on cluster
    buffer set j while building into set 1-j;
  * the host->device copy of the ids of batch k+1 (issued while graph k-1 is still running) and
  * the device->host copy of step k's loss run on a copy stream, fenced with events; node-id
    and loss buffers are double-buffered with the graphs (graph j reads ``nids[j]``, writes
    ``loss[j]``), so no copy ever touches a buffer a running graph uses.
The Python ``random`` call order (epoch-end shuffle, create_partition) is unchanged: the shuffle
for the next epoch still happens after the dispatch of the epoch's last step and before the next
dispatch (cluster_gcn/sampler.py:92, cluster_gcn_ist_distrib.py:398-404).

Callers: bench.py and the trainers; mirrors cluster_gcn_ist_distrib.py:398-417.
"""
import os

import torch

from . import ops
from .modules import GCN as _SageGCN, SAGE_PRE0
from .train import loss_and_backward, make_optimizer

_KEYS = ('feat', 'label', 'train_mask')
# resident CTAs per SM the next batch's layer-0 aggregation may take while the current batch trains
# (0 = no cap); A/B switches for the measurements in profiles/
PREP_CTAS_PER_SM = int(os.environ.get('GIST_PREP_CTAS', '0'))
TRAIN_FIRST = os.environ.get('GIST_TRAIN_FIRST', '0') != '0'
PREP_BALANCED = os.environ.get('GIST_PREP_BALANCED', '0') != '0'
# Fused tail of the step (pipelined 3xTF32 trainer): the Adam launch also writes the weights' 3xTF32 low
# halves and ticks the dropout clock, the cross entropy writes the loss into the trainer's slot — the
# step's dependency chain loses the loss copy, the clock tick and the weight split (adam -> copy -> tick
# -> [next graph] split -> first GEMM: 17.6 us from the end of Adam to the start of the first GEMM in
# the round-2 timeline, of which 2.4 us is the split's own work).
FUSED_TAIL = os.environ.get('GIST_FUSED_TAIL', '1') != '0'
STAGE_IDS_AHEAD = os.environ.get('GIST_STAGE_IDS_AHEAD', '1') != '0'


class GraphedClusterTrainer:
    def __init__(self, cluster_iter, model, lr, weight_decay, h2d='epoch', pipeline=True):
        assert h2d in ('epoch', 'step')
        self.it, self.model, self.h2d = cluster_iter, model, h2d
        self.pipeline = bool(pipeline)
        self.pre_aggregate = self.pipeline and isinstance(model, _SageGCN)
        g = cluster_iter.g
        self.dev = g.device
        self.n_pad = cluster_iter.max_batch_nodes()
        self.cap = max(cluster_iter.max_batch_edges(), 1)
        nbuf = 2 if self.pipeline else 1
        self.nids = [torch.full((self.n_pad,), -1, dtype=torch.int64, device=self.dev) for _ in range(nbuf)]
        self._loss2 = [torch.zeros(2, dtype=torch.float32, device=self.dev) for _ in range(nbuf)]   # (loss, 1 / #rows)
        self.loss = [b[0] for b in self._loss2]
        self.opt = make_optimizer(model.parameters(), lr, weight_decay)
        self.fused_tail = (FUSED_TAIL and self.pipeline and not TRAIN_FIRST
                           and ops.get_matmul_precision() == '3xtf32')
        if self.fused_tail:
            # one persistent low-half buffer per TMA-addressable weight, written by the Adam launch
            for p in model.parameters():
                if p.dim() == 2 and p.numel() > 0 and p.is_contiguous() and ops._tma_ok(p):
                    lo = torch.empty_like(p, memory_format=torch.contiguous_format).detach()
                    self.opt.lo_map[p] = ops.register_persistent_lo(p, lo)
        self.graphs = None
        self.clusters = [None, None]    # pipelined: the two cluster buffer sets
        self.k = 0                      # steps issued so far
        self.i = 0                      # next row of the epoch's id table to stage
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.replays = 0
        self.gist_launches_per_step = 0
        # persistent staging: two pinned [steps, n_pad] id tables (alternating per epoch, so a row
        # still being copied is never overwritten) and, for h2d='epoch', two device tables
        steps = len(cluster_iter)
        self._host_tab = [torch.empty((steps, self.n_pad), dtype=torch.int64).pin_memory() for _ in range(2)]
        self._dev_tab = ([torch.empty((steps, self.n_pad), dtype=torch.int64, device=self.dev) for _ in range(2)]
                         if h2d == 'epoch' else None)
        self._tab = 1
        # last asynchronous reader of each pinned table (the whole-table upload, or the latest row
        # copy): the host must not refill a table before that copy has executed — step() never
        # blocks the host, so it can run several epochs ahead of the GPU when epochs are short
        self._tab_read = [None, None]
        # two copy streams: uploads must never queue behind a loss readback, which waits for the
        # running graph to finish (one shared stream delayed the ids of batch k+1 — and with them
        # graph k+1 — until graph k had completed and its loss had been copied out)
        self._copy = torch.cuda.Stream(device=self.dev)
        self._copy_out = torch.cuda.Stream(device=self.dev)
        self._ev_ids = [torch.cuda.Event() for _ in range(nbuf)]      # ids landed in nids[j]
        self._ev_done = [torch.cuda.Event() for _ in range(nbuf)]     # graph j finished
        self._ev_ring = [torch.cuda.Event() for _ in range(nbuf)]     # loss[j] landed in the host ring
        self._ring = torch.zeros(nbuf, dtype=torch.float32).pin_memory()
        self._ring_busy = [False] * nbuf
        self._pending = None
        self._staged = [False, False]   # ids for the next use of nids[j] are already on their way
        g.is_symmetric()                # decided once, outside capture (it syncs)
        self._load_epoch()

    # ------------------------------------------------------------------ data --
    def _load_epoch(self):
        """Fill the other id table with the (re)shuffled epoch; h2d='epoch' uploads it whole."""
        self._tab ^= 1
        host = self._host_tab[self._tab]
        if self._tab_read[self._tab] is not None:
            self._tab_read[self._tab].synchronize()      # its previous contents have been consumed
        self.it.padded_epoch_ids(self.n_pad, out=host.numpy())
        if self.h2d == 'epoch':
            with torch.cuda.stream(self._copy):
                self._dev_tab[self._tab].copy_(host, non_blocking=True)
                self._mark_table_read()
            self.h2d_bytes += host.numel() * 8
        self.i = 0

    def _mark_table_read(self):
        """Record, on the copy stream, that every copy issued so far from the current pinned table
        has been enqueued (event reused per table: only the latest record matters)."""
        ev = self._tab_read[self._tab]
        if ev is None:
            ev = self._tab_read[self._tab] = torch.cuda.Event()
        ev.record(self._copy)

    def _next_row(self):
        if self.i >= len(self.it):
            self.it.end_epoch()
            self._load_epoch()
        row = (self._dev_tab if self.h2d == 'epoch' else self._host_tab)[self._tab][self.i]
        self.i += 1
        return row

    def _upload(self, j, after=None):
        """Copy the next batch's ids into nids[j] on the copy stream (after event `after`)."""
        row = self._next_row()
        with torch.cuda.stream(self._copy):
            if after is not None:
                self._copy.wait_event(after)
            self.nids[j].copy_(row, non_blocking=True)
            self._ev_ids[j].record(self._copy)
            if self.h2d == 'step':
                self._mark_table_read()                  # the row lives in pinned host memory
        if self.h2d == 'step':
            self.h2d_bytes += self.n_pad * 8

    # ------------------------------------------------------------------ step --
    def _build(self, nids, out=None):
        prev = out._cache.get(SAGE_PRE0) if out is not None else None
        sg = self.it.g.subgraph(nids, col_capacity=self.cap, ndata_keys=_KEYS, out=out, walk_capacity=self.cap)
        sg.seg_schedule()                   # segment schedule of the batch: built here, used by every SpMM
        sg.seg_schedule(transpose=True)
        if self.pre_aggregate:
            # layer 0 aggregates raw input features: no dependence on the weights, no gradient,
            # so it belongs to the batch-preparation branch (what the reference's use_pp idea is
            # after, but computed per batch on the batch subgraph: identical arithmetic), together
            # with that layer's dropout and 3xTF32 split, which K1 applies as it writes z
            # (row-per-warp kernel here: in the shadow of the training branch the scheduled kernels'
            # resident-CTA grids only compete with it — measured 0.281 vs 0.299 ms/step for the segment
            # kernel in round 1, 0.315 vs 0.340 for the shared-memory slab kernel in background mode in
            # round 2; GIST_PREP_BALANCED=1 opts in for A/B runs)
            sg._cache[SAGE_PRE0] = self.model.layers[0].prepare_input(sg, sg.ndata['feat'], out=prev,
                                                                     balanced=PREP_BALANCED,
                                                                     background=PREP_CTAS_PER_SM or (1 if PREP_BALANCED else 0))
        return sg

    def _train(self, cluster, j, before_update=None):
        """One training step on ``cluster``; the loss lands in loss[j] (written by the cross-entropy
        kernel itself).  ``before_update``: called between the backward pass and the optimizer."""
        self.opt.zero_grad(set_to_none=True)
        pred = self.model(cluster)          # (the dropout clock is ticked by the caller: auto_tick off)
        loss_and_backward(pred, cluster.ndata['label'], cluster.ndata['train_mask'], loss_out=self._loss2[j])
        if before_update is not None:
            before_update()
        self.opt.step()

    def _sync_lo(self):
        """Persistent weight low halves valid before the next step runs (a GIST dispatch, a restore or a
        load_state_dict wrote the weights since the last one): one split launch on the current stream."""
        if self.fused_tail and not ops.persistent_lo_valid():
            ops.refresh_persistent_lo()

    def capture(self):
        """Warm up on a real batch, capture, then restore parameters / optimizer state so the
        capture itself does not advance training."""
        from . import _lib
        params = list(self.model.parameters())
        saved = [p.detach().clone() for p in params]
        self.model.train()
        # the dropout clock (ops.DropoutState) is advanced by one node per captured step, outside
        # the region where the build / train branches run, so both read a stable value
        clock = ops.dropout_state(self.dev)
        auto_tick, clock.auto_tick = clock.auto_tick, False
        try:
            self._capture(clock, _lib)
        finally:
            clock.auto_tick = auto_tick
        self.k = 0
        self._staged = [False, False]
        with torch.no_grad():
            for p, q in zip(params, saved):
                p.copy_(q)
        self.reset_optimizer()
        return self

    def _capture(self, clock, _lib):
        main = torch.cuda.current_stream(self.dev)
        self._upload(0)                                  # batch 0
        main.wait_event(self._ev_ids[0])
        # the training branch is captured on a high-priority stream and the preparation
        # branch on a low-priority one: kernel nodes inherit the priority, so when both have
        # CTAs pending (the d = 602 aggregation of the next batch floods the chip) the block
        # scheduler serves the critical path first and the preparation fills the gaps.
        # The warm-up runs on the SAME stream as the capture: per-stream state the kernels keep
        # zeroed (queue heads, arrival counters) is then created here, eagerly, and not as fill
        # nodes at the head of every replayed step.
        lo_p, hi_p = torch.cuda.Stream.priority_range()
        s = torch.cuda.Stream(device=self.dev, priority=hi_p) if self.pipeline else torch.cuda.Stream(device=self.dev)
        s.wait_stream(main)
        with torch.cuda.stream(s):
            self._sync_lo()
            for _ in range(3):
                clock.tick()
                self._train(self._build(self.nids[0]), 0)
        main.wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self.graphs = []
        if self.pipeline:
            # prologue: batch 0 is built eagerly into buffer set 0; graph j trains on set j and
            # builds the NEXT batch (ids staged in nids[j] before the replay) into set 1 - j.
            # The dropout clock is advanced by the LAST node of every step (after the branches
            # join), for the next one: both branches are then roots of the graph and start with
            # the launch — with the tick as their common parent the training branch's first kernel
            # started ~8 us late at every replay (cross-branch dependency resolution).
            clock.tick()
            self.clusters[0] = self._build(self.nids[0].clone())
            clock.tick()                                 # the value step 0 runs with
            torch.cuda.synchronize(self.dev)
            # Each graph captures into its OWN memory pool.  Persistent state is created inside the
            # captures (buffer set 1, its segment schedule, counters, prepared layer-0 input); in a
            # shared pool the second capture would place such tensors on memory the first graph
            # still uses for its temporaries at every replay.
            train_stream = s
            for j in (0, 1):
                gph = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream(device=self.dev, priority=lo_p)
                l0 = _lib.launch_count()
                with torch.cuda.graph(gph, stream=train_stream):
                    cap_main = torch.cuda.current_stream(self.dev)
                    side.wait_stream(cap_main)

                    def prepare():
                        with torch.cuda.stream(side):
                            self.clusters[1 - j] = self._build(self.nids[j], out=self.clusters[1 - j])
                    # the training branch's nodes are created first: a replay hands its nodes to
                    # the GPU in creation order, a few microseconds apart at the head of the graph
                    if not TRAIN_FIRST:
                        prepare()
                    if self.fused_tail:
                        # the preparation branch (which reads the clock) joins BEFORE the optimizer, whose
                        # launch then ticks the clock: the step ends with the Adam node
                        self.opt.tick = clock.step
                        try:
                            self._train(self.clusters[j], j, before_update=lambda: cap_main.wait_stream(side))
                        finally:
                            self.opt.tick = None
                    else:
                        self._train(self.clusters[j], j)
                        if TRAIN_FIRST:
                            prepare()
                        cap_main.wait_stream(side)
                        clock.tick()
                self.gist_launches_per_step = _lib.launch_count() - l0
                self.graphs.append(gph)
        else:
            l0 = _lib.launch_count()
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                clock.tick()
                self._train(self._build(self.nids[0]), 0)
            self.gist_launches_per_step = _lib.launch_count() - l0   # this library's kernel nodes per replay
            self.graphs.append(gph)

    def reset_optimizer(self):
        """A fresh Adam, as the reference builds at every dispatch (…distrib.py:405-407), but in
        place: moments and step counters are zeroed so the captured graph's pointers stay valid."""
        self.opt.reset_state()
        if self.fused_tail:
            ops.refresh_persistent_lo()     # a dispatch may have written the weights through p.data

    def _issue(self):
        """Enqueue step k; returns the buffer index j whose loss[j] the step writes."""
        if self.graphs is None:
            self.capture()
        main = torch.cuda.current_stream(self.dev)
        self._sync_lo()
        if not self.pipeline:
            if self.k > 0:
                self._upload(0, after=self._ev_done[0])  # graph k-1 was the last reader of nids[0]
            main.wait_event(self._ev_ids[0])
            if self._ring_busy[0]:
                main.wait_event(self._ev_ring[0])        # loss[0] of step k-1 has been copied out
            self.graphs[0].replay()
            self._ev_done[0].record(main)
            j = 0
        else:
            j = self.k & 1
            # ids of batch k+1 (built during this step) go into nids[j], last read by graph k-2.
            # Normally they were staged one step ago (below); otherwise — first steps, first row of an
            # epoch — the copy is issued here.  (Fetching an epoch's FIRST row here, after any dispatch of
            # step k and before the next one, keeps the epoch-end shuffle in the reference's position in
            # the Python random stream.)
            if not self._staged[j]:
                self._upload(j, after=self._ev_done[j] if self.k >= 2 else None)
            self._staged[j] = False
            main.wait_event(self._ev_ids[j])
            if self._ring_busy[j]:
                main.wait_event(self._ev_ring[j])        # loss[j] of step k-2 has been copied out
            self.graphs[j].replay()
            self._ev_done[j].record(main)
            # Stage the ids graph k+1 will read (nids[1-j], last read by graph k-1) NOW, so the copy runs
            # while graph k does: issued at the head of step k+1 instead, its few microseconds sit between
            # two graphs, because a replay waits for its ids as a whole.  Rows of the current epoch only:
            # they are already in the id table, no shuffle is pulled forward.
            if STAGE_IDS_AHEAD and self.i < len(self.it):
                self._upload(1 - j, after=self._ev_done[1 - j] if self.k >= 1 else None)
                self._staged[1 - j] = True
        self.k += 1
        self.replays += 1
        return j

    def step(self):
        """One training step; returns the 0-d device loss tensor (valid until the next step but
        one is issued)."""
        return self.loss[self._issue()]

    # -------------------------------------------------------------- readback --
    def step_logged(self):
        """step() plus a device->host read of the step's loss EVERY step, without stalling the
        launch pipeline: the loss is copied into a pinned ring on the copy stream right behind
        the replay and the value returned is the PREVIOUS step's (None on the first call), whose
        copy has finished by now.  The reference's ``float(loss)`` (…distrib.py:416) feeds a
        running-loss log only, so a one-step lag changes nothing it computes; call ``drain()``
        after the last step for the final value."""
        j = self._issue()
        with torch.cuda.stream(self._copy_out):
            self._copy_out.wait_event(self._ev_done[j])
            self._ring[j:j + 1].copy_(self.loss[j].reshape(1), non_blocking=True)
            self._ev_ring[j].record(self._copy_out)
        self._ring_busy[j] = True
        self.d2h_bytes += 4
        if not self.pipeline:           # single buffer set: synchronous, like float(loss)
            self._ev_ring[j].synchronize()
            return float(self._ring[j])
        prev, self._pending = self._pending, j
        if prev is None:
            return None
        self._ev_ring[prev].synchronize()
        return float(self._ring[prev])

    def drain(self):
        """Loss of the last step issued through step_logged() (None if already consumed)."""
        if self._pending is None:
            return None
        self._ev_ring[self._pending].synchronize()
        v = float(self._ring[self._pending])
        self._pending = None
        return v
