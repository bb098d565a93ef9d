"""Thin Python wrappers + autograd Functions over the C-ABI kernels.

Every function here enqueues hand-written sm_100a kernels from
``csrc/libgist_b200.so`` on torch's current stream.  No CPU path exists:
non-CUDA inputs raise ``GistLibraryError``.
"""
import ctypes
import weakref
import os

import torch

from . import _lib
from ._lib import check, ptr, require_cuda, stream_ptr


# bench.py sets this to a list to time every SpMM launch with a CUDA-event pair on the
# launching stream (roofline.achieved); None = no instrumentation.
SPMM_PROFILE = None
# same for every tensor-core GEMM launch (tensor roofline of the ultra-wide config)
GEMM_PROFILE = None
GAT_PROFILE = None     # bench.py: list of dicts (event pair, n, D, rowptr) per K6 forward aggregation


def _mat(t, name):
    if t.dim() != 2 or t.dtype != torch.float32:
        raise ValueError('%s must be a 2-D float32 tensor, got %s %s' % (name, tuple(t.shape), t.dtype))
    if t.shape[1] > 1 and t.stride(1) != 1:
        t = t.contiguous()
    if t.shape[0] > 1 and t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t


def _ld(t):
    return t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))


def spmm_raw(rowptr, col, n_dst, n_src, X, out, *, src_scale=None, dst_scale=None, bias=None,
             addend=None, self_out=None, relu=False, flags=0, out_lo=None, self_lo=None, drop=None,
             drop_col0_out=0, drop_col0_self=0, schedule=None):
    """out[v] = act(dst_scale[v] * sum_{e in row v} src_scale[col e] * X[col e] + addend[v] + bias).

    ``out`` (and ``self_out``/``addend``) may be column-block views of wider
    row-major buffers; leading dimensions are taken from the strides.
    Extended epilogue: ``drop`` (a DropoutDesc) applies dropout to both outputs as they are
    written, ``out_lo`` / ``self_lo`` receive their 3xTF32 low halves.  ``schedule`` (a
    SegSchedule of the same rowptr, see GistGraph.seg_schedule) selects the segment-balanced
    kernel used for cluster batches."""
    require_cuda(rowptr, col, X, out, src_scale, dst_scale, bias, addend, self_out, out_lo, self_lo)
    d = X.shape[1]
    assert out.shape[0] == n_dst and out.shape[1] == d and X.shape[0] == n_src
    assert out.stride(1) == 1 or d == 1
    lib = _lib.load()
    f = flags | (_lib.SPMM_RELU if relu else 0) | (0 if SLAB_ENABLED else _lib.SPMM_SLAB_OFF)
    prof = SPMM_PROFILE
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    ex = None
    if schedule is not None and d > 384 and not (SLAB_ENABLED and not (flags & _lib.SPMM_SLAB_OFF) and _slab_ok(X, n_src)):
        schedule = None         # wide rows on the L2-gather segment kernel: per-segment overheads outweigh the balance (measured d = 602)
    if out_lo is not None or self_lo is not None or drop is not None or schedule is not None:
        sch = None
        if schedule is not None:
            assert schedule.n == n_dst
            ld_ws = (d + 3) // 4 * 4
            ws = torch.empty((schedule.max_segments, ld_ws), dtype=torch.float32, device=X.device)
            cnt = schedule.counters(max(1, (d + 7) // 8))       # one arrival counter per (feature chunk, row); 8-float chunks are the narrowest
            sch = _lib.SpmmSchedule(ptr(schedule.seg_ptr), ptr(schedule.seg_row), schedule.seg_len,
                                    schedule.max_segments, ptr(cnt), ptr(ws), ld_ws, ptr(schedule.seg_meta),
                                    _lib.SPMM_SCHED_PREFETCH if SEG_PREFETCH else 0)
        ex = _lib.SpmmEx(ptr(out_lo), _ld(out_lo) if out_lo is not None else 0,
                         ptr(self_lo), _ld(self_lo) if self_lo is not None else 0,
                         ctypes.pointer(drop) if drop is not None else None, drop_col0_out, drop_col0_self,
                         ctypes.pointer(sch) if sch is not None else None)
    st = lib.gist_spmm_csr_ex_f32(
        ptr(rowptr), ptr(col), n_dst, n_src, ptr(X), _ld(X), d, ptr(out), _ld(out),
        ptr(src_scale), ptr(dst_scale), ptr(bias),
        ptr(addend), _ld(addend) if addend is not None else 0,
        ptr(self_out), _ld(self_out) if self_out is not None else 0,
        f, ctypes.byref(ex) if ex is not None else None, stream_ptr(X.device))
    check(st, 'spmm_csr_ex_f32')
    if prof is not None:
        ev1.record()
        prof.append(dict(ev0=ev0, ev1=ev1, rowptr=rowptr, n_dst=n_dst, n_src=n_src, d=d,
                         scaled=(src_scale is not None) + (dst_scale is not None)))
    return out


SLAB_SMEM_BYTES = 224 * 1024 - (768 * 20 + 16384 * 2)      # csrc/spmm.cu kSlabSmemBudget
# Shared-memory slab kernel (spmm_slab_kernel): OFF by default.  Measured on the Reddit-shape step
# (profiles/README.md, round 2): 24.5 us per d = 256 launch against 20.1 us for the L2-gather segment
# kernel in isolation, and 30-48 us inside the replayed step, where a 1024-thread / 200 KB CTA has to
# wait for a whole SM to drain from the preparation branch's kernels.  GIST_SPMM_SLAB=1 opts in.
SLAB_ENABLED = os.environ.get('GIST_SPMM_SLAB', '0') != '0'


def _slab_ok(X, n_src):
    """The scheduled launch will stage column slabs of X in shared memory (spmm_slab_kernel): X rows
    16-byte aligned and an 8-float slab of all n_src rows within the shared-memory budget."""
    return (X.data_ptr() % 16 == 0 and _ld(X) % 4 == 0 and 0 < n_src * 32 <= SLAB_SMEM_BYTES and n_src < 65536)


# A/B switches of the segment kernel's per-item overheads (profiles/): packed segment records, queue prefetch
SEG_META = os.environ.get('GIST_SEG_META', '1') != '0'
SEG_PREFETCH = os.environ.get('GIST_SEG_PREFETCH', '0') != '0'     # measured slower: an item claimed early waits behind a long one (0.2502 vs 0.2334 ms/step)


class SegSchedule:
    """Segment schedule of one CSR (gist_spmm_schedule_t minus the per-launch workspace)."""

    def __init__(self, rowptr, n, capacity, seg_len=64):
        self.n, self.seg_len = n, seg_len
        self.max_segments = n + capacity // seg_len + 1
        dev = rowptr.device
        self.seg_ptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        self.seg_row = torch.empty(self.max_segments, dtype=torch.int32, device=dev)
        # packed per-segment records (one 16-byte load per warp item); the format holds < 2^20 segments
        self.seg_meta = (torch.empty((self.max_segments, 4), dtype=torch.int32, device=dev)
                         if SEG_META and self.max_segments < (1 << 20) and seg_len <= 4095 else None)
        # work-queue head + arrival counters (zeroed once; the kernel leaves them zero).  The first
        # stream that launches on this schedule owns the set allocated here; any other stream gets
        # its own (see counters()).  Sized for the widest scheduled launch of a Reddit-shape batch
        # (d = 602 in 8-float chunks: 76).
        self._counters = torch.zeros(1 + 76 * max(n, 1), dtype=torch.int32, device=dev)
        self._owner = None
        self._others = {}
        self.rebuild(rowptr)

    def rebuild(self, rowptr):
        """(Re)compute the schedule for the current contents of rowptr, in place."""
        lib = _lib.load()
        wsb = lib.gist_spmm_schedule_workspace_bytes(self.n)
        ws = torch.empty(max(wsb, 4), dtype=torch.uint8, device=rowptr.device)
        check(lib.gist_spmm_schedule_build_meta(ptr(rowptr), self.n, self.seg_len, ptr(self.seg_ptr), ptr(self.seg_row),
                                                ptr(self.seg_meta), self.max_segments, ptr(ws), wsb,
                                                stream_ptr(rowptr.device)), 'spmm_schedule_build_meta')

    def counters(self, chunks):
        """Zeroed work-queue head + arrival counters for `chunks` feature chunks (the kernel leaves
        them zero).  One set PER STREAM: launches on one stream are ordered, but the same structure
        is aggregated concurrently from parallel streams / graph branches (the m sub-networks of
        train_ist share the graph), and two running launches must never share a queue head.  The
        owner stream's set exists from construction, so a schedule first used inside a CUDA-graph
        capture adds no fill node to the captured step."""
        need = 1 + chunks * max(self.n, 1)
        dev = self.seg_ptr.device
        key = torch.cuda.current_stream(dev).cuda_stream
        if self._owner is None:
            self._owner = key
        if key == self._owner:
            if self._counters.numel() < need:
                self._counters = torch.zeros(need, dtype=torch.int32, device=dev)
            return self._counters
        cnt = self._others.get(key)
        if cnt is None or cnt.numel() < need:
            cnt = self._others[key] = torch.zeros(need, dtype=torch.int32, device=dev)
        return cnt


def degree_norm(rowptr, n, mode):
    require_cuda(rowptr)
    out = torch.empty(n, dtype=torch.float32, device=rowptr.device)
    check(_lib.load().gist_degree_norm_f32(ptr(rowptr), n, mode, ptr(out), stream_ptr(rowptr.device)),
          'degree_norm_f32')
    return out


def exclusive_scan(x):
    """int32 exclusive scan; returns n+1 entries (last = total)."""
    require_cuda(x)
    assert x.dtype == torch.int32 and x.dim() == 1 and x.is_contiguous()
    n = x.shape[0]
    lib = _lib.load()
    out = torch.empty(n + 1, dtype=torch.int32, device=x.device)
    wsb = lib.gist_scan_workspace_bytes(n)
    ws = torch.empty(max(wsb, 4), dtype=torch.uint8, device=x.device)
    check(lib.gist_exclusive_scan_i32(ptr(x), n, ptr(out), ptr(ws), wsb, stream_ptr(x.device)),
          'exclusive_scan_i32')
    return out


def _row_pitch(t):
    """(bytes per row, row pitch in bytes) of a tensor whose rows are contiguous runs of elements
    (any trailing shape; a 2-D matrix may carry a padded leading dimension), else None."""
    if t.dim() == 2 and (t.shape[1] <= 1 or t.stride(1) == 1) and (t.shape[0] <= 1 or t.stride(0) >= t.shape[1]):
        return t.shape[1] * t.element_size(), (t.stride(0) if t.shape[0] > 1 else max(t.shape[1], 1)) * t.element_size()
    if t.is_contiguous():
        rb = t.element_size()
        for s in t.shape[1:]:
            rb *= s
        return rb, rb
    return None


def pad_rows(t):
    """HBM layout of feature matrices: a 2-D fp32 matrix whose width is not a multiple of 4 floats
    (Reddit's 602 input features) is re-laid with a leading dimension rounded up to 4 floats and a
    16-byte-aligned base, and returned as the [n, d] view of that buffer (pad columns zero).  Rows
    are then 16-byte aligned: the SpMM gathers them with 128-bit loads and TMA addresses the matrix
    directly.  Anything else is returned unchanged."""
    if not (torch.is_tensor(t) and t.is_cuda and t.dim() == 2 and t.dtype == torch.float32):
        return t
    n, d = t.shape
    if d < 8 or (d % 4 == 0 and t.is_contiguous()) or _tma_ok(t):
        return t
    buf = torch.zeros((n, (d + 3) // 4 * 4), dtype=torch.float32, device=t.device)
    v = buf[:, :d]
    v.copy_(t)
    return v


def gather_rows(src, idx, out=None):
    """src[idx] along dim 0 for any dtype / trailing shape (ndata row-gather).  A 2-D fp32 source
    with padded rows (pad_rows) yields a result with padded rows."""
    require_cuda(src, idx, out)
    assert idx.dtype == torch.int64 and idx.dim() == 1
    rp = _row_pitch(src)
    if rp is None:
        src = src.contiguous()
        rp = _row_pitch(src)
    row_bytes, src_pitch = rp
    n = idx.shape[0]
    if out is None:
        if src.dim() == 2 and src.dtype == torch.float32 and src.shape[1] >= 8 and src.shape[1] % 4:
            out = torch.zeros((n, (src.shape[1] + 3) // 4 * 4), dtype=src.dtype, device=src.device)[:, :src.shape[1]]
        else:
            out = torch.empty((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    assert out.shape == (n,) + tuple(src.shape[1:]) and out.dtype == src.dtype
    op = _row_pitch(out)
    assert op is not None and op[0] == row_bytes, 'gather_rows: output rows must be contiguous runs'
    if row_bytes % 16 and src_pitch % 16 == 0 and op[1] % 16 == 0:
        # padded rows on both sides: move whole 16-byte vectors, pad columns included
        row_bytes = min((row_bytes + 15) // 16 * 16, src_pitch, op[1])
    check(_lib.load().gist_gather_rows(ptr(src), src_pitch, ptr(idx.contiguous()), n, ptr(out),
                                       op[1], row_bytes, stream_ptr(src.device)), 'gather_rows')
    return out


def slice_gather(src, ridx=None, cidx=None, out=None):
    """out[r, c] = src[ridx[r], cidx[c]] (None = identity). 1-D src is treated as a row."""
    require_cuda(src, ridx, cidx)
    one_d = src.dim() == 1
    s2 = src.unsqueeze(0) if one_d else src
    s2 = _mat(s2, 'src')
    nr = ridx.shape[0] if ridx is not None else s2.shape[0]
    nc = cidx.shape[0] if cidx is not None else s2.shape[1]
    if out is None:
        out = torch.empty((nr, nc), dtype=torch.float32, device=src.device)
    o2 = out.unsqueeze(0) if out.dim() == 1 else out
    assert tuple(o2.shape) == (nr, nc) and (o2.stride(1) == 1 or nc == 1)
    check(_lib.load().gist_slice_gather_f32(ptr(s2), _ld(s2), ptr(ridx), nr, ptr(cidx), nc, ptr(o2),
                                            _ld(o2), stream_ptr(src.device)), 'slice_gather_f32')
    note_raw_write(out)
    return out.reshape(nc) if one_d and out.dim() == 2 else out


def slice_scatter_(dst, src, ridx=None, cidx=None):
    """dst[ridx[r], cidx[c]] = src[r, c] in place (indices unique)."""
    require_cuda(dst, src, ridx, cidx)
    d2 = dst.unsqueeze(0) if dst.dim() == 1 else dst
    s2 = src.unsqueeze(0) if src.dim() == 1 else src
    s2 = _mat(s2, 'src')
    assert d2.dtype == torch.float32 and (d2.stride(1) == 1 or d2.shape[1] == 1)
    nr, nc = s2.shape
    assert nr == (ridx.shape[0] if ridx is not None else d2.shape[0])
    assert nc == (cidx.shape[0] if cidx is not None else d2.shape[1])
    check(_lib.load().gist_slice_scatter_f32(ptr(s2), _ld(s2), ptr(ridx), nr, ptr(cidx), nc, ptr(d2),
                                             _ld(d2), stream_ptr(dst.device)), 'slice_scatter_f32')
    note_raw_write(dst)
    return dst


def slice_multi(jobs, scatter):
    """Many slice gathers / scatters in ONE launch (gist_slice_multi_f32).
    jobs: list of (src, ridx, cidx, dst): gather  dst[r, c] = src[ridx[r], cidx[c]]
                                          scatter dst[ridx[r], cidx[c]] = src[r, c]   (None = identity)
    1-D tensors are rows.  `src` of a scatter may alias another rank's memory (peer-mapped buffer)."""
    if not jobs:
        return
    arr = (_lib.SliceJob * len(jobs))()
    dev = None
    for k, (src, ridx, cidx, dst) in enumerate(jobs):
        require_cuda(src, ridx, cidx, dst)
        s2 = src.unsqueeze(0) if src.dim() == 1 else src
        d2 = dst.unsqueeze(0) if dst.dim() == 1 else dst
        assert s2.dtype == torch.float32 and d2.dtype == torch.float32
        assert (s2.shape[1] <= 1 or s2.stride(1) == 1) and (d2.shape[1] <= 1 or d2.stride(1) == 1)
        dense = s2 if scatter else d2            # the side indexed densely: [n_rows, n_cols]
        nr, nc = dense.shape
        other = d2 if scatter else s2
        assert nr == (ridx.shape[0] if ridx is not None else other.shape[0]), (nr, other.shape)
        assert nc == (cidx.shape[0] if cidx is not None else other.shape[1]), (nc, other.shape)
        for ix in (ridx, cidx):
            assert ix is None or (ix.dtype == torch.int64 and ix.is_contiguous())
        arr[k] = _lib.SliceJob(s2.data_ptr(), _ld(s2), ridx.data_ptr() if ridx is not None else None, nr,
                               cidx.data_ptr() if cidx is not None else None, nc, d2.data_ptr(), _ld(d2))
        dev = dst.device
        note_raw_write(dst)
    check(_lib.load().gist_slice_multi_f32(1 if scatter else 0, len(jobs), arr, stream_ptr(dev)), 'slice_multi_f32')


def index_invert(idx, size, out=None):
    """inv[size] int32 with inv[idx[k]] = k and -1 elsewhere (gist_index_invert_i32)."""
    require_cuda(idx, out)
    assert idx.dtype == torch.int64 and idx.is_contiguous()
    if out is None:
        out = torch.empty(size, dtype=torch.int32, device=idx.device)
    assert out.dtype == torch.int32 and out.numel() == size and out.is_contiguous()
    check(_lib.load().gist_index_invert_i32(ptr(idx), idx.numel(), ptr(out), size, stream_ptr(idx.device)),
          'index_invert_i32')
    return out


def slice_scatter_rows_ok(dst, ridx, cidx):
    """Can gist_slice_scatter_rows_f32 serve this destination / slice (a 2-D slice of whole-sector rows)?"""
    return (dst.dim() == 2 and dst.is_cuda and dst.dtype == torch.float32 and dst.stride(1) == 1
            and dst.data_ptr() % 32 == 0 and dst.stride(0) % 8 == 0 and ridx is not None and cidx is not None)


def slice_scatter_rows_(jobs):
    """dst[ridx[r], cidx[c]] = src[r, c] for jobs (src, ridx, inv_col, dst) whose row sets are disjoint per
    destination, as ONE row-streaming launch (gist_slice_scatter_rows_f32): whole 32-byte sectors in ascending
    address order instead of 4-byte scatters.  ``inv_col`` = index_invert(cidx, dst.shape[1])."""
    if not jobs:
        return
    arr = (_lib.SliceRowsJob * len(jobs))()
    dev = None
    for k, (src, ridx, inv, dst) in enumerate(jobs):
        require_cuda(src, ridx, inv, dst)
        assert src.dim() == 2 and dst.dim() == 2 and src.dtype == torch.float32 and dst.dtype == torch.float32
        assert (src.shape[1] <= 1 or src.stride(1) == 1) and dst.stride(1) == 1
        assert inv.dtype == torch.int32 and inv.numel() == dst.shape[1] and inv.is_contiguous()
        assert ridx is None or (ridx.dtype == torch.int64 and ridx.is_contiguous() and ridx.numel() == src.shape[0])
        arr[k] = _lib.SliceRowsJob(src.data_ptr(), _ld(src), ridx.data_ptr() if ridx is not None else None, src.shape[0],
                                   src.shape[1], inv.data_ptr(), dst.data_ptr(), _ld(dst), dst.shape[1])
        dev = dst.device
        note_raw_write(dst)
    check(_lib.load().gist_slice_scatter_rows_f32(len(jobs), arr, stream_ptr(dev)), 'slice_scatter_rows_f32')


# --------------------------------------------------------------------------
# autograd
# --------------------------------------------------------------------------
class _ScaledSpMM(torch.autograd.Function):
    """Y = act(t ⊙ A (s ⊙ X) + bias); backward through K2 on the CSC."""

    @staticmethod
    def forward(ctx, g, X, src_scale, dst_scale, bias, relu):
        X = _mat(X, 'X')
        n = g.number_of_nodes()
        Y = torch.empty((n, X.shape[1]), dtype=torch.float32, device=X.device)
        spmm_raw(g.rowptr, g.col_buffer, n, n, X, Y, src_scale=src_scale, dst_scale=dst_scale,
                 bias=bias, relu=relu, schedule=g.seg_schedule())
        ctx.g, ctx.relu, ctx.has_bias = g, relu, bias is not None
        ctx.save_for_backward(src_scale, dst_scale, Y if relu else None)
        return Y

    @staticmethod
    def backward(ctx, dY):
        src_scale, dst_scale, Y = ctx.saved_tensors
        g = ctx.g
        dY = _mat(dY, 'dY')
        if ctx.relu:
            dY = dY * (Y > 0).to(dY.dtype)
        dbias = dY.sum(0) if (ctx.has_bias and ctx.needs_input_grad[4]) else None
        dX = None
        if ctx.needs_input_grad[1]:
            n = g.number_of_nodes()
            colptr, row = g.csc()
            dX = torch.empty_like(dY)
            # dX[u] = s[u] * sum_{u->v} t[v] dY[v]
            spmm_raw(colptr, row, n, n, dY, dX, src_scale=dst_scale, dst_scale=src_scale,
                     schedule=g.seg_schedule(transpose=True))
        return None, dX, None, None, dbias, None


def gspmm(g, X, src_scale=None, dst_scale=None, bias=None, relu=False):
    return _ScaledSpMM.apply(g, X, src_scale, dst_scale, bias, relu)


def copy_src_sum(g, X):
    """update_all(fn.copy_src, fn.sum): Y[v] = sum_{u->v} X[u]."""
    return _ScaledSpMM.apply(g, X, None, None, None, False)


class _SageConcat(torch.autograd.Function):
    """z = [h ‖ inv_deg ⊙ (A h)] written by ONE kernel into a [n, 2d] buffer
    (cluster_gcn/modules.py:222-227); backward dh = dz[:, :d] + Aᵀ(inv_deg ⊙ dz[:, d:])."""

    @staticmethod
    def forward(ctx, g, h):
        h = _mat(h, 'h')
        n, d = h.shape
        z = torch.empty((n, 2 * d), dtype=torch.float32, device=h.device)
        inv = g.inv_in_degree()
        spmm_raw(g.rowptr, g.col_buffer, n, n, h, z[:, d:], dst_scale=inv, self_out=z[:, :d],
                 schedule=g.seg_schedule())
        ctx.g = g
        return z

    @staticmethod
    def backward(ctx, dz):
        g = ctx.g
        dz = _mat(dz, 'dz')
        n, d2 = dz.shape
        d = d2 // 2
        colptr, row = g.csc()
        dh = torch.empty((n, d), dtype=torch.float32, device=dz.device)
        spmm_raw(colptr, row, n, n, dz[:, d:], dh, src_scale=g.inv_in_degree(), addend=dz[:, :d],
                 schedule=g.seg_schedule(transpose=True))
        return None, dh


def sage_concat(g, h):
    return _SageConcat.apply(g, h)


def sage_concat_into(g, h, out, out_lo=None, drop=None, balanced=True, background=0):
    """No-grad z = [h ‖ (A h) / in_deg] into a caller-owned [n, 2d] buffer; optionally with the
    layer's dropout applied as z is written (``drop``) and z's 3xTF32 low half (``out_lo``).
    ``balanced``: use the segment-scheduled kernel when the graph has a schedule.
    ``background`` (1..15, row-per-group kernel with an extended epilogue only): at most that many
    CTAs per SM (GIST_SPMM_BG_SHIFT), for a launch that runs beside a latency-critical branch."""
    h = _mat(h, 'h')
    n, d = h.shape
    assert out.shape == (n, 2 * d) and out.dtype == torch.float32 and out.stride(1) == 1
    spmm_raw(g.rowptr, g.col_buffer, n, n, h, out[:, d:], dst_scale=g.inv_in_degree(), self_out=out[:, :d],
             out_lo=out_lo[:, d:] if out_lo is not None else None,
             self_lo=out_lo[:, :d] if out_lo is not None else None,
             drop=drop, drop_col0_out=d, drop_col0_self=0, schedule=g.seg_schedule() if balanced else None,
             flags=(int(background) & 15) << _lib.SPMM_BG_SHIFT)
    return out


def _padded_empty(rows, cols, device):
    """[rows, cols] fp32 view of a buffer whose leading dimension is a multiple of 4 floats (TMA)."""
    return torch.empty((rows, (cols + 3) // 4 * 4), dtype=torch.float32, device=device)[:, :cols]


class SagePre:
    """A SAGE layer's prepared input: z = dropout([h ‖ mean-agg(h)]), its 3xTF32 low half (or None)
    and the dropout step the mask was drawn at (or None when no dropout was applied)."""
    __slots__ = ('z', 'z_lo', 'step_saved', 'dropped')

    def __init__(self, z, z_lo, step_saved, dropped):
        self.z, self.z_lo, self.step_saved, self.dropped = z, z_lo, step_saved, dropped


def sage_prepare(g, h, p_drop, stream_id, out=None, balanced=True, background=0):
    """Aggregation + concat (+ dropout, + 3xTF32 split) of a SAGE layer's input, no autograd.
    ``out``: a SagePre of the same shape to overwrite in place."""
    h = _mat(h, 'h')
    n, d = h.shape
    x3 = _MATMUL_PRECISION == '3xtf32'
    if out is None:
        z = _padded_empty(n, 2 * d, h.device)
        z_lo = _padded_empty(n, 2 * d, h.device) if x3 else None
        saved = torch.empty(1, dtype=torch.int64, device=h.device) if p_drop else None   # K1 writes it
        out = SagePre(z, z_lo, saved, bool(p_drop))
    desc = dropout_state(h.device).desc(p_drop, stream_id, step_saved=out.step_saved) if p_drop else None
    sage_concat_into(g, h, out.z, out.z_lo, desc, balanced=balanced, background=background)
    return out


# --------------------------------------------------------------------------
# fused dropout: counter-based mask shared by the kernels on both sides of nn.Dropout
# --------------------------------------------------------------------------
class DropoutState:
    """Per-device dropout clock.  A mask is a pure function of (seed, step, stream_id, row, col)
    (include/gist_b200.h, gist_dropout_t): ``step`` lives in device memory so captured CUDA graphs
    draw fresh masks at every replay, and is advanced once per model forward in training mode
    (``auto_tick``) or by whoever drives the step (GraphedClusterTrainer ticks before its branches
    fork).  The seed follows ``torch.manual_seed`` (``torch.initial_seed()`` at use)."""

    def __init__(self, device):
        self.device = device
        self.step = torch.zeros(1, dtype=torch.int64, device=device)
        self.auto_tick = True

    def tick(self):
        check(_lib.load().gist_counter_add_i64(ptr(self.step), 1, stream_ptr(self.device)), 'counter_add_i64')

    def desc(self, p, stream_id, step=None, step_saved=None):
        """ctypes gist_dropout_t reading ``step`` (default: the live clock)."""
        step = self.step if step is None else step
        return _lib.DropoutDesc(float(p), torch.initial_seed() & 0xFFFFFFFFFFFFFFFF, int(stream_id) & 0xFFFFFFFF,
                                step.data_ptr(), step_saved.data_ptr() if step_saved is not None else None)


_DROPOUT_STATES = {}
_drop_stream_counter = [0]


def dropout_state(device):
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _DROPOUT_STATES.get(idx)
    if st is None:
        st = _DROPOUT_STATES[idx] = DropoutState(torch.device('cuda', idx))
    return st


def new_dropout_stream():
    """A fresh stream id (one per dropout site; deterministic in construction order)."""
    _drop_stream_counter[0] += 1
    return _drop_stream_counter[0]


def dropout(x, desc, col0=0):
    """desc's mask applied to x as a stand-alone kernel (the same mask the fused kernels apply)."""
    require_cuda(x)
    x = _mat(x, 'x')
    out = torch.empty_like(x)
    check(_lib.load().gist_dropout_f32(ptr(x), _ld(x), x.shape[0], x.shape[1], col0, ptr(out), _ld(out),
                                       ctypes.byref(desc), stream_ptr(x.device)), 'dropout_f32')
    return out


# 3xTF32 low halves produced by the kernel that wrote a tensor (layer-norm / cross-entropy backward),
# handed to the nn.Linear backward that consumes it.  Keyed by data pointer; the entry holds the
# tensor itself, so the pointer cannot be recycled while the entry lives; consumed on first use.
_GRAD_LO = {}
# weight low halves split ahead of time by a trainer (GraphedClusterTrainer: on the side branch)
_WEIGHT_LO = {}


def _lo_put(t, lo):
    if len(_GRAD_LO) > 16:
        _GRAD_LO.clear()
    _GRAD_LO[t.data_ptr()] = (t, lo)


def _lo_take(t):
    e = _GRAD_LO.pop(t.data_ptr(), None)
    if e is not None and e[0].shape == t.shape and e[0].stride() == t.stride() and e[1].shape == t.shape:
        return e[1]
    return None


# Low halves that OUTLIVE a forward pass: a trainer that owns the parameter updates
# (GraphedClusterTrainer: the fused Adam launch writes tf32_lo(p) beside every updated p) registers one
# persistent buffer per weight; _weight_lo() then needs no split launch at the head of the step.
# Validity is checked at every use: an entry is served only while (a) no kernel of this library has
# written a registered weight through its raw pointer since the last refresh (note_raw_write: the K5
# slice kernels of a GIST dispatch) and (b) the parameter's Tensor._version is the one recorded at the
# last refresh (torch in-place ops: copy_, load_state_dict ...).  Otherwise the ordinary split runs.
# The owning trainer refreshes (one launch) before its next step.  Writes through ``p.data`` bypass
# both checks: a trainer refreshes at every optimizer reset (every GIST dispatch) for that reason.
class _PersistEntry:
    __slots__ = ('ref', 'lo', 'ver', 'shape', 'stride')

    def __init__(self, p, lo):
        self.ref, self.lo, self.ver = weakref.ref(p), lo, -1
        self.shape, self.stride = tuple(p.shape), tuple(p.stride())


_WEIGHT_LO_PERSIST = {}                 # data_ptr -> _PersistEntry
_PERSIST_EPOCH = [0, -1]                # [raw-write epoch, epoch the buffers were last refreshed at]


def register_persistent_lo(p, lo):
    """``p``: the parameter (the tensor object whose _version is watched); ``lo``: contiguous, same shape.
    Returns the buffer to write: ``lo``, or the one already registered for this very parameter (two trainers
    over one model must keep ONE low half current, not two of which only the newer is served)."""
    assert p.dim() == 2 and _tma_ok(p) and p.is_contiguous() and lo.shape == p.shape and lo.is_contiguous()
    e = _WEIGHT_LO_PERSIST.get(p.data_ptr())
    if e is not None and e.ref() is p and e.lo.shape == p.shape:
        return e.lo
    _WEIGHT_LO_PERSIST[p.data_ptr()] = _PersistEntry(p, lo)
    _PERSIST_EPOCH[1] = -1
    return lo


def unregister_persistent_lo(p):
    _WEIGHT_LO_PERSIST.pop(p.data_ptr(), None)


def note_raw_write(t):
    """A kernel of this library wrote tensor ``t`` through its raw pointer (no Tensor._version bump)."""
    if _WEIGHT_LO_PERSIST and t is not None and t.data_ptr() in _WEIGHT_LO_PERSIST:
        _PERSIST_EPOCH[0] += 1


def _persist_live():
    dead = [k for k, e in _WEIGHT_LO_PERSIST.items() if e.ref() is None]
    for k in dead:
        del _WEIGHT_LO_PERSIST[k]
    return list(_WEIGHT_LO_PERSIST.values())


def persistent_lo_valid():
    if _PERSIST_EPOCH[0] != _PERSIST_EPOCH[1]:
        return False
    return all(e.ref()._version == e.ver for e in _persist_live())


def refresh_persistent_lo():
    """Re-split every registered weight on the current stream (one launch per 16 weights)."""
    es = _persist_live()
    for i in range(0, len(es), 16):
        split_tf32_multi([e.ref().detach() for e in es[i:i + 16]], [e.lo for e in es[i:i + 16]])
    for e in es:
        e.ver = e.ref()._version
    _PERSIST_EPOCH[1] = _PERSIST_EPOCH[0]


def _persistent_lo(Wv):
    e = _WEIGHT_LO_PERSIST.get(Wv.data_ptr())
    if e is None or _PERSIST_EPOCH[0] != _PERSIST_EPOCH[1]:
        return None
    p = e.ref()
    if p is None or p._version != e.ver or tuple(Wv.shape) != e.shape or tuple(Wv.stride()) != e.stride:
        return None
    return e.lo


def split_tf32_multi(ws, los):
    """lo_i = tf32(w_i - trunc_tf32(w_i)) for up to 16 matrices in ONE launch (gist_split_tf32_multi_f32)."""
    n = len(ws)
    if n == 0:
        return
    assert n <= 16
    X, LO = (ctypes.c_void_p * n)(), (ctypes.c_void_p * n)()
    LDX, LDL = (ctypes.c_int64 * n)(), (ctypes.c_int64 * n)()
    R, C = (ctypes.c_int32 * n)(), (ctypes.c_int32 * n)()
    for i, (w, lo) in enumerate(zip(ws, los)):
        X[i], LO[i], LDX[i], LDL[i], R[i], C[i] = w.data_ptr(), lo.data_ptr(), _ld(w), _ld(lo), w.shape[0], w.shape[1]
    check(_lib.load().gist_split_tf32_multi_f32(n, X, LDX, R, C, LO, LDL, stream_ptr(ws[0].device)),
          'split_tf32_multi_f32')


def presplit_weights(params):
    """3xTF32 low halves of every TMA-addressable 2-D fp32 tensor in ``params`` with ONE launch
    (gist_split_tf32_multi_f32), left for the layers to pick up: each entry is consumed by the first
    _weight_lo() of that tensor, so nothing outlives the forward pass that asked for it.  Weights with a
    registered persistent low half (register_persistent_lo) are skipped."""
    _WEIGHT_LO.clear()
    ws = [p.data if isinstance(p, torch.nn.Parameter) else p for p in params]
    ws = [w for w in ws if w.dim() == 2 and w.numel() > 0 and _tma_ok(w) and _persistent_lo(w) is None][:16]
    if not ws:
        return
    n = len(ws)
    los = [_padded_empty(w.shape[0], w.shape[1], w.device) for w in ws]
    X, LO = (ctypes.c_void_p * n)(), (ctypes.c_void_p * n)()
    LDX, LDL = (ctypes.c_int64 * n)(), (ctypes.c_int64 * n)()
    R, C = (ctypes.c_int32 * n)(), (ctypes.c_int32 * n)()
    for i, (w, lo) in enumerate(zip(ws, los)):
        X[i], LO[i], LDX[i], LDL[i], R[i], C[i] = w.data_ptr(), lo.data_ptr(), _ld(w), _ld(lo), w.shape[0], w.shape[1]
    check(_lib.load().gist_split_tf32_multi_f32(n, X, LDX, R, C, LO, LDL, stream_ptr(ws[0].device)),
          'split_tf32_multi_f32')
    for w, lo in zip(ws, los):
        _WEIGHT_LO[w.data_ptr()] = (w, lo)


def clear_presplit():
    _WEIGHT_LO.clear()


def _weight_lo(Wv):
    lo = _persistent_lo(Wv)
    if lo is not None:
        return lo
    e = _WEIGHT_LO.pop(Wv.data_ptr(), None)
    if e is not None and e[1].shape == Wv.shape and e[0].stride() == Wv.stride():
        return e[1]
    return split_tf32(Wv)


# --------------------------------------------------------------------------
# K4: tensor-core GEMM (tcgen05, TF32 inputs / fp32 accumulate)
# --------------------------------------------------------------------------
# The product path: every dense contraction of the drop-in modules runs on the hand-written tcgen05
# kernel (K4) in its fp32-accurate 3xTF32 mode.  'fp32' (cuBLAS sgemm through torch) exists only as an
# explicit cross-check / debug mode that the parity tests switch on to compare the two back ends.
DEFAULT_MATMUL_PRECISION = '3xtf32'
_MATMUL_PRECISION = DEFAULT_MATMUL_PRECISION


def set_matmul_precision(mode):
    """'3xtf32' : (default) the hand-written tcgen05 kernel with split operands (x, x_lo): three TF32
               MMAs per K step, ~2^-20 relative per product -> fp32-level accuracy (meets the 1e-5
               parity tolerance) on the tensor cores;
    'tf32'   : the same kernel, single pass (10-bit mantissa inputs, fp32 accumulate, ~1e-3);
    'fp32'   : DEBUG / cross-check only — nn.Linear / matmul through cuBLAS sgemm via torch."""
    global _MATMUL_PRECISION
    assert mode in ('fp32', 'tf32', '3xtf32')
    _MATMUL_PRECISION = mode


def get_matmul_precision():
    return _MATMUL_PRECISION


def _tma_ok(t):
    return (t.dim() == 2 and t.dtype == torch.float32 and t.is_cuda and t.stride(1) == 1
            and t.data_ptr() % 16 == 0 and (t.shape[0] <= 1 or t.stride(0) % 4 == 0)
            and t.stride(0) >= t.shape[1])


def _tma_view(t):
    """`t` itself when TMA can address it (16-byte aligned base, leading dimension a multiple of
    4 floats), else a copy into a buffer whose rows are padded to a multiple of 4 floats (e.g. the
    41-wide logit gradient)."""
    if _tma_ok(t):
        return t
    rows, cols = t.shape
    buf = torch.empty((rows, (cols + 3) // 4 * 4), dtype=torch.float32, device=t.device)
    v = buf[:, :cols]
    v.copy_(t)
    return v


def split_tf32(x):
    """x_lo = tf32(x - trunc_tf32(x)) for a 2-D fp32 matrix, in a TMA-addressable buffer (the
    second operand half of the 3xTF32 GEMM; the tensor core itself takes trunc_tf32(x) from x)."""
    require_cuda(x)
    rows, cols = x.shape
    buf = torch.empty((rows, (cols + 3) // 4 * 4), dtype=torch.float32, device=x.device)
    lo = buf[:, :cols]
    check(_lib.load().gist_split_tf32_f32(ptr(x), _ld(x), rows, cols, None, 0, ptr(lo), _ld(lo),
                                          stream_ptr(x.device)), 'split_tf32_f32')
    return lo


# In-kernel split-K needs one zeroed uint32 per output tile that the kernel leaves zero: one array per
# (device, stream) — launches on a stream are ordered, launches on different streams (the weight-
# gradient branch beside the training branch) must not share counters.
GEMM_TILE_COUNTERS = 4096
# In-kernel split-K (last-arriver reduction) is OPT-IN: measured slower than the two-kernel form on
# every Reddit-shape GEMM (profiles/README.md round 2: e.g. dW0 24.0 vs 18.9 us, L0 fwd 21.3 vs 20.2 us,
# step 0.299 vs 0.280 ms) — the tile's last CTA folds S partial tiles alone while the second-pass kernel
# spreads the fold over the chip.  Row sums and the layer norm ride in the second pass instead.
FUSED_SPLITK = os.environ.get('GIST_GEMM_FUSED_SPLITK', '0') != '0'       # A/B switches for profiles/
FUSED_ROWSUM = os.environ.get('GIST_GEMM_FUSED_ROWSUM', '1') != '0'
BACKGROUND_DW = os.environ.get('GIST_GEMM_BACKGROUND_DW', '1') != '0'
DZ_FIRST = os.environ.get('GIST_DZ_FIRST', '1') != '0'
FUSED_LN = os.environ.get('GIST_GEMM_FUSED_LN', '1') != '0'
FUSED_LN_WIDE = os.environ.get('GIST_GEMM_FUSED_LN_WIDE', '1') != '0'     # rows of 129..256: in the split-K second pass
_GEMM_COUNTERS = {}


def _gemm_counters(device):
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    t = _GEMM_COUNTERS.get(key)
    if t is None:
        t = _GEMM_COUNTERS[key] = torch.zeros(GEMM_TILE_COUNTERS, dtype=torch.int32, device=device)
    return t


def gemm(A, B, *, a_mn=False, b_mn=False, bias=None, relu=False, out=None, flags=0, A_lo=None, B_lo=None,
         rowsum=False, ln=None, drop=None):
    """C[M,N] = op(A) @ op(B)^T (+bias) (ReLU) on the tcgen05 TF32 kernel (K4).

    a_mn=False: A is stored [M, K];  a_mn=True: A is stored [K, M] (MN-major, i.e. op(A) = A^T).
    b_mn=False: B is stored [N, K];  b_mn=True: B is stored [K, N].
    A_lo / B_lo (from split_tf32, both or neither) select the fp32-accurate 3xTF32 mode.
    Extended epilogue (gist_gemm_ex_f32):
      rowsum=True : also returns r[M] = sum_k op(A)[m, k]  (db of the dW = dy^T z contraction) -> (C, r)
      ln=(eps, relu) : LayerNorm(no affine)(+ReLU) over the rows of C when N <= 128, 3xTF32, K-major
                    operands -> (C_pre_norm, y, stats[M, 2]); raises GistLibraryError if it does not apply
      drop : a DropoutDesc — C *= the mask (see gemm_dropmask)
    Split-K runs as ONE kernel (last-arriver reduction) on this stream's tile counters.
    Raises if an operand is not TMA-addressable (see _tma_view)."""
    require_cuda(A, B, bias, out, A_lo, B_lo)
    if a_mn:
        K, M = A.shape
    else:
        M, K = A.shape
    if b_mn:
        K2, N = B.shape
    else:
        N, K2 = B.shape
    assert K == K2, (A.shape, B.shape, a_mn, b_mn)
    if out is None:
        out = _padded_empty(M, N, A.device) if ln is not None else torch.empty((M, N), dtype=torch.float32, device=A.device)
    assert tuple(out.shape) == (M, N) and (out.stride(1) == 1 or N == 1)
    lib = _lib.load()
    f = flags | (_lib.GEMM_RELU if relu else 0)
    x3 = A_lo is not None or B_lo is not None
    if x3:
        assert A_lo is not None and B_lo is not None and A_lo.shape == A.shape and B_lo.shape == B.shape
    dev = A.device
    ex = _lib.GemmEx()
    keep = []
    if FUSED_SPLITK and M * N > 0:
        cnt = _gemm_counters(dev)
        ex.tile_counters, ex.n_counters = cnt.data_ptr(), cnt.numel()
    rs = y = stats = None
    if rowsum:
        rs = torch.empty(M, dtype=torch.float32, device=dev)
        ex.rowsum = rs.data_ptr()
    if ln is not None:
        eps, ln_relu = ln
        y = _padded_empty(M, N, dev)
        stats = torch.empty((M, 2), dtype=torch.float32, device=dev)
        ex.ln_out, ex.ld_ln, ex.ln_stats = y.data_ptr(), _ld(y), stats.data_ptr()
        ex.ln_eps, ex.ln_flags = float(eps), (_lib.ACT_RELU if ln_relu else 0)
    if drop is not None:
        ex.drop = ctypes.pointer(drop)
        keep.append(drop)
    wsb = lib.gist_gemm_ex_workspace_bytes(M, N, K, f, 1 if x3 else 0, ctypes.byref(ex))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev) if wsb else None
    prof = GEMM_PROFILE
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    check(lib.gist_gemm_ex_f32(ptr(A), ptr(A_lo), _ld(A), _ld(A_lo) if x3 else 0, 1 if a_mn else 0,
                               ptr(B), ptr(B_lo), _ld(B), _ld(B_lo) if x3 else 0, 1 if b_mn else 0,
                               ptr(out), _ld(out), M, N, K, ptr(bias), f, ptr(ws), wsb, ctypes.byref(ex),
                               stream_ptr(dev)), 'gemm_ex_f32')
    if prof is not None:
        ev1.record()
        prof.append(dict(ev0=ev0, ev1=ev1, M=M, N=N, K=K, passes=3 if x3 else 1))
    if ln is not None:
        return out, y, stats
    if rowsum:
        return out, rs
    return out


def ln_fusable(M, N, x3, a_mn=False, b_mn=False):
    """Does K4's LayerNorm fusion apply (include/gist_b200.h gist_gemm_ex_t)?  N <= 128: one tile holds the
    row (GEMM epilogue or split-K second pass); 128 < N <= 256: the split-K second pass normalises the row
    (when K is not split the library runs the row-wise kernel behind the GEMM itself)."""
    return FUSED_LN and x3 and not a_mn and not b_mn and 0 < N <= (256 if FUSED_LN_WIDE else 128) and M > 0


def gemm_dropmask(A, B, drop, *, a_mn=False, b_mn=False, out=None, A_lo=None, B_lo=None):
    """C = mask(drop) ⊙ (op(A) op(B)^T): the dz contraction with the forward's dropout mask
    regenerated in the epilogue (gist_gemm_dropmask_f32)."""
    require_cuda(A, B, out, A_lo, B_lo)
    if a_mn:
        K, M = A.shape
    else:
        M, K = A.shape
    N = B.shape[1] if b_mn else B.shape[0]
    if out is None:
        out = _padded_empty(M, N, A.device)
    assert tuple(out.shape) == (M, N)
    check(_lib.load().gist_gemm_dropmask_f32(
        ptr(A), ptr(A_lo), _ld(A), _ld(A_lo) if A_lo is not None else 0, 1 if a_mn else 0,
        ptr(B), ptr(B_lo), _ld(B), _ld(B_lo) if B_lo is not None else 0, 1 if b_mn else 0,
        ptr(out), _ld(out), M, N, K, 0, ctypes.byref(drop), stream_ptr(A.device)), 'gemm_dropmask_f32')
    return out


def gemm_tn(A, B, bias=None, relu=False, out=None):
    """C[M,N] = A[M,K] @ B[N,K]^T (+bias) (ReLU): both operands K-major."""
    return gemm(A, B, bias=bias, relu=relu, out=out)


def transpose(x):
    """[rows, cols] -> [cols, rows] whose leading dimension is padded to a multiple of 4
    floats so the result is TMA-addressable."""
    require_cuda(x)
    x = _mat(x, 'x')
    rows, cols = x.shape
    ld = (rows + 3) // 4 * 4
    buf = torch.empty((cols, ld), dtype=torch.float32, device=x.device)
    check(_lib.load().gist_transpose_f32(ptr(x), _ld(x), rows, cols, ptr(buf), ld, stream_ptr(x.device)),
          'transpose_f32')
    return buf[:, :rows]


class _LinearTF32(torch.autograd.Function):
    """y = z W^T + b with all three contractions (y, dz, dW) on the tcgen05 kernel, no
    transposed copies: the kernel reads each operand K-major or MN-major as stored."""

    @staticmethod
    def forward(ctx, z, W, b):
        z = _tma_view(_mat(z, 'z'))
        Wv = _tma_view(W)
        y = gemm(z, Wv, bias=b)
        ctx.save_for_backward(z, Wv)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        z, W = ctx.saved_tensors
        dy = _tma_view(_mat(dy, 'dy'))
        dz = dW = db = None
        if ctx.needs_input_grad[0]:
            dz = gemm(dy, W, b_mn=True)                      # [n, out] x [out, in]: N = in, K = out
        if ctx.needs_input_grad[1]:
            dW = gemm(dy, z, a_mn=True, b_mn=True)           # dy^T z: M = out, N = in, K = n
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy)
        return dz, dW, db


class _Linear3xTF32(torch.autograd.Function):
    """y = z W^T + b at fp32 accuracy on the TF32 tensor cores.  Each of z, W, dy is split once
    (x_lo) and the pair is shared by the contractions that read it: z by y and dW, W by y and dz,
    dy by dz and dW — three split launches and three GEMMs per layer."""

    @staticmethod
    def forward(ctx, z, W, b):
        z = _tma_view(_mat(z, 'z'))
        Wv = _tma_view(W)
        z_lo, W_lo = split_tf32(z), split_tf32(Wv)
        y = gemm(z, Wv, bias=b, A_lo=z_lo, B_lo=W_lo)
        ctx.save_for_backward(z, z_lo, Wv, W_lo)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        z, z_lo, W, W_lo = ctx.saved_tensors
        dy = _mat(dy, 'dy')
        dy_lo = _lo_take(dy)
        if not _tma_ok(dy):
            dy, dy_lo = _tma_view(dy), None
        if dy_lo is None:
            dy_lo = split_tf32(dy)
        dz = dW = db = None
        if ctx.needs_input_grad[0]:
            dz = gemm(dy, W, b_mn=True, A_lo=dy_lo, B_lo=W_lo)
        if ctx.needs_input_grad[1]:
            dW = gemm(dy, z, a_mn=True, b_mn=True, A_lo=dy_lo, B_lo=z_lo)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy)
        return dz, dW, db


class _MatmulNN(torch.autograd.Function):
    """y = x @ W for W stored [in, out] (DGL GraphConv's layout, gcn/gcn.py:30-56) on the tcgen05
    kernel: W is read MN-major as stored, no transposed copy.  dx = dy W^T reads the same W
    K-major; dW = x^T dy reads both operands MN-major."""

    @staticmethod
    def forward(ctx, x, W):
        x = _tma_view(_mat(x, 'x'))
        Wv = _tma_view(W)
        x3 = _MATMUL_PRECISION == '3xtf32'
        x_lo = split_tf32(x) if x3 else None
        W_lo = _weight_lo(Wv) if x3 else None
        y = gemm(x, Wv, b_mn=True, A_lo=x_lo, B_lo=W_lo)
        ctx.save_for_backward(x, Wv, *((x_lo, W_lo) if x3 else ()))
        ctx.x3 = x3
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors[:2]
        x_lo, W_lo = ctx.saved_tensors[2:] if ctx.x3 else (None, None)
        dy = _tma_view(_mat(dy, 'dy'))
        dy_lo = split_tf32(dy) if ctx.x3 else None
        dx = dW = None
        if ctx.needs_input_grad[0]:
            dx = gemm(dy, W, A_lo=dy_lo, B_lo=W_lo)                               # [n,out] x [in,out]^T
        if ctx.needs_input_grad[1]:
            dW = gemm(x, dy, a_mn=True, b_mn=True, A_lo=x_lo, B_lo=dy_lo)         # x^T dy: M = in, N = out, K = n
        return dx, dW


def matmul(x, W):
    """x @ W with W stored [in, out]: cuBLAS sgemm through torch in 'fp32' mode, the tcgen05 kernel
    (K4) otherwise."""
    require_cuda(x, W)
    if _MATMUL_PRECISION == 'fp32':
        return x @ W
    return _MatmulNN.apply(x, W)


# --------------------------------------------------------------------------
# weight-gradient branch: dW / db never feed the rest of the backward pass (only the optimizer),
# so they run on a low-priority side stream and the activation-gradient chain
# (dz GEMM -> K2 -> layer-norm backward -> next dz GEMM ...) is the only thing on the main one.
# The side stream is joined by an autograd end-of-backward callback, so whoever reads .grad after
# loss.backward() is ordered behind it — under CUDA-graph capture this becomes a parallel branch.
# --------------------------------------------------------------------------
OVERLAP_WEIGHT_GRADS = True
_SIDE_STREAMS = {}


def _side_stream(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _SIDE_STREAMS.get(idx)
    if st is None:
        lo_p, _ = torch.cuda.Stream.priority_range()
        st = _SIDE_STREAMS[idx] = torch.cuda.Stream(device=idx, priority=lo_p)
    return st


def _weight_grad_branch(device, fn, keep, after=None):
    """Run fn() (returns tensors) on the side stream, forked from the current stream — or, with
    ``after`` (an event recorded on the current stream when fn's inputs were ready), from that earlier
    point: the caller can then enqueue its own latency-critical kernels FIRST and the branch afterwards
    without making the branch wait for them.  Registers the join.  Falls back to the current stream when
    overlap is disabled or outside a backward pass.

    ``keep``: every tensor fn() reads.  They were allocated on the main stream; the caching
    allocator would hand their memory to the next main-stream allocation as soon as autograd drops
    them, while the (low-priority, possibly long-delayed) side-stream kernels are still reading —
    so the join callback holds them until the main stream has been ordered behind the branch."""
    if not OVERLAP_WEIGHT_GRADS:
        return fn()
    cur = torch.cuda.current_stream(device)
    side = _side_stream(device)
    if side == cur:
        return fn()
    ev = torch.cuda.Event()
    held = [t for t in keep if t is not None]

    def join():
        cur.wait_event(ev)
        now = torch.cuda.current_stream(device)
        if now != cur:
            now.wait_event(ev)
        held.clear()
    try:
        torch.autograd.Variable._execution_engine.queue_callback(join)
    except RuntimeError:            # not inside a backward pass: nothing would join the branch
        return fn()
    if after is not None:
        side.wait_event(after)
    else:
        side.wait_stream(cur)
    with torch.cuda.stream(side):
        out = fn()
        ev.record(side)
    return out


def _ln_fwd_raw(x, eps, relu):
    """(y, stats) = act(LayerNorm(x)) rows, no affine (csrc/fused.cu ln_act_fwd_kernel)."""
    n, d = x.shape
    y = _padded_empty(n, d, x.device)
    stats = torch.empty((n, 2), dtype=torch.float32, device=x.device)
    check(_lib.load().gist_layernorm_act_fwd_f32(ptr(x), _ld(x), n, d, eps, _lib.ACT_RELU if relu else 0, ptr(y), _ld(y),
                                                 ptr(stats), stream_ptr(x.device)), 'layernorm_act_fwd_f32')
    return y, stats


def _ln_bwd_raw(dy, x, stats, relu, want_lo):
    """(dx, dx_lo) of act(LayerNorm(x)) (ln_act_bwd_kernel); dx_lo = 3xTF32 low half or None."""
    n, d = x.shape
    dx = _padded_empty(n, d, x.device)
    dx_lo = _padded_empty(n, d, x.device) if want_lo else None
    check(_lib.load().gist_layernorm_act_bwd_f32(ptr(dy), _ld(dy), ptr(x), _ld(x), ptr(stats), n, d,
                                                 _lib.ACT_RELU if relu else 0, ptr(dx), _ld(dx), ptr(dx_lo),
                                                 _ld(dx_lo) if want_lo else 0, stream_ptr(x.device)),
          'layernorm_act_bwd_f32')
    return dx, dx_lo


class _SageLinear(torch.autograd.Function):
    """h_out = act(LN(dropout([h ‖ (A h) / in_deg]) W^T + b)) — a whole IST SAGE layer
    (cluster_gcn/modules.py:222-236) as ONE autograd node over fused kernels:
      forward : K1 writes z with the dropout mask applied and (3xTF32) z_lo beside it -> K4, whose
                epilogue also applies the layer norm + ReLU when one tile holds the row (out <= 128;
                otherwise the row-wise kernel follows);
      backward: layer-norm backward -> dz = mask ⊙ (dy W) in the K4 epilogue (mask regenerated) -> K2 on
                the CSC gives dh; dW = dy^T z (K4, in-kernel split-K) with db = colsum(dy) from the same
                launch (row sums of dy^T on the tensor core).
    No dropout kernels, no masks in memory, no split passes over z / dy, no bias-gradient kernels."""

    @staticmethod
    def forward(ctx, g, h, W, b, p_drop, stream_id, pre, ln_eps, ln_relu):
        x3 = _MATMUL_PRECISION == '3xtf32'
        if pre is None:
            pre = sage_prepare(g, h.detach(), p_drop, stream_id)
        z, z_lo = pre.z, pre.z_lo
        if x3 and z_lo is None:
            z_lo = split_tf32(z)
        Wv = _tma_view(W)
        W_lo = _weight_lo(Wv) if x3 else None
        x_pre = stats = None
        if ln_eps is None:
            y = gemm(z, Wv, bias=b, A_lo=z_lo, B_lo=W_lo)
        elif ln_fusable(z.shape[0], Wv.shape[0], x3):
            x_pre, y, stats = gemm(z, Wv, bias=b, A_lo=z_lo, B_lo=W_lo, ln=(ln_eps, ln_relu))
        else:
            x_pre = gemm(z, Wv, bias=b, A_lo=z_lo, B_lo=W_lo, out=_padded_empty(z.shape[0], Wv.shape[0], z.device))
            y, stats = _ln_fwd_raw(x_pre, ln_eps, ln_relu)
        ctx.g, ctx.has_bias, ctx.x3 = g, b is not None, x3
        ctx.drop = (float(p_drop), int(stream_id)) if pre.dropped else None
        ctx.step_saved = pre.step_saved
        ctx.ln_relu = bool(ln_relu) if ln_eps is not None else None
        ctx.save_for_backward(z, z_lo, Wv, W_lo, x_pre, stats)
        return y

    @staticmethod
    def backward(ctx, dy):
        z, z_lo, W, W_lo, x_pre, stats = ctx.saved_tensors
        g = ctx.g
        dy = _mat(dy, 'dy')
        if ctx.ln_relu is not None:             # through the layer norm (+ ReLU) first
            dy, dy_lo = _ln_bwd_raw(dy, x_pre, stats, ctx.ln_relu, ctx.x3)
        else:
            dy_lo = _lo_take(dy) if ctx.x3 else None
        if not _tma_ok(dy):
            dy, dy_lo = _tma_view(dy), None
        if ctx.x3 and dy_lo is None:
            dy_lo = split_tf32(dy)
        n, d2 = z.shape
        d = d2 // 2
        dh = dW = db = None
        need_dW, need_db = ctx.needs_input_grad[2], ctx.has_bias and ctx.needs_input_grad[3]
        fused_db = need_db and need_dW and FUSED_ROWSUM

        # first layer (no dh): the main stream has nothing left to do, so a separate bias gradient runs
        # there, beside the dW GEMM, instead of behind it at the tail of the step
        db_here = need_db and not fused_db and not ctx.needs_input_grad[1]

        # a layer that still has its dz -> K2 -> layer-norm chain to run computes dW in the background
        # (a third of the SMs); the first layer's dW is the tail of the step and takes the whole chip
        small = 2.0 * dy.shape[1] * z.shape[1] * n < 4e9         # Reddit-shape dW: ~0.7 GFLOP; ultra-wide: 150 GFLOP
        bgf = _lib.GEMM_BACKGROUND if (ctx.needs_input_grad[1] and BACKGROUND_DW and small) else 0

        def weight_grads():
            if fused_db:        # db = colsum(dy) = row sums of dy^T: from the dW launch itself
                return gemm(dy, z, a_mn=True, b_mn=True, A_lo=dy_lo, B_lo=z_lo, rowsum=True, flags=bgf)
            return (gemm(dy, z, a_mn=True, b_mn=True, A_lo=dy_lo, B_lo=z_lo, flags=bgf) if need_dW else None,
                    colsum(dy) if need_db and not db_here else None)
        # The branch depends on dy only.  Its fork point is recorded NOW, but its kernels are enqueued
        # AFTER this layer's dz GEMM and K2 (DZ_FIRST): a replayed graph hands nodes to the GPU in creation
        # order, and a dW launch created first takes the SMs the latency-critical dz GEMM then waits for
        # (measured in the replayed step: dz 24 us beside an earlier-created dW, 10 us alone).
        dy_ready = None
        if DZ_FIRST and OVERLAP_WEIGHT_GRADS and ctx.needs_input_grad[1] and (need_dW or need_db):
            dy_ready = torch.cuda.Event()
            dy_ready.record(torch.cuda.current_stream(dy.device))
        if dy_ready is None and (need_dW or (need_db and not db_here)):
            dW, db = _weight_grad_branch(dy.device, weight_grads, (dy, dy_lo, z, z_lo))
        if db_here:
            db = colsum(dy)
        if ctx.needs_input_grad[1]:
            if ctx.drop is not None:
                desc = dropout_state(dy.device).desc(ctx.drop[0], ctx.drop[1], step=ctx.step_saved)
                dz = gemm_dropmask(dy, W, desc, b_mn=True, A_lo=dy_lo, B_lo=W_lo)
            else:
                dz = gemm(dy, W, b_mn=True, A_lo=dy_lo, B_lo=W_lo, out=_padded_empty(n, d2, dy.device))
            colptr, row = g.csc()
            dh = torch.empty((n, d), dtype=torch.float32, device=dy.device)
            # dh = dz[:, :d] + A^T (inv_deg ⊙ dz[:, d:])
            spmm_raw(colptr, row, n, n, dz[:, d:], dh, src_scale=g.inv_in_degree(), addend=dz[:, :d],
                     schedule=g.seg_schedule(transpose=True))
        if dy_ready is not None:
            dW, db = _weight_grad_branch(dy.device, weight_grads, (dy, dy_lo, z, z_lo), after=dy_ready)
        return None, dh, dW, db, None, None, None, None, None


def sage_linear(g, h, W, b, p_drop=0.0, stream_id=0, pre=None, ln=None):
    """Fused SAGE aggregation + dropout + linear (+ layer norm + ReLU: ``ln = (eps, relu)``) on the
    tensor-core path ('tf32' / '3xtf32')."""
    assert _MATMUL_PRECISION in ('tf32', '3xtf32')
    require_cuda(h, W, b)
    eps, relu = (float(ln[0]), bool(ln[1])) if ln is not None else (None, False)
    return _SageLinear.apply(g, h, W, b, float(p_drop), int(stream_id), pre, eps, relu)


@torch.no_grad()
def sage_project_first(g, h, W, b):
    """Inference-only form of a SAGE layer whose output is narrower than its input:
    [h ‖ D^-1 A h] W^T + b  ==  h W_l^T + b + D^-1 A (h W_r^T)   (W = [W_l ‖ W_r]),
    so the aggregation runs at the OUTPUT width (full Reddit-shape graph, layer 0: a 256-wide
    gather instead of a 602-wide one; last layer: 41 instead of 256).  One GEMM produces
    P = h [W_l; W_r]^T, one SpMM adds the self term and the bias in its epilogue."""
    require_cuda(h, W, b)
    h = _mat(h, 'h')
    n, d = h.shape
    out_f = W.shape[0]
    assert W.shape[1] == 2 * d
    W2 = torch.cat((W[:, :d], W[:, d:]), dim=0)                  # [2*out, d]
    if _MATMUL_PRECISION == 'fp32':
        P = h @ W2.t()
    else:
        hv, Wv = _tma_view(h), _tma_view(W2)
        x3 = _MATMUL_PRECISION == '3xtf32'
        P = gemm(hv, Wv, A_lo=split_tf32(hv) if x3 else None, B_lo=split_tf32(Wv) if x3 else None)
    y = torch.empty((n, out_f), dtype=torch.float32, device=h.device)
    spmm_raw(g.rowptr, g.col_buffer, n, n, P[:, out_f:], y, dst_scale=g.inv_in_degree(), addend=P[:, :out_f],
             bias=b, schedule=g.seg_schedule())
    return y


def linear(z, W, b=None):
    """F.linear(z, W, b) on the tcgen05 kernel (K4); cuBLAS only in the 'fp32' debug mode.  CUDA only."""
    require_cuda(z, W, b)
    if _MATMUL_PRECISION == '3xtf32':
        return _Linear3xTF32.apply(z, W, b)
    if _MATMUL_PRECISION == 'tf32':
        return _LinearTF32.apply(z, W, b)
    return torch.nn.functional.linear(z, W, b)


# --------------------------------------------------------------------------
# fused row-wise pieces of the training step (csrc/fused.cu)
# --------------------------------------------------------------------------
def colsum(x):
    """x.sum(0) for a 2-D fp32 matrix (bias gradient), fixed-order two-phase reduction."""
    require_cuda(x)
    x = _mat(x, 'x')
    n, d = x.shape
    lib = _lib.load()
    out = torch.empty(d, dtype=torch.float32, device=x.device)
    wsb = lib.gist_colsum_workspace_bytes(n, d)
    ws = torch.empty(max(wsb, 4), dtype=torch.uint8, device=x.device)
    check(lib.gist_colsum_f32(ptr(x), _ld(x), n, d, ptr(out), ptr(ws), wsb, stream_ptr(x.device)), 'colsum_f32')
    return out


class _LayerNormAct(torch.autograd.Function):
    """act(LayerNorm(x)) over the last dimension, no affine (modules.py:234-236), one kernel
    forward and one backward."""

    @staticmethod
    def forward(ctx, x, eps, relu):
        x = _mat(x, 'x')
        n, d = x.shape
        y = torch.empty((n, d), dtype=torch.float32, device=x.device)
        stats = torch.empty((n, 2), dtype=torch.float32, device=x.device)
        fl = _lib.ACT_RELU if relu else 0
        check(_lib.load().gist_layernorm_act_fwd_f32(ptr(x), _ld(x), n, d, eps, fl, ptr(y), _ld(y), ptr(stats),
                                                     stream_ptr(x.device)), 'layernorm_act_fwd_f32')
        ctx.save_for_backward(x, stats)
        ctx.fl = fl
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats = ctx.saved_tensors
        dy = _mat(dy, 'dy')
        n, d = x.shape
        x3 = _MATMUL_PRECISION == '3xtf32'
        dx = _padded_empty(n, d, x.device)
        dx_lo = _padded_empty(n, d, x.device) if x3 else None       # feeds the nn.Linear backward below
        check(_lib.load().gist_layernorm_act_bwd_f32(ptr(dy), _ld(dy), ptr(x), _ld(x), ptr(stats), n, d, ctx.fl,
                                                     ptr(dx), _ld(dx), ptr(dx_lo), _ld(dx_lo) if x3 else 0,
                                                     stream_ptr(x.device)), 'layernorm_act_bwd_f32')
        if x3:
            _lo_put(dx, dx_lo)
        return dx, None, None


def layer_norm_act(x, eps=1e-5, relu=False):
    """F.relu(F.layer_norm(x, (x.shape[-1],), eps=eps)) (or without the ReLU) for 2-D fp32 x."""
    require_cuda(x)
    return _LayerNormAct.apply(x, float(eps), bool(relu))


class _TensorLayerNorm(torch.autograd.Function):
    """F.layer_norm(x, x.shape): one mean / variance over the whole [n, d] tensor, no affine
    (gcn/gcn.py:65-66, cluster_gcn/modules.py:347-348)."""

    @staticmethod
    def forward(ctx, x, eps):
        x = _mat(x, 'x')
        n, d = x.shape
        lib = _lib.load()
        y = torch.empty((n, d), dtype=torch.float32, device=x.device)
        stats = torch.empty(2, dtype=torch.float32, device=x.device)
        wsb = lib.gist_tensor_layernorm_workspace_bytes(n, d)
        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=x.device)
        check(lib.gist_tensor_layernorm_fwd_f32(ptr(x), _ld(x), n, d, eps, ptr(y), _ld(y), ptr(stats), ptr(ws), wsb,
                                                stream_ptr(x.device)), 'tensor_layernorm_fwd_f32')
        ctx.save_for_backward(x, stats)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats = ctx.saved_tensors
        dy = _mat(dy, 'dy')
        n, d = x.shape
        lib = _lib.load()
        dx = torch.empty((n, d), dtype=torch.float32, device=x.device)
        wsb = lib.gist_tensor_layernorm_workspace_bytes(n, d)
        ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device=x.device)
        check(lib.gist_tensor_layernorm_bwd_f32(ptr(dy), _ld(dy), ptr(x), _ld(x), ptr(stats), n, d, ptr(dx), _ld(dx),
                                                ptr(ws), wsb, stream_ptr(x.device)), 'tensor_layernorm_bwd_f32')
        return dx, None


def tensor_layer_norm(x, eps=1e-5):
    """F.layer_norm(x, x.shape, eps=eps) for a 2-D fp32 matrix."""
    require_cuda(x)
    return _TensorLayerNorm.apply(x, float(eps))


class _MaskedCrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, mask):
        logits = _mat(logits, 'logits')
        n, C = logits.shape
        assert labels.dtype == torch.int64 and labels.shape == (n,)
        labels = labels.contiguous()
        if mask is not None:
            assert mask.dtype == torch.bool and mask.shape == (n,)
            mask = mask.contiguous()
        dev = logits.device
        lse = torch.empty(n, dtype=torch.float32, device=dev)
        row_loss = torch.empty(n, dtype=torch.float32, device=dev)
        out = torch.empty(2, dtype=torch.float32, device=dev)
        check(_lib.load().gist_masked_ce_fwd_f32(ptr(logits), _ld(logits), n, C, ptr(labels), ptr(mask), ptr(lse),
                                                 ptr(row_loss), ptr(out), stream_ptr(dev)), 'masked_ce_fwd_f32')
        ctx.save_for_backward(logits, labels, mask, lse, out)
        return out[0]

    @staticmethod
    def backward(ctx, gout):
        logits, labels, mask, lse, out = ctx.saved_tensors
        n, C = logits.shape
        ldd = (C + 3) // 4 * 4        # padded rows: the gradient feeds the TMA-addressed GEMMs
        buf = torch.empty((n, ldd), dtype=torch.float32, device=logits.device)
        x3 = _MATMUL_PRECISION == '3xtf32'
        buf_lo = torch.empty((n, ldd), dtype=torch.float32, device=logits.device) if x3 else None
        gout = gout.contiguous().float()
        check(_lib.load().gist_masked_ce_bwd_f32(ptr(logits), _ld(logits), n, C, ptr(labels), ptr(mask), ptr(lse),
                                                 ptr(out), ptr(gout), ptr(buf), ldd, ldd, ptr(buf_lo),
                                                 stream_ptr(logits.device)), 'masked_ce_bwd_f32')
        dl = buf[:, :C]
        if x3:
            _lo_put(dl, buf_lo[:, :C])
        return dl, None, None


def masked_cross_entropy(logits, labels, mask=None):
    """CrossEntropyLoss()(logits[mask], labels[mask]) (…distrib.py:413-414) as a 0-d tensor: two
    kernels forward, one backward, no boolean-index gather (hence no host sync)."""
    require_cuda(logits, labels, mask)
    return _MaskedCrossEntropy.apply(logits, labels, mask)


_CE_SYNC = {}


def _ce_sync(device):
    """Arrival counter of the fused CE kernel: one zeroed uint32 per stream (the kernel leaves it
    zero; two launches in flight on different streams must not share it)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    t = _CE_SYNC.get(key)
    if t is None:
        t = _CE_SYNC[key] = torch.zeros(1, dtype=torch.int32, device=device)
    return t


@torch.no_grad()
def masked_ce_loss_and_grad(logits, labels, mask=None, out=None):
    """(loss, dlogits) of ``CrossEntropyLoss()(logits[mask], labels[mask])`` in ONE launch — the
    trainers' replacement for ``loss = ...; loss.backward()`` (…distrib.py:413-415): call
    ``logits.backward(dlogits)`` to run the rest of the backward pass.  dlogits is the gradient for
    an upstream gradient of 1; its 3xTF32 low half is handed to the consuming linear backward.
    ``out``: a caller-owned fp32 tensor of 2 elements the kernel writes (loss, 1 / #rows) into — a
    trainer's persistent loss slot, so no copy node follows the step."""
    require_cuda(logits, labels, mask, out)
    logits = _mat(logits.detach(), 'logits')
    n, C = logits.shape
    assert n > 0
    assert labels.dtype == torch.int64 and labels.shape == (n,)
    labels = labels.contiguous()
    if mask is not None:
        assert mask.dtype == torch.bool and mask.shape == (n,)
        mask = mask.contiguous()
    dev = logits.device
    lib = _lib.load()
    ldd = (C + 3) // 4 * 4            # padded rows: the gradient feeds the TMA-addressed GEMMs
    buf = torch.empty((n, ldd), dtype=torch.float32, device=dev)
    x3 = _MATMUL_PRECISION == '3xtf32'
    buf_lo = torch.empty((n, ldd), dtype=torch.float32, device=dev) if x3 else None
    if out is None:
        out = torch.empty(2, dtype=torch.float32, device=dev)
    assert out.dtype == torch.float32 and out.shape == (2,) and out.is_contiguous()
    wsb = lib.gist_masked_ce_fused_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    check(lib.gist_masked_ce_fused_f32(ptr(logits), _ld(logits), n, C, ptr(labels), ptr(mask), ptr(buf), ldd, ldd,
                                       ptr(buf_lo), ptr(out), ptr(ws), wsb, ptr(_ce_sync(dev)), stream_ptr(dev)),
          'masked_ce_fused_f32')
    dl = buf[:, :C]
    if x3:
        _lo_put(dl, buf_lo[:, :C])
    return out[0], dl


# --------------------------------------------------------------------------
# K6: GAT edge-softmax + weighted aggregation (csrc/gat.cu)
# --------------------------------------------------------------------------
class _GatAggregate(torch.autograd.Function):
    """out[v] = sum_u softmax_u(leaky_relu(a_l.z[u] + a_r.z[v])) z[u]  for one attention head
    (cluster_gcn/modules.py:40-65).  attn is attn_fc.weight ([1, 2D] or [2D])."""

    @staticmethod
    def forward(ctx, g, z, attn, negative_slope):
        z = _mat(z, 'z')
        n, D = z.shape
        assert n == g.number_of_nodes() and attn.numel() == 2 * D
        attn = attn.reshape(-1).contiguous()
        dev = z.device
        lib = _lib.load()
        scores = torch.empty((n, 2), dtype=torch.float32, device=dev)
        out = torch.empty((n, D), dtype=torch.float32, device=dev)
        lse = torch.empty(n, dtype=torch.float32, device=dev)
        check(lib.gist_gat_scores_f32(ptr(z), _ld(z), n, D, ptr(attn), ptr(scores), stream_ptr(dev)),
              'gat_scores_f32')
        prof = GAT_PROFILE
        if prof is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        check(lib.gist_gat_aggregate_f32(ptr(g.rowptr), ptr(g.col_buffer), n, ptr(z), _ld(z), D, ptr(scores),
                                         negative_slope, ptr(out), _ld(out), ptr(lse), stream_ptr(dev)),
              'gat_aggregate_f32')
        if prof is not None:
            ev1.record()
            prof.append(dict(ev0=ev0, ev1=ev1, rowptr=g.rowptr, n=n, D=D))
        ctx.g, ctx.slope = g, negative_slope
        ctx.save_for_backward(z, attn, scores, lse, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        z, attn, scores, lse, out = ctx.saved_tensors
        g = ctx.g
        dout = _mat(dout, 'dout')
        n, D = z.shape
        dev = z.device
        lib = _lib.load()
        colptr, row = g.csc()
        dz = torch.empty((n, D), dtype=torch.float32, device=dev)
        dattn = torch.empty(2 * D, dtype=torch.float32, device=dev)
        wsb = lib.gist_gat_backward_workspace_bytes(n, D)
        ws = torch.empty(max(wsb, 4), dtype=torch.uint8, device=dev)
        check(lib.gist_gat_backward_f32(ptr(g.rowptr), ptr(g.col_buffer), ptr(colptr), ptr(row), n, ptr(z), _ld(z),
                                        D, ptr(scores), ptr(lse), ptr(attn), ctx.slope, ptr(out), _ld(out),
                                        ptr(dout), _ld(dout), ptr(dz), _ld(dz), ptr(dattn), ptr(ws), wsb,
                                        stream_ptr(dev)), 'gat_backward_f32')
        return None, dz, dattn, None


class _GatAggregateHeads(torch.autograd.Function):
    """All H heads of a MultiHeadGATLayer in one launch per kernel: z [n, H*D] (head h = columns
    [h D, (h+1) D), the heads' projections being ONE GEMM), attn [H, 2D] -> out [n, H*D]."""

    @staticmethod
    def forward(ctx, g, z, attn, heads, negative_slope):
        z = _mat(z, 'z')
        n, W = z.shape
        D = W // heads
        assert n == g.number_of_nodes() and W == D * heads and tuple(attn.shape) == (heads, 2 * D)
        attn = attn.contiguous()
        dev = z.device
        lib = _lib.load()
        scores = torch.empty((heads, n, 2), dtype=torch.float32, device=dev)
        out = _padded_empty(n, W, dev)
        lse = torch.empty((heads, n), dtype=torch.float32, device=dev)
        check(lib.gist_gat_scores_heads_f32(ptr(z), _ld(z), n, D, heads, ptr(attn), ptr(scores), stream_ptr(dev)),
              'gat_scores_heads_f32')
        prof = GAT_PROFILE
        if prof is not None:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        check(lib.gist_gat_aggregate_heads_f32(ptr(g.rowptr), ptr(g.col_buffer), n, ptr(z), _ld(z), D, heads, ptr(scores),
                                               negative_slope, ptr(out), _ld(out), ptr(lse), stream_ptr(dev)),
              'gat_aggregate_heads_f32')
        if prof is not None:
            ev1.record()
            prof.append(dict(ev0=ev0, ev1=ev1, rowptr=g.rowptr, n=n, D=D, heads=heads))
        ctx.g, ctx.slope, ctx.heads = g, negative_slope, heads
        ctx.save_for_backward(z, attn, scores, lse, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        z, attn, scores, lse, out = ctx.saved_tensors
        g, heads = ctx.g, ctx.heads
        dout = _mat(dout, 'dout')
        n, W = z.shape
        D = W // heads
        dev = z.device
        lib = _lib.load()
        colptr, row = g.csc()
        dz = _padded_empty(n, W, dev)
        dattn = torch.empty((heads, 2 * D), dtype=torch.float32, device=dev)
        wsb = lib.gist_gat_backward_heads_workspace_bytes(n, D, heads)
        ws = torch.empty(max(wsb, 4), dtype=torch.uint8, device=dev)
        check(lib.gist_gat_backward_heads_f32(ptr(g.rowptr), ptr(g.col_buffer), ptr(colptr), ptr(row), n, ptr(z), _ld(z),
                                              D, heads, ptr(scores), ptr(lse), ptr(attn), ctx.slope, ptr(out), _ld(out),
                                              ptr(dout), _ld(dout), ptr(dz), _ld(dz), ptr(dattn), ptr(ws), wsb,
                                              stream_ptr(dev)), 'gat_backward_heads_f32')
        return None, dz, dattn, None, None


def gat_aggregate_heads(g, z, attn, heads, negative_slope=0.01):
    """z [n, heads * D], attn [heads, 2D] -> [n, heads * D] (see _GatAggregateHeads)."""
    require_cuda(z, attn, g.rowptr)
    return _GatAggregateHeads.apply(g, z, attn, int(heads), float(negative_slope))


def gat_aggregate(g, z, attn, negative_slope=0.01):
    require_cuda(z, attn, g.rowptr)
    return _GatAggregate.apply(g, z, attn.reshape(-1), float(negative_slope))
