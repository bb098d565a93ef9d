"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN PYTHON FILES.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):      python oracle/gen_golden.py

How: the reference imports `dgl` (absent).  We put oracle/dgl_restated on
sys.path so `import dgl` binds to the CPU restatement, stub the plotting /
dataset-ingestion imports that are out of scope (matplotlib, AmazonDataset ->
tensorflow), and then import the reference modules *unmodified* from
/root/reference: cluster_gcn/modules.py, sampler.py, partition_utils.py,
cluster_gcn_ist_distrib.py, gcn/gcn.py, gcn/train_ist.py.  Their outputs on
small seeded inputs are the golden vectors.  So the goldens pin everything
ABOVE the DGL boundary to the reference's real code; DGL semantics themselves
stay unpinned (see oracle/README.md).

The multi-rank dispatch/sync golden runs the reference's DistributedGNNWrapper
on m real processes over the gloo backend.
"""
import os
import random
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = '/root/reference'
OUT = os.path.join(ROOT, 'tests', 'golden')


def bind_reference(subdir):
    """Make `import dgl`, `import modules`, ... resolve to restated DGL + reference files."""
    for p in (os.path.join(HERE, 'dgl_restated'), os.path.join(REF, subdir), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    mpl = types.ModuleType('matplotlib')
    plt = types.ModuleType('matplotlib.pyplot')
    for name in ('plot', 'title', 'savefig', 'xlabel', 'ylabel', 'close'):
        setattr(plt, name, lambda *a, **k: None)
    mpl.pyplot = plt
    sys.modules.setdefault('matplotlib', mpl)
    sys.modules.setdefault('matplotlib.pyplot', plt)
    amz = types.ModuleType('AmazonDataset')
    amz.AmazonDataset = type('AmazonDataset', (), {})
    sys.modules.setdefault('AmazonDataset', amz)


def small_graph(n, nnz, seed, loops):
    rng = np.random.RandomState(seed)
    src = rng.randint(0, n, nnz)
    dst = rng.randint(0, n - 3, nnz)          # last 3 nodes: no in-edges (unless loops)
    src[:4], dst[:4] = src[4:8], dst[4:8]     # duplicate edges
    if loops:
        src = np.concatenate([src, np.arange(n)])
        dst = np.concatenate([dst, np.arange(n)])
    return src.astype(np.int64), dst.astype(np.int64)


def sd_np(module):
    return {k: v.detach().numpy().copy() for k, v in module.state_dict().items()}


# --------------------------------------------------------------------------
def gen_sage(out):
    """cluster_gcn/modules.py GCN (ISTSAGELayer stack): logits + CE gradients."""
    import dgl
    import modules
    cases = [dict(n=60, nnz=500, fin=12, hid=16, ncls=5, L=2, ln=True, seed=1),
             dict(n=90, nnz=900, fin=7, hid=8, ncls=3, L=1, ln=False, seed=2),
             dict(n=70, nnz=800, fin=10, hid=12, ncls=4, L=3, ln=True, seed=3)]
    for ci, c in enumerate(cases):
        src, dst = small_graph(c['n'], c['nnz'], c['seed'], loops=False)
        g = dgl.DGLGraph((src, dst), num_nodes=c['n'])
        torch.manual_seed(c['seed'])
        x = torch.randn(c['n'], c['fin'])
        y = torch.randint(0, c['ncls'], (c['n'],))
        g.ndata['feat'] = x
        model = modules.GCN(c['fin'], c['hid'], c['ncls'], c['L'], F.relu, 0.3, c['ln'],
                            False, False, 1, True)
        model.eval()
        logits = model(g)
        loss = F.cross_entropy(logits, y)
        loss.backward()
        p = 'sage%d_' % ci
        out[p + 'src'], out[p + 'dst'], out[p + 'n'] = src, dst, np.int64(c['n'])
        out[p + 'x'], out[p + 'y'] = x.numpy(), y.numpy()
        out[p + 'cfg'] = np.array([c['fin'], c['hid'], c['ncls'], c['L'], int(c['ln'])])
        out[p + 'logits'] = logits.detach().numpy()
        out[p + 'loss'] = np.float64(loss.item())
        for k, v in sd_np(model).items():
            out[p + 'param.' + k] = v
        for k, v in model.named_parameters():
            out[p + 'grad.' + k] = v.grad.numpy().copy()
    # one layer, in isolation, relu + layernorm, plus the norm vector
    src, dst = small_graph(40, 300, 9, loops=False)
    g = dgl.DGLGraph((src, dst), num_nodes=40)
    torch.manual_seed(9)
    layer = modules.ISTSAGELayer(6, 10, 0.0, True, activation=F.relu)
    x = torch.randn(40, 6)
    out['layer_src'], out['layer_dst'] = src, dst
    out['layer_x'] = x.numpy()
    out['layer_w'] = layer.linear.weight.detach().numpy()
    out['layer_b'] = layer.linear.bias.detach().numpy()
    out['layer_out'] = layer(g, x).detach().numpy()
    out['layer_norm'] = layer.get_norm(g).numpy()


def gen_graphconv(out):
    """gcn/gcn.py GCN and cluster_gcn BaselineGCN on restated GraphConv."""
    import dgl
    import modules
    sys.path.insert(0, os.path.join(REF, 'gcn'))
    import gcn as ref_gcn
    cases = [dict(n=80, nnz=600, fin=30, hid=8, ncls=4, L=1, si=False, so=False, k=1, seed=4),
             dict(n=80, nnz=600, fin=24, hid=16, ncls=3, L=2, si=False, so=True, k=4, seed=5),
             dict(n=64, nnz=500, fin=16, hid=16, ncls=5, L=2, si=True, so=True, k=2, seed=6)]
    for ci, c in enumerate(cases):
        src, dst = small_graph(c['n'], c['nnz'], c['seed'], loops=True)
        g = dgl.DGLGraph((src, dst), num_nodes=c['n'])
        torch.manual_seed(c['seed'])
        model = ref_gcn.GCN(g, c['fin'], c['hid'], c['ncls'], c['L'], F.relu, 0.5, True,
                            c['si'], c['so'], c['k'])
        with torch.no_grad():
            for l in model.layers:
                l.bias.uniform_(-0.3, 0.3)
        fin_eff = c['fin'] // c['k'] if c['si'] else c['fin']
        x = torch.randn(c['n'], fin_eff)
        y = torch.randint(0, c['ncls'], (c['n'],))
        model.eval()
        logits = model(x)
        loss = F.cross_entropy(logits, y)
        loss.backward()
        p = 'gc%d_' % ci
        out[p + 'src'], out[p + 'dst'], out[p + 'n'] = src, dst, np.int64(c['n'])
        out[p + 'x'], out[p + 'y'] = x.numpy(), y.numpy()
        out[p + 'cfg'] = np.array([c['fin'], c['hid'], c['ncls'], c['L'], int(c['si']), int(c['so']), c['k']])
        out[p + 'logits'] = logits.detach().numpy()
        for k, v in sd_np(model).items():
            out[p + 'param.' + k] = v
        for k, v in model.named_parameters():
            out[p + 'grad.' + k] = v.grad.numpy().copy()
    # BaselineGCN (cluster_gcn/modules.py:316-349)
    src, dst = small_graph(50, 400, 7, loops=True)
    g = dgl.DGLGraph((src, dst), num_nodes=50)
    torch.manual_seed(7)
    x = torch.randn(50, 9)
    g.ndata['feat'] = x
    model = modules.BaselineGCN(9, 12, 4, 2, F.relu, 0.5, True)
    model.eval()
    out['base_src'], out['base_dst'], out['base_x'] = src, dst, x.numpy()
    out['base_logits'] = model(g).detach().numpy()
    for k, v in sd_np(model).items():
        out['base_param.' + k] = v


def gen_partition(out):
    """create_partition from cluster_gcn_ist_distrib.py for (seed, m, size) triples."""
    import cluster_gcn_ist_distrib as ref
    triples = [(0, 2, 16), (3, 8, 256), (3, 4, 64), (7, 1, 8), (11, 8, 2048)]
    out['cp_triples'] = np.array(triples)
    for t, (seed, m, size) in enumerate(triples):
        random.seed(seed)
        first = ref.create_partition(m, size)
        second = ref.create_partition(m, size)     # stream continues
        for tag, part in (('a', first), ('b', second)):
            out['cp%d%s_idx' % (t, tag)] = np.stack([p[0].numpy() for p in part])
            out['cp%d%s_full' % (t, tag)] = np.stack([p[1].numpy() for p in part])


def gen_cluster_iter(out):
    """sampler.py ClusterIter + partition_utils.get_subgraph on restated DGL."""
    import dgl
    import sampler as ref_sampler
    from gist_b200 import synth
    ds = synth.make('reddit', seed=1, device='cpu', scale=0.004, feat_dim=8)
    g = dgl.DGLGraph((ds.src, ds.dst), num_nodes=ds.num_nodes)
    g.ndata['feat'] = ds.feat
    g.ndata['label'] = ds.label
    g.ndata['train_mask'] = ds.train_mask
    g.ndata['_part'] = ds.part
    train_nid = np.nonzero(ds.train_mask.numpy())[0].astype(np.int64)
    psize, bs = int(ds.part.max()) + 1, 2
    random.seed(5)
    it = ref_sampler.ClusterIter('', g, psize, bs, train_nid, use_pp=False)
    out['ci_src'], out['ci_dst'], out['ci_n'] = ds.src.numpy(), ds.dst.numpy(), np.int64(ds.num_nodes)
    out['ci_part'], out['ci_train_nid'] = ds.part.numpy(), train_nid
    out['ci_feat'], out['ci_label'] = ds.feat.numpy(), ds.label.numpy()
    out['ci_cfg'] = np.array([psize, bs, 5])
    k = 0
    for epoch in range(2):
        for batch in it:
            nid = batch.ndata[dgl.NID].numpy()
            s, d = batch.edges()
            key = np.sort(d.numpy() * len(nid) + s.numpy())
            out['ci_b%d_nid' % k] = nid
            out['ci_b%d_edges' % k] = key            # canonical: sorted dst*n_b+src
            out['ci_b%d_feat0' % k] = batch.ndata['feat'][:, 0].numpy()
            k += 1
    out['ci_nbatches'] = np.int64(k)
    # what the python-random stream yields next (pins the call order)
    out['ci_next_random'] = np.float64(random.random())


def _wrapper_worker(rank, m, port, cfg, q):
    bind_reference('cluster_gcn')
    import torch.distributed as dist
    import cluster_gcn_ist_distrib as ref
    fin, hid, ncls, L, seed = cfg
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=m)
    args = types.SimpleNamespace(rank=rank, num_subnet=m, n_hidden=hid, n_layers=L, dropout=0.0,
                                 use_layernorm=True)
    w = ref.DistributedGNNWrapper(args, None, fin, ncls, torch.device('cpu'))
    rec = {}
    if rank == 0:
        for k, v in sd_np(w.base_model).items():
            rec['base0.' + k] = v
    w.ini_sync_dispatch_model()
    for l, layer_parts in enumerate(w.current_partition):
        rec['part0.%d' % l] = np.stack([p[0].numpy() for p in layer_parts])
    for k, v in sd_np(w.sub_model).items():
        rec['sub0.' + k] = v
    # deterministic stand-in for local training
    with torch.no_grad():
        for li, lyr in enumerate(w.sub_model.layers):
            lyr.linear.weight.data = lyr.linear.weight.data * 1.25 + 0.01 * (rank + 1) * (li + 1)
            lyr.linear.bias.data = lyr.linear.bias.data - 0.125 * (rank + 1)
    for k, v in sd_np(w.sub_model).items():
        rec['trained.' + k] = v
    dist.barrier()
    w.sync_model()
    rec['lastbias_after_sync'] = w.sub_model.layers[-1].linear.bias.detach().numpy().copy()
    if rank == 0:
        for k, v in sd_np(w.base_model).items():
            rec['base1.' + k] = v
    dist.barrier()
    w.dispatch_model()
    for l, layer_parts in enumerate(w.current_partition):
        rec['part1.%d' % l] = np.stack([p[0].numpy() for p in layer_parts])
    for k, v in sd_np(w.sub_model).items():
        rec['sub1.' + k] = v
    q.put((rank, rec))
    dist.barrier()
    dist.destroy_process_group()


def gen_wrapper(out):
    """DistributedGNNWrapper.{ini_sync_dispatch_model, sync_model, dispatch_model} run by
    the reference on m gloo processes."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    for ci, (m, cfg) in enumerate([(2, (10, 8, 3, 2, 3)), (4, (6, 16, 5, 3, 0))]):
        q = ctx.Queue()
        procs = [ctx.Process(target=_wrapper_worker, args=(r, m, 29610 + ci, cfg, q)) for r in range(m)]
        for p in procs:
            p.start()
        recs = dict(q.get(timeout=300) for _ in range(m))
        for p in procs:
            p.join(60)
        out['w%d_cfg' % ci] = np.array((m,) + cfg)
        for r in range(m):
            for k, v in recs[r].items():
                out['w%d_r%d_%s' % (ci, r, k)] = v


def gen_train_ist(out):
    """gcn/train_ist.py main(): run the reference trainer on a tiny synthetic dataset
    and record what it splits and merges (patched torch.randperm / load_state_dict
    only OBSERVE)."""
    bind_reference('gcn')
    import dgl
    import networkx as nx
    import gcn as ref_gcn
    sys.modules['models'] = ref_gcn            # train_ist.py:13 imports a module that does not exist
    import train_ist as ref

    def loader(args):
        rng = np.random.RandomState(0)
        n, f, c = 60, 12, 3
        gnx = nx.gnm_random_graph(n, 150, seed=0)
        feats = rng.rand(n, f).astype(np.float32)
        labels = rng.randint(0, c, n)
        mask = np.zeros(n, dtype=np.uint8)
        tr, va, te = mask.copy(), mask.copy(), mask.copy()
        tr[:20], va[20:40], te[40:] = 1, 1, 1
        return types.SimpleNamespace(graph=gnx, features=feats, labels=labels, train_mask=tr,
                                     val_mask=va, test_mask=te, num_labels=c)
    dgl.data._LOADER = loader
    ref.load_data = loader

    for ci, (si, so, L) in enumerate([('False', 'True', 2), ('True', 'True', 2), ('False', 'False', 2),
                                      ('True', 'False', 1)]):
        log = {'perms': [], 'loads': []}
        orig_randperm = torch.randperm
        orig_load = ref_gcn.GCN.load_state_dict
        instances = []
        orig_init = ref_gcn.GCN.__init__

        def randperm(n, *a, **k):
            p = orig_randperm(n, *a, **k)
            log['perms'].append(p.clone())
            return p

        def init(self, *a, **k):
            orig_init(self, *a, **k)
            instances.append(self)

        def load(self, sd, *a, **k):
            log['loads'].append((len(instances), self is instances[0],
                                 {kk: vv.detach().clone() for kk, vv in sd.items()},
                                 {kk: vv.detach().clone() for kk, vv in instances[0].state_dict().items()},
                                 [{kk: vv.detach().clone() for kk, vv in inst.state_dict().items()}
                                  for inst in instances[-4:]]))
            return orig_load(self, sd, *a, **k)

        torch.randperm = randperm
        ref_gcn.GCN.__init__ = init
        ref_gcn.GCN.load_state_dict = load
        try:
            torch.manual_seed(ci)
            args = types.SimpleNamespace(dataset='cora', use_ist='True', iter_per_site=2, num_subnet=4,
                                         dropout=0.0, split_output=so, split_input=si, gpu=-1, lr=0.01,
                                         n_epochs=4, n_hidden=8, n_layers=L, weight_decay=5e-4,
                                         self_loop='True', use_layernorm='True', use_random_proj='False')
            ref.main(args)
        finally:
            torch.randperm = orig_randperm
            ref_gcn.GCN.__init__ = orig_init
            ref_gcn.GCN.load_state_dict = orig_load
        # loads: per round 4 sub-model loads then 1 main-model load (the merge)
        p = 'ti%d_' % ci
        out[p + 'cfg'] = np.array([int(si == 'True'), int(so == 'True'), L, 4, 8, 12, 3])
        for i, perm in enumerate(log['perms']):
            out[p + 'perm%d' % i] = perm.numpy()
        out[p + 'nperm'] = np.int64(len(log['perms']))
        main_loads = [x for x in log['loads'] if x[1]]
        sub_loads = [x for x in log['loads'] if not x[1]]
        out[p + 'nrounds'] = np.int64(len(main_loads))
        for r in range(len(main_loads)):
            for k, v in sub_loads[4 * r][3].items():      # full model when round r was split
                out[p + 'r%d_main.%s' % (r, k)] = v.numpy()
        for r, (_, _, merged, _, subs) in enumerate(main_loads):
            for k, v in merged.items():
                out[p + 'r%d_merged.%s' % (r, k)] = v.numpy()
            for s, sd in enumerate(subs):
                for k, v in sd.items():
                    out[p + 'r%d_trained%d.%s' % (r, s, k)] = v.numpy()
        for i, (_, _, sd, _, _) in enumerate(sub_loads):
            for k, v in sd.items():
                out[p + 'split%d.%s' % (i, k)] = v.numpy()
        out[p + 'nsplit'] = np.int64(len(sub_loads))


def gen_train_ist_data(out):
    """The tiny dataset gen_train_ist's loader hands to gcn/train_ist.py (so the GPU trainer test
    does not depend on networkx reproducing the same random graph): directed edge list after the
    reference's self-loop step (train_ist.py:110-113), features, labels, masks."""
    import networkx as nx
    rng = np.random.RandomState(0)
    n, f, c = 60, 12, 3
    gnx = nx.gnm_random_graph(n, 150, seed=0)
    feats = rng.rand(n, f).astype(np.float32)
    labels = rng.randint(0, c, n)
    gnx.remove_edges_from(nx.selfloop_edges(gnx))
    gnx.add_edges_from(zip(gnx.nodes(), gnx.nodes()))
    e = np.asarray(list(gnx.to_directed().edges()), dtype=np.int64)
    out['src'], out['dst'] = e[:, 0].copy(), e[:, 1].copy()
    out['features'], out['labels'] = feats, labels.astype(np.int64)
    out['n'], out['num_labels'] = np.int64(n), np.int64(c)


def gen_gat(out):
    """cluster_gcn/modules.py GATLayer (one head) run unmodified on the restated DGL's UDF /
    degree-bucketing path: output and gradients.  MultiHeadGATLayer / GAT as committed cannot
    run (scalar mean), so the multi-head golden stacks the reference's own
    GATLayer heads and takes the intended mean over the head axis."""
    import dgl
    import modules
    cases = [dict(n=50, nnz=400, fin=9, D=8, seed=11, loops=True),
             dict(n=70, nnz=900, fin=12, D=20, seed=12, loops=False),
             dict(n=40, nnz=300, fin=5, D=3, seed=13, loops=True)]
    for ci, c in enumerate(cases):
        src, dst = small_graph(c['n'], c['nnz'], c['seed'], loops=c['loops'])
        g = dgl.DGLGraph((src, dst), num_nodes=c['n'])
        torch.manual_seed(c['seed'])
        layer = modules.GATLayer(c['fin'], c['D'])
        x = torch.randn(c['n'], c['fin'], requires_grad=True)
        wy = torch.randn(c['n'], c['D'])
        y = layer(g, x)
        (y * wy).sum().backward()
        p = 'gat%d_' % ci
        out[p + 'src'], out[p + 'dst'], out[p + 'n'] = src, dst, np.int64(c['n'])
        out[p + 'x'], out[p + 'wy'] = x.detach().numpy(), wy.numpy()
        out[p + 'fc'] = layer.fc.weight.detach().numpy()
        out[p + 'attn'] = layer.attn_fc.weight.detach().numpy()
        out[p + 'out'] = y.detach().numpy()
        out[p + 'dx'] = x.grad.numpy().copy()
        out[p + 'dfc'] = layer.fc.weight.grad.numpy().copy()
        out[p + 'dattn'] = layer.attn_fc.weight.grad.numpy().copy()
    # two layers x two heads, intended head mean + ELU (modules.py:93-98)
    src, dst = small_graph(60, 600, 14, loops=True)
    g = dgl.DGLGraph((src, dst), num_nodes=60)
    torch.manual_seed(14)
    l0 = [modules.GATLayer(7, 6) for _ in range(2)]
    l1 = [modules.GATLayer(6, 4)]
    x = torch.randn(60, 7)
    h = x
    for heads in (l0, l1):
        h = F.elu(torch.stack([hd(g, h) for hd in heads]).mean(dim=0))
    out['mh_src'], out['mh_dst'], out['mh_n'], out['mh_x'] = src, dst, np.int64(60), x.numpy()
    for li, heads in enumerate((l0, l1)):
        for hi, hd in enumerate(heads):
            out['mh_fc_%d_%d' % (li, hi)] = hd.fc.weight.detach().numpy()
            out['mh_attn_%d_%d' % (li, hi)] = hd.attn_fc.weight.detach().numpy()
    out['mh_out'] = h.detach().numpy()


def gen_graphsage(out):
    """cluster_gcn/modules.py:100-189 — GraphSAGELayer (affine LayerNorm, optional bias, use_pp) and the
    GraphSAGE container; plus BaselineGCN (:316-349) with gradients."""
    import dgl
    import modules
    # single layers: (bias, use_pp, use_lynorm, train-mode?) — dropout 0 so train mode is deterministic
    lcases = [dict(bias=True, pp=False, ln=True, train=False, act=True, seed=11),
              dict(bias=False, pp=False, ln=True, train=True, act=True, seed=12),
              dict(bias=True, pp=True, ln=True, train=True, act=True, seed=13),      # use_pp: no aggregation
              dict(bias=True, pp=True, ln=False, train=False, act=False, seed=14)]    # use_pp in eval aggregates
    for ci, c in enumerate(lcases):
        src, dst = small_graph(48, 380, c['seed'], loops=False)
        g = dgl.DGLGraph((src, dst), num_nodes=48)
        torch.manual_seed(c['seed'])
        fin, fout = 6, 9
        layer = modules.GraphSAGELayer(fin, fout, F.relu if c['act'] else None, 0.0, bias=c['bias'],
                                       use_pp=c['pp'], use_lynorm=c['ln'])
        if c['ln']:
            with torch.no_grad():
                layer.lynorm.weight.uniform_(0.5, 1.5)
                layer.lynorm.bias.uniform_(-0.3, 0.3)
        layer.train(c['train'])
        skip_agg = c['pp'] and c['train']
        x = torch.randn(48, 2 * fin if skip_agg else fin, requires_grad=True)
        y = layer(g, x)
        wy = torch.randn(48, fout)
        (y * wy).sum().backward()
        p = 'gl%d_' % ci
        out[p + 'src'], out[p + 'dst'] = src, dst
        out[p + 'cfg'] = np.array([fin, fout, int(c['bias']), int(c['pp']), int(c['ln']), int(c['train']), int(c['act'])])
        out[p + 'x'], out[p + 'wy'], out[p + 'out'] = x.detach().numpy(), wy.numpy(), y.detach().numpy()
        out[p + 'dx'] = x.grad.numpy().copy()
        for k, v in sd_np(layer).items():
            out[p + 'param.' + k] = v
        for k, v in layer.named_parameters():
            out[p + 'grad.' + k] = v.grad.numpy().copy()
    # GraphSAGE containers
    for ci, c in enumerate([dict(n=70, nnz=700, fin=10, hid=12, ncls=4, L=2, pp=False, seed=21),
                            dict(n=56, nnz=500, fin=8, hid=8, ncls=3, L=1, pp=False, seed=22)]):
        src, dst = small_graph(c['n'], c['nnz'], c['seed'], loops=False)
        g = dgl.DGLGraph((src, dst), num_nodes=c['n'])
        torch.manual_seed(c['seed'])
        x = torch.randn(c['n'], c['fin'])
        y = torch.randint(0, c['ncls'], (c['n'],))
        g.ndata['feat'] = x
        model = modules.GraphSAGE(c['fin'], c['hid'], c['ncls'], c['L'], F.relu, 0.4, c['pp'])
        with torch.no_grad():
            for l in model.layers[:-1]:
                l.lynorm.weight.uniform_(0.5, 1.5)
                l.lynorm.bias.uniform_(-0.3, 0.3)
        model.eval()
        logits = model(g)
        F.cross_entropy(logits, y).backward()
        p = 'gs%d_' % ci
        out[p + 'src'], out[p + 'dst'], out[p + 'n'] = src, dst, np.int64(c['n'])
        out[p + 'x'], out[p + 'y'] = x.numpy(), y.numpy()
        out[p + 'cfg'] = np.array([c['fin'], c['hid'], c['ncls'], c['L'], int(c['pp'])])
        out[p + 'logits'] = logits.detach().numpy()
        for k, v in sd_np(model).items():
            out[p + 'param.' + k] = v
        for k, v in model.named_parameters():
            out[p + 'grad.' + k] = v.grad.numpy().copy()
    # BaselineGCN forward + CE gradients, with and without the whole-tensor layer norm
    for ci, c in enumerate([dict(n=50, nnz=400, fin=9, hid=12, ncls=4, L=2, ln=True, seed=31),
                            dict(n=44, nnz=300, fin=20, hid=6, ncls=3, L=1, ln=False, seed=32)]):
        src, dst = small_graph(c['n'], c['nnz'], c['seed'], loops=True)
        g = dgl.DGLGraph((src, dst), num_nodes=c['n'])
        torch.manual_seed(c['seed'])
        x = torch.randn(c['n'], c['fin'])
        y = torch.randint(0, c['ncls'], (c['n'],))
        g.ndata['feat'] = x
        model = modules.BaselineGCN(c['fin'], c['hid'], c['ncls'], c['L'], F.relu, 0.5, c['ln'])
        with torch.no_grad():
            for l in model.layers:
                l.bias.uniform_(-0.3, 0.3)
        model.eval()
        logits = model(g)
        F.cross_entropy(logits, y).backward()
        p = 'bg%d_' % ci
        out[p + 'src'], out[p + 'dst'], out[p + 'n'] = src, dst, np.int64(c['n'])
        out[p + 'x'], out[p + 'y'] = x.numpy(), y.numpy()
        out[p + 'cfg'] = np.array([c['fin'], c['hid'], c['ncls'], c['L'], int(c['ln'])])
        out[p + 'logits'] = logits.detach().numpy()
        for k, v in sd_np(model).items():
            out[p + 'param.' + k] = v
        for k, v in model.named_parameters():
            out[p + 'grad.' + k] = v.grad.numpy().copy()


def main():
    os.makedirs(OUT, exist_ok=True)
    bind_reference('cluster_gcn')
    which = sys.argv[1:] or ['sage', 'graphconv', 'partition', 'cluster_iter', 'wrapper', 'train_ist', 'train_ist_data', 'gat', 'graphsage']
    gens = dict(sage=gen_sage, graphconv=gen_graphconv, partition=gen_partition,
                cluster_iter=gen_cluster_iter, wrapper=gen_wrapper, train_ist=gen_train_ist, train_ist_data=gen_train_ist_data, gat=gen_gat, graphsage=gen_graphsage)
    for name in which:
        out = {}
        gens[name](out)
        path = os.path.join(OUT, name + '.npz')
        np.savez_compressed(path, **out)
        print('wrote %s (%d arrays, %.1f KB)' % (path, len(out), os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
