"""dgl.data shims: only what the reference trainers import at module scope."""


def register_data_args(parser):
    parser.add_argument('--dataset', type=str, required=False, default='cora')


_LOADER = None


def load_data(args):
    if _LOADER is None:
        raise RuntimeError('oracle dgl.data.load_data: set dgl.data._LOADER to a synthetic loader')
    return _LOADER(args)


class RedditDataset:  # imported by gcn/train_ist.py:12, never used
    pass


class PPIDataset:
    pass
