"""Message / reduce descriptors mirroring ``dgl.function`` for the two builtins
the reference uses: ``fn.copy_src(src='h', out='m')`` and ``fn.sum(msg='m',
out='h')`` (cluster_gcn/modules.py:136-137, :224-225; sampler.py:64-66)."""


class CopySrc:
    def __init__(self, src, out):
        self.src, self.out = src, out


class Sum:
    def __init__(self, msg, out):
        self.msg, self.out = msg, out


def copy_src(src, out):
    return CopySrc(src, out)


copy_u = copy_src


def sum(msg, out):  # noqa: A001 - same name as dgl.function.sum
    return Sum(msg, out)
