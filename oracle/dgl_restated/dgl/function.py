"""dgl.function builtins used by the reference (copy_src + sum only)."""


class CopyMessage:
    def __init__(self, src, out):
        self.src, self.out = src, out


class SumReduce:
    def __init__(self, msg, out):
        self.msg, self.out = msg, out


def copy_src(src, out):
    return CopyMessage(src, out)


copy_u = copy_src


def sum(msg, out):  # noqa: A001 - mirrors dgl.function.sum
    return SumReduce(msg, out)
