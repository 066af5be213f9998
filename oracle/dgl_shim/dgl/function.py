"""``dgl.function`` built-ins used by the reference (descriptors only; executed by DGLGraph)."""
import operator


class _Binary:
    def __init__(self, lhs, rhs, out, op):
        self.lhs, self.rhs, self.out, self.op = lhs, rhs, out, op


class _UMulE:
    def __init__(self, u, e, out):
        self.u, self.e, self.out = u, e, out


class _CopyE:
    def __init__(self, e, out):
        self.e, self.out = e, out


class _Sum:
    def __init__(self, msg, out):
        self.msg, self.out = msg, out


def u_add_v(lhs, rhs, out):
    return _Binary(lhs, rhs, out, operator.add)


def u_mul_e(u, e, out):
    return _UMulE(u, e, out)


def copy_e(e, out):
    return _CopyE(e, out)


def sum(msg, out):  # noqa: A001  (mirrors dgl.function.sum)
    return _Sum(msg, out)
