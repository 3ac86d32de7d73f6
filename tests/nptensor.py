"""A NumPy stand-in for the `Tensor` dictionary, used ONLY by the CPU tests to exercise the product's host logic
(tensor_ops_b200.top / expr / nn TOp construction) without a GPU.  Deliberately independent of oracle/."""
import numpy as np

from tensor_ops_b200 import expr as E


class NpT:
    def __init__(self, a):
        self.a = np.asarray(a, dtype=np.float64)

    @property
    def shape(self): return self.a.shape

    @staticmethod
    def liftT(f, xs, like=None):
        e = E.trace(f, len(xs))
        code, consts = E.compile_expr(e)
        return NpT(run_bytecode(code, consts, [x.a for x in xs], xs[0].a.shape if xs else like.a.shape))

    @staticmethod
    def gmul(lM, lO, lN, x, y):
        perm = list(range(lO - 1, -1, -1)) + list(range(lO, lO + lN))
        return NpT(np.tensordot(x.a, np.transpose(y.a, perm), axes=lO))

    @staticmethod
    def sumT(xs):
        acc = xs[0].a
        for x in xs[1:]:
            acc = acc + x.a
        return NpT(acc)

    @staticmethod
    def scaleT(a, x): return NpT(a * x.a)
    @staticmethod
    def transp(x): return NpT(np.transpose(x.a))
    @staticmethod
    def sumRows(x): return NpT(x.a.sum(axis=0))
    @staticmethod
    def broadcastRows(n, row): return NpT(np.broadcast_to(row.a, (n,) + row.a.shape).copy())
    @staticmethod
    def konst(shape, v, like): return NpT(np.full(shape, v))


def run_bytecode(code, consts, xs, shape):
    """Python model of the device interpreter (csrc/kernels.cu k_lift) — checks the bytecode the host emits."""
    from tensor_ops_b200._lib import OPCODES as OP
    inv = {v: k for k, v in OP.items()}
    st = []
    for ins in code:
        op, arg = inv[ins >> 16], ins & 0xffff
        if op == "VAR": st.append(np.asarray(xs[arg], dtype=np.float64))
        elif op == "CONST": st.append(np.full(shape, consts[arg], dtype=np.float64))
        elif op in ("ADD", "SUB", "MUL", "DIV", "MAX", "MIN", "POW"):
            b = st.pop(); a = st.pop()
            st.append({"ADD": a + b, "SUB": a - b, "MUL": a * b, "DIV": a / b, "MAX": np.maximum(a, b), "MIN": np.minimum(a, b), "POW": a ** b}[op])
        else:
            a = st.pop()
            st.append({"NEG": -a, "EXP": np.exp(a), "LOG": np.log(a), "RECIP": 1 / a, "SQRT": np.sqrt(a), "TANH": np.tanh(a),
                       "ABS": np.abs(a), "SIGNUM": np.sign(a), "LOGISTIC": 1 / (1 + np.exp(-a)), "SIN": np.sin(a), "COS": np.cos(a)}[op])
    assert len(st) == 1
    return np.broadcast_to(st[0], shape).copy()
