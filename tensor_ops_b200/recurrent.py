"""Stateful (recurrent) networks over TOps: the mirror of src/TensorOps/Learn/NeuralNet/Recurrent.hs.

Everything is built from the same TOp combinators as the reference — `fullyConnected`, `(~*~)`, `unroll`, `rollup` — so a
sequence gradient (`netGrad`, back-propagation through time by unrolling the step TOp n times) runs every tensor method on
the device through the C ABI.  There is no fused kernel for this path yet: it is the generic evaluator.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

from . import nn
from . import top as TO
from .tensor import Context, CuTensor, default_context


@dataclass
class Network:
    """`N { _nOp :: TOp ('[i] : ss ++ ps) ('[o] : ss), _nState :: Prod t ss, _nParams :: Prod t ps }` (Recurrent.hs:69-75)."""
    op: TO.TOp
    state: List
    params: List


def fullyConnected_(act: nn.Activation) -> TO.TOp:
    """The step TOp `fc` of `fullyConnected` (Recurrent.hs:114-124) on (x, h, W', W, b): output y = W x + W' h + b, new state act(y)."""
    return (TO.secondOp(1, TO.firstOp(TO.swap() >> TO.matVec(), 2) >> TO.firstOp(TO.swap(), 1))
            >> TO.firstOp(TO.swap() >> TO.matVec(), 2)
            >> TO.add3()
            >> TO.duplicate()
            >> TO.secondOp(1, act.op()))


def fullyConnected(i: int, o: int, act: nn.Activation, seed: int, ctx: Optional[Context] = None) -> Network:
    """`fullyConnected` (Recurrent.hs:97-125): s, w, w', b ~ N(0, 0.5^2) drawn on the device; parameters (w', w, b)."""
    ctx = ctx or default_context()
    s = ctx.rand_normal((o,), 0.0, 0.5, seed * 4 + 1)
    w = ctx.rand_normal((o, i), 0.0, 0.5, seed * 4 + 2)
    w_ = ctx.rand_normal((o, o), 0.0, 0.5, seed * 4 + 3)
    b = ctx.rand_normal((o,), 0.0, 0.5, seed * 4 + 4)
    return Network(fullyConnected_(act), [s], [w_, w, b])


def stateless(n: nn.Network) -> Network:
    """`stateless` (Recurrent.hs:127-137)."""
    return Network(n.op, [], list(n.params))


def then_act(n: Network, f: TO.TOp) -> Network:
    """`(*~)` (Recurrent.hs:262-267)."""
    return Network(n.op >> TO.firstOp(f, len(n.state)), n.state, n.params)


def compose(n1: Network, n2: Network) -> Network:
    """`(~*~)` (Recurrent.hs:178-233): states ss2 ++ ss1, parameters ps1 ++ ps2."""
    s1, p1, s2, p2 = len(n1.state), len(n1.params), len(n2.state), len(n2.params)
    o = (TO.secondOp(1, TO.firstOp(TO.swap_(s2, s1 + p1), p2))
         >> TO.firstOp(n1.op, s2 + p2)
         >> TO.secondOp(1, TO.swap_(s1, s2 + p2))
         >> TO.firstOp(n2.op, s1))
    return Network(o, n2.state + n1.state, n1.params + n2.params)


def runNetwork(n: Network, x, T=None):
    """`runNetwork` (Recurrent.hs:235-244): (output, network carrying the new state)."""
    out = TO.runTOp(n.op, [x] + n.state + n.params, T)
    return out[0], Network(n.op, out[1:], n.params)


def unroll(nS: int, nP: int, o: TO.TOp, n: int) -> TO.TOp:
    """`unroll` (Recurrent.hs:392-431): TOp (Replicate n '[i] ++ ss ++ ps) (ss ++ Replicate n '[o])."""
    if n == 0:
        return TO.take(nS, nS + nP)
    m = n - 1
    step = TO.fanout(o, TO.drop(1 + nS, 1 + nS + nP)) >> TO.swap_(1, nS + nP)
    return TO.secondOp(m, step) >> TO.firstOp(unroll(nS, nP, o, m), 1)


def rollup(loss: TO.TOp, n: int) -> TO.TOp:
    """`rollup` (Recurrent.hs:434-463)."""
    if n == 0:
        return TO.konst([()], 0.0)
    if n == 1:
        return loss
    m = n - 1
    return TO.secondOp(m, TO.firstOp(loss, m) >> TO.swap_(1, m)) >> TO.firstOp(rollup(loss, m), 1) >> TO.add()


def netGrad(loss: TO.TOp, xs: Sequence, ys: Sequence, n: Network, T=None):
    """`netGrad` (Recurrent.hs:277-324): BPTT by unrolling; returns (input grads in the order of `reverse xs`, state grads,
    parameter grads) exactly as the reference does."""
    steps, nS, nP = len(xs), len(n.state), len(n.params)
    unrolled = unroll(nS, nP, n.op, steps) >> TO.drop(nS, nS + steps)
    full = TO.firstOp(unrolled, steps) >> rollup(loss, steps)
    grad = TO.gradTOp(full, list(xs)[::-1] + n.state + n.params + list(ys), T)[:steps + nS + nP]
    return grad[:steps], grad[steps:steps + nS], grad[steps + nS:]


def trainNetwork(loss: TO.TOp, rS: float, rP: float, xs: Sequence[CuTensor], ys: Sequence[CuTensor], n: Network) -> Network:
    """`trainNetwork'` (Recurrent.hs:326-352): s' = s - rS*gS, p' = p - rP*gP (fused SGD kernel)."""
    _, gS, gP = netGrad(loss, xs, ys, n)
    return Network(n.op, nn.sgd_step(n.state, gS, rS) if n.state else [], nn.sgd_step(n.params, gP, rP))
