"""Auto-encoders over TOps: the mirror of src/TensorOps/Learn/NeuralNet/AutoEncoder.hs.

An `Encoder` is a pair of feed-forward Networks; everything is expressed with the same TOp combinators as the reference
(`firstOp duplicate >>> secondOp o >>> swap >>> loss`), so the per-sample functions run every tensor method on the device.
The `*Batched` functions evaluate a whole batch with parameter gradients summed over samples; for ffLayer chains with a
recognised loss they reach the fused kernels (`tops_mlp_fwd_grad` with the batch as its own target).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

from . import nn
from . import top as TO
from .tensor import CuTensor


@dataclass
class Encoder:
    """`E { eEncoder :: Network t i o, eDecoder :: Network t o i }` (AutoEncoder.hs:36-39)."""
    enc: nn.Network
    dec: nn.Network


def encoderNet(e: Encoder) -> nn.Network:
    """`encoderNet (E e d) = e >>> d` (AutoEncoder.hs:80-84)."""
    return nn.net_compose(e.enc, e.dec)


def encode(e: Encoder, x):
    """AutoEncoder.hs:41-47."""
    return nn.runNetwork(e.enc, x)


def decode(e: Encoder, h):
    """AutoEncoder.hs:49-55."""
    return nn.runNetwork(e.dec, h)


def encodeDecode(e: Encoder, x):
    """AutoEncoder.hs:57-62."""
    return nn.runNetwork(encoderNet(e), x)


def _encoder_loss_op(loss: TO.TOp, net: nn.Network) -> TO.TOp:
    """firstOp duplicate >>> secondOp @'[ '[i] ] o >>> swap >>> loss (AutoEncoder.hs:72-78, 129-138)."""
    return TO.firstOp(TO.duplicate(), len(net.params)) >> TO.secondOp(1, net.op) >> TO.swap() >> loss


def testEncoder(loss: TO.TOp, e: Encoder, x, T=None):
    """AutoEncoder.hs:64-78: the reconstruction loss of one sample (a rank-0 tensor)."""
    net = encoderNet(e)
    return TO.runTOp(_encoder_loss_op(loss, net), [x] + net.params, T)[0]


def encGrad(loss: TO.TOp, x, e: Encoder, T=None) -> Tuple[List, List]:
    """`encGrad` (AutoEncoder.hs:110-142): (encoder parameter gradients, decoder parameter gradients)."""
    net = encoderNet(e)
    gr = TO.gradTOp(_encoder_loss_op(loss, net), [x] + net.params, T)[1:]
    nE = len(e.enc.params)
    return gr[:nE], gr[nE:]


def trainEncoder(loss: TO.TOp, r: float, x: CuTensor, e: Encoder) -> Encoder:
    """`trainEncoder` (AutoEncoder.hs:86-108): p' = p - r*g on both halves (fused SGD kernel)."""
    gE, gD = encGrad(loss, x, e)
    return Encoder(nn.Network(e.enc.op, nn.sgd_step(e.enc.params, gE, r), e.enc.layers),
                   nn.Network(e.dec.op, nn.sgd_step(e.dec.params, gD, r), e.dec.layers))


# ------------------------------------------------------------------ batched (SURVEY §8-d semantics)
def encGradBatched(loss: TO.TOp, X: CuTensor, e: Encoder):
    """encGrad for every row of X, gradients summed over rows: returns (loss_sum, enc grads, dec grads).  The parameter
    gradients of `firstOp duplicate >>> secondOp o >>> swap >>> loss` equal those of `netGrad loss x x (e >>> d)`."""
    ls, _, grads = nn.netGradBatched(loss, X, X, encoderNet(e), want_dx=False)
    nE = len(e.enc.params)
    return ls, grads[:nE], grads[nE:]


def trainEncoderBatched(loss: TO.TOp, r: float, X: CuTensor, e: Encoder) -> Encoder:
    _, gE, gD = encGradBatched(loss, X, e)
    return Encoder(nn.Network(e.enc.op, nn.sgd_step(e.enc.params, gE, r), e.enc.layers),
                   nn.Network(e.dec.op, nn.sgd_step(e.dec.params, gD, r), e.dec.layers))
