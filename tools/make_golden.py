"""Generates tests/golden/*.npz from the CPU oracle (fp64).  The reference ships no golden vectors (test/Spec.hs is a
stub) and cannot be run here (no GHC), so these fixtures pin the ORACLE's behaviour (regressions) and give the GPU tests
size-stable targets; they are not outputs of the reference itself.   Usage: python tools/make_golden.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tensor_ops_oracle as O

out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
os.makedirs(out, exist_ok=True)

# 1. ffLayer' >>> logistic, fwd + VJP, reference init distributions (FeedForward.hs:206-207, Dots.hs:63)
rng = np.random.default_rng(2024)
B, i, o = 24, 20, 12
X = rng.uniform(-1, 1, (B, i)); W = rng.normal(0, 0.5, (o, i)); b = rng.normal(0, 0.5, o); dA = rng.normal(size=(B, o))
A, dX, dW, db = O.fflayer_logistic_per_sample(X, W, b, dA)
np.savez(os.path.join(out, "fflayer_logistic.npz"), X=X, W=W, b=b, dA=dA, A=A, dX=dX, dW=dW, db=db)

# 2. config-1 plumbing: Dots 2->16->1 logistic/squaredError, netGrad on 8 samples + 200 SGD steps
rng = np.random.default_rng(7)
net = O.genNet(2, [(16, O.actLogistic)], 1, O.actLogistic, rng)
xs = rng.uniform(-1, 1, (200, 2)); ys = np.array([[O.dots_target(v)] for v in xs])
g = [O.netGrad(O.squaredError(), xs[s], ys[s], net) for s in range(8)]
params0 = [p.copy() for p in net.params]
n2 = net
for s in range(200):
    n2 = O.trainNetwork(O.squaredError(), 1.0, xs[s], ys[s], n2)
np.savez(os.path.join(out, "dots_2_16_1.npz"), xs=xs, ys=ys, W0=params0[0], b0=params0[1], W1=params0[2], b1=params0[3],
         dx=np.stack([t[0] for t in g]), dW0=np.stack([t[1] for t in g]), db0=np.stack([t[2] for t in g]),
         dW1=np.stack([t[3] for t in g]), db1=np.stack([t[4] for t in g]),
         W0_200=n2.params[0], b0_200=n2.params[1], W1_200=n2.params[2], b1_200=n2.params[3])

# 3. MNIST-shaped head at reduced width: 30 -> 16 -> 12 -> 10, logistic/logistic/softmax + crossEntropy
rng = np.random.default_rng(11)
net = O.genNet(30, [(16, O.actLogistic), (12, O.actLogistic)], 10, O.softmax, rng)
B = 16
X = rng.uniform(0, 1, (B, 30)); Y = np.eye(10)[rng.integers(0, 10, B)]
Ws, bs = net.params[0::2], net.params[1::2]
# non-saturating weights so the naive softmax stays finite and informative
Ws = [w / np.sqrt(w.shape[1]) for w in Ws]
A, L, dXm, dWs, dbs = O.mlp_dense_fwd_grad(X, Ws, bs, ["logistic", "logistic", "softmax"], "crossEntropy", Y)
np.savez(os.path.join(out, "mlp_softmax_ce.npz"), X=X, Y=Y, A=A, L=L, dX=dXm,
         **{f"W{l}": Ws[l] for l in range(3)}, **{f"b{l}": bs[l] for l in range(3)},
         **{f"dW{l}": dWs[l] for l in range(3)}, **{f"db{l}": dbs[l] for l in range(3)})

# 4. general contraction, config-5 interpretation at reduced size: inner (LS (LS LZ)) (LS LZ) >>> sumRows, fwd + VJP
rng = np.random.default_rng(5)
x = rng.normal(size=(6, 5, 4)); y = rng.normal(size=(4, 7)); ct = rng.normal(size=(5, 7))
op = O.op_gmul(2, 1, 1) >> O.op_sumRows()
z = O.runTOp(op, [x, y])[0]
dx, dy = O.gradTOp_(op, [x, y], [ct])
# and a reversed-order double contraction
x2 = rng.normal(size=(3, 4, 5)); y2 = rng.normal(size=(5, 4, 6)); ct2 = rng.normal(size=(3, 6))
op2 = O.op_gmul(1, 2, 1)
z2 = O.runTOp(op2, [x2, y2])[0]
dx2, dy2 = O.gradTOp_(op2, [x2, y2], [ct2])
np.savez(os.path.join(out, "gmul_rank3.npz"), x=x, y=y, ct=ct, z=z, dx=dx, dy=dy, x2=x2, y2=y2, ct2=ct2, z2=z2, dx2=dx2, dy2=dy2)
print("wrote", sorted(os.listdir(out)))
