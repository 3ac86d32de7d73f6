"""CPU tests of the product's host-side logic: symbolic ElemT (tracing, differentiation, bytecode), the TOp algebra and
the network builders — run at a NumPy stand-in `Tensor` dictionary (tests/nptensor.py) and compared with the oracle."""
import numpy as np
import pytest

from oracle import tensor_ops_oracle as O
from tensor_ops_b200 import expr as E
from tensor_ops_b200 import nn, top as TO
from tensor_ops_b200._lib import ACT_LOGISTIC, ACT_SOFTMAX, OPCODES

from nptensor import NpT, run_bytecode


def test_logistic_is_fused_to_one_opcode():
    code, consts = E.compile_expr(E.trace(E.logistic, 1))
    assert [c >> 16 for c in code] == [OPCODES["VAR"], OPCODES["LOGISTIC"]]


@pytest.mark.parametrize("f", [E.logistic, E.logistic_, lambda x: E.exp(x) * 2 - 1 / x, lambda x: E.tanh(x) ** 2 + E.sqrt(x * x + 1),
                               lambda x: E.log(x * x + 3) / (1 + x * x)])
def test_bytecode_matches_direct_evaluation_and_derivative(f):
    x = np.linspace(-2.0, 2.5, 41) + 0.013
    e = E.trace(f, 1)
    code, consts = E.compile_expr(e)
    np.testing.assert_allclose(run_bytecode(code, consts, [x], x.shape), E.evaluate(e, [x]), rtol=1e-12)
    d = E.diff(e, 0)
    h = 1e-6
    fd = (E.evaluate(e, [x + h]) - E.evaluate(e, [x - h])) / (2 * h)
    np.testing.assert_allclose(E.evaluate(d, [x]), fd, rtol=2e-6, atol=1e-8)


def test_binary_lift_and_gradlift():
    rng = np.random.default_rng(0)
    a, b = NpT(rng.normal(size=(5,))), NpT(rng.normal(size=(5,)))
    op = TO.zip(lambda x, y: x * y + E.exp(x - y))
    (z,) = TO.runTOp(op, [a, b])
    np.testing.assert_allclose(z.a, a.a * b.a + np.exp(a.a - b.a), rtol=1e-12)
    ct = NpT(rng.normal(size=(5,)))
    ga, gb = TO.gradTOp_(op, [a, b], [ct])
    np.testing.assert_allclose(ga.a, ct.a * (b.a + np.exp(a.a - b.a)), rtol=1e-12)
    np.testing.assert_allclose(gb.a, ct.a * (a.a - np.exp(a.a - b.a)), rtol=1e-12)


@pytest.mark.parametrize("loss_name", ["squaredError", "crossEntropy"])
def test_product_top_algebra_matches_oracle_netgrad(loss_name):
    rng = np.random.default_rng(4)
    out_act_o, out_act_p = (O.actLogistic, nn.actLogistic) if loss_name == "squaredError" else (O.softmax, nn.actSoftmax)
    onet = O.genNet(6, [(5, O.actLogistic), (4, O.actLogistic)], 3, out_act_o, rng)
    x = rng.uniform(-1, 1, 6)
    y = rng.uniform(0, 1, 3) if loss_name == "squaredError" else np.eye(3)[2]
    oloss = O.squaredError() if loss_name == "squaredError" else O.crossEntropy()
    want = O.netGrad(oloss, x, y, onet)
    # the product's network over the same parameters (params order: W,b of the first layer first — FeedForward.hs:82-90)
    pnet = nn.networkFromParams([NpT(p) for p in onet.params], [nn.actLogistic, nn.actLogistic, out_act_p])
    assert pnet.layers == [ACT_LOGISTIC, ACT_LOGISTIC, ACT_LOGISTIC if loss_name == "squaredError" else ACT_SOFTMAX]
    ploss = nn.squaredError() if loss_name == "squaredError" else nn.crossEntropy()
    got = TO.gradTOp(TO.then_first(pnet.op, ploss), [NpT(x)] + pnet.params + [NpT(y)], NpT)[:-1]
    for g, w in zip(got, want):
        np.testing.assert_allclose(g.a, w, rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(TO.runTOp(pnet.op, [NpT(x)] + pnet.params)[0].a, O.runNetwork(onet, x), rtol=1e-12)


def test_routing_combinators():
    x = NpT([1.0, 2.0, 3.0])
    op = TO.fanout(TO.scale(2.0), TO.scale(3.0))
    assert [t.a.tolist() for t in TO.runTOp(op, [x])] == [[2, 4, 6], [3, 6, 9]]
    (g,) = TO.gradTOp_(op, [x], [NpT(np.ones(3)), NpT(np.ones(3))])
    assert g.a.tolist() == [5, 5, 5]
    op2 = TO.par(TO.scale(2.0), TO.negate())
    assert [t.a.tolist() for t in TO.gradTOp_(op2, [x, x], [x, x])] == [[2, 4, 6], [-1, -2, -3]]
    op3 = TO.secondOp(1, TO.scale(5.0))
    assert [t.a.tolist() for t in TO.runTOp(op3, [x, x])] == [[1, 2, 3], [5, 10, 15]]
    sh = TO.shuffle([1, 1, 0], 3)
    gs = TO.gradTOp_(sh, [x, x, x], [x, x, x])
    assert [t.a.tolist() for t in gs] == [[1, 2, 3], [2, 4, 6], [0, 0, 0]]
    assert [t.a.tolist() for t in TO.runTOp(TO.drop(1, 2), [NpT([9.0]), x])] == [[1, 2, 3]]


def test_gmul_vjp_matches_oracle_for_rank3():
    rng = np.random.default_rng(1)
    x, y, ct = rng.normal(size=(3, 4, 5)), rng.normal(size=(5, 4, 6)), rng.normal(size=(3, 6))
    want = O.gradTOp_(O.op_gmul(1, 2, 1), [x, y], [ct])
    got = TO.gradTOp_(TO.gmul(1, 2, 1), [NpT(x), NpT(y)], [NpT(ct)])
    for g, w in zip(got, want):
        np.testing.assert_allclose(g.a, w, rtol=1e-12)


def test_saved_activations_make_a_chain_linear_in_forwards():
    """`g3 xs ds = g1 xs (g2 (f1 xs) ds)` (Types.hs:155) nests a forward of the prefix in every `>>>`; with the
    saved-activations scope a chain of n ops runs each composed prefix ONCE per gradient evaluation, and the values are the
    reference's (checked against the oracle's own, un-memoised, chain rule)."""
    from oracle import tensor_ops_oracle as O
    from nptensor import NpT
    rng = np.random.default_rng(3)
    n = 12
    x = NpT(rng.uniform(0.2, 1.5, 5))
    op, oop = TO.scale(1.1), O.op_scale(1.1)
    for k in range(n - 1):
        op = op >> TO.scale(1.0 + 0.01 * k) if k % 2 else op >> TO.map(lambda v: 0.25 * v * v + 0.5)
        oop = oop >> O.op_scale(1.0 + 0.01 * k) if k % 2 else oop >> O.op_map(lambda v: 0.25 * v * v + 0.5)
    ct = NpT(rng.normal(size=5))
    TO.forward_evals = 0
    (g,) = TO.gradTOp_(op, [x], [ct])
    linear = TO.forward_evals
    assert linear <= n, linear                      # one forward per composed prefix
    np.testing.assert_allclose(g.a, O.gradTOp_(oop, [x.a], [ct.a])[0], rtol=1e-12)
    # without the scope the same call tree re-runs every prefix at every level: quadratic
    TO.forward_evals = 0
    op.grad_(NpT, [x], [ct])
    assert TO.forward_evals >= n * (n - 1) // 2, TO.forward_evals
