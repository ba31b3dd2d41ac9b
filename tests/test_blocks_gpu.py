"""f4 on the GPU: CrossNet / CrossNetV2 / CompressedInteractionNet / DIN_Attention / MultiHeadTargetAttention of
recbox_b200.blocks (every Linear / 1x1 convolution on the tcgen05 GEMM) against the outputs and gradients of the reference's
own modules on the same seeded init and inputs (tests/golden/blocks.npz, minted by oracle/make_golden.py)."""
import pytest
import torch

from helpers import assert_close
from test_oracle_golden import load

from recbox_b200 import blocks

pytestmark = pytest.mark.gpu
DEV = "cuda"
B, F_, D = 48, 6, 8


def _cases(g):
    E, target, hist, mask = (g[k].to(DEV) for k in ("E", "target", "hist", "mask"))
    flat = E.flatten(1)
    return {
        "crossnet": (lambda: blocks.CrossNet(F_ * D, 2), lambda m, a: m(a[0]), [flat]),
        "crossnetv2": (lambda: blocks.CrossNetV2(F_ * D, 3), lambda m, a: m(a[0]), [flat]),
        "cin": (lambda: blocks.CompressedInteractionNet(F_, [8, 4], output_dim=1), lambda m, a: m(a[0]), [E]),
        "din": (lambda: blocks.DIN_Attention(embedding_dim=D, attention_units=[16], hidden_activations="ReLU"),
                lambda m, a: m(a[0], a[1], mask), [target, hist]),
        "din_softmax": (lambda: blocks.DIN_Attention(embedding_dim=D, attention_units=[16, 8], hidden_activations="ReLU", use_softmax=True),
                        lambda m, a: m(a[0], a[1], mask), [target, hist]),
        "mhta": (lambda: blocks.MultiHeadTargetAttention(input_dim=D, attention_dim=16, num_heads=2),
                 lambda m, a: m(a[0], a[1], mask), [target, hist]),
    }


@pytest.mark.parametrize("tag", ["crossnet", "crossnetv2", "cin", "din", "din_softmax", "mhta"])
def test_block_matches_reference(tag):
    g = load("blocks")
    mk, call, args = _cases(g)[tag]
    m = mk()
    init = {k[len(tag) + 6:]: v for k, v in g.items() if k.startswith(tag + ".init.")}
    assert sorted(m.state_dict()) == sorted(init)
    m.load_state_dict(init)
    m.to(DEV).train()
    ins = [a.detach().clone().requires_grad_(True) for a in args]
    y = call(m, ins)
    assert_close(y, g[tag + ".y"], rtol=1e-5, atol_scale=1e-5, what=tag + ".y")
    (y * g[tag + ".w"].to(DEV)).sum().backward()
    for i, a in enumerate(ins):
        assert_close(a.grad, g["%s.din%d" % (tag, i)], rtol=1e-5, atol_scale=1e-5, what="%s.din%d" % (tag, i))
    # absolute floor from the module's largest gradient: the last bias of a softmax-normalised score has a mathematically
    # zero gradient (both sides hold round-off there)
    scale = max(float(g["%s.grad.%s" % (tag, k)].abs().max()) for k, _ in m.named_parameters())
    for k, p in m.named_parameters():
        want = g["%s.grad.%s" % (tag, k)].double()
        err = (p.grad.double().cpu() - want).abs()
        assert bool((err <= 1e-5 * want.abs() + 1e-5 * scale).all()), "%s.grad.%s: max err %.3e (scale %.3e)" % (tag, k, float(err.max()), scale)
