"""a13: the tcgen05 fp32 GEMM (csrc/gemm.cu) through the C ABI against float64 CPU matmuls of the same inputs.

Bar (BASELINE north star): logits and grads within 1e-5 rel of the fp32 reference -> the 3xTF32 path is held to
|got - want| <= 1e-5 * (|want| + max|want|) against the float64 product; the plain-TF32 option to 4e-3."""
import pytest
import torch

from recbox_b200 import ops
from recbox_b200._lib import RbxError

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _check(got, want, tol, what):
    want = want.to(torch.float64)
    err = (got.double().cpu() - want).abs()
    bound = tol * (want.abs() + want.abs().max())
    bad = err > bound
    assert not bool(bad.any()), "%s: %d / %d outside %g (max err %.3e at %s, scale %.3e)" % (
        what, int(bad.sum()), bad.numel(), tol, float(err.max()), tuple(int(i) for i in (err == err.max()).nonzero()[0]),
        float(want.abs().max()))


def _operands(M, N, K, a_mn, b_mn, seed):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(M, K, generator=g)
    B = torch.randn(N, K, generator=g) * 0.1
    a = (A.t().contiguous() if a_mn else A).to(DEV)
    b = (B.t().contiguous() if b_mn else B).to(DEV)
    return A, B, a, b


SHAPES = [(128, 16, 32), (128, 208, 64), (256, 400, 624), (1000, 400, 400), (300, 624, 400), (77, 50, 100), (513, 256, 36),
          (4096, 400, 624)]


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_3xtf32_matches_float64(M, N, K, a_mn, b_mn):
    if a_mn and M % 4:
        M += 4 - M % 4          # stored [K, M]: the pitch must be a multiple of 4 floats
    if b_mn and N % 4:
        N += 4 - N % 4
    A, B, a, b = _operands(M, N, K, a_mn, b_mn, 7 * M + N + K)
    want = A.double() @ B.double().t()
    _check(ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn), want, 1e-5, "3xTF32 %s" % ((M, N, K, a_mn, b_mn),))
    _check(ops.gemm(a, b, a_mn=a_mn, b_mn=b_mn, precision=1), want, 4e-3, "TF32 %s" % ((M, N, K, a_mn, b_mn),))


def test_linear_relu_forward_and_backward_epilogues():
    """The three GEMMs of one hidden layer of MLP_Block (mlp_block.py:43-61) and its backward, with their fused epilogues."""
    M, Kin, Nout = 2048, 624, 400
    g = torch.Generator().manual_seed(3)
    H = torch.randn(M, Kin, generator=g)
    W = torch.randn(Nout, Kin, generator=g) * 0.05
    bias = torch.randn(Nout, generator=g)
    dZ = torch.randn(M, Nout, generator=g)
    Hd, Wd, bd, dZd = H.to(DEV), W.to(DEV), bias.to(DEV), dZ.to(DEV)
    Z = H.double() @ W.double().t() + bias.double()
    _check(ops.gemm(Hd, Wd, bias=bd, relu=True), Z.clamp_min(0), 1e-5, "relu(H W^T + b)")
    _check(ops.gemm(Hd, Wd, bias=bd), Z, 1e-5, "H W^T + b")
    # dH = (dZ W) * (H > 0): reduction over Nout, W read as stored ([Nout, Kin] = [K, N] with N contiguous)
    dH = (dZ.double() @ W.double()) * (H > 0)
    _check(ops.gemm(dZd, Wd, b_mn=True, mask=Hd), dH, 1e-5, "(dZ W) * (H > 0)")
    # dW = dZ^T H: reduction over the batch (split-K), both operands read as stored
    dW = dZ.double().t() @ H.double()
    _check(ops.gemm(dZd, Hd, a_mn=True, b_mn=True), dW, 1e-5, "dZ^T H")
    acc = torch.ones(Nout, Kin, device=DEV)
    _check(ops.gemm(dZd, Hd, a_mn=True, b_mn=True, out=acc, accumulate=True), dW + 1, 1e-5, "dW accumulate")
    _check(ops.colsum(dZd), dZ.double().sum(0), 1e-5, "colsum")


def test_head_layer_skinny_shapes():
    """Linear(hidden, 1): y = H w^T + b, dH = dy w, dw = dy^T H (memory-bound SIMT kernels behind the same entry point)."""
    M, K = 3000, 400
    g = torch.Generator().manual_seed(5)
    H, w, b, dy = torch.randn(M, K, generator=g), torch.randn(1, K, generator=g) * 0.1, torch.randn(1, generator=g), torch.randn(M, 1, generator=g)
    Hd, wd, bd, dyd = H.to(DEV), w.to(DEV), b.to(DEV), dy.to(DEV)
    _check(ops.gemm(Hd, wd, bias=bd), H.double() @ w.double().t() + b.double(), 1e-5, "head fwd")
    _check(ops.gemm(dyd, wd, b_mn=True, mask=Hd), (dy.double() @ w.double()) * (H > 0), 1e-5, "head dH")
    _check(ops.gemm(dyd, Hd, a_mn=True, b_mn=True), dy.double().t() @ H.double(), 1e-5, "head dw")


def test_strided_views_and_errors():
    g = torch.Generator().manual_seed(9)
    big = torch.randn(256, 640, generator=g).to(DEV)
    a = big[:, :624]                                  # pitch 640, 624 columns
    W = (torch.randn(400, 624, generator=g) * 0.1).to(DEV)
    out = torch.zeros(256, 512, device=DEV)
    ops.gemm(a, W, out=out[:, :400])
    _check(out[:, :400], a.double().cpu() @ W.double().cpu().t(), 1e-5, "strided")
    assert float(out[:, 400:].abs().sum()) == 0.0
    with pytest.raises(RbxError):
        ops.gemm(a, W[:, :620])
    with pytest.raises(RbxError):
        ops.gemm(a.cpu(), W.cpu())
    with pytest.raises(RbxError):
        ops.gemm(big[:, 1:625], W)                    # not 16-byte aligned
