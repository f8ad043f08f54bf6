"""W8A8 family parity on the GPU: BIT-EXACT against the CPU oracle (kernels/ref semantics)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from tinychatengine_b200.runtime import Context

    c = Context(0)
    yield c
    c.close()


def rnd8(shape, seed):
    rng = np.random.default_rng(seed)
    return rng.integers(-127, 128, shape, dtype=np.int8)


# (variant of the C ABI, oracle variant) ; alpha/beta from the reference's op tests (test_ops.cc:179)
ALPHA, BETA = 0.00050354, 0.0213013


@pytest.mark.parametrize("M,N,K", [(1, 768, 768), (108, 768, 768), (1, 4096, 4096), (7, 1000, 136), (33, 64, 4096), (3, 40, 50), (1, 1031, 11008), (3, 1000, 2064), (4, 64, 528), (2, 9, 16384)])
def test_linear_variants_bit_exact(ctx, M, N, K):
    from oracle import capi

    A, B = rnd8((M, K), 1), rnd8((N, K), 2)
    b8 = rnd8((N,), 3)
    bf = np.random.default_rng(4).standard_normal(N).astype(np.float32)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    # int8 out with int8 bias (W8A8B8O8Linear) and its ReLU flavour (q_min = 0)
    for qmin in (-128, 0):
        got = ctx.w8a8_matmul(0, dA, dB, torch.from_numpy(b8).cuda(), ALPHA, BETA, qmin, 127).cpu().numpy()
        assert np.array_equal(got, capi.int8_matmul(0, A, B, b8, None, ALPHA, BETA, qmin, 127))
    got = ctx.w8a8_matmul(1, dA, dB, None, ALPHA, 0.0).cpu().numpy()
    assert np.array_equal(got, capi.int8_matmul(2, A, B, alpha=ALPHA))
    got = ctx.w8a8_matmul(2, dA, dB, torch.from_numpy(bf).cuda(), ALPHA, 0.0).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), capi.int8_matmul(4, A, B, biasf=bf, alpha=ALPHA).view(np.uint32))
    got = ctx.w8a8_matmul(3, dA, dB, None, ALPHA, 0.0).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), capi.int8_matmul(6, A, B, alpha=ALPHA).view(np.uint32))


def test_saturation_and_ties(ctx):
    """large alpha drives outputs into the clamp; alpha = 0.5 with odd accumulators produces exact .5 ties
    (round half away from zero, std::round)."""
    from oracle import capi

    M, N, K = 4, 64, 32
    A, B, b8 = rnd8((M, K), 9), rnd8((N, K), 10), rnd8((N,), 11)
    dA, dB, db = (torch.from_numpy(t).cuda() for t in (A, B, b8))
    for alpha, beta in ((0.05, 1.0), (0.5, 0.5), (1.0 / 1024, 0.5)):
        got = ctx.w8a8_matmul(0, dA, dB, db, alpha, beta).cpu().numpy()
        want = capi.int8_matmul(0, A, B, b8, None, alpha, beta)
        assert np.array_equal(got, want), (alpha, beta)
    assert np.abs(capi.int8_matmul(0, A, B, b8, None, 0.05, 1.0).astype(int)).max() >= 127


@pytest.mark.parametrize("heads,T,d", [(12, 64, 64), (32, 512, 128), (32, 129, 128)])
def test_batched_attention_matmuls_bit_exact(ctx, heads, T, d):
    """BMM_S8T_S8N_F32T (QK^T) and BMM_S8T_S8N_S8T (PV) at sqlen 1: the *_batch flavours."""
    from oracle import capi

    q = rnd8((heads, d), 21)
    Kc = rnd8((heads, T, d), 22)
    dq, dK = torch.from_numpy(q).cuda(), torch.from_numpy(Kc).cuda()
    got = ctx.w8a8_matmul(3, dq, dK, None, ALPHA, 0.0, batch=True).cpu().numpy()
    want = capi.int8_matmul(7, q, Kc, alpha=ALPHA)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    p = rnd8((heads, T), 23)
    Vt = rnd8((heads, d, T), 24)
    got = ctx.w8a8_matmul(1, torch.from_numpy(p).cuda(), torch.from_numpy(Vt).cuda(), None, 0.0031, 0.0, batch=True).cpu().numpy()
    assert np.array_equal(got, capi.int8_matmul(3, p, Vt, alpha=0.0031))


def test_golden_fixture_on_gpu(ctx, golden_dir):
    """the committed outputs of the reference's own kernels/ref build"""
    g = np.load(golden_dir / "kernels_generic.npz")
    A, B, Bb, b8, bf = (g[k] for k in ("i8_A", "i8_B", "i8_Bb", "i8_b8", "i8_bf"))
    alpha, beta = float(g["i8_alpha"]), float(g["i8_beta"])
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    assert np.array_equal(ctx.w8a8_matmul(0, T(A), T(B), T(b8), alpha, beta).cpu().numpy(), g["i8_C0"])
    assert np.array_equal(ctx.w8a8_matmul(0, T(A), T(B), T(b8), alpha, beta, 0, 127).cpu().numpy(), g["i8_C1"])
    assert np.array_equal(ctx.w8a8_matmul(1, T(A), T(B), None, alpha, 0.0).cpu().numpy(), g["i8_C2"])
    assert np.array_equal(ctx.w8a8_matmul(1, T(A), T(Bb), None, alpha, 0.0, batch=True).cpu().numpy(), g["i8_C3"])
    assert np.array_equal(ctx.w8a8_matmul(2, T(A), T(B), T(bf), alpha, 0.0).cpu().numpy(), g["i8_C4"])
    assert np.array_equal(ctx.w8a8_matmul(3, T(A), T(B), None, alpha, 0.0).cpu().numpy(), g["i8_C6"])
    assert np.array_equal(ctx.w8a8_matmul(3, T(A), T(Bb), None, alpha, 0.0, batch=True).cpu().numpy(), g["i8_C7"])
    assert np.array_equal(ctx.w8a8_matmul(0, T(A), T(B), T(b8), 0.05, 1.0).cpu().numpy(), g["i8_Csat"])
