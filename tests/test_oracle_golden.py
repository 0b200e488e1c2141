"""Pin the CPU oracle against fixtures produced by executing the reference's own code
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import oracle
from conftest import load_golden


@pytest.mark.parametrize("key,layout,tr", [
    ("W_side16_group16_linear", (4, 4), "linear"), ("W_side16_group16_cos", (4, 4), "cos"),
    ("W_side8_group4_exp", (4, 4), "exp"), ("W_side8_group16_gaussian", (2, 2), "gaussian"),
    ("W_side8_group4_local", (4, 4), "local"), ("W_side14_group49_linear", (2, 2), "linear"),
])
def test_block_distance_2d(key, layout, tr):
    ref = load_golden("blockdist2d")[key]
    W = oracle.block_distance_matrix(layout, tr)
    assert W.shape == ref.shape
    torch.testing.assert_close(W, ref, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("layout,tr", [((3, 5, 10), "linear"), ((2, 2, 3), "cos"), ((4, 1, 1), "linear"),
                                       ((2, 3, 2), "exp"), ((2, 2, 2), "local"), ((2, 2, 2), "gaussian")])
def test_block_distance_3d(layout, tr):
    ref = load_golden("blockdist3d")["W_" + "x".join(map(str, layout)) + "_" + tr]
    W = oracle.block_distance_matrix(layout, tr)
    torch.testing.assert_close(W, ref, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", ["a_dit_s2", "a_qknorm"])
def test_variant_a_core(name):
    g = load_golden(name)
    out = oracle.blockmix_fwd(g["q"], g["k"], g["v"], g["W"], eps=g["eps"], normalize=True)
    assert oracle.err_ratio(g["out"], out) < 2e-6
    if "kv" in g:
        torch.testing.assert_close(oracle.blockmix_summaries(g["k"], g["v"]), g["kv"], rtol=1e-5, atol=1e-5)


def test_variant_a_quirk_is_reproduced():
    """The textbook normaliser q_{i,t} . sum_j W[i,j] ksum_j must NOT match the reference."""
    g = load_golden("a_dit_s2")
    q, k, v, W = g["q"], g["k"], g["v"], g["W"]
    kv = torch.einsum("ij,zjxy->zixy", W, torch.matmul(k.transpose(-2, -1), v))
    den_textbook = torch.einsum("bitd,bid->bit", q, torch.einsum("ij,bjd->bid", W.float(), k.sum(-2))) + g["eps"]
    textbook = torch.matmul(q, kv) / den_textbook.unsqueeze(-1)
    assert oracle.err_ratio(g["out"], textbook) > 1e-3


@pytest.mark.parametrize("name", ["b_norm", "b_nonorm"])
def test_variant_b_core(name):
    g = load_golden(name)
    out = oracle.blockmix_fwd(g["q"], g["k"], g["v"], g["W"], eps=g["eps"], normalize=bool(g["normalize_out"]),
                              q_rope=g["q_rope"], k_rope=g["k_rope"])
    assert oracle.err_ratio(g["out"], out) < 2e-6


@pytest.mark.parametrize("name", ["b_norm", "b_nonorm"])
def test_rope_wan(name):
    g = load_golden(name)
    d = g["q_tok"].shape[-1]
    freqs = oracle.rope_freqs_wan(d)
    roped = oracle.rope_apply_wan(g["q_tok"], tuple(int(x) for x in g["grid"]), freqs)
    torch.testing.assert_close(roped, g["q_rope_tok"], rtol=1e-6, atol=1e-6)


def _cfg1_inputs(g):
    B, T, H, K, V = (int(x) for x in g["shape"])
    gen = torch.Generator().manual_seed(int(g["seed"]))
    q = torch.randn(B, T, H, K, generator=gen)
    k = torch.randn(B, T, H, K, generator=gen)
    v = torch.randn(B, T, H, V, generator=gen)
    mm = torch.clamp(torch.rand(32, 32, generator=gen), 1e-5, 1).tril()
    assert abs(q.double().sum().item() - float(g["q_sum"])) < 1e-6, "torch RNG stream changed; regenerate fixtures"
    return q, k, v, mm


@pytest.mark.parametrize("name", ["c_t256", "c_t200_ragged", "c_cfg1", "c_kv_128_256"])
def test_variant_c_chunk(name):
    g = load_golden(name)
    if name == "c_cfg1":
        q, k, v, mm = _cfg1_inputs(g)
    else:
        q, k, v, mm = g["q"], g["k"], g["v"], g["mm"]
    o = oracle.causal_chunk_fwd(q, k, v, mm)
    assert o.shape == g["o"].shape
    assert oracle.err_ratio(g["o"], o) < 2e-6
    if q.shape[1] <= 256:
        closed = oracle.causal_closed_form(q, k, v, mm).float()
        assert oracle.err_ratio(g["o"], closed) < 1e-5


@pytest.mark.parametrize("name", ["c_recurrent_t48", "c_recurrent_k64"])
def test_variant_c_recurrent_first_chunk(name):
    g = load_golden(name)
    o, state = oracle.recurrent_first_chunk_fwd(g["q"], g["k"], g["v"], g["mm"])
    assert state is None
    assert oracle.err_ratio(g["o"], o) < 1e-5          # reference recurrent form
    assert oracle.err_ratio(g["o_chunk"], o) < 2e-6    # reference chunk form
    if "S" in g:
        assert float(g["S"].abs().max()) == 0.0        # the reference's "final state" is all zeros


def test_variant_c_needs_enough_mixing_rows():
    q = torch.randn(1, 64 * 33, 1, 8)
    with pytest.raises(IndexError):
        oracle.causal_chunk_fwd(q, q, q, torch.ones(32, 32).tril())


def test_properties_linear_attention_limits():
    """SURVEY section 4 property tests: W=ones -> global linear attention; W=I -> block-local."""
    torch.manual_seed(0)
    q, k, v = torch.rand(2, 4, 8, 16), torch.rand(2, 4, 8, 16), torch.randn(2, 4, 8, 16)
    M = 4
    out = oracle.blockmix_fwd(q, k, v, torch.ones(M, M), normalize=False)
    glob = torch.matmul(q.reshape(2, 32, 16), torch.matmul(k.reshape(2, 32, 16).transpose(-2, -1), v.reshape(2, 32, 16)))
    torch.testing.assert_close(out.reshape(2, 32, 16), glob, rtol=1e-4, atol=1e-4)
    out = oracle.blockmix_fwd(q, k, v, torch.eye(M), normalize=False)
    loc = torch.matmul(q, torch.matmul(k.transpose(-2, -1), v))
    torch.testing.assert_close(out, loc, rtol=1e-5, atol=1e-5)
    # causal with mm = tril(ones) is plain causal linear attention
    T = 128
    q, k, v = torch.randn(1, T, 2, 16), torch.randn(1, T, 2, 16), torch.randn(1, T, 2, 8)
    o = oracle.causal_chunk_fwd(q, k, v, torch.ones(2, 2).tril())
    mask = torch.tril(torch.ones(T, T))
    ref = torch.einsum("bhts,bshv->bthv", torch.einsum("bthk,bshk->bhts", q, k) * mask, v) * 16 ** -0.5
    torch.testing.assert_close(o, ref, rtol=1e-4, atol=1e-4)
