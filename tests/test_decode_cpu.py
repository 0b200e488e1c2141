"""Incremental evaluation of the causal operator (mhla_b200/decode.py) and the variable-length helpers of the NLP layer,
on CPU against the oracle: prefill + decode in arbitrary pieces must equal the operator on the whole sequence."""
import pytest
import torch

import oracle
from mhla_b200.decode import causal_with_state, get_unpad_data, pad_input
from mhla_b200.modules.nlp import Cache, RotaryEmbedding, ShortConvolution


@pytest.mark.parametrize("pieces", [[200], [64, 1, 1, 1], [70, 58, 1, 63, 8], [1] * 70, [130, 70]])
def test_prefill_and_decode_equal_full_sequence(pieces):
    g = torch.Generator().manual_seed(0)
    B, H, K, V, T = 2, 2, 16, 24, sum(pieces)
    q, k = torch.randn(B, T, H, K, generator=g), torch.randn(B, T, H, K, generator=g)
    v = torch.randn(B, T, H, V, generator=g)
    mm = torch.clamp(torch.rand(32, 32, generator=g), 1e-5, 1).tril()
    ref = oracle.causal_chunk_fwd(q, k, v, mm)
    state, outs, a = None, [], 0
    for n in pieces:
        o, state = causal_with_state(q[:, a:a + n], k[:, a:a + n], v[:, a:a + n], mm.view(32, 32, 1, 1, 1, 1), state)
        outs.append(o)
        a += n
    assert state.seen_tokens == T and state.S.shape[2] == T // 64 and state.k_tail.shape[1] == T % 64
    assert oracle.err_ratio(ref, torch.cat(outs, dim=1)) < 1e-5


def test_decode_needs_enough_mixing_rows():
    q = torch.randn(1, 64 * 2, 1, 8)
    _, st = causal_with_state(q, q, q, torch.ones(2, 2).tril(), None)
    with pytest.raises(IndexError):
        causal_with_state(q[:, :1], q[:, :1], q[:, :1], torch.ones(2, 2).tril(), st)


def test_unpad_pad_roundtrip_and_cache():
    mask = torch.tensor([[1, 1, 1, 0, 0], [1, 1, 1, 1, 1], [1, 0, 0, 0, 0]])
    idx, cu, mx = get_unpad_data(mask)
    assert cu.tolist() == [0, 3, 8, 9] and mx == 5
    x = torch.arange(15.).view(3, 5, 1)
    packed = x.reshape(15, 1)[idx]
    back = pad_input(packed, idx, 3, 5)
    assert torch.equal(back, x * mask.unsqueeze(-1))
    c = Cache()
    assert c.get_seq_length(0) == 0 and len(c) == 0
    c.update(recurrent_state="s0", conv_state=None, layer_idx=0, offset=7)
    c.update(recurrent_state="s1", conv_state=None, layer_idx=1, offset=7)
    c.update(recurrent_state="s0b", layer_idx=0, offset=1)
    assert c.get_seq_length(0) == 8 and c[0]["recurrent_state"] == "s0b" and c[1]["recurrent_state"] == "s1"


def test_rotary_and_short_conv_respect_sequence_borders():
    torch.manual_seed(0)
    lens = [5, 9, 3]
    cu = torch.tensor([0, 5, 14, 17], dtype=torch.int32)
    rot = RotaryEmbedding(8)
    q, k = torch.randn(1, 17, 2, 8), torch.randn(1, 17, 2, 8)
    qp, kp = rot(q, k, cu_seqlens=cu)
    conv = ShortConvolution(6, 4, bias=True)
    x = torch.randn(1, 17, 6)
    yp, _ = conv(x, cu_seqlens=cu)
    a = 0
    for n in lens:
        qs, ks = rot(q[:, a:a + n], k[:, a:a + n])
        torch.testing.assert_close(qp[:, a:a + n], qs)
        torch.testing.assert_close(kp[:, a:a + n], ks)
        ys, _ = conv(x[:, a:a + n])
        torch.testing.assert_close(yp[:, a:a + n], ys, rtol=1e-5, atol=1e-6)
        a += n
    # decode: the conv state continues the window, the rotary offset continues the positions
    y_full, _ = conv(x)
    y1, st = conv(x[:, :10], output_final_state=True)
    y2, _ = conv(x[:, 10:], cache=st)
    torch.testing.assert_close(torch.cat([y1, y2], dim=1), y_full, rtol=1e-5, atol=1e-6)
    q1, _ = rot(q[:, :10], k[:, :10])
    q2, _ = rot(q[:, 10:], k[:, 10:], seqlen_offset=10)
    torch.testing.assert_close(torch.cat([q1, q2], dim=1), rot(q, k)[0])
