"""Oracle self-consistency for the learned alignment embedding (CPU only)."""
import numpy as np

from oracle import align as oalign
from oracle import embed as oembed


def test_gram_form_equals_direct_distance():
    p = oembed.make_embed_params()
    a, b = oalign.synth_swings(1, 40, 33, seed=3)
    fa, fb = oembed.embed(a[0], p), oembed.embed(b[0], p)
    direct = np.sqrt(((fa[:, None, :].astype(np.float64) - fb[None, :, :]) ** 2).sum(-1))
    c = oembed.embed_cost(a[0], b[0], p)
    assert c.shape == (40, 33) and c.dtype == np.float32
    assert np.abs(c - direct).max() / direct.max() < 1e-4          # fp32 cancellation in the Gram form


def test_self_alignment_is_the_diagonal():
    p = oembed.make_embed_params()
    a, _ = oalign.synth_swings(1, 50, 50, seed=4)
    total, path, c = oembed.align_embed_ref(a[0], a[0], p)
    assert np.all(np.diag(c) < 2e-2 * c.max())                     # |f|^2 + |f|^2 - 2 f.f cancels to ~0, not exactly
    assert np.array_equal(path, np.stack([np.arange(50), np.arange(50)], 1).astype(np.int32))


def test_bf16_emulation_is_close_and_rounding_is_exact():
    p = oembed.make_embed_params()
    a, b = oalign.synth_swings(1, 64, 48, seed=5)
    c32 = oembed.embed_cost(a[0], b[0], p)
    c16 = oembed.embed_cost(a[0], b[0], p, emulate_bf16=True)
    assert np.abs(c16 - c32).max() / c32.max() < 1e-2
    x = np.array([1.0, 1.00390625, 1.005859375, -3.1415926, 1e-30, 65504.0], np.float32)
    import torch
    assert np.array_equal(oembed.bf16_round(x), torch.from_numpy(x).to(torch.bfloat16).float().numpy())


def test_blob_layout():
    p = oembed.make_embed_params(7)
    blob = oembed.pack_embed_blob(p)
    assert blob.size == 34 * 128 + 128 + 128 * 128 + 128
    assert np.array_equal(blob[:34 * 128].reshape(34, 128), p["W1"])
    assert np.array_equal(blob[-128:], p["b2"])
