"""Host side of the convolution kernels: the bf16 weight image `models/conv3d.prepare_weight` builds must be exactly the layout
include/dlsg_b200.h documents (the kernel stages it with one bulk copy per pipeline stage and hands it to the tensor core as is)."""
import torch

from dl_swin_gan_b200.models.conv3d import prepare_weight


def pieces(w):
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    return hi, lo


def test_plain_and_stacked_images_follow_the_documented_index_formula():
    g = torch.Generator().manual_seed(0)
    cout, cin = 64, 96
    w = torch.randn(cout, cin, 3, 3, 3, generator=g)
    hi, lo = pieces(w)
    plain = prepare_weight(w, 32)                 # [tap][ci/32][h][(ci%32)/8][co/8][co%8][ci%8]
    stacked = prepare_weight(w, 32 | 0x100)       # [tap][ci/32][(ci%32)/8][h][co/8][co%8][ci%8]
    assert plain.dtype == torch.bfloat16 and plain.shape == (27, 3, 2, 4, 8, 8, 8) and stacked.shape == (27, 3, 4, 2, 8, 8, 8)
    for (tap, ci, co) in [(0, 0, 0), (13, 37, 21), (26, 95, 63), (5, 64, 8)]:
        for h, piece in enumerate((hi, lo)):
            want = piece.reshape(cout, cin, 27)[co, ci, tap]
            assert plain[tap, ci // 32, h, (ci % 32) // 8, co // 8, co % 8, ci % 8] == want
            assert stacked[tap, ci // 32, (ci % 32) // 8, h, co // 8, co % 8, ci % 8] == want
    # head + remainder reproduce the weight to 16 mantissa bits
    rec = hi.float() + lo.float()
    assert float((rec - w).abs().max() / w.abs().max()) < 2 ** -16


def test_first_layer_image_gathers_k_over_taps_and_pads_with_zeros():
    g = torch.Generator().manual_seed(1)
    for cin in (2, 4):
        cout = 32
        w = torch.randn(cout, cin, 3, 3, 3, generator=g)
        hi, _ = pieces(w)
        img = prepare_weight(w, 32)               # [k/32][h][(k%32)/8][co/8][co%8][k%8], k = tap*Cin + ci
        stages = -(-27 * cin // 32)
        assert img.shape == (stages, 2, 4, cout // 8, 8, 8)
        for (tap, ci, co) in [(0, 0, 0), (26, cin - 1, 31), (9, 1, 17)]:
            k = tap * cin + ci
            assert img[k // 32, 0, (k % 32) // 8, co // 8, co % 8, k % 8] == hi.reshape(cout, cin, 27)[co, ci, tap]
        flat = img[:, 0].permute(2, 3, 0, 1, 4).reshape(cout, stages * 32)          # [co][k]
        assert bool((flat[:, 27 * cin:] == 0).all())


def test_transposed_and_patch_weights_use_their_own_tap_count():
    g = torch.Generator().manual_seed(2)
    wt = torch.randn(32, 64, 4, 4, 4, generator=g)           # ConvTranspose3d weight [Cin, Cout, p, p, p]
    img = prepare_weight(wt, 32, transposed=True)
    assert img.shape == (64, 1, 2, 4, 8, 8, 8)
    hi, _ = pieces(wt.transpose(0, 1).contiguous())
    tap, ci, co = 37, 5, 60
    assert img[tap, 0, 0, ci // 8, co // 8, co % 8, ci % 8] == hi.reshape(64, 32, 64)[co, ci, tap]


def test_frame_windows_chain_through_the_denoisers_layers():
    """The window a layer must produce grows by one frame per 3-tap consumer and is clipped to the volume by the kernel wrapper."""
    from dl_swin_gan_b200.models.conv3d import _window
    from dl_swin_gan_b200.models.layers import grow_frames
    import pytest
    assert grow_frames(None, 3) is None
    keep = (6, 26)                                   # ResNet, 2 residual blocks: 32 padded frames, the crop keeps 6 .. 25
    chain = [grow_frames(keep, k) for k in (5, 3, 1, 0)]     # init layer, block 1, block 2, last layer
    assert chain == [(1, 31), (3, 29), (5, 27), (6, 26)]
    assert _window(None, 32) == (0, 32) and _window((1, 31), 32) == (1, 31)
    assert _window((-2, 40), 32) == (0, 32)          # beyond the volume: clipped, never an error
    with pytest.raises(ValueError):
        _window((10, 10), 32)
