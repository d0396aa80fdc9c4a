"""Integer identities the crop-warp kernel (mindpose_b200/csrc/warp_affine.cu) relies on,
checked exhaustively / on random data with numpy.  They are what makes its shortcuts
bit-exact with OpenCV's fixed-point pipeline (restated in oracle/warp.py); the kernel itself
is compared with the oracle and with cv2 goldens in tests/test_topdown_gpu.py."""
import numpy as np

from oracle import warp


def test_opencv_weights_are_products_of_5_bit_fractions():
    """rint(w * 32768) of the float32 bilinear weights = 32 * (32-fx | fx) * (32-fy | fy), so
    (sum(w15 * p) + 16384) >> 15 == (sum(w10 * p) + 512) >> 10."""
    tab = warp.bilinear_weight_table().astype(np.int64)       # [fy * 32 + fx, (tl, tr, bl, br)]
    fy, fx = np.divmod(np.arange(1024), 32)
    w10 = np.stack([(32 - fx) * (32 - fy), fx * (32 - fy), (32 - fx) * fy, fx * fy], axis=1)
    assert np.array_equal(tab, 32 * w10)
    rng = np.random.RandomState(0)
    p = rng.randint(0, 256, size=(1024, 4)).astype(np.int64)
    assert np.array_equal(((tab * p).sum(1) + 16384) >> 15, ((w10 * p).sum(1) + 512) >> 10)


def test_folded_weight_pairs_give_the_four_tap_sum():
    """The kernel keeps the weight pair wg = (32-fx) | fx << 16, scales BOTH halves by the row
    weight in one 32-bit multiply (no carry between the halves: products <= 1024) and chains
    two 16-bit x 8-bit two-way dot products per channel (dp2a)."""
    fx, fy = np.meshgrid(np.arange(32, dtype=np.uint64), np.arange(32, dtype=np.uint64))
    fx, fy = fx.reshape(-1), fy.reshape(-1)
    wg = (32 - fx) | (fx << 16)
    for row_w in (32 - fy, fy):
        w = (wg * row_w) & 0xFFFFFFFF
        assert np.array_equal(w & 0xFFFF, (32 - fx) * row_w)      # low half: left tap
        assert np.array_equal(w >> 16, fx * row_w)                # high half: right tap
        assert (w & 0xFFFF).max() <= 1024 and (w >> 16).max() <= 1024
    rng = np.random.RandomState(1)
    for _ in range(8):
        p = rng.randint(0, 256, size=(4, 1024)).astype(np.uint64)  # tl, tr, bl, br
        wt, wu = wg * (32 - fy), wg * fy
        dp2a = lambda w, a, b, c: (w & 0xFFFF) * a + (w >> 16) * b + c   # noqa: E731
        got = dp2a(wu, p[2], p[3], dp2a(wt, p[0], p[1], 512)) >> 10
        want = ((32 - fx) * (32 - fy) * p[0] + fx * (32 - fy) * p[1]
                + (32 - fx) * fy * p[2] + fx * fy * p[3] + 512) >> 10
        assert np.array_equal(got, want)
        assert got.max() <= 255


def test_column_deltas_are_monotone_so_a_quad_is_bounded_by_its_end_pixels():
    """adelta[x] = rint(m00 * x * 1024) is monotone in x for either sign of m00; the kernel
    tests the first and the last pixel of a quad for the whole quad."""
    rng = np.random.RandomState(2)
    x = np.arange(1024, dtype=np.float64)
    for m00 in np.concatenate([rng.uniform(-4, 4, 200), [0.0, 1e-9, -1e-9, 0.5, -0.5]]):
        ad = np.rint(m00 * x * 1024.0)
        d = np.diff(ad)
        assert (d >= 0).all() or (d <= 0).all()
    # and so is everything derived from it by adding a constant and shifting right
    ad = np.rint(-1.37 * x * 1024.0).astype(np.int64)
    sx = ((12345 + ad) >> 5) >> 5
    q = sx[: 1024 // 4 * 4].reshape(-1, 4)
    assert (np.minimum(q[:, 0], q[:, 3]) == q.min(1)).all()
    assert (np.maximum(q[:, 0], q[:, 3]) == q.max(1)).all()


def test_three_word_packing_of_four_pixels_from_one_neighbour():
    """Column-per-lane mapping (PC_WARP_COLUMN_MAP): lane k of a group of
    four holds pixel p_k = r | g << 8 | b << 16 and its right neighbour's; word k of the
    group's 12 output bytes is (p_k >> 8k) | (p_{k+1} << (24 - 8k)) for k = 0, 1, 2."""
    rng = np.random.RandomState(3)
    rgb = rng.randint(0, 256, size=(1000, 4, 3), dtype=np.uint8)
    p = (rgb[..., 0].astype(np.uint64) | (rgb[..., 1].astype(np.uint64) << 8)
         | (rgb[..., 2].astype(np.uint64) << 16))
    words = np.stack([((p[:, k] >> (8 * k)) | (p[:, k + 1] << (24 - 8 * k))) & 0xFFFFFFFF
                      for k in range(3)], axis=1).astype("<u4")
    assert np.array_equal(words.view(np.uint8).reshape(1000, 12), rgb.reshape(1000, 12))
