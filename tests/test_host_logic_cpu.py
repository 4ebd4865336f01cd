"""CPU checks of arithmetic claims the CUDA kernels rely on, restated in numpy float32 (IEEE, correctly rounded like the device's
__fdiv_rn / __fmul_rn / __fadd_rn). No GPU, no oracle: these are properties of the reference's own texture-coordinate arithmetic
(denoiser.frag:82-86: textureLod(s, gl_FragCoord.xy / viewportSize, 0))."""
import numpy as np
import pytest

f32 = np.float32

SIZES = [1, 2, 3, 7, 16, 141, 250, 512, 640, 1080, 1366, 1920, 2160, 2560, 3840, 4096, 4320, 7680, 8192, 16384]


def _centre_axis(size: int, viewport: int):
    """csrc/k_streaming.cu: centreAxis for every texel x of one axis -> (exact, weight, i0, i1)."""
    x = np.arange(size, dtype=np.int64)
    c = x.astype(f32) + f32(0.5)
    p = ((c / f32(viewport)).astype(f32) * f32(size)).astype(f32)
    u = (p - f32(0.5)).astype(f32)
    fl = np.floor(u)
    w = (u - fl).astype(f32)
    i = fl.astype(np.int64)
    i0, i1 = np.clip(i, 0, size - 1), np.clip(i + 1, 0, size - 1)
    return x, c, p, (w == 0) & (i0 == x), w, i0, i1


@pytest.mark.parametrize("size", SIZES)
def test_centre_tap_texel_test_is_conservative(size):
    """K6 + K7 (denoiseFinalGatherKernel) decides "the centre tap is the texel itself" with fl(fl((x + .5) / V) * S) == x + .5 instead
    of building the whole bilinear footprint. Claim: whenever that test says yes, the footprint is (x, weight 0); with viewport == image
    size (the frame's only case) the two agree on every texel; with a different viewport the only disagreement is a tap clamped onto
    the last texel, which then takes the general path and fetches that texel."""
    for viewport in sorted({size, 2 * size, max(size // 2, 1), size + 1, max(size - 1, 1)}):
        x, c, p, exact, w, i0, i1 = _centre_axis(size, viewport)
        fast = p == c
        assert not np.any(fast & ~exact), (size, viewport)
        if viewport == size:
            assert np.array_equal(fast, exact), (size, viewport)
        else:
            differ = exact & ~fast
            # only clamped footprints: both taps are the texel itself
            assert np.all((i0[differ] == x[differ]) & (i1[differ] == x[differ]) & (w[differ] == 0)), (size, viewport)


def test_centre_tap_is_not_always_the_texel():
    """Why K6 cannot be a copy: on a few per cent of the columns of a 4K / 8K viewport the coordinate rounds to x +- ulp and the bilinear
    unit blends up to 2^-24 * W of a neighbour in."""
    for size, lo, hi in ((3840, 0.005, 0.08), (7680, 0.005, 0.08), (2160, 0.0, 0.08), (4320, 0.0, 0.08)):
        _, _, _, exact, w, _, _ = _centre_axis(size, size)
        frac = 1.0 - exact.mean()
        assert lo <= frac <= hi, (size, frac)
        wi = w[~exact].astype(np.float64)
        assert np.all(np.minimum(wi, 1.0 - wi) <= size * 2.0 ** -23), size  # x +- one ulp of the coordinate, never a real blend


def test_floor_by_magic_constant_matches_floor():
    """k_gather_fast.cu: footprint() takes floor(u) of u in [-1, 2^22) as the low mantissa bits of (u + 1.5 * 2^23) rounded DOWN.
    numpy has no directed rounding, so the check is done in float64 (exact for these magnitudes) and rounded down by hand."""
    rng = np.random.default_rng(7)
    u = np.concatenate([rng.uniform(-1.0, 8192.0, 200000), np.array([-1.0, -0.5, -2.0 ** -24, 0.0, 0.5, 1.0, 4095.999, 7679.0])]).astype(f32)
    magic = 12582912.0
    s = u.astype(np.float64) + magic                      # exact in float64
    ulp = 1.0                                             # spacing of float32 in [2^23, 2^24)
    t = np.floor(s / ulp) * ulp                           # round toward -inf onto the float32 grid
    assert np.all((t >= 2.0 ** 23) & (t < 2.0 ** 24))
    got = t.astype(f32).view(np.int32) - np.int32(0x4B400000)
    assert np.array_equal(got, np.floor(u.astype(np.float64)).astype(np.int32))
    frac = u.astype(np.float64) - (t - magic)
    assert np.all((frac >= 0.0) & (frac < 1.0))


def test_atan2_polynomial_error_bound():
    """k_gather_fast.cu: atan2Poly (8-term odd minimax polynomial on [0, 1] + octant reduction) — the kernel's comment claims 1.2e-7 rad
    for the polynomial; with the fp32 roundings of the evaluation the result stays within 5e-7 rad of atan2, i.e. two ulps at pi."""
    coeff = [-0.004054398275911808, 0.021862303838133812, -0.055911313742399216, 0.09642116725444794, -0.13908594846725464,
             0.1994655728340149, -0.33329859375953674, 0.9999993443489075]
    rng = np.random.default_rng(11)
    ang = rng.uniform(-np.pi, np.pi, 400000)
    rad = np.exp(rng.uniform(-6.0, 6.0, ang.size))
    x, y = (rad * np.cos(ang)).astype(f32), (rad * np.sin(ang)).astype(f32)
    keep = (x != 0) | (y != 0)
    x, y = x[keep], y[keep]
    ax, ay = np.abs(x), np.abs(y)
    mn, mx = np.minimum(ax, ay), np.maximum(ax, ay)
    t = (mn.astype(np.float64) / mx.astype(np.float64)).astype(f32)  # the kernel multiplies by rcp.approx (1 ulp): same bound class
    s = (t.astype(np.float64) * t.astype(np.float64)).astype(f32)
    p = np.full(t.shape, f32(coeff[0]), dtype=f32)
    for c in coeff[1:]:
        p = (p.astype(np.float64) * s.astype(np.float64) + c).astype(f32)  # fmaf
    r = (p.astype(np.float64) * t.astype(np.float64)).astype(f32)
    r = np.where(ay > ax, (f32(1.57079632679489662) - r).astype(f32), r)
    r = np.where(x < 0, (f32(3.14159265358979324) - r).astype(f32), r)
    r = np.copysign(r, y)
    err = np.abs(r.astype(np.float64) - np.arctan2(y.astype(np.float64), x.astype(np.float64)))
    assert float(err.max()) < 5e-7, float(err.max())
    # the polynomial itself, evaluated without rounding
    tt = np.linspace(0.0, 1.0, 200001)
    ss = tt * tt
    pp = np.full(tt.shape, coeff[0])
    for c in coeff[1:]:
        pp = pp * ss + c
    assert float(np.abs(pp * tt - np.arctan(tt)).max()) < 1.3e-7
