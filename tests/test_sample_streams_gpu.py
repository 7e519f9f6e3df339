"""GPU: sample streams (hc_pt_set_sample_streams).  S generators per pixel, pass p of a pixel draws from stream p mod S (generator index
k*W*H + pixel), so up to S consecutive passes are independent and share ONE wavefront - what keeps the machine full on a small frame or on
a GPU that owns 1/G of the tiles.  The reference's OpenCL layer has the same freedom (one RandomGen per slot of its ray block,
randGenState[MEGABLOCKSIZE], GPUOCLLayer.cpp:131); the oracle here is the reference's CPU integrator driven with the same stream rule
(oracle/ref_driver.cpp, ref_render_set_streams).  Checked: parity with the reference integrators, and that the image depends on
(seed, S) only - not on how many passes shared a wavefront, how the passes were split over calls, or the tile split."""
import numpy as np
import pytest

from tests import scenes

pytestmark = pytest.mark.gpu
MISPT, PT, QMC = 2, 0, 3


def _rel_rmse(a, b):
    return float(np.sqrt(((a - b)**2).mean())/max(np.sqrt((b**2).mean()), 1e-12))


def _sum(layer, scn, integ, passes, seed, streams, limit=0, calls=None):
    layer.SetSampleStreams(streams, limit)
    layer.LoadScene(scn)
    layer.InitPathTracing(seed)
    for n in (calls or [passes]):
        layer.TracingPass(integ, n)
    assert abs(layer.GetSPP() - passes) < 1e-4
    return layer.GetSumImage()[..., :3].copy()


@pytest.mark.parametrize("integ,kind,tol", [(PT, 0, 1e-5), (MISPT, 2, 1e-4)])
def test_streams_match_the_reference_integrators(layer, ref, integ, kind, tol):
    scn = scenes.cornell(96, 64)
    try:
        got = _sum(layer, scn, integ, 7, 55, 4)              # wavefronts of 4 and 3 passes
        assert layer.GroupPasses() == 4
    finally:
        layer.SetSampleStreams(1)
    rs = ref.scene(scn)
    try:
        want, npass = rs.render(kind, 55, 7, streams=4)
        one, _ = rs.render(kind, 55, 7)
    finally:
        rs.close()
    assert npass == 7 and np.isfinite(got).all()
    assert _rel_rmse(got, want[..., :3]) <= tol, _rel_rmse(got, want[..., :3])
    assert _rel_rmse(got, one[..., :3]) > 1e-3               # ... and it is a different sample set from the single-stream one


def test_image_depends_on_seed_and_streams_only(layer):
    scn = scenes.cornell(96, 64)
    n = 96*64
    try:
        a = _sum(layer, scn, MISPT, 7, 9, 4)                            # 4 + 3 passes per wavefront
        b = _sum(layer, scn, MISPT, 7, 9, 4, limit=n)                   # one pass per wavefront (streams taken in turn)
        c = _sum(layer, scn, MISPT, 7, 9, 4, limit=2*n)                 # 2 + 2 + 2 + 1
        d = _sum(layer, scn, MISPT, 7, 9, 4, calls=[3, 1, 3])           # wavefronts that start in the middle of the stream cycle
        assert np.array_equal(a, b) and np.array_equal(a, c) and np.array_equal(a, d)
        total = np.zeros_like(a)
        for r in range(3):                                              # three emulated ranks, interleaved tiles
            layer.SetTiles(32, r, 3)
            total += _sum(layer, scn, MISPT, 7, 9, 4)
        layer.SetTiles(32, 0, 1)
        assert np.array_equal(total, a)
        p = _sum(layer, scn, PT, 5, 9, 8)
        q = _sum(layer, scn, PT, 5, 9, 8, limit=n)
        assert np.array_equal(p, q)
    finally:
        layer.SetTiles(32, 0, 1)
        layer.SetSampleStreams(1)


def test_first_pass_of_any_stream_count_is_the_single_generator_rule(layer):
    scn = scenes.cornell(64, 64)
    try:
        a = _sum(layer, scn, MISPT, 1, 4, 1)
        b = _sum(layer, scn, MISPT, 1, 4, 4)                            # stream 0 = the generators of the single-stream rule
        assert np.array_equal(a, b)
    finally:
        layer.SetSampleStreams(1)


def test_qmc_streams_match_the_reference_and_do_not_depend_on_the_grouping(layer, ref):
    """MISPT-QMC with sample streams: one generator per SAMPLE index and stream, the Sobol index still runs over pass * W * H + sample, so passes of a
    stream cycle are independent here as well.  Samples land on arbitrary pixels and add with float atomics (order-dependent rounding), hence tolerances:
    1e-3 against the reference integrator with the same stream rule (as for one stream), 1e-5 between groupings and the emulated sample partition."""
    scn = scenes.cornell(96, 64, two_lights=True, dof=True)
    n = 96*64
    try:
        a = _sum(layer, scn, QMC, 6, 11, 4)                               # wavefronts of 4 and 2 passes
        b = _sum(layer, scn, QMC, 6, 11, 4, limit=n)                      # one pass per wavefront
        c = _sum(layer, scn, QMC, 6, 11, 4, calls=[1, 3, 2])
        assert _rel_rmse(b, a) <= 1e-5 and _rel_rmse(c, a) <= 1e-5
        total = np.zeros_like(a)
        for r in range(3):                                                # sample indices i = r (mod 3)
            layer.SetTiles(32, r, 3)
            layer.SetSampleStreams(4)
            layer.LoadScene(scn)
            layer.InitPathTracing(11)
            layer.TracingPass(QMC, 6)
            total += layer.GetSumImage()[..., :3]
        layer.SetTiles(32, 0, 1)
        assert _rel_rmse(total, a) <= 1e-5
        one = _sum(layer, scn, QMC, 6, 11, 1)
    finally:
        layer.SetTiles(32, 0, 1)
        layer.SetSampleStreams(1)
    rs = ref.scene(scn)
    try:
        want, npass = rs.render(3, 11, 6, streams=4)
    finally:
        rs.close()
    assert npass == 6 and _rel_rmse(a, want[..., :3]) <= 1e-3, _rel_rmse(a, want[..., :3])
    assert _rel_rmse(a, one) > 1e-2                                        # a different sample set from the single-stream one


def test_bad_arguments(layer):
    from hydracore_b200 import HcError
    for s in (0, 65, -1):
        with pytest.raises(HcError):
            layer.SetSampleStreams(s)
    layer.SetSampleStreams(1)
