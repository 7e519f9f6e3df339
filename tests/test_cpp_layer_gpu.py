"""GPU: the C++ drop-in — GPUCUDALayer : IHWLayer with MemoryStorageCUDA : IMemoryStorage — driven through the reference's own
virtual interface the way RenderDriverRTE drives a layer (tests/layerapi.py -> hydracore_b200/cpp/layer_harness.cpp).

The layer library is built against the reference's IHWLayer.h / IMemoryStorage.h and links the reference's own base-class code
(IHWLayerDataAssembler.cpp: PrepareEngineGlobals/Tables, SetAllPODLights, SetCamMatrices; MemoryStorageCPU.cpp: Update/GetTable), so
the EngineGlobals blob these tests render from is assembled by REFERENCE code, not by our Python packer.  Checks:
  * images through IHWLayer == the golden images of the reference CPU integrators (same tolerances as test_path_gpu.py)
  * the blob assembled by the reference base class == the blob hydracore_b200/scene.py packs (header fields, tables modulo offsets)
  * MemoryStorageCUDA keeps the IMemoryStorage contract (offsets in 16-byte blocks, host mirror only for geom / textures, size_t(-1))."""
import json
import os

import numpy as np
import pytest

from tests import scenes

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def consts():
    from hydracore_b200.layout import C
    d = dict(json.load(open(os.path.join(G, "ref_consts.json"))))     # enums dumped from the reference headers
    d.update(C)                                                        # + the EngineGlobals field offsets of csrc/hc_layout.h
    return d


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(G, "images.npz"))


def _make(scn, consts):
    from tests import layerapi
    if not layerapi.CppLayer.available():
        pytest.skip("hydracore_b200/cpp/_build/libhydra_cuda_layer.so not present (built where /root/reference exists; travels as a binary)")
    lay = layerapi.CppLayer(scn.width, scn.height)
    layerapi.load_scene_like_render_driver(lay, scn, consts)
    return lay


def _rel_rmse(a, b):
    return float(np.sqrt(((a - b)**2).mean())/max(np.sqrt((b**2).mean()), 1e-12))


@pytest.mark.parametrize("name,kw", [("cornell", dict()), ("cornell_two_lights_dof", dict(two_lights=True, dof=True))])
def test_ihwlayer_mispt_and_pt_match_reference_golden(consts, golden, name, kw):
    scn = scenes.cornell(64, 64, **kw)
    lay = _make(scn, consts)
    assert not lay.StoreCPUData()                       # no CPU integrator behind the layer
    assert "B200" in lay.GetDeviceName() or "NVIDIA" in lay.GetDeviceName()
    assert lay.DeviceCount() >= 1
    lay.InitPathTracing(777)
    lay.TracingPasses(3)                                 # default integrator = what CPUExpLayer instantiates: MISPT
    assert abs(lay.GetSPP() - 3.0) < 1e-6
    got = lay.GetHDRImage()[..., :3]*np.float32(3)
    want = golden[name + "_mispt_sum3"]
    close = (np.abs(got - want) <= 1e-5*np.maximum(np.abs(want), 1e-3)).all(-1)
    assert close.mean() >= 0.999 and _rel_rmse(got, want) <= 1e-4
    # HRT_STUPID_PT_MODE selects IntegratorStupidPT, as RenderDriverRTE::UpdateSettings would (RenderDriverRTE.cpp:340)
    vi, vf, fl = lay.GetAllFlagsAndVars()
    lay.SetAllFlagsAndVars(vi, vf, fl | consts["HRT_STUPID_PT_MODE"])
    lay.InitPathTracing(777)
    lay.TracingPasses(3)
    got = lay.GetHDRImage()[..., :3]*np.float32(3)
    assert _rel_rmse(got, golden[name + "_pt_sum3"]) <= 1e-5
    st = lay.GetRaysStat()
    assert st["raysPerSec"] > 0 and st["samplesPerSec"] > 0
    lay.close()


def test_ihwlayer_qmc_and_equivalence_with_c_abi_path(consts, golden, layer):
    scn = scenes.cornell(64, 64)
    lay = _make(scn, consts)
    lay.CallNamedFunc("integrator", "qmc")
    lay.InitPathTracing(777)
    lay.TracingPasses(3)
    got = lay.GetHDRImage()[..., :3]*np.float32(3)
    assert _rel_rmse(got, golden["cornell_qmc_sum3"]) <= 1e-3
    # same scene through the plain C ABI with the blob packed by scene.py: identical pixels (tables differ only in offsets)
    lay.CallNamedFunc("integrator", "mispt")
    lay.InitPathTracing(5)
    lay.TracingPasses(2)
    a = lay.GetHDRImage()
    layer.LoadScene(scn)
    layer.InitPathTracing(5)
    layer.TracingPass(2, 2)
    b = layer.GetHDRImage()
    assert np.array_equal(a, b)
    ldr = lay.GetLDRImage()
    assert np.array_equal(ldr, layer.GetLDRImage())
    lay.ClearAccumulatedColor()
    assert lay.GetSPP() == 0.0 and lay.GetHDRImage().max() == 0.0
    lay.close()


@pytest.mark.parametrize("which", ["cornell", "delta_lights_with_sun"])
def test_reference_assembled_globals_equal_python_packer(consts, which):
    """IHWLayerDataAssembler.cpp (reference code, compiled in place) vs hydracore_b200/scene.py::_pack_globals."""
    C = consts
    scn = scenes.cornell(64, 48, two_lights=True, dof=True) if which == "cornell" else scenes.cornell_spot_and_direct_lights(64, 48, True)
    lay = _make(scn, consts)
    ref = lay.EngineGlobalsBlob()
    ours = scn.globals_blob
    rb, ob = ref.view(np.uint8), ours.view(np.uint8)

    def same(off, nbytes):
        return np.array_equal(rb[off:off + nbytes], ob[off:off + nbytes])

    for k in ("EG_mProjInverse", "EG_mWorldViewInverse", "EG_mProj", "EG_mWorldView"):
        assert same(C[k], 64), k
    assert same(C["EG_rmQMC"], 64)
    assert same(C["EG_varsF"], 256)
    vi_r = ref[C["EG_varsI"]//4:C["EG_varsI"]//4 + 64]
    vi_o = ours[C["EG_varsI"]//4:C["EG_varsI"]//4 + 64]
    assert np.array_equal(vi_r, vi_o)
    assert same(C["EG_camForward"], 36)
    assert same(C["EG_imagePlaneDist"], 4)
    for k in ("EG_lightsNum", "EG_skyLightId", "EG_sunNumber", "EG_materialsTableSize", "EG_geometryTableSize", "EG_lightSelectorTableSizeRev"):
        assert ref[C[k]//4] == ours[C[k]//4], k
    assert ref[C["EG_sunNumber"]//4] == (8 if which == "delta_lights_with_sun" else 0)      # the first soft directional light, MAX_SUN_NUM times
    assert same(C["EG_suns"], 8*512)
    assert same(C["EG_m_essGgx2017Table"], 64*64*2) or scn.ms_tables is None
    nl = ref[C["EG_lightsNum"]//4]
    lr, lo = ref[C["EG_lightsOffset"]//4], ours[C["EG_lightsOffset"]//4]
    assert np.array_equal(ref[lr:lr + nl*128], ours[lo:lo + nl*128])
    sr, so = ref[C["EG_lightSelectorTableOffsetRev"]//4], ours[C["EG_lightSelectorTableOffsetRev"]//4]
    assert np.array_equal(ref[sr:sr + nl + 1], ours[so:so + nl + 1])
    lay.close()


def test_memory_storage_cuda_contract(consts):
    from tests import layerapi
    if not layerapi.CppLayer.available():
        pytest.skip("layer library not present")
    lay = layerapi.CppLayer(32, 32)
    for name in ("textures", "textures_aux", "geom", "materials", "pdfs"):
        lay.CreateMemStorage(name, 1 << 16)
    assert lay.StorageUpdate("geom", 0, np.ones(64, np.uint8)) == 0 and lay.StorageUpdate("textures", 1, np.ones(64, np.uint8)) == 0
    assert lay.StorageInfo("geom")[2] and lay.StorageInfo("textures")[2]                 # host mirrors (GPUOCLData.cpp:56-63)
    assert not lay.StorageInfo("materials")[2] and not lay.StorageInfo("pdfs")[2]        # device only: GetBegin() == nullptr
    a = np.arange(100, dtype=np.uint8)
    assert lay.StorageUpdate("materials", 0, a) == 0                                      # offsets are in 16-byte blocks
    assert lay.StorageUpdate("materials", 3, a) == 7                                      # 100 B -> 7 blocks
    assert lay.StorageUpdate("materials", 0, a[:50]) == 0                                 # fits in place
    assert lay.StorageUpdate("materials", 0, np.zeros(200, np.uint8)) == 14               # does not fit: appended
    assert lay.StorageInfo("materials")[0] == (14 + 13)*16
    assert lay.StorageUpdate("materials", 9, np.zeros(1 << 17, np.uint8)) == -1           # beyond the reservation: -1, no exception
    with pytest.raises(layerapi.LayerError):
        lay.CreateMemStorage("no_such_storage", 64)
    lay.close()


def test_ihwlayer_sky_dome_scene_equals_c_abi_path(consts, layer):
    """A scene with a sky-dome light (pdf table in the "pdfs" storage, skyLightId found by the reference's SetAllPODLights) through IHWLayer."""
    scn = scenes.open_box_under_sky(64, 64, True)
    lay = _make(scn, consts)
    lay.InitPathTracing(9)
    lay.TracingPasses(2)
    a = lay.GetHDRImage()
    layer.LoadScene(scn)
    layer.InitPathTracing(9)
    layer.TracingPass(2, 2)
    assert a[..., :3].mean() > 0.1 and np.array_equal(a, layer.GetHDRImage())
    lay.close()


def test_ihwlayer_sample_streams_one_call_carries_several_passes(consts, layer):
    """CallNamedFunc("sample_streams", "4"): one BeginTracingPass / EndTracingPass is one wavefront of four passes (GetSPP says so) and the image is
    the C-ABI path's with the same streams."""
    scn = scenes.cornell(64, 64)
    lay = _make(scn, consts)
    lay.CallNamedFunc("sample_streams", "4")
    lay.InitPathTracing(9)
    lay.TracingPasses(2)
    assert abs(lay.GetSPP() - 8.0) < 1e-4
    a = lay.GetHDRImage()
    try:
        layer.SetSampleStreams(4)
        layer.LoadScene(scn)
        layer.InitPathTracing(9)
        layer.TracingPass(2, 8)
        assert np.array_equal(a, layer.GetHDRImage())
    finally:
        layer.SetSampleStreams(1)
    lay.close()


def test_ihwlayer_two_trees_with_alpha_table_equals_c_abi_path(consts, layer):
    """SetAllBVH4 with a two-tree ConvertionResult (tree 1 = meshes with opacity maps + pTriangleAlpha, GPUOCLData.cpp:103-116) through IHWLayer."""
    scn = scenes.cornell_with_cutout(64, 64)
    lay = _make(scn, consts)
    lay.CallNamedFunc("shadow_trees", "0")        # like the session's C-ABI layer: shadow rays as the CPU integrators trace them
    lay.InitPathTracing(9)
    lay.TracingPasses(2)
    a = lay.GetHDRImage()
    layer.LoadScene(scn)
    layer.InitPathTracing(9)
    layer.TracingPass(2, 2)
    assert a[..., :3].mean() > 0.05 and np.array_equal(a, layer.GetHDRImage())
    lay.close()


def test_ihwlayer_normal_mapped_scene_equals_c_abi_path(consts, layer):
    """Normal maps through IHWLayer: images in the "textures_aux" storage, the aux texture table assembled by the reference's PrepareEngineTables."""
    scn = scenes.cornell_normal_mapped(64, 64)
    lay = _make(scn, consts)
    lay.InitPathTracing(9)
    lay.TracingPasses(2)
    a = lay.GetHDRImage()
    layer.LoadScene(scn)
    layer.InitPathTracing(9)
    layer.TracingPass(2, 2)
    assert a[..., :3].mean() > 0.05 and np.array_equal(a, layer.GetHDRImage())
    lay.close()


def test_ihwlayer_remap_lists_equal_c_abi_path(consts, layer):
    """SetAllRemapLists / SetAllInstIdToRemapId (IHWLayer.h:122-123) through the virtual interface."""
    scn = scenes.cornell_remap_lists(64, 64)
    lay = _make(scn, consts)
    lay.InitPathTracing(9)
    lay.TracingPasses(2)
    a = lay.GetHDRImage()
    layer.LoadScene(scn)
    layer.InitPathTracing(9)
    layer.TracingPass(2, 2)
    assert a[..., :3].mean() > 0.05 and np.array_equal(a, layer.GetHDRImage())
    lay.close()


def test_shared_image_accumulation_like_the_reference_worker_mode(consts):
    """ContribToExternalImageAccumulator (GPUOCLLayerOther.cpp:365-430): two render processes (here: two layers with different seeds, full frame
    each) add their SUM buffers into one shared image and hand over their sample counts; the device buffer starts over afterwards."""
    from tests import layerapi
    scn = scenes.cornell(48, 48)
    shared = layerapi.SharedImage(48, 48)
    sums = []
    for seed, passes in ((11, 2), (12, 3)):
        lay = _make(scn, consts)
        lay.InitPathTracing(seed)
        lay.TracingPasses(passes)
        sums.append(lay.GetHDRImage()*np.float32(passes))
        lay.ContribToExternalImageAccumulator(shared)
        assert lay.GetSPP() == 0.0 and lay.GetSPPContrib() == float(passes) and lay.GetHDRImage().max() == 0.0
        lay.close()
    img, spp, cnt, locked = shared.read()
    assert spp == 5.0 and cnt == 2 and not locked
    want = sums[0] + sums[1]
    assert np.abs(img[..., :3] - want[..., :3]).max() <= 1e-5*max(1.0, float(want[..., :3].max()))
    shared.close()
