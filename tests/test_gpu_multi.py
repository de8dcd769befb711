"""The multi-GPU renderer of the C ABI (forma_renderer_multi_*): one process, tile-row bands,
one worker per device. With one device it must behave like forma_renderer_render; with two or
more (when the box has them) the frame assembled from the devices' bands must equal the
oracle's, in host memory and in the first device's memory, while the bands are rebalanced from
frame to frame and every device only keeps its band's geometry."""
import numpy as np
import pytest

import synth
from forma_b200.binding import RGBA, Color, Rect

pytestmark = pytest.mark.gpu

W, H = 1920, 1080
CLEAR = Color(1.0, 1.0, 1.0, 0.0)


def _scene(api, seed=41, n=1500):
    comp = api.Composition()
    synth.random_mixed(api, comp, n, W, H, seed)
    return comp


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def expected(oracle_api):
    buf = np.zeros(W * H * 4, np.uint8)
    oracle_api.Renderer(0).render(_scene(oracle_api), buf, W, H, RGBA, CLEAR)
    return buf


@pytest.mark.parametrize("n_dev", [1, 2, 4, 8])
def test_multi_renderer_host_frame_matches_oracle(cuda_api, expected, n_dev):
    if _device_count() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    m = cuda_api.MultiRenderer(list(range(n_dev)))
    comp = _scene(cuda_api)
    for frame in range(3):  # frame 0: equal bands; later frames: bands balanced on the previous frame's row costs
        buf = np.zeros(W * H * 4, np.uint8)
        t = m.render(comp, buf, W, H, RGBA, CLEAR)
        assert np.array_equal(buf, expected), f"{n_dev} devices, frame {frame}"
        bounds, ms = m.bands()
        assert bounds[0] == 0 and bounds[-1] == (H + 15) // 16 and all(a <= b for a, b in zip(bounds, bounds[1:]))
    assert t.n_segments > 0


@pytest.mark.parametrize("n_dev", [1, 2, 4, 8])
def test_multi_renderer_device_frame_and_crop(cuda_api, oracle_api, expected, n_dev):
    import torch
    if _device_count() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    m = cuda_api.MultiRenderer(list(range(n_dev)))
    comp = _scene(cuda_api)
    fb = torch.zeros(W * H * 4, dtype=torch.uint8, device="cuda:0")
    for _ in range(2):
        fb.zero_()
        m.render_device(comp, fb.data_ptr(), W, H, RGBA, CLEAR)
        for d in range(n_dev):
            torch.cuda.synchronize(d)
        assert np.array_equal(fb.cpu().numpy(), expected)
    # A cropped render only touches the crop's tiles.
    crop = Rect((256, 1024), (128, 700))
    got = np.full(W * H * 4, 7, np.uint8)
    m.render(comp, got, W, H, RGBA, CLEAR, crop)
    want = np.full(W * H * 4, 7, np.uint8)
    oracle_api.Renderer(0).render(_scene(oracle_api), want, W, H, RGBA, CLEAR, crop)
    assert np.array_equal(got, want)


def test_multi_renderer_rejects_bad_device_lists(cuda_api):
    from forma_b200.binding import FormaError
    with pytest.raises(FormaError):
        cuda_api.MultiRenderer([])
    with pytest.raises(FormaError):
        cuda_api.MultiRenderer([0, 0])
    with pytest.raises(FormaError):
        cuda_api.MultiRenderer([99])


def test_band_render_keeps_only_the_bands_geometry(cuda_api, cuda_renderer, oracle_api, unsliced):
    """A render cropped to a band uploads only the inserts that can reach the band (h2d bytes
    well below the whole composition's), and frames rendered afterwards with other bands or the
    whole frame are still exact."""
    comp = _scene(cuda_api, seed=43, n=3000)
    r = cuda_api.Renderer(0)
    want = np.zeros(W * H * 4, np.uint8)
    oracle_api.Renderer(0).render(_scene(oracle_api, seed=43, n=3000), want, W, H, RGBA, CLEAR)
    b0 = r.counters()["h2d_bytes"]
    got = np.zeros(W * H * 4, np.uint8)
    r.render(comp, got, W, H, RGBA, CLEAR, Rect((0, W), (0, 128)))
    band_bytes = r.counters()["h2d_bytes"] - b0
    comp.evict()
    b1 = r.counters()["h2d_bytes"]
    r.render(comp, got, W, H, RGBA, CLEAR, Rect((0, W), (128, H)))  # a different band: re-uploaded for it
    rest_bytes = r.counters()["h2d_bytes"] - b1
    assert np.array_equal(got, want)
    comp.evict()
    b2 = r.counters()["h2d_bytes"]
    whole = np.zeros(W * H * 4, np.uint8)
    r.render(comp, whole, W, H, RGBA, CLEAR)
    whole_bytes = r.counters()["h2d_bytes"] - b2
    assert np.array_equal(whole, want)
    assert band_bytes < 0.5 * whole_bytes and rest_bytes < whole_bytes
    # narrowing inside the resident band needs no upload at all; widening re-uploads
    b3 = r.counters()["h2d_bytes"]
    r.render(comp, whole, W, H, RGBA, CLEAR, Rect((0, W), (256, 512)))
    assert r.counters()["h2d_bytes"] - b3 < 4096
