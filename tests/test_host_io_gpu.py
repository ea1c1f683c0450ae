"""Host <-> device transfers of the library (csrc/host_io.cu): the pinned bounce ring in both directions against plain
copies, at sizes around the ring geometry (4 slots of 32 MB) and with a dtype change on the device."""
import numpy as np
import pytest
import torch

from chrysalis_b200 import device as dv

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nbytes", [8 << 20, (32 << 20) - 8, (32 << 20) + 8, 3 * (32 << 20), 5 * (32 << 20) + 4104])
def test_upload_and_download_through_the_ring_are_exact(nbytes):
    rng = np.random.default_rng(nbytes)
    a = rng.integers(0, 2 ** 63 - 1, nbytes // 8, dtype=np.int64)
    d = dv.upload_array(a)
    assert torch.equal(d.cpu(), torch.from_numpy(a))
    back = dv.to_host(d)
    assert back.dtype == a.dtype and np.array_equal(back, a)
    # an upload queued right after a download reuses the ring's slots
    d2 = dv.upload_array(back[::-1].copy())
    assert torch.equal(d2.cpu(), torch.from_numpy(a[::-1].copy()))


def test_to_host_shapes_and_dtypes():
    for dt in (torch.float32, torch.float64):
        t = torch.randn((700_000, 12), dtype=dt, device="cuda")
        h = dv.to_host(t)
        assert h.shape == (700_000, 12) and np.array_equal(h, t.cpu().numpy())
    v = torch.randn((1_000_000, 6), device="cuda")[:, :5]          # non-contiguous view
    assert np.array_equal(dv.to_host(v), v.cpu().numpy())
    f32 = np.random.default_rng(0).normal(size=(700_000, 30)).astype(np.float32)
    assert torch.equal(dv.upload_array(f32, torch.float64).cpu(), torch.from_numpy(f32).to(torch.float64))
