"""One process driving two GPUs (SURVEY.md 8(e) allows it beside one process per GPU): launch attributes, occupancy and
scratch buffers of the library are cached per device, and a batch made for cuda:1 runs there whatever the current
device is.  Needs two GPUs (gpurun --gpus 2); skipped otherwise."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_devices_one_process_bit_identical():
    from gym_cloth_b200 import lib as L
    from gym_cloth_b200.batched import BatchedCloth
    rng = np.random.RandomState(5)
    acts = rng.uniform(-0.8, 0.8, size=(6, 4))
    res = {}
    # f64 needs > 48 KB of shared memory (cudaFuncSetAttribute per device); the sliced launch needs the slot count
    for dev in ("cuda:0", "cuda:1"):
        for dt in (torch.float64, torch.float32):
            torch.cuda.set_device(0)                       # the current device is NOT the batch's device for cuda:1
            bc = BatchedCloth(L.default_params(), 6, dtype=dt, device=dev)
            host = {"coverage": np.zeros(6), "sim_steps": np.zeros(6, np.int32), "flags": np.zeros(6, np.int32)}
            bc.step_host(acts, host)
            bc.step_actions(torch.from_numpy(acts).to(dev, dt))
            torch.cuda.synchronize(dev)
            assert bc.pos.device == torch.device(dev)
            res[(dev, dt)] = (bc.pos.cpu().numpy(), bc.coverage.cpu().numpy(), host["sim_steps"].copy())
    for dt in (torch.float64, torch.float32):
        a, b = res[("cuda:0", dt)], res[("cuda:1", dt)]
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        assert (a[2] > 0).all()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sliced_launch_on_second_device():
    from gym_cloth_b200 import lib as L
    from gym_cloth_b200.batched import BatchedCloth
    lib = L.lib()
    rng = np.random.RandomState(6)
    acts = rng.uniform(-0.8, 0.8, size=(12, 4))
    out = []
    for dev in ("cuda:0", "cuda:1"):
        torch.cuda.set_device(0)
        bc = BatchedCloth(L.default_params(), 12, dtype=torch.float32, device=dev)
        lib.clothb200_debug_set_slicing(3, 50)
        try:
            bc.step_actions(torch.from_numpy(acts).to(dev, torch.float32)); torch.cuda.synchronize(dev)
        finally:
            lib.clothb200_debug_set_slicing(0, 0)
        out.append(bc.pos.cpu().numpy())
    assert np.array_equal(out[0], out[1])
