"""The peer-to-peer transport primitives on ONE GPU (local memory stands in for the peer mapping; the 2+ GPU runs are in
test_multigpu_gpu.py): lgcu_copy_rows, the flag kernels, and lgcu_exchange — one fused exchange step (signal -> wait -> copy -> last CTA
acknowledges) — against the same step made of the separate calls, over several frames as under CUDA-graph replay."""
import ctypes as C

import numpy as np
import pytest

from legitengine_b200 import abi

pytestmark = pytest.mark.gpu


def _ptrs(addrs):
    return (C.c_void_p * max(len(addrs), 1))(*addrs), len(addrs)


def test_fused_exchange_step_equals_the_separate_calls():
    import torch

    lib = abi.load_lgcu()
    dev = "cuda:0"
    rng = np.random.default_rng(3)
    # 70 slabs would need two copy launches; the fused step takes at most 64: use 40 slabs of ragged sizes (multiples of 16 bytes)
    sizes = [16 * int(v) for v in rng.integers(1, 3000, size=40)]
    sizes[5] = 0  # empty slabs are skipped
    src = torch.from_numpy(rng.integers(0, 255, size=sum(sizes) + 64, dtype=np.uint8)).to(dev)
    dst_a, dst_b = torch.zeros_like(src), torch.zeros_like(src)
    flags = torch.zeros(64, dtype=torch.int32, device=dev)  # [0..7] signalBefore, [8..15] signalAfter, [32] frame counter, [33] done counter
    base = flags.data_ptr()
    counter, done = base + 4 * 32, base + 4 * 33

    def slabs(dst):
        items, off = [], 0
        for n in sizes:
            items.append(abi.RowCopy(src.data_ptr() + off, dst.data_ptr() + off, n))
            off += n
        return (abi.RowCopy * len(items))(*items), len(items)

    before, after = _ptrs([base + 4 * i for i in range(8)]), _ptrs([base + 4 * (8 + i) for i in range(8)])
    none = _ptrs([])
    open_step = abi.ExchangeDesc(none[0], 0, none[0], 0, 0, None, 0, none[0], 0, C.c_void_p(counter), C.c_void_p(done), 1)
    cp = slabs(dst_a)
    # the step waits on the very flags it signals (a rank whose sources are itself), with lag 0
    step = abi.ExchangeDesc(before[0], before[1], before[0], before[1], 0, cp[0], cp[1], after[0], after[1], C.c_void_p(counter), C.c_void_p(done), 0)
    stream = torch.cuda.current_stream().cuda_stream
    for frame in (1, 2, 3):
        dst_a.zero_()
        assert lib.lgcu_exchange(C.byref(open_step), C.c_void_p(stream)) == 0
        assert lib.lgcu_exchange(C.byref(step), C.c_void_p(stream)) == 0
        torch.cuda.synchronize()
        f = flags.cpu().numpy()
        assert f[32] == frame and (f[:16] == frame).all() and f[33] == 0, f[:34]
        assert torch.equal(dst_a[: sum(sizes)], src[: sum(sizes)]) and int(dst_a[sum(sizes):].sum()) == 0
    # the same with the separate calls
    cpb = slabs(dst_b)
    assert lib.lgcu_frame_counter_bump(C.c_void_p(counter), C.c_void_p(stream)) == 0
    assert lib.lgcu_signal_flags(before[0], before[1], C.c_void_p(counter), C.c_void_p(stream)) == 0
    assert lib.lgcu_wait_flags(before[0], before[1], C.c_void_p(counter), 0, C.c_void_p(stream)) == 0
    assert lib.lgcu_copy_rows(cpb[0], cpb[1], C.c_void_p(stream)) == 0
    assert lib.lgcu_signal_flags(after[0], after[1], C.c_void_p(counter), C.c_void_p(stream)) == 0
    torch.cuda.synchronize()
    assert torch.equal(dst_a, dst_b) and (flags.cpu().numpy()[:16] == 4).all()
    # malformed steps are refused before anything is launched
    bad = abi.ExchangeDesc(none[0], 0, none[0], 0, 0, cp[0], cp[1], none[0], 0, C.c_void_p(counter), C.c_void_p(done), 1)  # bump + copies
    assert lib.lgcu_exchange(C.byref(bad), C.c_void_p(stream)) == abi.LGCU_ERR_INVALID_ARGUMENT
    bad = abi.ExchangeDesc(none[0], 0, none[0], 0, 0, None, 0, after[0], after[1], C.c_void_p(counter), None, 0)  # acknowledgement without a done counter
    assert lib.lgcu_exchange(C.byref(bad), C.c_void_p(stream)) == abi.LGCU_ERR_INVALID_ARGUMENT
