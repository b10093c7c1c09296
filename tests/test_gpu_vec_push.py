"""spmv_b200_vec_push -- the all-gather of a slice written against peer memory -- on ONE device: the "peers" are other
buffers of the same GPU, so the copy must be exact for every alignment of the slice (row ranges of the nnz-balanced
partition start anywhere), every length around the vector width, and must not touch a byte outside the slice.  The
multi-GPU use of the same kernel is checked by tools/dist_check.py (tests/test_gpu_distributed.py) and by the parity
block of bench.py --gpus N (mode allgather_peer_kernel)."""
import pytest

from sparsematrixvectormultiplication_b200 import device

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("npeers", [1, 3, 7])
def test_vec_push_copies_exactly_the_slice(npeers):
    import torch
    torch.manual_seed(7)
    total = 1 << 16
    src = torch.rand(total, dtype=torch.float64, device="cuda")
    lengths = list(range(0, 14)) + [255, 256, 257, 4099, 50001]
    for offset in range(0, 5):
        for n in lengths:
            dsts = [torch.full((total,), -1.0, dtype=torch.float64, device="cuda") for _ in range(npeers)]
            device.vec_push(src[offset: offset + n] if n else src[offset: offset + 1], n,
                            [d.data_ptr() + 8 * offset for d in dsts], ctas=(3 if n < 300 else 0))
            torch.cuda.synchronize()
            for d in dsts:
                assert torch.equal(d[offset: offset + n], src[offset: offset + n]), (offset, n)
                assert bool((d[:offset] == -1.0).all()) and bool((d[offset + n:] == -1.0).all()), (offset, n)


def test_vec_push_scalar_path_for_differently_aligned_targets():
    import torch
    src = torch.arange(1000, dtype=torch.float64, device="cuda")
    dst = torch.zeros(1003, dtype=torch.float64, device="cuda")
    device.vec_push(src, 1000, [dst.data_ptr() + 8 * 3])      # source 32-byte aligned, target 24 bytes past
    torch.cuda.synchronize()
    assert torch.equal(dst[3:], src) and bool((dst[:3] == 0).all())


def test_vec_push_rejects_bad_arguments():
    import torch
    from sparsematrixvectormultiplication_b200 import _native as N
    src = torch.zeros(8, dtype=torch.float64, device="cuda")
    with pytest.raises(N.SpmvError):
        device.vec_push(src, 8, [0])                          # NULL target
    with pytest.raises(N.SpmvError):
        device.vec_push(src, 8, [src.data_ptr()] * 8)         # more than SPMV_B200_MAX_PEERS targets
