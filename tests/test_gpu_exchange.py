"""spmv_b200_mail_exchange on one device (world = 1: nobody to talk to, so the kernel is just the fixed-order sum of the
flat product's partials): the single-CTA form (up to 8192 partials) and the multi-CTA form (chunk sums + last-CTA ticket)
against a float64 sum computed by torch, deterministic from launch to launch, the ticket counter back at zero, and the
two-launch iteration built on it against the plain product + norm."""
import ctypes as C

import pytest

from sparsematrixvectormultiplication_b200 import _native as N
from sparsematrixvectormultiplication_b200 import device, synth

pytestmark = pytest.mark.gpu


def _mail(sync):
    import torch
    box = torch.zeros(N.MAILBOX_BYTES // 8, dtype=torch.int64, device="cuda")
    m = N.Mail()
    m.world, m.rank, m.iteration = 1, 0, 0
    m.box[0] = box.data_ptr()
    m.counter = sync.data_ptr()
    m.status = sync.data_ptr() + 4
    return m, box


@pytest.mark.parametrize("count", [1, 5, 1023, 8192, 8193, 20000, 65536, 300001])
def test_exchange_sum_matches_torch_and_is_deterministic(count):
    import torch
    torch.manual_seed(count)
    src = torch.rand(count, dtype=torch.float64, device="cuda") + 0.5
    sync = torch.zeros(2, dtype=torch.int32, device="cuda")
    mail, _box = _mail(sync)
    outs = []
    for _ in range(3):
        partials = src.clone()                      # consumed by the multi-CTA form
        out = torch.zeros(2, dtype=torch.float64, device="cuda")
        device.mail_exchange(partials, count, mail, out)
        torch.cuda.synchronize()
        outs.append(out.cpu())
        assert int(sync[0].item()) == 0 and int(sync[1].item()) == 0      # ticket counter reset, no timeout
    ref = float(src.sum().item())
    assert abs(float(outs[0][0]) - ref) <= 1e-13 * ref
    assert abs(float(outs[0][1]) - ref ** -0.5) <= 1e-13 * ref ** -0.5
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])    # fixed order: bitwise repeatable


def test_two_launch_iteration_on_a_matrix_large_enough_for_the_multi_cta_sum():
    import torch
    n = 140                                             # 2.7 M rows -> 10 719 partials: three CTAs
    A = device.DeviceCSR.synth(synth.SYNTH_LAP3D, n)
    rows = n ** 3
    assert A.flat_partials_count() > 8192
    xs = [torch.ones(rows, dtype=torch.float64, device="cuda"), torch.empty(rows, dtype=torch.float64, device="cuda")]
    partials = torch.zeros(A.flat_partials_count(), dtype=torch.float64, device="cuda")
    scale = torch.ones(2, dtype=torch.float64, device="cuda")
    sync = torch.zeros(2, dtype=torch.int32, device="cuda")
    mail, _box = _mail(sync)
    x_ref = torch.ones(rows, dtype=torch.float64, device="cuda")
    y_ref = torch.empty(rows, dtype=torch.float64, device="cuda")
    for k in range(6):
        cur, nxt = k & 1, (k & 1) ^ 1
        A.spmv_fused_flat(xs[cur].data_ptr(), xs[nxt].data_ptr(), inv_norm=scale.data_ptr() + 8 if k else None, partials=partials)
        mail.iteration = k
        device.mail_exchange(partials, partials.numel(), mail, scale)
        A.spmv(x_ref, y_ref)                            # the plain iteration: y = A x; x = y / |y|
        lam_ref = float(torch.linalg.vector_norm(y_ref).item())
        x_ref = y_ref / lam_ref
        torch.cuda.synchronize()
        lam = float(scale[0].item()) ** 0.5
        assert abs(lam - lam_ref) <= 1e-12 * lam_ref, (k, lam, lam_ref)
    v = xs[0] / (float(scale[0].item()) ** 0.5)
    assert float((v - x_ref).abs().max().item()) <= 1e-12 * float(x_ref.abs().max().item())
    assert int(sync[0].item()) == 0 and int(sync[1].item()) == 0
