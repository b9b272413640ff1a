"""GPU: the prologue-free ("lean") producers of lb_conv_fwd -- tile-mask table, indices read straight from the neighbour
table, TMA tile loads for identity rows -- against the per-tile-prologue path of the same kernel (LB_CONV_NO_LEAN).
Both run the same MMA sequence, so outputs must be bit-identical; the prologue path itself is checked against the
oracle in test_gpu_conv.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def random_table(n, k, seed, p_lo=0.05, p_hi=0.6):
    g = torch.Generator(device="cuda").manual_seed(seed)
    prob = torch.linspace(p_lo, p_hi, k, device="cuda")[torch.randperm(k, device="cuda", generator=g)]
    valid = torch.rand((k, n), device="cuda", generator=g) < prob[:, None]
    idx = torch.randint(0, n, (k, n), device="cuda", generator=g, dtype=torch.int32)
    return torch.where(valid, idx, torch.full_like(idx, -1))


@pytest.mark.parametrize("n,k,ld_pad", [(1000, 27, 0), (70001, 27, 3), (513, 8, 1), (128, 27, 0)])
def test_tile_masks_vs_numpy(n, k, ld_pad):
    from lidal_b200 import engine
    full = torch.full((k, n + ld_pad), -1, dtype=torch.int32, device="cuda")
    table = full[:, :n]
    table.copy_(random_table(n, k, seed=n, p_lo=0.0005, p_hi=0.05))
    got = engine.tile_masks_of(table).cpu().numpy().view(np.uint32)
    t = table.cpu().numpy() >= 0
    groups = (n + 127) // 128
    want = np.zeros(groups, np.uint32)
    for g in range(groups):
        rows = t[:, g * 128:(g + 1) * 128].any(1)
        want[g] = sum(1 << j for j in range(k) if rows[j])
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n,k,cin,cout,res,slice_in", [
    (200000, 27, 32, 32, True, False), (200000, 27, 96, 96, True, True), (200000, 27, 128, 96, False, False),
    (200000, 27, 64, 64, True, False), (200000, 27, 256, 128, False, False), (200000, 8, 96, 96, False, False),
    (20000, 27, 128, 128, True, False), (20000, 27, 256, 256, True, False), (20000, 8, 256, 128, False, True),
    (333, 27, 64, 32, True, False), (30000, 27, 192, 128, False, False)])
def test_lean_gather_bit_identical(n, k, cin, cout, res, slice_in):
    from lidal_b200 import engine
    g = torch.Generator().manual_seed(n + cin)
    nbr = random_table(n, k, seed=cin + cout)
    sm = engine._mask_sorted(nbr)
    assert sm.tile_masks is not None
    wide = torch.randn(n, cin + 32, generator=g).cuda().bfloat16()
    x = wide[:, 32:] if slice_in else wide[:, :cin].contiguous()
    conv = engine._Conv((torch.randn(k, cin, cout, generator=g) * 0.05).cuda(), None, relu=True)
    residual = torch.randn(n, cout, generator=g).cuda().bfloat16() if res else None
    lean = conv(x, sm, n, residual=residual)
    torch.cuda.synchronize()
    base = conv(x, sm, n, residual=residual, no_lean=True)
    torch.cuda.synchronize()
    assert torch.equal(lean.view(torch.int16), base.view(torch.int16))
    # and against a plain fp32 evaluation of the same gather-GEMM (independent of either kernel)
    if n <= 30000:
        xf, w = x.float(), conv.w.float()              # packed weight [k, cout, cin]
        want = torch.zeros(n, cout, device="cuda")
        for j in range(k):
            rows = nbr[j].long()
            ok = rows >= 0
            want[ok] += xf[rows[ok]] @ w[j].t()
        if residual is not None:
            want += residual.float()
        want = want.relu()
        err = float((lean.float() - want).norm() / want.norm())
        assert err < 1e-2, err


@pytest.mark.parametrize("n,cin,cout,res,f32out", [
    (200000, 32, 256, True, False), (200000, 128, 96, True, False), (200000, 96, 32, False, True), (50000, 384, 256, False, False),
    (131, 64, 128, False, False), (200000, 256, 128, True, False)])
def test_lean_identity_rows_by_tma(n, cin, cout, res, f32out):
    """1x1 layers (nbr = None): A operand by TMA tile loads vs the cp.async path; input is a column slice of a wider buffer."""
    from lidal_b200 import engine
    g = torch.Generator().manual_seed(cin * 7 + cout)
    wide = torch.randn(n, cin + 64, generator=g).cuda().bfloat16()
    x = wide[:, 64:]
    bn = torch.nn.BatchNorm1d(cout).cuda().eval()
    with torch.no_grad():
        bn.running_mean.normal_(generator=None)
        bn.running_var.uniform_(0.5, 1.5)
    conv = engine._Conv((torch.randn(1, cin, cout, generator=g) * 0.05).cuda(), bn, relu=True)
    residual = torch.randn(n, cout, generator=g).cuda().bfloat16() if res else None
    od = torch.float32 if f32out else None
    lean = conv(x, None, n, residual=residual, relu_first=res, out_dtype=od)
    torch.cuda.synchronize()
    base = conv(x, None, n, residual=residual, relu_first=res, out_dtype=od, no_lean=True)
    torch.cuda.synchronize()
    assert torch.equal(lean, base)
    want = (x.float() @ conv.w[0].float().t()) * conv.scale + conv.shift
    want = (want.relu() + residual.float()) if res else want.relu()
    err = float((lean.float() - want).norm() / want.norm())
    assert err < 1e-2, err


def test_sort_emits_the_same_tile_masks_as_the_table_pass():
    """lb_kmap_sort_by_mask_tm derives the tile masks from the per-row masks it already holds; lb_kmap_tile_masks re-reads
    the sorted table: both routes must agree."""
    from lidal_b200 import engine
    for n, k in ((50000, 27), (4097, 8), (100, 27)):
        sm = engine._mask_sorted(random_table(n, k, seed=n + k, p_lo=0.001, p_hi=0.3))
        assert torch.equal(sm.tile_masks, engine.tile_masks_of(sm[0]))
