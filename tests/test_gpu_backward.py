"""GPU parity of the backward kernels (C ABI) against torch autograd on the same fp32 / bf16-rounded inputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import dicow_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from ts_asr_whisper_b200 import ops as _ops
    return _ops


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp(min=1e-6)).item()


@pytest.mark.parametrize("d,T,B,fddt,ln", [(1280, 300, 2, True, True), (384, 77, 3, False, True), (128, 50, 2, True, False),
                                           (1280, 1501, 1, True, True)])
def test_layernorm_fddt_backward(ops, d, T, B, fddt, ln):
    g = torch.Generator(device=DEV).manual_seed(d + T)
    rows = B * T
    x = torch.randn(rows, d, device=DEV, generator=g)
    d1 = (torch.randn(rows, d, device=DEV, generator=g) * 0.3).bfloat16()
    d2 = (torch.randn(rows, d, device=DEV, generator=g) * 0.3).bfloat16()
    stno = torch.softmax(3 * torch.randn(B, 4, T, device=DEV, generator=g), dim=1)
    fw = (torch.rand(4, d, device=DEV, generator=g) + 0.5).requires_grad_(True)
    fb = (torch.randn(4, d, device=DEV, generator=g) * 0.1).requires_grad_(True)
    gam = (torch.rand(d, device=DEV, generator=g) + 0.5).requires_grad_(True)
    bet = (torch.randn(d, device=DEV, generator=g) * 0.1).requires_grad_(True)
    dy = (torch.randn(rows, d, device=DEV, generator=g) * 0.5).bfloat16()
    gin = torch.randn(rows, d, device=DEV, generator=g) * 0.2
    xs = (x + d1.float() + d2.float()).requires_grad_(True)
    xp = xs
    if fddt:
        m = stno.permute(0, 2, 1).reshape(rows, 4)
        xp = sum((xs * fw[c] + fb[c]) * m[:, c:c + 1] for c in range(4))
    loss = (xp * gin).sum()
    if ln:
        loss = loss + (F.layer_norm(xp, (d,), gam, bet, 1e-5) * dy.float()).sum()
    loss.backward()
    gout = torch.full((rows, d), float("nan"), device=DEV)
    gout_b = torch.empty(rows, d, device=DEV, dtype=torch.bfloat16)
    dgam, dbet = torch.zeros(d, device=DEV), torch.zeros(d, device=DEV)
    dfw, dfb = torch.zeros(4, d, device=DEV), torch.zeros(4, d, device=DEV)
    ops.layernorm_fddt_bwd(x, gout, dy=dy if ln else None, g_in=gin, gamma=gam.detach() if ln else None, delta1=d1, delta2=d2,
                           T=T, stno=stno if fddt else None, fddt_w=fw.detach() if fddt else None,
                           fddt_b=fb.detach() if fddt else None, g_out_bf16=gout_b, dgamma=dgam, dbeta=dbet, dfddt_w=dfw,
                           dfddt_b=dfb)
    torch.cuda.synchronize()
    assert rel(gout, xs.grad) < 1e-4 and rel(gout_b, xs.grad) < 1e-2
    if ln:
        assert rel(dgam, gam.grad) < 1e-3 and rel(dbet, bet.grad) < 1e-3
    if fddt:
        assert rel(dfw, fw.grad) < 1e-3 and rel(dfb, fb.grad) < 1e-3


@pytest.mark.parametrize("form,M,N,K", [(1, 640, 1536, 384), (2, 640, 1536, 384), (2, 3000, 5120, 1280), (2, 777, 1000, 264),
                                        (0, 12000, 1280, 640)])  # form 0: the library's choice at the fine-tune step's row count
def test_colsum_and_gelu_epilogues(ops, form, M, N, K):
    """form 1 = single-CTA kernel, 2 = CTA pairs (both outputs / the dgelu product leave through TMA stores)"""
    g = torch.Generator(device=DEV).manual_seed(1)
    X = (torch.randn(777, 1003, device=DEV, generator=g)).bfloat16()
    out = torch.ones(1003, device=DEV)
    ops.colsum(X, out, alpha=0.5)
    torch.cuda.synchronize()
    assert rel(out, 1 + 0.5 * X.float().sum(0)) < 1e-3
    # 16-byte vector path: aligned widths, a column block of a fused projection (row stride 3 N), ragged row count
    Y = torch.randn(5001, 3 * 1280, device=DEV, generator=g).bfloat16()
    for cols in (slice(0, 1280), slice(2560, 3840)):
        out = torch.zeros(1280, device=DEV)
        ops.colsum(Y[:, cols], out, alpha=0.125)
        torch.cuda.synchronize()
        assert rel(out, 0.125 * Y[:, cols].float().sum(0)) < 1e-3
    # fc1 training forward saves the pre-activation; the fc2 dgrad multiplies by gelu'(pre)
    A = (torch.randn(M, K, device=DEV, generator=g) * 0.5).bfloat16()
    W = (torch.randn(N, K, device=DEV, generator=g) * 0.08).bfloat16()
    b = torch.randn(N, device=DEV, generator=g) * 0.3
    h = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    pre = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.gemm(A, W, h, epilogue=ops.EPI_GELU_SAVE_BF16, bias=b, aux=pre, flags=form)
    ref_pre = A.float() @ W.float().t() + b
    torch.cuda.synchronize()
    assert rel(pre, ref_pre) < 1e-2 and rel(h, F.gelu(ref_pre)) < 1e-2
    W2 = (torch.randn(K, N, device=DEV, generator=g) * 0.05).bfloat16()  # fc2 weight [d, ffn]
    dY = (torch.randn(M, K, device=DEV, generator=g) * 0.4).bfloat16()
    dpre = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.gemm(dY, W2, dpre, epilogue=ops.EPI_DGELU_BF16, flags=ops.GEMM_W_T | form, aux=pre)
    pf = pre.float().requires_grad_(True)
    (F.gelu(pf) * (dY.float() @ W2.float())).sum().backward()
    torch.cuda.synchronize()
    assert rel(dpre, pf.grad) < 1e-2


def test_conv_dgrad_col2im(ops):
    """dgrad of Conv1d(k3, p1, stride 2) = GEMM against the forward weight (MN-major) + col2im"""
    B, T, C, Co = 2, 150, 128, 256
    g = torch.Generator(device=DEV).manual_seed(2)
    w = (torch.randn(Co, C, 3, device=DEV, generator=g) * 0.05)
    wk = w.permute(0, 2, 1).reshape(Co, 3 * C).contiguous().bfloat16()  # tap-major GEMM weight [Co, 3 C]
    T_out = (T + 2 - 3) // 2 + 1
    dY = (torch.randn(B, T_out, Co, device=DEV, generator=g) * 0.3).bfloat16()
    dcol = torch.empty(B * T_out, 3 * C, device=DEV, dtype=torch.bfloat16)
    ops.gemm(dY.view(B * T_out, Co), wk, dcol, epilogue=ops.EPI_BIAS_BF16, flags=ops.GEMM_W_T)
    dx = torch.empty(B, T, C, device=DEV, dtype=torch.bfloat16)
    ops.conv1d_col2im(dcol, dx, B=B, T=T, T_out=T_out, C_in=C, stride=2, dx_batch_stride=T * C, dx_row_stride=C)
    xin = torch.zeros(B, C, T, device=DEV, requires_grad=True)
    y = F.conv1d(xin, wk.float().view(Co, 3, C).permute(0, 2, 1), stride=2, padding=1)
    y.backward(dY.float().transpose(1, 2))
    torch.cuda.synchronize()
    assert rel(dx, xin.grad.transpose(1, 2)) < 2e-2


@pytest.mark.parametrize("red", ["mean", "sum"])
def test_ctc_backward(ops, red):
    rng = np.random.default_rng(3)
    B, T, V1 = 4, 60, 301
    lg = (torch.from_numpy(rng.normal(size=(B, T, V1)).astype(np.float32)) * 2).to(DEV)
    lab = torch.full((B, 16), -100, dtype=torch.int64)
    for b, n in enumerate([16, 5, 0, 9]):
        lab[b, :n] = torch.from_numpy(rng.integers(0, 12, size=n))  # small alphabet: repeated labels
    lab = lab.to(DEV)
    loss, dl = ops.ctc_loss_fwd_bwd(lg, lab, reduction=red, loss_scale=0.3)
    x = lg.clone().requires_grad_(True)
    ref = orc.ctc_loss(x.cpu(), lab.cpu(), reduction=red) if False else F.ctc_loss(
        F.log_softmax(x, -1).transpose(0, 1), lab, torch.full((B,), T), (lab >= 0).sum(-1), blank=V1 - 1, reduction=red,
        zero_infinity=True)
    (0.3 * ref).backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - ref.item()) < 2e-4 * max(1, abs(ref.item()))
    assert float(dl[:, :, V1:].abs().max()) == 0.0
    e = rel(dl[:, :, :V1], x.grad)
    print(f"ctc grad rel err ({red}): {e:.3e}")
    assert e < 1e-2
    # infeasible alignment (more labels than frames) -> zero_infinity: zero gradient
    lg2 = lg[:, :8].contiguous()
    loss2, dl2 = ops.ctc_loss_fwd_bwd(lg2, lab, reduction=red)
    x2 = lg2.clone().requires_grad_(True)
    F.ctc_loss(F.log_softmax(x2, -1).transpose(0, 1), lab, torch.full((B,), 8), (lab >= 0).sum(-1), blank=V1 - 1,
               reduction=red, zero_infinity=True).backward()
    torch.cuda.synchronize()
    assert rel(dl2[:, :, :V1], x2.grad) < 1e-2


@pytest.mark.parametrize("soft", [True, False])
def test_softlabel_ce_backward(ops, soft):
    from oracle import synth
    rng = np.random.default_rng(4)
    R, V, TSB, NTS = 40, 300, 262, 38
    logits = (torch.from_numpy(rng.normal(size=(R, V)).astype(np.float32)) * 2)
    labels = torch.from_numpy(synth.make_labels("ceb", 4, 10, V, 257, TSB, prefix=(259, 260))).reshape(-1)
    upp = labels.clone()
    upp[::3] = torch.where(upp[::3] >= 0, (upp[::3] + 3) % 250, upp[::3])
    x = logits.clone().requires_grad_(True)
    ref = orc.decoder_loss(x.view(4, 10, V), labels.view(4, 10), upp.view(4, 10), TSB if soft else None, NTS)
    (0.7 * ref).backward()
    n_valid = float((labels != -100).sum()) if soft else float(R)
    sm = orc.timestamp_smoothing(NTS).to(DEV) if soft else None
    dl = ops.softlabel_ce_bwd(logits.to(DEV), labels.to(DEV), upp.to(DEV), ts_begin=TSB, smoothing=sm, soft_mode=soft,
                              scale=0.7 / n_valid)
    torch.cuda.synchronize()
    e = rel(dl[:, :V].cpu(), x.grad)
    print(f"ce grad rel err (soft={soft}): {e:.3e}")
    assert e < 1e-2 and float(dl[:, V:].abs().max()) == 0.0


@pytest.mark.parametrize("rows,cols,ld", [(300, 128, 128), (1500, 1280, 2560), (7, 6, 8)])
def test_gate_bwd(ops, rows, cols, ld):
    """SE-DiCoW gate backward (layers.py:79-93,168): dupd = tanh(g) * G, dgate = (1 - tanh^2 g) * sum(G * upd)"""
    gen = torch.Generator(device="cuda").manual_seed(3)
    G = torch.randn(rows, ld, generator=gen, device="cuda")[:, :cols]
    upd = torch.randn(rows, ld, generator=gen, device="cuda").bfloat16()[:, :cols]
    gate = torch.tensor([0.37], device="cuda")
    dgate = torch.tensor([0.25], device="cuda")  # accumulates
    out = ops.gate_bwd(G, upd, gate, dgate)
    th = torch.tanh(gate.double())
    ref = (th * G.double())
    ref_dg = 0.25 + ((1 - th * th) * (G.double() * upd.double()).sum()).item()
    assert (out.double() - ref).abs().max().item() <= 1e-2 * ref.abs().max().item()  # bf16 output rounding
    assert abs(dgate.item() - ref_dg) <= 1e-3 * max(1.0, abs(ref_dg))
    assert torch.equal(ops.gate_bwd(G, upd, gate, None), out)


@pytest.mark.parametrize("rows,T,d", [(600, 200, 384), (96, 48, 128)])
def test_fddt_full_scatter(rows, T, d):
    """backward of the mask-weighted sum of full-matrix FDDT (FDDT.py:52-62): dY[:, c d:(c+1) d] = mask_c * G, bit-exact"""
    from ts_asr_whisper_b200 import ops
    gen = torch.Generator(device="cpu").manual_seed(3)
    G = torch.randn(rows, d, generator=gen).cuda()
    stno = torch.rand(rows // T, 4, T, generator=gen).cuda()
    dY = torch.empty(rows, 4 * d, dtype=torch.bfloat16, device="cuda")
    ops.fddt_full_scatter(G, stno, dY, T=T)
    m = stno.permute(0, 2, 1).reshape(rows, 4)                       # [rows, class]
    ref = (m[:, :, None] * G[:, None, :]).to(torch.bfloat16).reshape(rows, 4 * d)
    assert torch.equal(dY, ref)


@pytest.mark.parametrize("rows,cols,cols_out", [(37, 51867, 51872), (5, 7, 8), (300, 1280, 1280), (3, 33, 40)])
def test_cast_bf16_padded(ops, rows, cols, cols_out):
    """fp32 [rows, cols] (odd row lengths: unaligned rows) -> bf16 [rows, cols_out] with zeroed padding columns"""
    g = torch.Generator(device=DEV).manual_seed(rows)
    buf = torch.randn(rows, cols + 3, device=DEV, generator=g)
    src = buf[:, :cols]  # row stride cols + 3
    out = ops.cast_bf16_padded(src, cols_out)
    torch.cuda.synchronize()
    assert out.shape == (rows, cols_out) and out.dtype == torch.bfloat16
    assert torch.equal(out[:, :cols], src.bfloat16())
    assert not out[:, cols:].float().abs().any()
