"""Parity of the CUDA SENSE path (through SenseModel -> ctypes -> C ABI -> sm_100a kernels) on a B200.

Bar (BASELINE.json north_star): forward / adjoint within 1e-5 relative L2 of the reference in fp32.
Checked three ways: against the committed outputs of the unmodified reference (tests/golden), against
the numpy oracle on seeded inputs over the radix and direct-DFT kernels, and at BASELINE.json's full
cine size through size-independent properties (adjointness, fused == composed, linearity, delta -> plane wave).
"""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import sense_oracle as so
from oracle.golden_inputs import CASES

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import __graft_entry__ as ge
    ge.build()
    return torch.device("cuda:0")


def make_model(maps, w, dev):
    from dl_swin_gan_b200.mri import SenseModel
    tw = None if w is None else torch.from_numpy(w).to(dev)
    return SenseModel(torch.from_numpy(maps).to(dev), weights=tw)


def to_np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(params=["sampled-lines", "dense"])
def kernel_path(request, monkeypatch):
    """Masks with empty ky lines take the sampled-line kernels by default; every parity case also runs on the dense kernels."""
    monkeypatch.setenv("DLSG_SAMPLED_LINES", "force" if request.param == "sampled-lines" else "0")   # force: also for E > 1
    return request.param


@pytest.mark.parametrize("name", list(CASES))
def test_matches_reference_golden_vectors(golden_cases, dev, name, kernel_path):
    inp, ref = golden_cases[name]
    A = make_model(inp["maps"], inp["weights"], dev)
    assert A.fft.ndims == int(ref["ndims"])
    x, k, aty = (torch.from_numpy(inp[n]).to(dev) for n in ("x", "k", "aty"))
    step = float(ref["step"])
    got = {
        "fwd": A(x), "adj": A(k, adjoint=True), "normal": A.normal(x), "dc": A.dc_step(x, aty, step),
        "composed": A(A(x), adjoint=True),
    }
    for key in ("fwd", "adj", "normal", "dc"):
        assert rel_l2(to_np(got[key]), ref[f"{key}_c64"]) < TOL, (name, key)
    assert rel_l2(to_np(got["composed"]), ref["normal_c64"]) < TOL


SWEEP = [
    # B, E, C, T, Y, X, weights
    (2, 1, 8, 5, 192, 160, "mask"),      # cine radices (6, 5), E = 1 as config_recon.yaml
    (1, 2, 8, 20, 192, 160, "mask"),     # one full cine slice, E = 2 (config 1 / 3)
    (1, 2, 4, 3, 128, 128, "none"),
    (1, 1, 3, 2, 160, 160, "full"),
    (1, 2, 2, 2, 256, 256, "mask"),      # perfusion matrix
    (1, 1, 2, 2, 320, 320, "mask"),
    (1, 2, 3, 2, 96, 64, "full"),
    (1, 3, 2, 2, 64, 32, "mask"),        # E = 3: two accumulator groups
    (1, 2, 3, 2, 100, 90, "mask"),       # direct-DFT kernels (axes not 32*R)
    (1, 1, 2, 2, 191, 160, "full"),      # direct columns + radix rows
    (1, 1, 2, 2, 192, 150, "full"),      # radix-ineligible rows -> direct rows and direct columns
    (3, 1, 2, 1, 33, 31, "none"),
]


@pytest.mark.parametrize("case", SWEEP)
def test_matches_oracle_on_seeded_inputs(dev, case, kernel_path):
    B, E, C, T, Y, X, wkind = case
    rng = np.random.default_rng(11)
    maps, x, mask = so.synthetic_problem(B, E, C, T, Y, X, seed=1000)
    w = {"none": None, "mask": mask, "full": rng.standard_normal((B, C, T, Y, X)).astype(np.float32)}[wkind]
    k = (rng.standard_normal((B, C, T, Y, X)) + 1j * rng.standard_normal((B, C, T, Y, X))).astype(np.complex64)
    aty = (rng.standard_normal((B, E, T, Y, X)) + 1j * rng.standard_normal((B, E, T, Y, X))).astype(np.complex64)
    A = make_model(maps, w, dev)
    xd, kd, ad = (torch.from_numpy(a).to(dev) for a in (x, k, aty))
    m128 = maps.astype(np.complex128)
    assert rel_l2(to_np(A(xd)), so.sense_forward(x.astype(np.complex128), m128, w)) < TOL
    assert rel_l2(to_np(A(kd, adjoint=True)), so.sense_adjoint(k.astype(np.complex128), m128, w)) < TOL
    assert rel_l2(to_np(A.normal(xd)), so.sense_normal(x.astype(np.complex128), m128, w)) < TOL
    ref_dc = so.dc_step(x.astype(np.complex128), aty.astype(np.complex128), m128, w, -2.0)
    assert rel_l2(to_np(A.dc_step(xd, ad, -2.0)), ref_dc) < TOL


def test_sampled_line_path_is_selected_and_agrees_with_dense(dev, monkeypatch):
    """12x k-t mask at the cine size: the plan is built, the sampled-line kernels run, and every operator agrees with the
    dense kernels to fp32 rounding; per-coil real weights, a frame without samples, and a dense mask (falls back)."""
    from dl_swin_gan_b200.mri import SenseModel
    B, E, C, T, Y, X = 3, 1, 8, 6, 192, 160
    g = torch.Generator(device=dev).manual_seed(5)

    def cn(*s):
        return torch.view_as_complex(torch.randn(*s, 2, generator=g, device=dev))

    maps = cn(B, E, C, 1, Y, X) / C ** 0.5
    x, k, aty = cn(B, E, T, Y, X), cn(B, C, T, Y, X), cn(B, E, T, Y, X)
    lines = (torch.rand(B, 1, T, Y, 1, generator=g, device=dev) < 1 / 12).float()
    lines[1, :, 2] = 0                                                    # a frame with no sampled line at all
    mask = lines.expand(B, 1, T, Y, X).contiguous()
    mask[..., : X // 4] = 0
    percoil = mask * torch.randn(B, C, T, Y, X, generator=g, device=dev)  # real-valued weights that differ per coil
    for w in (mask, percoil):
        outs = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("DLSG_SAMPLED_LINES", mode)
            A = SenseModel(maps, weights=w)
            outs[mode] = [A(x), A(k, adjoint=True), A.normal(x), A.dc_step(x, aty, -2.0), A.normal_reg(x, 0.3)]
            plan = A._line_plans.get((B, E, C, T, Y, X))
            if mode == "1":
                counts = (w != 0).any(dim=1).any(dim=-1).sum(-1)          # [B, T]
                assert plan is not None and plan[1] == int(counts.max())
                assert torch.equal(plan[0][: B * T].cpu(), counts.flatten().to(torch.int32).cpu())
            else:
                assert plan is None
        for a, b in zip(outs["1"], outs["0"]):
            assert rel_l2(to_np(a), to_np(b)) < 2e-6
        monkeypatch.setenv("DLSG_SAMPLED_LINES", "1")
        monkeypatch.setenv("DLSG_WORKSPACE_MB", "0.001")                   # one slice per workspace chunk: bit-identical results
        A = SenseModel(maps, weights=w)
        for a, b in zip([A(x), A(k, adjoint=True), A.dc_step(x, aty, -2.0)], [outs["1"][0], outs["1"][1], outs["1"][3]]):
            assert torch.equal(a, b)
        monkeypatch.delenv("DLSG_WORKSPACE_MB")
        assert not outs["1"][0][(w == 0).expand(B, C, T, Y, X)].any()     # forward: unsampled samples are exact zeros
    # 35 % of the lines sampled: still the sampled-line path (larger line buffer, one column CTA per SM)
    w40 = (torch.rand(B, 1, T, Y, 1, generator=g, device=dev) < 0.35).float().expand(B, 1, T, Y, X).contiguous()
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("DLSG_SAMPLED_LINES", mode)
        A = SenseModel(maps, weights=w40)
        res[mode] = [A(x), A(k, adjoint=True), A.dc_step(x, aty, -2.0)]
        assert (A._line_plans.get((B, E, C, T, Y, X)) is not None) == (mode == "1")
    for a, b in zip(res["1"], res["0"]):
        assert rel_l2(to_np(a), to_np(b)) < 2e-6
    # weights edited in place after first use: the plan follows (the tensor's version counter is part of the cache key)
    monkeypatch.setenv("DLSG_SAMPLED_LINES", "1")
    wmut = mask.clone()
    A = SenseModel(maps, weights=wmut)
    before = A(x)
    wmut[:, :, :, 7, X // 2:] = 1.0                                        # a line that was (almost surely) empty becomes sampled
    after = A(x)
    monkeypatch.setenv("DLSG_SAMPLED_LINES", "0")
    assert rel_l2(to_np(after), to_np(SenseModel(maps, weights=wmut)(x))) < 2e-6
    assert after[:, :, :, 7, X // 2:].abs().sum() > 0 and not torch.equal(before, after)
    monkeypatch.setenv("DLSG_SAMPLED_LINES", "1")
    dense = (torch.rand(B, 1, T, Y, 1, generator=g, device=dev) < 0.7).float().expand(B, 1, T, Y, X).contiguous()
    A = SenseModel(maps, weights=dense)
    A(x)
    assert A._line_plans[(B, E, C, T, Y, X)] is None                      # more than half of the lines sampled: dense kernels


@pytest.mark.parametrize("E", [1, 2])
def test_column_kernel_generations_agree(dev, monkeypatch, E):
    """Second-generation column kernels (packed butterflies, pruned radix stage, TMA-staged maps; E <= 2) against the
    first generation (DLSG_LINES_KERNELS=v1) and the dense kernels, reference k-t mask (VDktMaskFunc, 16 lines of 192),
    perfusion-like 256 x 256 included (R2 = 16)."""
    from dl_swin_gan_b200.mri import SenseModel
    from dl_swin_gan_b200.mri.subsample import VDktMaskFunc
    for (B, C, T, Y, X, acc) in [(2, 8, 6, 192, 160, 12), (1, 5, 3, 256, 256, 8), (1, 3, 2, 160, 128, 10), (1, 2, 2, 320, 320, 12), (1, 2, 2, 128, 128, 4)]:
        g = torch.Generator(device=dev).manual_seed(31 + Y)

        def cn(*s):
            return torch.view_as_complex(torch.randn(*s, 2, generator=g, device=dev))

        maps = cn(B, E, C, 1, Y, X) / (E * C) ** 0.5
        x, k, aty = cn(B, E, T, Y, X), cn(B, C, T, Y, X), cn(B, E, T, Y, X)
        f = VDktMaskFunc((acc, acc), sim_partial_kx=0.25, sim_partial_ky=0.25)
        mask = torch.cat([f((1, 1, T, Y, X), seed=3 + b) for b in range(B)], 0).to(dev)
        outs = {}
        for mode in ("gen2", "v1", "dense"):
            monkeypatch.setenv("DLSG_SAMPLED_LINES", "0" if mode == "dense" else "force")
            monkeypatch.setenv("DLSG_LINES_KERNELS", "v1" if mode == "v1" else "")
            A = SenseModel(maps, weights=mask)
            outs[mode] = [A(x), A(k, adjoint=True), A.normal(x), A.dc_step(x, aty, -2.0), A.normal_reg(x, 0.3)]
            assert (A._line_plans.get((B, E, C, T, Y, X)) is not None) == (mode != "dense")
        for a, b, c in zip(outs["gen2"], outs["v1"], outs["dense"]):
            assert rel_l2(to_np(a), to_np(c)) < 2e-6 and rel_l2(to_np(b), to_np(c)) < 2e-6
    monkeypatch.delenv("DLSG_LINES_KERNELS")


def test_perfusion_shape_at_full_length_matches_fp64(dev):
    """BASELINE configs[3]: 2 map sets, 32 coils, 40 frames, 256 x 256 (one slice) through the sampled-line kernels (R2 = 16,
    E = 2) against the reference expression evaluated in complex128 with torch.fft on the same device."""
    from dl_swin_gan_b200.mri import SenseModel
    from dl_swin_gan_b200.mri.subsample import VDktMaskFunc
    B, E, C, T, Y, X = 1, 2, 32, 40, 256, 256
    g = torch.Generator(device=dev).manual_seed(77)

    def cn(*s):
        return torch.view_as_complex(torch.randn(*s, 2, generator=g, device=dev))

    maps = cn(B, E, C, 1, Y, X) / (E * C) ** 0.5
    x, aty = cn(B, E, T, Y, X), cn(B, E, T, Y, X)
    mask = VDktMaskFunc((8, 8), sim_partial_kx=0.25, sim_partial_ky=0.25)((1, 1, T, Y, X), seed=9).to(dev)
    A = SenseModel(maps, weights=mask)
    k = A(x)
    assert A._line_plans.get((B, E, C, T, Y, X)) is not None
    m128, w = maps.to(torch.complex128), mask.to(torch.float64)

    def fwd(v):
        return w * torch.fft.fftn((v.unsqueeze(2) * m128).sum(1), dim=(-2, -1), norm="ortho")

    def adj(v):
        return (torch.fft.ifftn(w * v, dim=(-2, -1), norm="ortho").unsqueeze(1) * m128.conj()).sum(2)

    def rel(a, b):
        return float((a.to(torch.complex128) - b).norm() / b.norm())

    x128 = x.to(torch.complex128)
    k_ref = fwd(x128)
    assert rel(k, k_ref) < TOL
    assert rel(A(k, adjoint=True), adj(k.to(torch.complex128))) < TOL
    assert rel(A.dc_step(x, aty, -2.0), x128 - 2.0 * (adj(k_ref) - aty.to(torch.complex128))) < TOL


def test_adjoint_reads_pinned_host_kspace_in_place(dev, monkeypatch):
    """A(y, adjoint=True) with y in pinned host memory == the same with y copied to the device (bit for bit);
    pageable host memory, dense weights and the dense kernels refuse instead of falling back to a CPU path."""
    from dl_swin_gan_b200.mri import SenseModel
    B, E, C, T, Y, X = 2, 1, 4, 5, 192, 160
    maps, x, mask = so.synthetic_problem(B, E, C, T, Y, X, seed=21)
    rng = np.random.default_rng(2)
    k = (rng.standard_normal((B, C, T, Y, X)) + 1j * rng.standard_normal((B, C, T, Y, X))).astype(np.complex64)
    A = make_model(maps, mask, dev)
    hk = torch.from_numpy(k).pin_memory()
    got = A(hk, adjoint=True)
    assert got.is_cuda
    assert torch.equal(got, A(hk.to(dev), adjoint=True))
    assert rel_l2(to_np(got), so.sense_adjoint(k.astype(np.complex128), maps.astype(np.complex128), mask)) < TOL
    maps2, _, _ = so.synthetic_problem(B, 2, C, T, Y, X, seed=22)                  # two map sets through the same entry
    got2 = make_model(maps2, mask, dev)(hk, adjoint=True)
    assert rel_l2(to_np(got2), so.sense_adjoint(k.astype(np.complex128), maps2.astype(np.complex128), mask)) < TOL
    with pytest.raises(RuntimeError):
        A(torch.from_numpy(k), adjoint=True)                   # pageable memory: no CPU fallback
    with pytest.raises(ValueError):
        A(torch.from_numpy(np.ascontiguousarray(k.transpose(0, 1, 2, 4, 3))).pin_memory().transpose(3, 4), adjoint=True)   # strided view of a pinned tensor: .contiguous() would be pageable
    # the kernel reads the pinned buffer after the call returns: a temporary must stay alive until then
    from dl_swin_gan_b200.mri import transforms as tr
    tmp = A(torch.from_numpy(k).pin_memory(), adjoint=True)    # the pinned temporary is dropped by the caller right here
    scratch = [torch.from_numpy(k * 0 + 7).pin_memory() for _ in range(3)]      # would recycle a freed pinned block
    assert torch.equal(tmp, got) and len(tr._HOST_INFLIGHT) >= 1
    del scratch
    with pytest.raises(RuntimeError):
        make_model(maps, np.ones_like(mask), dev)(hk, adjoint=True)      # every line sampled: dense kernels need a device copy
    monkeypatch.setenv("DLSG_SAMPLED_LINES", "0")
    with pytest.raises(RuntimeError):
        make_model(maps, mask, dev)(hk, adjoint=True)


def test_repeated_launches_are_bit_identical(dev, monkeypatch):
    """No atomics and no order-dependent reductions in this path: 12 repeats of every operator give the same bits on both
    kernel families (a shared-memory race or a missing barrier shows up here as run-to-run differences)."""
    from dl_swin_gan_b200.mri import SenseModel
    B, E, C, T, Y, X = 6, 1, 8, 20, 192, 160
    g = torch.Generator(device=dev).manual_seed(9)

    def cn(*s):
        return torch.view_as_complex(torch.randn(*s, 2, generator=g, device=dev))

    maps, x, k, aty = cn(B, E, C, 1, Y, X), cn(B, E, T, Y, X), cn(B, C, T, Y, X), cn(B, E, T, Y, X)
    mask = (torch.rand(B, 1, T, Y, 1, generator=g, device=dev) < 1 / 8).float().expand(B, 1, T, Y, X).contiguous()
    for mode in ("1", "0"):
        monkeypatch.setenv("DLSG_SAMPLED_LINES", mode)
        A = SenseModel(maps, weights=mask)
        first = [A(x), A(k, adjoint=True), A.normal(x), A.dc_step(x, aty, -2.0)]
        for _ in range(12):
            again = [A(x), A(k, adjoint=True), A.normal(x), A.dc_step(x, aty, -2.0)]
            for a, b in zip(first, again):
                assert torch.equal(a, b), mode


def test_workspace_chunking_does_not_change_results(dev, monkeypatch):
    B, E, C, T, Y, X = 3, 2, 4, 5, 64, 96
    maps, x, mask = so.synthetic_problem(B, E, C, T, Y, X, seed=3)
    A = make_model(maps, mask, dev)
    xd = torch.from_numpy(x).to(dev)
    monkeypatch.setenv("DLSG_WORKSPACE_MB", "4096")
    full = [to_np(A(xd)), to_np(A.normal(xd))]
    for mb in ("0.0001", "0.5", "1.5"):          # one (b,t) set per chunk ... a few frames ... > one slice
        monkeypatch.setenv("DLSG_WORKSPACE_MB", mb)
        assert np.array_equal(to_np(A(xd)), full[0])
        assert np.array_equal(to_np(A.normal(xd)), full[1])


def test_full_size_properties_cine_batch(dev):
    """config_recon.yaml shape, 8 slices of the 64-slice batch: properties that need no CPU oracle run."""
    B, E, C, T, Y, X = 8, 1, 8, 20, 192, 160
    g = torch.Generator(device=dev).manual_seed(1000)

    def cn(*s):
        return torch.view_as_complex(torch.randn(*s, 2, generator=g, device=dev))

    maps = cn(B, E, C, 1, Y, X) / (E * C) ** 0.5
    mask = (torch.rand(B, 1, T, Y, 1, generator=g, device=dev) < 1 / 12).float().expand(B, 1, T, Y, X).contiguous()
    from dl_swin_gan_b200.mri import SenseModel
    A = SenseModel(maps, weights=mask)
    x, y = cn(B, E, T, Y, X), cn(B, C, T, Y, X)
    Ax, Ahy = A(x), A(y, adjoint=True)
    lhs, rhs = torch.vdot(y.flatten(), Ax.flatten()), torch.vdot(Ahy.flatten(), x.flatten())
    assert abs(lhs - rhs) / abs(lhs) < 1e-5                              # <y, Ax> == <A^H y, x>
    assert rel_l2(to_np(A.normal(x)), to_np(A(Ax, adjoint=True))) < 2e-6   # fused == composed
    x2 = cn(B, E, T, Y, X)
    lin = A(2.0 * x - 0.5j * x2)
    assert rel_l2(to_np(lin), to_np(2.0 * Ax - 0.5j * A(x2))) < 2e-6      # linearity
    aty = cn(B, E, T, Y, X)
    dc = A.dc_step(x, aty, -2.0)
    assert rel_l2(to_np(dc), to_np(x - 2.0 * (A(Ax, adjoint=True) - aty))) < 2e-6
    # unmasked operator with unit maps: a delta at (y0, x0) becomes a plane wave of modulus 1/sqrt(YX)
    ones = torch.ones(1, 1, 1, 1, Y, X, dtype=torch.complex64, device=dev)
    F = SenseModel(ones)
    d = torch.zeros(1, 1, 1, Y, X, dtype=torch.complex64, device=dev)
    d[0, 0, 0, 5, 7] = 1
    kd = F(d)[0, 0, 0]
    ky, kx = torch.meshgrid(torch.arange(Y, device=dev), torch.arange(X, device=dev), indexing="ij")
    wave = torch.exp(-2j * torch.pi * (5 * ky / Y + 7 * kx / X)) / (Y * X) ** 0.5
    assert rel_l2(to_np(kd), to_np(wave)) < TOL   # the fp32 torch plane wave itself carries ~1e-6
    # parseval: ortho transform with unit maps preserves energy
    z = cn(1, 1, 1, Y, X)
    assert abs(torch.linalg.norm(F(z)) / torch.linalg.norm(z) - 1) < 1e-5


@pytest.mark.parametrize("E", [1, 2])          # E = 1 with a mask: the sampled-line kernels; E = 2: the dense ones
def test_autograd_vjp_is_the_adjoint(dev, E):
    B, C, T, Y, X = 1, 3, 2, 64, 32
    maps, x, mask = so.synthetic_problem(B, E, C, T, Y, X, seed=5)
    A = make_model(maps, mask, dev)
    xd = torch.from_numpy(x).to(dev).requires_grad_()
    g = torch.from_numpy(so.synthetic_problem(B, C, 1, T, Y, X, seed=6)[1]).to(dev)   # [B,C,T,Y,X]
    (gx,) = torch.autograd.grad(A(xd), xd, grad_outputs=g)
    assert rel_l2(to_np(gx), to_np(A(g, adjoint=True))) < 1e-6
    kd = g.clone().requires_grad_()
    gi = torch.from_numpy(x).to(dev)
    (gk,) = torch.autograd.grad(A(kd, adjoint=True), kd, grad_outputs=gi)
    assert rel_l2(to_np(gk), to_np(A(gi))) < 1e-6
    # fused DC step: gradients w.r.t. x, ATy and a learnable CUDA step (FIX_STEP_SIZE False)
    step = torch.tensor([-2.0], device=dev, requires_grad=True)
    aty = torch.from_numpy(so.synthetic_problem(B, E, 1, T, Y, X, seed=8)[1]).to(dev).requires_grad_()
    x1 = torch.from_numpy(x).to(dev).requires_grad_()
    out = A.dc_step(x1, aty, step)
    x2 = torch.from_numpy(x).to(dev).requires_grad_()
    aty2 = aty.detach().clone().requires_grad_()
    step2 = step.detach().clone().requires_grad_()
    ref = x2 + step2 * (A(A(x2), adjoint=True) - aty2)              # the reference's expression, our operator
    go = torch.randn_like(out)
    got = torch.autograd.grad(out, (x1, aty, step), grad_outputs=go)
    want = torch.autograd.grad(ref, (x2, aty2, step2), grad_outputs=go)
    for a, b in zip(got, want):
        assert rel_l2(to_np(a), to_np(b)) < 1e-5
    assert (A._line_plans.get((B, E, C, T, Y, X)) is not None) == (E <= 2)


def test_fft_module_matches_reference_vectors(dev):
    import os
    from conftest import GOLDEN
    from dl_swin_gan_b200.mri import FFT
    ref = np.load(os.path.join(GOLDEN, "sense_misc.npz"))
    img = torch.from_numpy(ref["img"]).to(dev)
    f2, f1 = FFT(2), FFT(1)
    assert rel_l2(to_np(f2(img)), ref["fft2"]) < TOL
    assert rel_l2(to_np(f2(img, adjoint=True)), ref["ifft2"]) < TOL
    assert rel_l2(to_np(f2(img, centered=True)), ref["fft2_centered"]) < TOL
    assert rel_l2(to_np(f2(img, adjoint=True, centered=True)), ref["ifft2_centered"]) < TOL
    assert rel_l2(to_np(f1(img)), ref["fft1"]) < TOL
    big = torch.view_as_complex(torch.randn(3, 2, 192, 160, 2, device=dev))
    assert rel_l2(to_np(f2(big)), to_np(torch.fft.fftn(big, dim=(-2, -1), norm="ortho"))) < TOL


def test_error_behaviour(dev):
    from dl_swin_gan_b200.mri import SenseModel
    maps = torch.view_as_complex(torch.randn(1, 2, 3, 1, 16, 16, 2, device=dev))
    A = SenseModel(maps)
    with pytest.raises(AssertionError):
        A(torch.randn(1, 2, 2, 16, 16, device=dev))                                 # not complex (transforms.py:104)
    with pytest.raises(ValueError):
        A(torch.view_as_complex(torch.randn(1, 3, 2, 16, 16, 2, device=dev)))       # wrong emaps count
    with pytest.raises(TypeError):
        A(torch.view_as_complex(torch.randn(1, 2, 2, 16, 16, 2, device=dev, dtype=torch.float64)))
    with pytest.raises(AssertionError):
        SenseModel(torch.randn(1, 2, 3, 1, 16, 16, device=dev))                    # transforms.py:70
    empty = torch.zeros(1, 2, 0, 16, 16, dtype=torch.complex64, device=dev)
    assert A(empty).shape == (1, 3, 0, 16, 16)


def test_unrolled_dc_path_matches_cpu_port(dev):
    """A^H y + 5 data-consistency updates (unrolled.py:96-117 without the CNN) through the PGD mirror:
    NRMSE against the torch-CPU port of the reference path (north_star: full reconstructions within 1e-3 NRMSE)."""
    from dl_swin_gan_b200.mri import SenseModel
    from dl_swin_gan_b200.unrolled import ProximalGradientDescent
    from oracle import sense_oracle_torch as st
    B, E, C, T, Y, X = 2, 2, 8, 6, 192, 160
    maps, x, mask = so.synthetic_problem(B, E, C, T, Y, X, seed=1000)
    tm, tx, tw = torch.from_numpy(maps), torch.from_numpy(x), torch.from_numpy(mask)
    y = st.forward(tx, tm, tw)
    want = st.dc_path(y, tm, tw, 5, -0.5)
    pgd = ProximalGradientDescent(num_unrolls=5, step_init=-0.5).to(dev)
    A = SenseModel(tm.to(dev), weights=tw.to(dev))
    with torch.no_grad():
        got = pgd(y.to(dev), A)
        # the same loop written the reference's way (two operator calls per update) must agree with the fused entry
        ATy = A(y.to(dev), adjoint=True)
        xi = ATy
        for _ in range(5):
            xi = xi + pgd.step_size * (A(A(xi), adjoint=True) - ATy)
    nrmse = rel_l2(got.cpu().numpy(), want.numpy())
    assert nrmse < 1e-5, nrmse
    assert rel_l2(xi.cpu().numpy(), want.numpy()) < 1e-5


def test_perfusion_shape_properties(dev):
    """BASELINE config 4 shape (E=2, C=32, T=40, 256x256; 6 frames here): several coil groups per CTA, radix (8, 8)."""
    B, E, C, T, Y, X = 1, 2, 32, 6, 256, 256
    g = torch.Generator(device=dev).manual_seed(1004)

    def cn(*s):
        return torch.view_as_complex(torch.randn(*s, 2, generator=g, device=dev))

    from dl_swin_gan_b200.mri import SenseModel
    maps = cn(B, E, C, 1, Y, X) / (E * C) ** 0.5
    mask = (torch.rand(B, 1, T, Y, 1, generator=g, device=dev) < 0.15).float().expand(B, 1, T, Y, X).contiguous()
    A = SenseModel(maps, weights=mask)
    x, y = cn(B, E, T, Y, X), cn(B, C, T, Y, X)
    Ax, Ahy = A(x), A(y, adjoint=True)
    lhs, rhs = torch.vdot(y.flatten(), Ax.flatten()), torch.vdot(Ahy.flatten(), x.flatten())
    assert abs(lhs - rhs) / abs(lhs) < 1e-5
    assert rel_l2(to_np(A.normal(x)), to_np(A(Ax, adjoint=True))) < 2e-6
    # against torch.fft on the same device (the reference's own arithmetic, transforms.py:84-98)
    from oracle import sense_oracle_torch as st
    assert rel_l2(to_np(Ax), to_np(st.forward(x, maps, mask))) < TOL
    assert rel_l2(to_np(Ahy), to_np(st.adjoint(y, maps, mask))) < TOL
    aty = cn(B, E, T, Y, X)
    want = x - 0.7 * (st.adjoint(st.forward(x, maps, mask), maps, mask) - aty)
    assert rel_l2(to_np(A.dc_step(x, aty, -0.7)), to_np(want)) < TOL


def test_side_stream_and_cuda_graph_replay(dev):
    """The C ABI enqueues on the caller's stream only: results on a side stream match, and a DC step captured into a
    CUDA graph (step size read from device memory) replays correctly with new inputs."""
    B, E, C, T, Y, X = 2, 1, 4, 3, 64, 96
    maps, x, mask = so.synthetic_problem(B, E, C, T, Y, X, seed=21)
    A = make_model(maps, mask, dev)
    xd = torch.from_numpy(x).to(dev)
    want = A(xd)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        got = A(xd)
    side.synchronize()
    assert torch.equal(got, want)

    aty = A(want, adjoint=True)
    step = torch.tensor([-1.5], device=dev)
    static_x = xd.clone()
    A.dc_step(static_x, aty, step)                      # warm-up outside capture (lazy preparation, allocator)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_out = A.dc_step(static_x, aty, step)
    for scale, s in ((1.0, -1.5), (0.5, -0.25)):
        static_x.copy_(xd * scale)
        step.fill_(s)
        graph.replay()
        torch.cuda.synchronize()
        ref = static_x + s * (A(A(static_x), adjoint=True) - aty)
        assert rel_l2(to_np(static_out), to_np(ref)) < 2e-6


@pytest.mark.parametrize("name", ["tiny_odd", "pow2_32"])
def test_conjugate_gradient_and_hqs_match_reference(golden_cases, dev, name):
    """MoDL / HQS path: the reference's ConjugateGradient on A^H A + lamda I (algorithms.py:11-73,
    unrolledswin.py:139-172), here on the fused normal_reg entry; vectors from the unmodified reference."""
    import os
    from conftest import GOLDEN
    from dl_swin_gan_b200.mri import ConjugateGradient
    from dl_swin_gan_b200.unrolled import HalfQuadraticSplitting
    ref = np.load(os.path.join(GOLDEN, "sense_misc.npz"))
    inp, _ = golden_cases[name]
    A = make_model(inp["maps"], inp["weights"], dev)
    x, k, aty = (torch.from_numpy(inp[n]).to(dev) for n in ("x", "k", "aty"))
    lam = torch.tensor([0.1], device=dev)
    with torch.no_grad():
        ATy = A(k, adjoint=True)
        assert rel_l2(to_np(A.normal_reg(x, lam)), to_np(A(A(x), adjoint=True) + lam * x)) < 2e-6
        got = ConjugateGradient(lambda m: A.normal_reg(m, lam), 6)(x, ATy + lam * aty)
        assert rel_l2(to_np(got), ref[f"cg6_{name}"]) < 1e-4
        hqs = HalfQuadraticSplitting(num_unrolls=2, num_cg_iter=4, lamda_init=0.1).to(dev)
        assert rel_l2(to_np(hqs(k, A)), ref[f"hqs2x4_{name}"]) < 1e-4
    # gradient of the regularised normal operator w.r.t. a learnable lamda and x
    lam_g = torch.tensor([0.1], device=dev, requires_grad=True)
    xg = x.clone().requires_grad_()
    out = A.normal_reg(xg, lam_g)
    go = torch.randn_like(out)
    gx, gl = torch.autograd.grad(out, (xg, lam_g), go)
    x2 = x.clone().requires_grad_()
    l2 = lam_g.detach().clone().requires_grad_()
    gx2, gl2 = torch.autograd.grad(A(A(x2), adjoint=True) + l2 * x2, (x2, l2), go)
    assert rel_l2(to_np(gx), to_np(gx2)) < 1e-5 and rel_l2(to_np(gl), to_np(gl2)) < 1e-4


def test_kt_preprocessing_matches_reference(dev):
    """fftmod / get_mask / time_average / sliding_window kernels and the DataTransform chain (reconstruct.py:132-171)
    against outputs of the reference functions (tests/golden/sense_misc.npz)."""
    import os
    from conftest import GOLDEN
    from dl_swin_gan_b200.mri import utils as U
    from dl_swin_gan_b200.data import DataTransform
    ref = np.load(os.path.join(GOLDEN, "sense_misc.npz"))
    assert np.array_equal(to_np(U.fftmod(torch.from_numpy(ref["img_even"].copy()).to(dev))), ref["fftmod"])
    series = torch.from_numpy(ref["series"]).to(dev)
    assert np.array_equal(to_np(U.get_mask(series)), ref["get_mask"])
    assert rel_l2(to_np(U.sliding_window(series, 2, 5)), ref["sliding_window5"]) < 1e-6
    for slwin, tag in ((False, "plain"), (True, "slwin")):
        k, m, mask, init, scale = DataTransform(slwin_init=slwin, device=dev)(ref["dt_kspace_in"], ref["dt_maps_in"])
        assert np.array_equal(to_np(mask), ref[f"dt_{tag}_mask"])
        assert np.array_equal(to_np(m), ref[f"dt_{tag}_maps"])
        assert abs(float(scale) - float(ref[f"dt_{tag}_scale"])) / float(ref[f"dt_{tag}_scale"]) < 1e-5
        assert rel_l2(to_np(k), ref[f"dt_{tag}_kspace"]) < 1e-5
        assert rel_l2(to_np(init), ref[f"dt_{tag}_init"]) < 1e-5
    tav = U.time_average(U.fftmod(torch.from_numpy(ref["dt_kspace_in"].copy()).to(dev).unsqueeze(0)), dim=2)
    assert rel_l2(to_np(tav), ref["time_average"]) < 1e-6
    big = torch.view_as_complex(torch.randn(2, 8, 20, 192, 160, 2, device=dev))
    big[:, :, ::3] = 0
    want = torch.cat([torch.roll(big, 2 - i, 2).narrow(2, 0, 5).sum(2, keepdim=True) /
                      ((torch.roll(big, 2 - i, 2).narrow(2, 0, 5).abs() > 1e-12).float().sum(2, keepdim=True) + 1e-6)
                      for i in range(20)], dim=2)
    assert rel_l2(to_np(U.sliding_window(big, 2, 5)), to_np(want)) < 1e-6


def test_max_lines_promise_skips_the_read_back_and_flags_a_broken_promise():
    """SenseModel(max_lines=...) builds the sampled-line plan without reading counts back: same results when the promise holds,
    dropped lines + a flag (no out-of-bounds write) when it does not."""
    from dl_swin_gan_b200.mri import SenseModel
    from dl_swin_gan_b200.mri.subsample import VDktMaskFunc
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    B, E, C, T, Y, X = 2, 1, 4, 3, 192, 160
    maps = torch.view_as_complex(torch.randn(B, E, C, 1, Y, X, 2, generator=g, device=dev))
    x = torch.view_as_complex(torch.randn(B, E, T, Y, X, 2, generator=g, device=dev))
    mask = torch.cat([VDktMaskFunc((12, 12), 0.25, 0.25)((1, 1, T, Y, X), seed=7 + b) for b in range(B)], 0).to(dev)
    n_lines = int((mask.sum(-1) > 0).sum(-1).max())
    ref = SenseModel(maps, weights=mask)
    want = ref.dc_step(x, ref(ref(x), adjoint=True), -2.0)
    A = SenseModel(maps, weights=mask, max_lines=n_lines + 3)
    got = A.dc_step(x, A(A(x), adjoint=True), -2.0)
    assert A._line_plans and all(p is not None and p[1] == n_lines + 3 for p in A._line_plans.values())
    assert not A.lines_overflowed()
    assert float((got - want).abs().max() / want.abs().max()) < 1e-6
    # broken promise: fewer lines than the mask has -> the extra lines count as unsampled, flagged, nothing crashes
    A2 = SenseModel(maps, weights=mask, max_lines=max(1, n_lines - 4))
    k2 = A2(x)
    torch.cuda.synchronize()
    assert A2.lines_overflowed() and torch.isfinite(torch.view_as_real(k2)).all()
    with pytest.raises(ValueError):
        SenseModel(maps, weights=mask, max_lines=0)
