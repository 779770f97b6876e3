"""-m gpu: parity of the training path (forward that keeps activations, backward kernels, fused AdamW + EMA) with
autograd through the CPU oracle (fp32) on the same seeded inputs and weights, and with the committed gradients of the
unmodified reference (tests/golden/dit_tiny_grads_p*.npz, written by oracle/make_golden.py).

Tolerance: gradients are products of bf16 tensor-core GEMMs with fp32 accumulation; per-parameter relative error
(Frobenius) <= 3e-2 against fp32 autograd, loss relative error <= 1e-2 (north_star's per-forward bound).
"""
import numpy as np
import pytest
import torch

from oracle import ldmae_oracle as O

pytestmark = pytest.mark.gpu
GRAD_TOL = 3e-2


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    prev = torch.is_grad_enabled()          # other test modules switch autograd off process-wide
    torch.set_grad_enabled(True)
    yield
    torch.set_grad_enabled(prev)


def _rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _oracle_grads(spec, sd, x1, t, x0, y):
    leaves = {k: v.clone().requires_grad_(v.is_floating_point() and k != "pos_embed" and not k.startswith("feat_rope"))
              for k, v in sd.items()}
    terms = O.training_losses(lambda xt, tt, y: O.dit_forward(leaves, spec, xt, tt, y), x1, t, x0, y=y)
    terms["loss"].mean().backward()
    return terms, {k: v.grad for k, v in leaves.items() if v.grad is not None}


def _our_grads(m, x1, t, x0, y):
    from ldmae_b200.transport import create_transport
    tr = create_transport("Linear", "velocity", None, None, None, use_cosine_loss=False, use_lognorm=True)
    tr.sample = lambda x1_, *a, **k: (t.to(x1_), x0.to(x1_), x1_)       # inject the random draws (SURVEY 8c pitfall 3)
    m.zero_grad(set_to_none=True)
    with torch.enable_grad():
        terms = tr.training_losses(m, x1, dict(y=y))
        terms["loss"].mean().backward()
    torch.cuda.synchronize()
    return terms, {k: p.grad for k, p in m.named_parameters() if p.grad is not None}


def _compare(ours, ref, tol=GRAD_TOL):
    assert set(ours) == set(ref), set(ours) ^ set(ref)
    bad = []
    worst = ("", 0.0)
    for k in sorted(ref):
        e = _rel(ours[k], ref[k])
        if e > worst[1]:
            worst = (k, e)
        if not e < tol:
            bad.append((k, e, float(ref[k].norm())))
    print(f"worst gradient rel err {worst[1]:.3e} ({worst[0]})")
    assert not bad, f"{len(bad)} gradients off: {bad[:12]}"


def _tiny(patch, seed, input_size=8, **flags):
    from ldmae_b200.models.lightningdit import LightningDiT
    spec = O.DiTSpec(depth=2, hidden_size=128, patch_size=patch, num_heads=2, input_size=input_size, in_channels=16, num_classes=10,
                     **flags)
    m = LightningDiT(input_size=input_size, patch_size=patch, in_channels=16, hidden_size=128, depth=2, num_heads=2, num_classes=10,
                     use_qknorm=spec.use_qknorm, use_swiglu=True, use_rope=spec.use_rope, use_rmsnorm=True, wo_shift=spec.wo_shift,
                     learn_sigma=spec.learn_sigma)
    sd = O.synth_dit_state(spec, seed)
    m.load_state_dict(sd, strict=True)
    return spec, sd, m.cuda().eval()        # eval: no label dropout, the draws are injected


@pytest.mark.parametrize("patch,flags", [(1, {}), (2, {}), (1, dict(use_qknorm=False)), (1, dict(wo_shift=True)),
                                         (1, dict(use_rope=False)), (1, dict(learn_sigma=True)),
                                         (1, dict(input_size=6)),      # T = 36: ragged against every tile size (rows, keys, slabs)
                                         (1, dict(input_size=14))])    # T = 196: several ragged blocks per sample
def test_tiny_backward_matches_oracle_autograd(patch, flags):
    flags = dict(flags)
    S = flags.pop("input_size", 8)
    spec, sd, m = _tiny(patch, 21 + patch, input_size=S, **flags)
    g = torch.Generator().manual_seed(5 + patch)
    B = 5
    x1 = torch.randn(B, 16, S, S, generator=g)
    x0 = torch.randn(B, 16, S, S, generator=g)
    t = torch.rand(B, generator=g)
    y = torch.randint(0, 10, (B,), generator=g)
    ref_terms, ref = _oracle_grads(spec, sd, x1, t, x0, y)
    terms, ours = _our_grads(m, x1.cuda(), t, x0, y.cuda())
    assert _rel(terms["loss"], ref_terms["loss"].detach()) < 1e-2
    assert _rel(terms["pred"], ref_terms["pred"].detach()) < 1e-2
    _compare(ours, ref)


@pytest.mark.parametrize("patch", [1, 2])
def test_tiny_backward_matches_reference_golden(golden_dir, patch):
    """Gradients of the UNMODIFIED reference (autograd through LDMAE/models/lightningdit.py on CPU, fp32)."""
    from gpu_util import load_npz
    g = load_npz(golden_dir, f"dit_tiny_grads_p{patch}.npz")
    spec, sd, m = _tiny(patch, int(g["seed"]))
    x1, t, x0, y = (torch.from_numpy(g[k]) for k in ("x1", "t", "x0", "y"))
    terms, ours = _our_grads(m, x1.cuda(), t, x0, y.cuda())
    assert _rel(terms["loss"], g["loss"]) < 1e-2
    ref = {k[len("grad."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("grad.")}
    _compare(ours, ref)


@pytest.mark.parametrize("hidden,heads,patch,inp", [(1024, 16, 2, 16), (1536, 24, 1, 8), (1792, 28, 2, 16)])
def test_wider_registry_widths_forward_and_backward(hidden, heads, patch, inp):
    """Widths of the other registry entries with head_dim 64 (L: 1024, 1p0B: 1536, 1p6B: 1792; lightningdit.py:498-531) at
    depth 2: the row-statistics slots, the column panels of the backward kernels and the GEMM tails all depend on D."""
    from ldmae_b200.models.lightningdit import LightningDiT
    spec = O.DiTSpec(depth=2, hidden_size=hidden, patch_size=patch, num_heads=heads, input_size=inp, in_channels=16, num_classes=10)
    m = LightningDiT(input_size=inp, patch_size=patch, in_channels=16, hidden_size=hidden, depth=2, num_heads=heads, num_classes=10,
                     use_qknorm=True, use_swiglu=True, use_rope=True, use_rmsnorm=True)
    sd = O.synth_dit_state(spec, 77)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(hidden)
    B = 3
    x1 = torch.randn(B, 16, inp, inp, generator=g); x0 = torch.randn(B, 16, inp, inp, generator=g)
    t = torch.rand(B, generator=g); y = torch.randint(0, 10, (B,), generator=g)
    with torch.no_grad():
        out = m(x1.cuda(), t.cuda(), y.cuda())
        ref_out = O.dit_forward(sd, spec, x1, t, y)
    assert _rel(out, ref_out) < 1e-2
    ref_terms, ref = _oracle_grads(spec, sd, x1, t, x0, y)
    terms, ours = _our_grads(m, x1.cuda(), t, x0, y.cuda())
    assert _rel(terms["loss"], ref_terms["loss"].detach()) < 1e-2
    _compare(ours, ref)


def test_xl_head_dim_72_backward_matches_oracle_autograd():
    """LightningDiT-XL geometry (width 1152, 16 heads of 72) at depth 2: the wide-head training path (128-column head slots,
    head-norm / RoPE Jacobians over the real head_dim, gradients un-padded into the reference's parameter shapes)."""
    from ldmae_b200.models.lightningdit import LightningDiT
    spec = O.DiTSpec(depth=2, hidden_size=1152, patch_size=1, num_heads=16, input_size=16, in_channels=16, num_classes=10)
    m = LightningDiT(input_size=16, patch_size=1, in_channels=16, hidden_size=1152, depth=2, num_heads=16, num_classes=10,
                     use_qknorm=True, use_swiglu=True, use_rope=True, use_rmsnorm=True)
    sd = O.synth_dit_state(spec, 92)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(8)
    B = 3
    x1 = torch.randn(B, 16, 16, 16, generator=g); x0 = torch.randn(B, 16, 16, 16, generator=g)
    t = torch.rand(B, generator=g); y = torch.randint(0, 10, (B,), generator=g)
    ref_terms, ref = _oracle_grads(spec, sd, x1, t, x0, y)
    terms, ours = _our_grads(m, x1.cuda(), t, x0, y.cuda())
    assert _rel(terms["loss"], ref_terms["loss"].detach()) < 1e-2
    _compare(ours, ref)


def test_b1_backward_matches_oracle_autograd():
    """LightningDiT-B/1 at the benchmark shape (T = 1024, D = 768, 12 blocks), batch 2."""
    from ldmae_b200.models.lightningdit import LightningDiT_models
    spec = O.DiTSpec.named("LightningDiT-B/1", input_size=32, in_channels=16)
    sd = O.synth_dit_state(spec, 3)
    m = LightningDiT_models["LightningDiT-B/1"](input_size=32, in_channels=16, use_qknorm=True, use_swiglu=True, use_rope=True,
                                                use_rmsnorm=True)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(9)
    B = 2
    x1 = torch.randn(B, 16, 32, 32, generator=g)
    x0 = torch.randn(B, 16, 32, 32, generator=g)
    t = torch.rand(B, generator=g)
    y = torch.randint(0, 1000, (B,), generator=g)
    ref_terms, ref = _oracle_grads(spec, sd, x1, t, x0, y)
    terms, ours = _our_grads(m, x1.cuda(), t, x0, y.cuda())
    assert _rel(terms["loss"], ref_terms["loss"].detach()) < 1e-2
    # the embedding table only has gradient rows for the two drawn labels: compare it as a whole like the others
    _compare(ours, ref, tol=4e-2)


def test_fused_adamw_ema_matches_torch():
    from ldmae_b200 import _lib
    n = 100_003 * 4
    g = torch.Generator().manual_seed(1)
    p0 = torch.randn(n, generator=g)
    ema0 = p0.clone()
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p_ref], lr=2e-4, betas=(0.9, 0.95), weight_decay=0.01)
    ema_ref = ema0.clone()
    p = p0.clone().cuda(); m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda"); ema = ema0.clone().cuda()
    for step in range(1, 4):
        grad = torch.randn(n, generator=g)
        p_ref.grad = grad.clone()
        opt.step()
        ema_ref.mul_(0.9999).add_(p_ref.data, alpha=1 - 0.9999)
        gd = (grad * 2).cuda()                              # grad_scale 0.5 undoes the factor (all-reduce average)
        _lib.check(_lib.lib().ldmae_adamw_ema_step(_lib.ptr(p), _lib.ptr(gd), _lib.ptr(m), _lib.ptr(v), _lib.ptr(ema), n, 2e-4, 0.9,
                                                   0.95, 1e-8, 0.01, step, 0.9999, 0.5, _lib.stream_ptr()), "adamw")
    torch.cuda.synchronize()
    torch.testing.assert_close(p.cpu(), p_ref.data, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(ema.cpu(), ema_ref, rtol=1e-5, atol=1e-6)


def test_fused_trainer_step_matches_autograd_plus_torch_adamw():
    """FusedTrainer.step (flat buffers, fused AdamW + EMA) == oracle autograd gradients + torch.optim.AdamW + update_ema."""
    from ldmae_b200.training import FusedTrainer
    spec, sd, m = _tiny(1, 31)
    g = torch.Generator().manual_seed(77)
    B = 4
    x1 = torch.randn(B, 16, 8, 8, generator=g); x0 = torch.randn(B, 16, 8, 8, generator=g)
    t = torch.rand(B, generator=g); y = torch.randint(0, 10, (B,), generator=g)
    _, ref_grads = _oracle_grads(spec, sd, x1, t, x0, y)
    ref_params = {k: torch.nn.Parameter(v.clone()) for k, v in sd.items() if k in ref_grads}
    opt = torch.optim.AdamW(list(ref_params.values()), lr=1e-3, betas=(0.9, 0.95), weight_decay=0.0)
    for k, p in ref_params.items():
        p.grad = ref_grads[k].clone()
    opt.step()
    tr = FusedTrainer(m, lr=1e-3, betas=(0.9, 0.95), weight_decay=0.0, ema_decay=0.99)
    with torch.no_grad():
        loss = tr.step(x1.cuda(), y.cuda(), t=t.cuda(), x0=x0.cuda())
    torch.cuda.synchronize()
    assert torch.isfinite(loss).all()
    # first Adam step moves every weight by lr * sign(grad) (up to eps): compare the update direction and size
    agree, total = 0, 0
    for k, p in m.named_parameters():
        if k not in ref_params:
            continue
        du = (p.detach().cpu() - sd[k]).flatten()
        dr = (ref_params[k].detach() - sd[k]).flatten()
        big = ref_grads[k].flatten().abs() > 1e-6           # sign of tiny gradients is not stable in bf16
        agree += int((torch.sign(du[big]) == torch.sign(dr[big])).sum()); total += int(big.sum())
        assert float(du.abs().max()) <= 1e-3 * 1.001
    assert agree / total > 0.97, agree / total
    ema = tr.ema_state_dict()
    k = "blocks.0.attn.qkv.weight"
    torch.testing.assert_close(ema[k].cpu(), 0.99 * sd[k] + 0.01 * dict(m.named_parameters())[k].detach().cpu(), rtol=1e-5, atol=1e-7)
    # the next forward uses the updated weights (bf16 copies re-packed)
    with torch.no_grad():
        out = m(x1.cuda(), t.cuda(), y.cuda())
    new_sd = {k_: v.detach().cpu() for k_, v in m.state_dict().items()}
    ref_out = O.dit_forward(new_sd, spec, x1, t, y)
    assert _rel(out, ref_out) < 1e-2


def test_fused_trainer_cosine_loss_and_gradient_clipping():
    """The loop's two optional terms (train_accum.py:216-218 with use_cosine_loss, :235-238 with optimizer.max_grad_norm):
    the flat gradient of (cos_loss.mean() + mse.mean()) against oracle autograd, and the clipped gradient / reported norm
    against torch.nn.utils.clip_grad_norm_ on the oracle's gradients."""
    from ldmae_b200.training import FusedTrainer
    from ldmae_b200.transport import create_transport
    spec, sd, m = _tiny(1, 61)
    g = torch.Generator().manual_seed(78)
    B = 4
    x1 = torch.randn(B, 16, 8, 8, generator=g); x0 = torch.randn(B, 16, 8, 8, generator=g)
    t = torch.rand(B, generator=g); y = torch.randint(0, 10, (B,), generator=g)
    leaves = {k: v.clone().requires_grad_(v.is_floating_point() and k != "pos_embed" and not k.startswith("feat_rope"))
              for k, v in sd.items()}
    terms = O.training_losses(lambda xt, tt, y: O.dit_forward(leaves, spec, xt, tt, y), x1, t, x0, y=y)
    cos = (1 - torch.nn.functional.cosine_similarity(terms["pred"], terms["ut"], dim=1)).mean(dim=(1, 2))
    (cos.mean() + terms["loss"].mean()).backward()
    ref = {k: v.grad for k, v in leaves.items() if v.grad is not None}
    tr = FusedTrainer(m, lr=1e-3, transport=create_transport("Linear", "velocity", None, None, None, use_cosine_loss=True,
                                                             use_lognorm=True))
    with torch.no_grad():
        loss, _ = tr.loss_and_grad(x1.cuda(), y.cuda(), t.cuda(), x0.cuda())
    torch.cuda.synchronize()
    assert _rel(loss, terms["loss"].detach()) < 1e-2 and _rel(tr.last_cos_loss, cos.detach()) < 1e-2
    _compare({k: tr.grad_of(k) for k in tr.names}, ref)
    # clipping: threshold at half of the true norm
    ref_params = [torch.nn.Parameter(sd[k].clone()) for k in ref]
    for p_, k in zip(ref_params, ref):
        p_.grad = ref[k].clone()
    total = float(torch.nn.utils.clip_grad_norm_(ref_params, 1e9))
    tr.max_grad_norm = 0.5 * total
    before = {k: sd[k].clone() for k in tr.names}
    tr.optimizer_step()
    torch.cuda.synchronize()
    assert abs(float(tr.last_grad_norm) - total) < 2e-2 * total
    assert abs(float(tr.grad.norm()) - 0.5 * total) < 2e-2 * total           # the buffer was scaled in place
    moved = max(float((dict(m.named_parameters())[k].detach().cpu() - before[k]).abs().max()) for k in tr.names)
    assert 0 < moved <= 1e-3 * 1.001                                          # first AdamW step: |update| <= lr


def test_gradient_accumulation_equals_the_full_batch():
    """FusedTrainer micro-batches (train_accum.py gradient_accumulation_steps): the accumulated flat gradient of 2 micro-batches
    equals the gradient of the full batch (same draws), up to fp32 summation order."""
    from ldmae_b200.training import FusedTrainer
    g = torch.Generator().manual_seed(5)
    B = 6
    x1 = torch.randn(B, 16, 8, 8, generator=g).cuda(); x0 = torch.randn(B, 16, 8, 8, generator=g).cuda()
    t = torch.rand(B, generator=g).cuda(); y = torch.randint(0, 10, (B,), generator=g).cuda()
    _, _, m = _tiny(1, 41)
    tr = FusedTrainer(m, lr=1e-3)
    with torch.no_grad():
        tr.loss_and_grad(x1, y, t, x0)
        full = tr.grad.clone()
        for i in range(2):
            sl = slice(3 * i, 3 * i + 3)
            tr.loss_and_grad(x1[sl], y[sl], t[sl], x0[sl], accumulate=i > 0, loss_scale=0.5)
    torch.cuda.synchronize()
    assert _rel(tr.grad, full) < 2e-3


def test_fused_trainer_reduces_the_loss_on_a_fixed_batch():
    """End-to-end sanity of the training loop (forward, backward, AdamW, EMA, weight re-pack): 40 optimizer steps on one fixed
    batch with fixed draws must drive the flow-matching loss down, and the EMA weights must trail the live ones."""
    from ldmae_b200.training import FusedTrainer
    _, sd, m = _tiny(1, 51)
    g = torch.Generator().manual_seed(3)
    B = 8
    x1 = torch.randn(B, 16, 8, 8, generator=g).cuda(); x0 = torch.randn(B, 16, 8, 8, generator=g).cuda()
    t = torch.rand(B, generator=g).cuda(); y = torch.randint(0, 10, (B,), generator=g).cuda()
    tr = FusedTrainer(m, lr=2e-3, betas=(0.9, 0.95), weight_decay=0.0, ema_decay=0.9)
    losses = []
    with torch.no_grad():
        for _ in range(40):
            losses.append(float(tr.step(x1, y, t=t, x0=x0).mean()))
    assert all(l == l for l in losses)                      # finite
    assert losses[-1] < 0.6 * losses[0], (losses[0], losses[-1])
    k = "blocks.0.mlp.w12.weight"
    live = dict(m.named_parameters())[k].detach()
    ema = tr.ema_state_dict()[k]
    d_live = float((live.cpu() - sd[k]).norm()); d_ema = float((ema.cpu() - sd[k]).norm())
    assert 0 < d_ema < d_live                                # the EMA moved, but less than the live weights


def test_second_forward_before_backward_fails_loudly():
    """The library keeps the activations of ONE training forward per model; a backward whose forward is no longer the latest
    (two grad-enabled forwards, or any inference call in between) must raise instead of returning wrong gradients."""
    from ldmae_b200 import _lib
    spec, sd, m = _tiny(1, 12)
    g = torch.Generator().manual_seed(1)
    xa, xb = torch.randn(2, 16, 8, 8, generator=g).cuda(), torch.randn(2, 16, 8, 8, generator=g).cuda()
    t = torch.rand(2, generator=g).cuda(); y = torch.randint(0, 10, (2,), generator=g).cuda()
    a = m(xa, t, y)
    b = m(xb, t, y)
    with pytest.raises(_lib.LdmaeError, match="ONE training forward"):
        (a.sum() + b.sum()).backward()
    m.zero_grad(set_to_none=True)
    c = m(xa, t, y)
    with torch.no_grad():
        m(xb, t, y)                                   # inference call in between overwrites the shared workspace
    with pytest.raises(_lib.LdmaeError, match="ONE training forward"):
        c.sum().backward()
    # the C ABI itself refuses too (callers that bypass the autograd node)
    L, h = _lib.lib(), m._handle
    out = torch.empty_like(xa)
    _lib.check(L.ldmae_dit_train_forward(h, _lib.ptr(xa), _lib.ptr(t), _lib.ptr(y), _lib.ptr(out), 2, _lib.stream_ptr()))
    _lib.check(L.ldmae_dit_forward(h, _lib.ptr(xb), _lib.ptr(t), 0.0, _lib.ptr(y), _lib.ptr(out), 2, 2, _lib.stream_ptr()))
    assert L.ldmae_dit_backward(h, _lib.ptr(out), 2, _lib.stream_ptr()) != 0
    assert b"one training forward" in L.ldmae_last_error()
    # and the regular order still works afterwards
    m.zero_grad(set_to_none=True)
    m(xa, t, y).sum().backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)


def test_out_of_range_class_label_is_flagged_not_read():
    """nn.Embedding would device-assert (lightningdit.py:146-169); the library clamps the index, never reads or writes
    outside the table, and reports the bad label."""
    from ldmae_b200 import _lib
    from ldmae_b200.models.lightningdit import LightningDiT
    from ldmae_b200.pipeline import SamplingJob
    from ldmae_b200.tokenizer import models_mae
    m = LightningDiT(input_size=8, patch_size=1, in_channels=16, hidden_size=128, depth=2, num_heads=2, num_classes=10,
                     class_dropout_prob=0.0, use_qknorm=True, use_swiglu=True, use_rope=True, use_rmsnorm=True).cuda().eval()
    assert m.y_embedder.embedding_table.weight.shape[0] == 10          # no null row without label dropout
    x = torch.randn(2, 16, 8, 8).cuda(); t = torch.rand(2).cuda()
    with torch.no_grad():
        m(x, t, torch.tensor([3, 9]).cuda())                            # in range: fine
        L, h = _lib.lib(), m._handle
        _lib.check(L.ldmae_dit_check_labels(h, _lib.stream_ptr()))
        m(x, t, torch.tensor([3, 10]).cuda())                           # 10 is outside a 10-row table
        assert L.ldmae_dit_check_labels(h, _lib.stream_ptr()) != 0
        assert b"class label outside" in L.ldmae_last_error()
        m(x, t, torch.tensor([0, 1]).cuda())                            # the flag was reported once and cleared
        _lib.check(L.ldmae_dit_check_labels(h, _lib.stream_ptr()))
    vae = models_mae.mae_for_ldmae_f8d16_prev(ldmae_mode=True, no_cls=True, kl_loss_weight=True, smooth_output=True, img_size=64)
    with pytest.raises(ValueError, match="null-class row"):
        SamplingJob(m, vae.cuda().eval(), num_steps=4, cfg_scale=4.0)   # CFG needs the null row this model does not have


def test_flow_prepare_and_loss_kernels_vs_reference_expressions():
    """The fused trainer-input kernel (flip select, posterior sample, normalise, xt / ut) and the fused loss kernel against the
    reference's expressions evaluated by PyTorch on the same device (img_latent_dataset.py:76-94, tokenizer/util/misc.py:74-96,
    transport.py:136-166,195, path.py:114-136)."""
    from ldmae_b200.tokenizer.models_mae import DiagonalGaussianDistribution
    from ldmae_b200.training import FusedTrainer
    spec, sd, m = _tiny(1, 12)
    tr = FusedTrainer(m)
    g = torch.Generator(device="cuda").manual_seed(5)
    B, C, S = 6, 16, 8
    mom = torch.randn(B, 2 * C, S, S, device="cuda", generator=g)
    mom[:, C:] = mom[:, C:] * 3.0 - 1.0
    mom[0, C] = 45.0; mom[1, C + 1] = -80.0                                   # both clamps of the log-variance
    momf = torch.randn(B, 2 * C, S, S, device="cuda", generator=g)
    flip = torch.tensor([0, 1, 1, 0, 1, 0], device="cuda", dtype=torch.bool)
    eps = torch.randn(B, C, S, S, device="cuda", generator=g)
    x0 = torch.randn(B, C, S, S, device="cuda", generator=g)
    t = torch.rand(B, device="cuda", generator=g)
    mean = torch.linspace(-0.3, 0.4, C, device="cuda").view(1, C, 1, 1); std = torch.linspace(0.6, 1.7, C, device="cuda").view(1, C, 1, 1)
    mult = 0.75
    xt, ut, x1 = tr.prepare(None, t=t, x0=x0, moments=mom, moments_flip=momf, flip=flip, eps_post=eps, latent_mean=mean,
                            latent_std=std, latent_multiplier=mult, want_x1=True)
    src = torch.where(flip.view(B, 1, 1, 1), momf, mom)
    post = DiagonalGaussianDistribution(src)
    feat = ((post.mean + post.std * eps) - mean) / std * mult
    tt = t.view(B, 1, 1, 1)
    torch.testing.assert_close(x1, feat, rtol=2e-6, atol=1e-6)
    assert torch.equal(xt, tt * x1 + (1 - tt) * x0)                           # same roundings as the eager expression
    assert torch.equal(ut, x1 - x0)
    # posterior mode + no normalisation; ready latents
    xt2, ut2, x12 = tr.prepare(None, t=t, x0=x0, moments=mom, want_x1=True)
    assert torch.equal(x12, mom[:, :C])
    xt3, ut3, _ = tr.prepare(x12, t=t, x0=x0)
    assert torch.equal(xt3, xt2) and torch.equal(ut3, ut2)
    # loss + output gradient
    from ldmae_b200 import _lib
    out = torch.randn(B, C, S, S, device="cuda", generator=g)
    loss = torch.empty(B, device="cuda"); dout = torch.empty_like(out)
    _lib.check(_lib.lib().ldmae_flow_loss(_lib.ptr(out), _lib.ptr(ut), _lib.ptr(loss), _lib.ptr(dout), 0.5, B, C * S * S, _lib.stream_ptr()))
    want = ((out - ut) ** 2).mean(dim=(1, 2, 3))
    torch.testing.assert_close(loss, want, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(dout, (out - ut) * (2.0 * 0.5 / (C * S * S * B)), rtol=1e-6, atol=1e-9)


def test_step_from_moments_equals_step_on_prepared_latents():
    """FusedTrainer.step_from_moments (dataset features in) == FusedTrainer.step on the latents the dataset would have produced."""
    import copy
    from ldmae_b200.tokenizer.models_mae import DiagonalGaussianDistribution
    from ldmae_b200.training import FusedTrainer
    spec, sd, m = _tiny(1, 12)
    m2 = copy.deepcopy(m)
    g = torch.Generator(device="cuda").manual_seed(8)
    B, C, S = 4, 16, 8
    mom = torch.randn(B, 2 * C, S, S, device="cuda", generator=g); momf = torch.randn(B, 2 * C, S, S, device="cuda", generator=g)
    flip = torch.tensor([1, 0, 0, 1], device="cuda", dtype=torch.bool)
    eps = torch.randn(B, C, S, S, device="cuda", generator=g); x0 = torch.randn(B, C, S, S, device="cuda", generator=g)
    t = torch.rand(B, device="cuda", generator=g); y = torch.randint(0, 10, (B,), device="cuda", generator=g)
    mean = torch.full((1, C, 1, 1), 0.1, device="cuda"); std = torch.full((1, C, 1, 1), 1.3, device="cuda")
    a, b = FusedTrainer(m), FusedTrainer(m2)
    la = a.step_from_moments(mom, momf, y, latent_mean=mean, latent_std=std, latent_multiplier=1.0, flip=flip, eps_post=eps, t=t, x0=x0)
    src = torch.where(flip.view(B, 1, 1, 1), momf, mom)
    post = DiagonalGaussianDistribution(src)
    x1 = ((post.mean + post.std * eps) - mean) / std * 1.0
    lb = b.step(x1, y, t=t, x0=x0)
    torch.cuda.synchronize()
    torch.testing.assert_close(la, lb, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(a.flat, b.flat, rtol=1e-5, atol=1e-7)
