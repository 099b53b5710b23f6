"""runtime.TrainStep on a GPU: the captured step (multi-tensor weight refresh, side-stream weight gradients,
flat gradient buckets, optional fused clip + Adam) against the plain eager autograd step of the same model."""
import pytest
import torch

from fs2b200 import sub
from oracle import synth
from tests.util_parity import cuda_batch, disable_dropout, rel_err

pytestmark = pytest.mark.gpu


def _build(seed=0, **cfg_over):
    M = sub("lightning.model")
    cfg = synth.model_cfg(encoder_layer=2, decoder_layer=2, **cfg_over)
    model = M.FastSpeech2(cfg)
    model.load_state_dict(synth.init_state_dict(model.state_dict(), seed))
    return disable_dropout(model.cuda().train()), M.FastSpeech2Loss(cfg), cfg


@pytest.mark.parametrize("use_graph", [False, True])
def test_train_step_matches_eager_step(use_graph):
    rt = sub("runtime")
    batch = synth.make_batch(B=4, src_len=(10, 40), dur=synth.uniform_dur(1, 8), seed=21)
    # eager reference: per-call weight casts, torch-allocated gradients
    model_e, loss_fn, _ = _build()
    b = cuda_batch(batch)
    out = model_e(b[2], b[3], *b[4:12], lang_args=b[12])
    losses_e = loss_fn(b[:-1], out)
    losses_e[0].backward()
    grads_e = {k: p.grad.detach().clone() for k, p in model_e.named_parameters() if p.grad is not None}
    # captured / bucketed step on an identical model
    model_s, loss_fn_s, _ = _build()
    step = rt.TrainStep(model_s, loss_fn_s, batch, use_graph=use_graph)
    for _ in range(2):  # replays are idempotent without an optimizer (BatchNorm running stats aside)
        got = step.step_e2e(batch).clone()
    for a, r in zip(got.tolist(), [float(l) for l in losses_e]):
        assert abs(a - r) <= 1e-4 * abs(r), (a, r)
    # The activation / input-gradient chain is bit-reproducible (fixed-order BatchNorm reductions, no atomics); only
    # gradients that are themselves accumulated with fp32 atomics (split-K weight gradients, column sums) differ,
    # by their summation order (~1e-7, tools/debug_determinism.py).  A wrong cached weight, a stream race or a
    # lost gradient shows up as O(1).
    gmax = max(float(g.norm()) for g in grads_e.values())
    for k, p in model_s.named_parameters():
        if k not in grads_e or float(grads_e[k].norm()) < 1e-3 * gmax:
            continue
        assert rel_err(p.main_grad, grads_e[k]) < 1e-3, (k, rel_err(p.main_grad, grads_e[k]))


def test_train_step_with_fused_adam_learns_and_counts_steps():
    rt = sub("runtime")
    batch = synth.make_batch(B=4, src_len=(10, 40), dur=synth.uniform_dur(1, 8), seed=22)
    model, loss_fn, _ = _build(seed=1)
    buckets = rt.GradBuckets(model.parameters(), device=torch.device("cuda"))
    cfg = {"scheduler_type": "const", "optimizer": {"betas": [0.9, 0.98], "eps": 1e-9, "weight_decay": 0.0,
                                                   "grad_clip_thresh": 1.0, "warm_up_step": 0, "anneal_steps": [],
                                                   "anneal_rate": 1.0, "lr": 1e-3}}
    opt = rt.FusedAdam(buckets, train_config=cfg)
    w0 = model.mel_linear.weight.detach().clone()
    step = rt.TrainStep(model, loss_fn, batch, use_graph=True, buckets=buckets, optimizer=opt)
    # constructing the step (two eager warm-up bodies + capture) must leave the training state untouched
    assert int(opt.step_dev) == 0 and float(opt.exp_avg.abs().sum()) == 0.0
    assert torch.equal(model.mel_linear.weight.detach(), w0)
    assert int(model.postnet.convolutions[0][1].num_batches_tracked) == 0
    first = float(step.step_e2e(batch)[0])
    for _ in range(20):
        last = float(step.step_e2e(batch)[0])
    assert int(opt.step_dev) == 21
    assert int(model.postnet.convolutions[0][1].num_batches_tracked) == 21
    assert not torch.equal(model.mel_linear.weight.detach(), w0)
    assert last < 0.9 * first, (first, last)  # same batch every step: the loss must go down
    assert float(opt.grad_norm) > 0


def test_prefetched_batches_are_the_ones_that_run():
    """prefetch_batch() stages the NEXT batch on a copy stream; run() must consume exactly that batch."""
    rt = sub("runtime")
    b1 = synth.make_batch(B=4, src_len=(10, 40), dur=synth.uniform_dur(1, 8), seed=31)
    model, loss_fn, _ = _build()
    step = rt.TrainStep(model, loss_fn, b1, use_graph=True)
    ref1 = step.step_e2e(b1).clone()
    # same padded shape, different contents: permute the utterances and perturb the mel targets
    perm = torch.tensor([2, 0, 3, 1])
    b2 = list(b1)
    for i in (2, 3, 4, 6, 7, 9, 10, 11, 12):
        b2[i] = b1[i][perm].clone()
    b2[6] = b2[6] + 0.5
    ref2 = step.step_e2e(tuple(b2)).clone()
    assert abs(float(ref2[1]) - float(ref1[1])) > 1e-3  # the mel loss moved
    step.prefetch_batch(b1)
    step.run()
    got1 = step.read_losses().clone()
    step.prefetch_batch(tuple(b2))
    step.run()
    got2 = step.read_losses().clone()
    assert torch.allclose(got1, ref1, rtol=1e-4) and torch.allclose(got2, ref2, rtol=1e-4)


def test_fused_adam_first_step_moves_every_weight_against_its_gradient():
    """Conv1d weight gradients live in [Co][k][Ci] order inside the flat bucket and FusedAdam re-homes the weights
    in the same order.  First Adam step: delta = -lr * g / (|g| + eps) = -lr * sign(g) element by element, so a layout
    mismatch between value and gradient would show up as ~50 % sign agreement."""
    rt = sub("runtime")
    batch = synth.make_batch(B=4, src_len=(10, 40), dur=synth.uniform_dur(1, 8), seed=23)
    model_e, loss_fn, _ = _build(seed=2)
    b = cuda_batch(batch)
    out = model_e(b[2], b[3], *b[4:12], lang_args=b[12])
    loss_fn(b[:-1], out)[0].backward()
    model, loss_fn2, _ = _build(seed=2)
    before = {k: p.detach().clone() for k, p in model.named_parameters()}
    buckets = rt.GradBuckets(model.parameters(), device=torch.device("cuda"))
    lr = 1e-3
    opt = rt.FusedAdam(buckets, lr=lr, betas=(0.9, 0.98), eps=1e-12, weight_decay=0.0, max_grad_norm=0.0,
                       scheduler_type="none")
    step = rt.TrainStep(model, loss_fn2, batch, use_graph=False, buckets=buckets, optimizer=opt)
    step.run()
    torch.cuda.synchronize()
    assert int(opt.step_dev) == 1
    grads_e = dict((k, p.grad) for k, p in model_e.named_parameters())
    for k in ("decoder.layer_stack.0.pos_ffn.w_1.weight", "postnet.convolutions.1.0.conv.weight",
              "variance_adaptor.duration_predictor.conv_layer.conv1d_1.conv.weight", "mel_linear.weight"):
        p = dict(model.named_parameters())[k]
        delta = p.detach() - before[k]
        g = grads_e[k]
        big = g.abs() > 0.05 * g.abs().mean()  # elements whose sign is not rounding noise
        agree = ((delta < 0) == (g > 0))[big].float().mean().item()
        assert agree > 0.97, (k, agree)
        assert torch.allclose(delta.abs()[big], torch.full_like(delta[big], lr), rtol=2e-2), k
    # state_dict still has the reference's shapes, whatever the storage order
    assert model.state_dict()["decoder.layer_stack.0.pos_ffn.w_1.weight"].shape == (1024, 256, 9)


def test_shape_bucketed_step_cache_reuses_graphs_and_matches_the_padded_batch():
    """runtime.TrainStepCache: batches with different (max_src_len, max_mel_len) that fall into one bucket share ONE
    captured graph; the result equals a fresh step on the bucket-padded batch (which is what the reference computes
    for that padded batch), and a batch of another bucket triggers exactly one more capture."""
    rt = sub("runtime")
    model, loss_fn, _ = _build()
    cache = rt.TrainStepCache(model, loss_fn, bucket=(32, 128), max_graphs=2)
    b1 = synth.make_batch(B=4, src_len=(10, 28), dur=synth.uniform_dur(1, 4), seed=41)
    b2 = synth.make_batch(B=4, src_len=(12, 30), dur=synth.uniform_dur(1, 4), seed=42)
    assert (int(b1[5]), int(b1[8])) != (int(b2[5]), int(b2[8])) and cache.key_for(b1) == cache.key_for(b2)
    l1 = cache.run(b1).clone()
    g1 = cache.buckets.flat.clone()
    l2 = cache.run(b2).clone()
    assert cache.captures == 1 and len(cache.steps) == 1
    assert not torch.allclose(l1, l2)
    # the same batch again reproduces itself bit-exactly through the shared graph (activations deterministic;
    # atomically accumulated weight gradients to fp32 round-off)
    l1b = cache.run(b1).clone()
    assert torch.equal(l1, l1b) and rel_err(cache.buckets.flat, g1) < 1e-5
    # against an independent eager step on the explicitly padded batch
    model_e, loss_fn_e, _ = _build()
    key = cache.key_for(b1)
    padded = rt.pad_batch(b1, key[1], key[2])
    assert padded[3].shape[1] == key[1] and padded[6].shape[1] == key[2] and padded[8] == key[2]
    pb = cuda_batch(padded)
    out = model_e(pb[2], pb[3], *pb[4:12], lang_args=pb[12])
    le = loss_fn_e(pb[:-1], out)
    for a, r in zip(l1.tolist(), [float(x) for x in le]):
        assert abs(a - r) <= 1e-4 * abs(r), (a, r)
    # another bucket: one more capture; a third bucket evicts the least recently used graph (max_graphs = 2)
    b3 = synth.make_batch(B=4, src_len=(40, 60), dur=synth.uniform_dur(2, 5), seed=43)
    cache.run(b3)
    assert cache.captures == 2 and len(cache.steps) == 2
    b4 = synth.make_batch(B=4, src_len=(70, 90), dur=synth.uniform_dur(2, 5), seed=44)
    cache.run(b4)
    assert cache.captures == 3 and len(cache.steps) == 2 and cache.key_for(b1) not in cache.steps


def test_weight_cache_refresh_matches_per_tensor_cast_and_pack():
    """csrc/weight_prep.cu (one launch for every bf16 operand copy) against torch: Linear weights are cast, Conv1d
    weights [Co][Ci][k] become [Co][k][Cpad] with zero channel padding, Q|K|V weights / biases are gathered."""
    ops = sub("ops")
    MHA = sub("transformer.SubLayers").MultiHeadAttention
    model, _, _ = _build(seed=3)
    wc = ops.WeightCache(model)
    for t in wc.map.values():
        t.fill_(7.0)
    wc.refresh()
    torch.cuda.synchronize()
    n = 0
    for m in model.modules():
        if isinstance(m, torch.nn.Conv1d):
            d = wc.lookup("conv", m.weight)
            if d is None:
                continue
            Co, Ci, k = m.weight.shape
            ref = torch.zeros_like(d)
            ref[:, :, :Ci] = m.weight.detach().permute(0, 2, 1).to(torch.bfloat16)
            assert torch.equal(d, ref), (Co, Ci, k)
            n += 1
        elif isinstance(m, MHA):
            w = wc.lookup("qkv", m.w_qs.weight)
            b = wc.lookup("bqkv", m.w_qs.bias)
            assert torch.equal(w, torch.cat([m.w_qs.weight, m.w_ks.weight, m.w_vs.weight]).detach().to(torch.bfloat16))
            assert torch.equal(b, torch.cat([m.w_qs.bias, m.w_ks.bias, m.w_vs.bias]).detach())
            n += 1
        elif isinstance(m, torch.nn.Linear):
            d = wc.lookup("lin", m.weight)
            if d is not None:
                assert torch.equal(d, m.weight.detach().to(torch.bfloat16))
                n += 1
    assert n > 20


def test_lagged_loss_readback_returns_the_previous_step():
    """TrainStep.read_losses_lagged: the host picks the losses up one step late and never waits for the step it
    has just launched; flat batch buffers: prefetch -> run moves a whole batch with one copy per hop."""
    rt = sub("runtime")
    batch = synth.make_batch(B=4, src_len=(10, 40), dur=synth.uniform_dur(1, 8), seed=23)
    other = list(batch)  # same padded shape, other targets
    other[9] = batch[9] * 0.5 + 0.25
    other = tuple(other)
    model, loss_fn, _ = _build(seed=2)
    step = rt.TrainStep(model, loss_fn, batch, use_graph=True)
    la = step.step_e2e(batch).clone()
    lb = step.step_e2e(other).clone()
    assert not torch.allclose(la, lb)
    step.prefetch_batch(batch)
    step.run()
    assert step.read_losses_lagged() is None
    step.prefetch_batch(other)
    step.run()
    r = step.read_losses_lagged()
    assert torch.allclose(r, la, rtol=1e-5), (r, la)
    step.prefetch_batch(batch)
    step.run()
    r = step.read_losses_lagged()
    assert torch.allclose(r, lb, rtol=1e-5), (r, lb)
