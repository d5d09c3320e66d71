"""Data-parallel training step for the denoising network: the hot loop of the reference's
train_denoising_syn.py:169-184, as one kernel-only CUDA program per step.

    im_noisy, im_gt, sigma_gt  (host or device, NCHW fp32)
      -> forward (SNet, RNet)            virnet_b200.engine.DenoiseEngine.forward
      -> fused ELBO (loss + d_mu, d_sigma)          vk_elbo_denoise
      -> backward (dgrad / wgrad kernels)           DenoiseEngine.backward
      -> [world_size > 1] NCCL all-reduce (sum) of the flat fp32 gradient buffer over NVLink
      -> per-sub-network grad-norm -> clip -> Adam  vk_adam_clip_step   (averaging folded in)

Semantics follow the reference: each rank computes the mean loss over its local batch, DDP
averages gradients over ranks, clip_grad_norm_ runs per sub-network (RNet / SNet) on the averaged
gradients, Adam(lr, betas=(0.9, 0.999), eps=1e-8) without weight decay; the LR is not rescaled
with the world size.  There is no host synchronisation inside a step.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch
import torch.distributed as dist

from . import dp
from . import lib as _l
from . import ops
from .loss.ELBO_simple import _digamma, sisr_draws
from .loss.resize_right import downsample_matrix


def _clip_groups(net, eng, spec):
    """Contiguous ranges of the flat parameter buffer per sub-network: [(key, max_norm)] -> device descriptor table."""
    names = [n for n, _ in net.named_parameters()]
    groups = []
    for key, max_norm in spec:
        idx = [i for i, n in enumerate(names) if n.lower().startswith(key)]
        assert idx and idx == list(range(idx[0], idx[-1] + 1)), "sub-network parameters must be contiguous"
        begin = eng.flat_offsets[idx[0]]
        end = eng.flat_offsets[idx[-1] + 1] if idx[-1] + 1 < len(names) else eng.flat_total
        groups.append((begin, end, float(max_norm)))
    arr = (_l.vk_adam_group * len(groups))()
    for i, (b, e, m) in enumerate(groups):
        arr[i].begin, arr[i].end, arr[i].max_norm = b, e, m
    dev = eng.flat_params.device
    table = torch.frombuffer(bytearray(bytes(memoryview(arr))), dtype=torch.uint8).to(dev)
    return table, len(groups), max(e - b for b, e, _ in groups)


def _bias_corrections(betas, step):
    """(1 - beta1^step, sqrt(1 - beta2^step)) exactly as vk_adam_clip_step computes them: double arithmetic on the betas
    rounded to fp32 (the kernel argument type), so the graph-replayed update equals the eager one bit for bit."""
    import struct
    b1, b2 = (struct.unpack("f", struct.pack("f", b))[0] for b in betas)
    return 1.0 - b1 ** step, (1.0 - b2 ** step) ** 0.5


class DenoiseTrainer:
    def __init__(self, net, lr=1e-4, clip_grad_R=1e3, clip_grad_S=1e2, alpha0=24.5, eps2=1e-6, betas=(0.9, 0.999),
                 adam_eps=1e-8, process_group=None, deterministic=None):
        """deterministic=True: split-K weight gradients are reduced in a fixed order (engine.deterministic) — with the
        atomic-free loss / norm / bias reductions every step is then run-to-run bit-identical, eager or graph-replayed,
        at the cost of one extra pass over the partial sums (a few % of the step)."""
        self.net = net
        self.engine = net.engine()
        if deterministic is not None:
            self.engine.deterministic = bool(deterministic)
        self.engine._ensure_flat()
        eng = self.engine
        dev = eng.flat_params.device
        self.lr, self.betas, self.adam_eps = lr, betas, adam_eps
        self.alpha0, self.eps2 = float(alpha0), float(eps2)
        self.digamma_am1 = _digamma(self.alpha0 - 1.0)
        self.pg = process_group
        self.world = dp.world_size(process_group)
        dp.broadcast_flat_params(eng.flat_params, 0, process_group)
        self.step_count = 0
        self.exp_avg = torch.zeros_like(eng.flat_params)
        self.exp_avg_sq = torch.zeros_like(eng.flat_params)
        # clip groups: contiguous ranges of the flat buffer (parameters() order is SNet..., RNet...)
        groups = []
        names = [n for n, _ in net.named_parameters()]
        for key, max_norm in (("snet", clip_grad_S), ("rnet", clip_grad_R)):
            idx = [i for i, n in enumerate(names) if key in n.lower()]
            assert idx == list(range(idx[0], idx[-1] + 1)), "sub-network parameters must be contiguous"
            begin = eng.flat_offsets[idx[0]]
            end = eng.flat_offsets[idx[-1] + 1] if idx[-1] + 1 < len(names) else eng.flat_total
            groups.append((begin, end, float(max_norm)))
        arr = (_l.vk_adam_group * len(groups))()
        for i, (b, e, m) in enumerate(groups):
            arr[i].begin, arr[i].end, arr[i].max_norm = b, e, m
        self.group_names = ["SNet", "RNet"]
        self._groups_dev = torch.frombuffer(bytearray(bytes(memoryview(arr))), dtype=torch.uint8).to(dev)
        self._ngroups = len(groups)
        self._max_group = max(e - b for b, e, _ in groups)
        self._sq_ws = ops.adam_ws(len(groups), dev)
        self.grad_norms = torch.zeros(len(groups), device=dev, dtype=torch.float32)
        self._acc3 = ops.elbo_ws(dev)
        self.losses = torch.zeros(4, device=dev, dtype=torch.float32)
        self._d_mu = self._d_sigma = None
        self._staging = {}
        self._copy_stream, self._ev = None, None
        self._pf_free = self._pf_ready = self._pf_used = self._pf_key = None
        self._pf_next = 0
        self._graph = self._graph_key = self._g_in = None
        self._hyper = torch.zeros(3, device=dev, dtype=torch.float32)
        self._hyper_ring = [(torch.zeros(3, dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(8)]
        self._hyper_used = [False] * 8
        # world > 1: the gradient all-reduce runs bucket by bucket on a side stream, overlapped with the rest of the
        # backward pass (dp.BucketedGradSync); VIRNET_B200_OVERLAP_ALLREDUCE=0 restores the single
        # blocking all-reduce after the backward
        self._sync = None
        if self.world > 1 and os.environ.get("VIRNET_B200_OVERLAP_ALLREDUCE", "1") != "0" and dev.type == "cuda":
            self._sync = dp.BucketedGradSync(eng, process_group)

    # -- host -> device staging (pinned host buffers are the caller's) --------------------------------------
    # The copies run on a side stream: the noisy patches are needed first (the weight packing of this step
    # overlaps their copy), the clean patches and the variance map only at the loss, a forward pass later.
    def _stage(self, name, t, stream):
        buf = self._staging.get(name)
        if buf is None or buf.shape != t.shape:
            buf = torch.empty(t.shape, device=self.engine.flat_params.device, dtype=torch.float32)
            self._staging[name] = buf
        with torch.cuda.stream(stream):
            buf.copy_(t, non_blocking=True)
        return buf

    def prefetch(self, im_noisy, im_gt, sigma_gt):
        """Start copying the NEXT batch (pinned host tensors) while the current step computes: what a
        DataLoader with pin_memory + non_blocking copies does for the reference loop (train_denoising_syn.py:169-171).
        The following step() call with the same host tensors picks the device copies up."""
        eng = self.engine
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=eng.flat_params.device)
            self._ev = [torch.cuda.Event() for _ in range(3)]
        if self._pf_free is None:
            self._pf_free = [torch.cuda.Event(), torch.cuda.Event()]
            self._pf_ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._pf_used = [False, False]
        k = self._pf_next
        self._pf_next ^= 1
        cs = self._copy_stream
        if self._pf_used[k]:
            cs.wait_event(self._pf_free[k])     # the step that last read this staging set has passed its loss
        bufs = [self._stage(f"pf{k}.{nm}", t, cs) for nm, t in (("noisy", im_noisy), ("gt", im_gt), ("sg", sigma_gt))]
        self._pf_ready[k].record(cs)
        self._pf_key = (k, im_noisy.data_ptr(), im_gt.data_ptr(), sigma_gt.data_ptr(), bufs)

    def step_real(self, im_noisy, im_gt, var_window=7, mixup=None, lr: Optional[float] = None):
        """The real-noise trainer's iteration (train_denoising_real.py:158-176): MixUp of the (clean, noisy) pairs,
        variance-map prior from the squared error under a var_window Gaussian, then the same step.  `mixup`: a
        virnet_b200.datasets.data_tools.MixUp_AUG (None = no MixUp)."""
        from .utils.util_denoising import noise_estimate_fun
        dev = self.engine.flat_params.device
        im_noisy, im_gt = [t if t.is_cuda else t.to(dev, non_blocking=True) for t in (im_noisy, im_gt)]
        if mixup is not None:
            im_gt, im_noisy = mixup.aug(im_gt, im_noisy)
        sigma_gt = noise_estimate_fun(im_noisy, im_gt, var_window)
        return self.step(im_noisy, im_gt, sigma_gt, lr=lr)

    def step(self, im_noisy, im_gt, sigma_gt, lr: Optional[float] = None):
        """One optimisation step; returns the device tensor [loss, lh, kl_gauss, kl_Igamma]
        of the local batch (no host sync)."""
        eng = self.engine
        main = torch.cuda.current_stream()
        ev_late = None
        pf_set = None
        if (self._pf_key is not None and not im_noisy.is_cuda
                and self._pf_key[1:4] == (im_noisy.data_ptr(), im_gt.data_ptr(), sigma_gt.data_ptr())):
            pf_set, x, gt, sg = self._pf_key[0], *self._pf_key[4]
            self._pf_key = None
            main.wait_event(self._pf_ready[pf_set])
        elif not (im_noisy.is_cuda and im_gt.is_cuda and sigma_gt.is_cuda):
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=eng.flat_params.device)
                self._ev = [torch.cuda.Event() for _ in range(3)]
            cs = self._copy_stream
            self._ev[2].record(main)            # the previous step has finished reading the staging buffers
            cs.wait_event(self._ev[2])
            x = im_noisy if im_noisy.is_cuda else self._stage("noisy", im_noisy, cs)
            self._ev[0].record(cs)
            gt = im_gt if im_gt.is_cuda else self._stage("gt", im_gt, cs)
            sg = sigma_gt if sigma_gt.is_cuda else self._stage("sigma_gt", sigma_gt, cs)
            self._ev[1].record(cs)
            ev_late = self._ev[1]
            eng._ensure_flat()
            eng._ensure_packed()                # runs while the first copy is in flight
            main.wait_event(self._ev[0])
        else:
            x, gt, sg = im_noisy, im_gt, sigma_gt
        self.step_count += 1
        self._device_step(x, gt, sg, ev_late, pf_set, self.lr if lr is None else lr, None)
        return self.losses

    def _device_step(self, x, gt, sg, ev_late, pf_set, lr, hyper_dev):
        """Everything of a step that runs on the device, for inputs already resident: shared by the eager path and by the
        CUDA-graph capture (hyper_dev: per-step Adam scalars in device memory instead of kernel arguments)."""
        self._device_fwd_bwd(x, gt, sg, ev_late, pf_set, overlap=True)
        self._device_update(lr, hyper_dev, reduced=self._reduced_in_backward)

    def _device_fwd_bwd(self, x, gt, sg, ev_late, pf_set, overlap):
        """forward + fused ELBO + backward; with `overlap` (and world > 1) the gradient buckets are all-reduced on the
        side stream while the backward is still running, and are complete when this returns (in stream order)."""
        eng = self.engine
        main = torch.cuda.current_stream()
        mu, sigma = eng.forward(x, save=True)
        if self._d_mu is None or self._d_mu.shape != mu.shape:
            self._d_mu, self._d_sigma = torch.empty_like(mu), torch.empty_like(sigma)
        if ev_late is not None:
            main.wait_event(ev_late)
        # beta0 = alpha0 * sigma_gt (train_denoising_syn.py:172) is folded into the loss kernel
        ops.elbo_denoise(mu, sigma, x, gt, sg, beta0_scale=self.alpha0, eps2=self.eps2, alpha0=self.alpha0,
                         digamma_am1=self.digamma_am1, d_mu=self._d_mu, d_sigma=self._d_sigma, acc3=self._acc3,
                         out4=self.losses)
        if pf_set is not None:
            self._pf_free[pf_set].record(main)   # inputs are not read after the loss kernel
            self._pf_used[pf_set] = True
        # bucket-wise overlap pays when the backward is long enough to hide 5 all-reduces behind; at the reference's
        # 2 patches per GPU the extra launches cost more than they hide (2 GPUs: 3.9 vs 3.3 ms eager) -> one blocking call
        big = x.shape[0] * x.shape[2] * x.shape[3] >= 8 * 128 * 128
        eng.grad_sync = self._sync if (overlap and big) else None
        self._reduced_in_backward = eng.grad_sync is not None
        try:
            eng.backward(self._d_mu, self._d_sigma)
        finally:
            eng.grad_sync = None
        self.last_mu, self.last_sigma = mu, sigma

    def _device_update(self, lr, hyper_dev, reduced):
        """[all-reduce unless the buckets were reduced during the backward] -> per-sub-network norm, clip, Adam."""
        eng = self.engine
        grad_scale = 1.0 / self.world if reduced else dp.all_reduce_flat_grads(eng.flat_grads, self.pg)
        if hyper_dev is None:
            ops.adam_clip_step(eng.flat_params, eng.flat_grads, self.exp_avg, self.exp_avg_sq, self._groups_dev,
                               self._ngroups, self._max_group, self._sq_ws, grad_scale=grad_scale, lr=lr,
                               beta1=self.betas[0], beta2=self.betas[1], eps=self.adam_eps, step=self.step_count,
                               norms_out=self.grad_norms)
        else:
            ops.adam_clip_step_dev(eng.flat_params, eng.flat_grads, self.exp_avg, self.exp_avg_sq, self._groups_dev,
                                   self._ngroups, self._max_group, self._sq_ws, hyper_dev, grad_scale=grad_scale,
                                   beta1=self.betas[0], beta2=self.betas[1], eps=self.adam_eps, norms_out=self.grad_norms)
        eng.mark_params_dirty()

    # -- CUDA-graph replay of the whole step ---------------------------------------------------------------------
    # At the reference's own batch (16 patches over 8 GPUs = 2 per GPU) a step is ~190 kernel launches of a few
    # microseconds each and the Python / driver enqueue (about 3 ms) is the bound; one graph launch removes it.
    def _capture(self, shapes):
        """world == 1: the whole step is one graph.  world > 1: forward + ELBO + backward (ending with the workspace ->
        parameter-layout conversion) is the graph; the NCCL all-reduce and the clip + Adam launch stay outside it —
        three enqueues per step instead of ~190 (NCCL inside a capture did not complete in a 2-GPU trial)."""
        eng = self.engine
        dev = eng.flat_params.device
        multi = self.world > 1
        self._g_in = [torch.zeros(s, device=dev, dtype=torch.float32) for s in shapes]
        state = [t.clone() for t in (eng.flat_params, self.exp_avg, self.exp_avg_sq)]
        self._hyper.copy_(torch.tensor([0.0, 1.0, 1.0]))     # warm-up / capture run with lr = 0

        def body():
            if multi:
                self._device_fwd_bwd(*self._g_in, None, None, overlap=False)
            else:
                self._device_fwd_bwd(*self._g_in, None, None, overlap=False)
                self._device_update(0.0, self._hyper, reduced=False)

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up outside capture (lazy allocations, attributes)
            for _ in range(2):
                eng.mark_params_dirty()
                body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(dev)
        eng.mark_params_dirty()                             # the captured step always re-packs the weights
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, capture_error_mode="thread_local" if multi else "global"):
            body()
        for t, sv in zip((eng.flat_params, self.exp_avg, self.exp_avg_sq), state):
            t.copy_(sv)                                     # warm-up steps must not count as training
        self._graph, self._graph_key = graph, tuple(shapes)

    def step_graph(self, im_noisy, im_gt, sigma_gt, lr: Optional[float] = None):
        """step() with the device work replayed from a CUDA graph: inputs (host or device) are copied into static
        buffers, then ONE graph launch replays forward + ELBO + backward (+ clip + Adam when world == 1, with the
        per-step Adam scalars read from device memory).  With world > 1 the gradient all-reduce and the clip + Adam
        kernel are enqueued after the graph (see _capture)."""
        shapes = (tuple(im_noisy.shape), tuple(im_gt.shape), tuple(sigma_gt.shape))
        if self._graph is None or self._graph_key != shapes:
            self._capture(shapes)
        for dst, src in zip(self._g_in, (im_noisy, im_gt, sigma_gt)):
            dst.copy_(src, non_blocking=True)
        self.step_count += 1
        if self.world > 1:
            self._graph.replay()
            self._device_update(self.lr if lr is None else lr, None, reduced=False)
            # the captured forward re-packs the weights unconditionally: nothing else to mark
            return self.losses
        k = self.step_count % len(self._hyper_ring)
        host, ev = self._hyper_ring[k]
        if self._hyper_used[k]:
            ev.synchronize()                                # its previous copy has been consumed (8 steps ago)
        host[0] = self.lr if lr is None else lr
        host[1], host[2] = _bias_corrections(self.betas, self.step_count)
        self._hyper.copy_(host, non_blocking=True)
        ev.record()
        self._hyper_used[k] = True
        self._graph.replay()
        return self.losses


class SISRTrainer:
    """The hot loop of the reference's train_SISR.py:197-229 as one kernel-only CUDA program per step:

        im_hr, im_lr, kinfo_gt, sigma_prior
          -> forward (SNet, KNet, SFT-modulated RNet)      DenoiseEngine.forward_sr
          -> SISR negative ELBO + gradients                 vk_elbo_sisr
          -> backward                                       DenoiseEngine.backward_sr
          -> [world_size > 1] NCCL all-reduce of the flat gradient buffer
          -> clip_grad_norm_ per sub-network (R / S / K) -> Adam   vk_adam_clip_step

    The loss's random draws come from torch's CUDA generator in the reference's order (loss/ELBO_simple.py)."""

    def __init__(self, net, sf, lr=1e-4, clip_grad_R=5e2, clip_grad_S=1e2, clip_grad_K=5e2, var_window=9, kappa0=50.0,
                 r2=1e-4, eps2=1e-5, k_size=21, penalty_K=(0.02, 2), kernel_shift=False, downsampler="Bicubic",
                 betas=(0.9, 0.999), adam_eps=1e-8, process_group=None, deterministic=None):
        """deterministic=True: bit-reproducible training for per-sample-constant conditioning (the shipped configurations) —
        split-K weight gradients in ordered slabs (engine.deterministic) and the fixed-order forms of the small per-sample
        kernels (vk_sft_bwd_det, vk_sft_mlp_bwd_batched_det, vk_ca_layer_bwd_det, vk_knet_head_wgrad_det); the loss
        (vk_elbo_sisr) never uses atomics.  Per-pixel conditioning maps (noise_avg=False) still scatter with atomics."""
        self.net, self.sf = net, int(sf)
        self.engine = eng = net.engine()
        if deterministic is not None:
            eng.deterministic = bool(deterministic)
        eng._ensure_flat()
        dev = eng.flat_params.device
        self.lr, self.betas, self.adam_eps = lr, betas, adam_eps
        self.alpha0 = 0.5 * float(var_window) ** 2                       # train_SISR.py:180
        self.kappa0, self.r2, self.eps2, self.k_size = float(kappa0), float(r2), float(eps2), int(k_size)
        self.penalty_K, self.shift, self.downsampler = tuple(penalty_K), bool(kernel_shift), downsampler
        self.pg = process_group
        self.world = dp.world_size(process_group)
        dp.broadcast_flat_params(eng.flat_params, 0, process_group)
        self.step_count = 0
        self.exp_avg = torch.zeros_like(eng.flat_params)
        self.exp_avg_sq = torch.zeros_like(eng.flat_params)
        self.group_names = ["RNet", "SNet", "KNet"]
        self._loss_ws = {}                 # scratch of vk_elbo_sisr, owned by this trainer (never freed: graphs capture it)
        self._groups_dev, self._ngroups, self._max_group = _clip_groups(
            net, eng, (("rnet", clip_grad_R), ("snet", clip_grad_S), ("knet", clip_grad_K)))
        self._sq_ws = ops.adam_ws(self._ngroups, dev)
        self.grad_norms = torch.zeros(self._ngroups, device=dev, dtype=torch.float32)
        self.losses = None

    def step(self, im_hr, im_lr, kinfo_gt, sigma_prior, lr: Optional[float] = None, draws=None):
        """One optimisation step; returns the device tensor [loss, lh, kl_rnet, kl_snet, kl_knet, kl_k0, kl_k1, kl_k2]."""
        dev = self.engine.flat_params.device
        im_hr, im_lr, kinfo_gt, sigma_prior = [t if t.is_cuda else t.to(dev, non_blocking=True)
                                               for t in (im_hr, im_lr, kinfo_gt, sigma_prior)]
        if draws is None:
            draws = self._draw(im_hr)
        self.step_count += 1
        return self._device_step(im_hr, im_lr, kinfo_gt, sigma_prior, draws, self.lr if lr is None else lr, None)

    def _draw(self, im_hr):
        """The loss's random draws (shapes are known before the forward pass: mu has the shape of im_hr)."""
        return sisr_draws(im_hr[:, 0, 0, :1], im_hr, self.kappa0)

    def _device_step(self, im_hr, im_lr, kinfo_gt, sigma_prior, draws, lr, hyper_dev):
        eng = self.engine
        dev = eng.flat_params.device
        n = im_lr.shape[0]
        mu, kinfo, sigma = eng.forward_sr(im_lr, self.sf, save=True)
        H, W = mu.shape[2], mu.shape[3]
        rh = downsample_matrix(H, self.sf, self.downsampler, dev)
        rw = downsample_matrix(W, self.sf, self.downsampler, dev)
        sp = sigma_prior.float().expand(n, *sigma_prior.shape[1:]).reshape(n, -1)
        center = self.k_size // 2 + 0.5 * (self.sf - self.k_size % 2) if self.shift else float(self.k_size // 2)
        terms, kernel, d_mu, d_sigma, d_kinfo = ops.elbo_sisr(
            mu, im_hr.contiguous(), im_lr.contiguous(), sigma.reshape(n), kinfo, kinfo_gt.contiguous().float(),
            sp.mean(1).contiguous(), sp.log().mean(1).contiguous(), draws[0].contiguous(),
            draws[1].reshape(n).contiguous(), draws[2].contiguous(), rh, rw, k_size=self.k_size, center=float(center),
            alpha0=self.alpha0, digamma_am1=_digamma(self.alpha0 - 1.0), kappa0=self.kappa0, r2=self.r2, eps2=self.eps2,
            pk0=float(self.penalty_K[0]), pk1=float(self.penalty_K[1]), ws_cache=self._loss_ws)
        eng.backward_sr(d_mu, d_kinfo, d_sigma)
        grad_scale = dp.all_reduce_flat_grads(eng.flat_grads, self.pg)
        if hyper_dev is None:
            ops.adam_clip_step(eng.flat_params, eng.flat_grads, self.exp_avg, self.exp_avg_sq, self._groups_dev,
                               self._ngroups, self._max_group, self._sq_ws, grad_scale=grad_scale, lr=lr,
                               beta1=self.betas[0], beta2=self.betas[1], eps=self.adam_eps, step=self.step_count,
                               norms_out=self.grad_norms)
        else:
            ops.adam_clip_step_dev(eng.flat_params, eng.flat_grads, self.exp_avg, self.exp_avg_sq, self._groups_dev,
                                   self._ngroups, self._max_group, self._sq_ws, hyper_dev, grad_scale=grad_scale,
                                   beta1=self.betas[0], beta2=self.betas[1], eps=self.adam_eps, norms_out=self.grad_norms)
        eng.mark_params_dirty()
        self.losses, self.last_kernel = terms, kernel
        self.last_mu, self.last_kinfo, self.last_sigma = mu, kinfo, sigma
        return terms

    # -- CUDA-graph replay (see DenoiseTrainer.step_graph); the loss's random draws are made OUTSIDE the graph ------
    def step_graph(self, im_hr, im_lr, kinfo_gt, sigma_prior, lr: Optional[float] = None, draws=None):
        if self.world > 1:
            raise NotImplementedError("step_graph with world_size > 1 (NCCL all-reduce inside the capture) is not validated: "
                                      "use step() for data-parallel runs")
        eng = self.engine
        dev = eng.flat_params.device
        ins = (im_hr, im_lr, kinfo_gt, sigma_prior)
        shapes = tuple(tuple(t.shape) for t in ins)
        if getattr(self, "_graph", None) is None or self._graph_key != shapes:
            self._g_in = [torch.zeros(sh, device=dev, dtype=torch.float32) for sh in shapes]
            self._g_in[2][:, :2] = 1.0                       # kernel variances must be positive during the dry runs
            self._g_in[3].fill_(1e-3)
            self._g_draws = [torch.ones(im_hr.shape[0], 2, device=dev), torch.zeros(im_hr.shape[0], 1, device=dev),
                             torch.zeros(tuple(im_hr.shape), device=dev)]
            self._hyper = torch.tensor([0.0, 1.0, 1.0], device=dev)
            self._hyper_ring = [(torch.zeros(3).pin_memory(), torch.cuda.Event()) for _ in range(8)]
            self._hyper_used = [False] * 8
            state = [t.clone() for t in (eng.flat_params, self.exp_avg, self.exp_avg_sq)]
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    eng.mark_params_dirty()
                    self._device_step(*self._g_in, self._g_draws, 0.0, self._hyper)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(dev)
            eng.mark_params_dirty()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._g_out = self._device_step(*self._g_in, self._g_draws, 0.0, self._hyper)
            for t, sv in zip((eng.flat_params, self.exp_avg, self.exp_avg_sq), state):
                t.copy_(sv)
            self._graph, self._graph_key = graph, shapes
        for dst, src in zip(self._g_in, ins):
            dst.copy_(src, non_blocking=True)
        if draws is None:
            draws = self._draw(self._g_in[0])
        for dst, src in zip(self._g_draws, draws):
            dst.copy_(src.reshape(dst.shape), non_blocking=True)
        self.step_count += 1
        k = self.step_count % len(self._hyper_ring)
        host, ev = self._hyper_ring[k]
        if self._hyper_used[k]:
            ev.synchronize()
        host[0] = self.lr if lr is None else lr
        host[1], host[2] = _bias_corrections(self.betas, self.step_count)
        self._hyper.copy_(host, non_blocking=True)
        ev.record()
        self._hyper_used[k] = True
        self._graph.replay()
        return self._g_out
