"""Training side of the ViT-CNN module: what ``net(data, data2)`` in train mode, ``loss.backward()``
and ``optimizer.step()`` do in the reference's loop (model_utils.py:908-945), computed by
libvitcnn.so.

* ``train_forward``  - autograd bridge used by ``ViTCNN.forward`` when ``model.training``: the
  forward (BatchNorm with batch statistics) and the whole backward run in the library; torch
  only routes the returned gradients to the ``nn.Parameter``s, so the reference's own loop
  (criterion = nn.CrossEntropyLoss, optimizer = optim.Adam) works unchanged.
* ``Trainer``        - the fast path: patches gathered on the device from the rasters, loss,
  backward, (NCCL all-reduce of one flat fp32 gradient bucket), Adam - no torch operator on
  the step except the collective.
* ``train``          - the reference's ``train()``: same signature, same results (model_utils.py:854-1045).

Parameters are kept in ONE flat fp32 buffer (the nn.Parameters are views into it, in the
canonical order of include/vitcnn.h) so that packing, all-reduce and Adam are single launches.
"""
from __future__ import annotations

import ctypes
import weakref

import numpy as np
import torch
import torch.nn as nn

from . import _lib

_BLOCK_KEYS = ("norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias",
               "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias")


def param_names() -> list:
    """Canonical parameter order (include/vitcnn.h, struct vc_train)."""
    names = []
    for stem in ("hsi_stem.0", "hsi_stem.1", "hsi_stem.2", "lidar_stem.0", "lidar_stem.1", "lidar_stem.2", "fusion"):
        names += [f"{stem}.conv.weight", f"{stem}.conv.bias", f"{stem}.bn.weight", f"{stem}.bn.bias"]
    names += ["cls_token", "pos_embed"]
    for l in (0, 1):
        names += [f"blocks.{l}.{k}" for k in _BLOCK_KEYS]
    names += ["norm.weight", "norm.bias", "head.weight", "head.bias"]
    return names


def flat_offsets(numels) -> tuple:
    """Element offset of every tensor in the flat parameter / gradient bucket (each tensor
    starts on a 16-byte boundary) and the bucket length."""
    offs, o = [], 0
    for n in numels:
        offs.append(o)
        o += (int(n) + 3) // 4 * 4
    return offs, o


def flatten_grads(model, device=None) -> torch.Tensor:
    """Gradients of ``model`` (any module with this package's state_dict keys) as one flat fp32
    bucket in the canonical order -- the layout the all-reduce and the Adam kernel work on."""
    named = dict(model.named_parameters())
    ps = [named[k] for k in param_names()]
    offs, total = flat_offsets([p.numel() for p in ps])
    flat = torch.zeros(total, dtype=torch.float32, device=device or ps[0].device)
    for p, o in zip(ps, offs):
        if p.grad is not None:
            flat[o:o + p.numel()] = p.grad.reshape(-1)
    return flat


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the flat bucket over the ranks (NCCL on GPUs, gloo in the CPU tests) and divide by the
    world size: the data-parallel gradient of the mean of the per-rank losses."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            flat.div_(world)
    return flat


def shard_batch(n_global: int, rank: int, world: int) -> slice:
    """Contiguous shard of a global batch for one rank (sizes differ by at most one)."""
    base, rem = divmod(n_global, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def _bn_layers(model):
    return [model.hsi_stem[0].bn, model.hsi_stem[1].bn, model.hsi_stem[2].bn, model.lidar_stem[0].bn,
            model.lidar_stem[1].bn, model.lidar_stem[2].bn, model.fusion.bn]


class TrainState:
    """Flat parameter / gradient buffers, the ``vc_train`` descriptor and per-batch-size
    workspaces of one model on one device."""

    def __init__(self, model):
        self.model = model
        self.names = param_names()
        named = dict(model.named_parameters())
        self.params = [named[k] for k in self.names]
        assert len(self.params) == 58 and len(named) == 58
        self.device = self.params[0].device
        if self.device.type != "cuda":
            raise RuntimeError("ViTCNN trains on a CUDA device only (no CPU path)")
        self.offsets, o = flat_offsets([p.numel() for p in self.params])
        self.numel = o
        self.flat = torch.zeros(o, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros(o, dtype=torch.float32, device=self.device)
        with torch.no_grad():
            for p, off in zip(self.params, self.offsets):
                view = self.flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
        P, K = model.patch_size, model.num_classes
        self.segments = self._blob_segments(P, K)
        st = _lib.VcTrain()
        st.C1, st.C2, st.P, st.K = model.n_bands, model.n_bands2, P, K
        st.params, st.grads = self.flat.data_ptr(), self.grads.data_ptr()
        for i, off in enumerate(self.offsets):
            st.off[i] = off
        for i, bn in enumerate(_bn_layers(model)):
            st.bn_running_mean[i] = bn.running_mean.data_ptr()
            st.bn_running_var[i] = bn.running_var.data_ptr()
            st.bn_num_batches[i] = bn.num_batches_tracked.data_ptr()
        bn0 = model.hsi_stem[0].bn
        st.bn_eps, st.bn_momentum = bn0.eps, bn0.momentum
        st.blob_segments, st.n_blob_segments = self.segments.data_ptr(), self.segments.shape[0]
        # dropout masks are a hash of (seed word, sample, site, element); the word advances on the device
        self.drop_seed = torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int64).to(torch.int32).to(self.device)
        st.dropout, st.drop_seed = float(model.dropout), self.drop_seed.data_ptr()
        self.struct = st
        self._bn_ptrs = [b.running_mean.data_ptr() for b in _bn_layers(model)]
        self.ws = {}
        self.on_evict = []      # callbacks(batch size) run when a per-batch-size workspace is dropped
        self.pending = None     # batch size of a forward whose backward has not run yet
        self.generation = 0     # stamp of the last training forward: its activations are the ones in the workspace

    def __deepcopy__(self, memo):   # copies / pickles of the module rebuild their own state lazily
        return None

    def __reduce__(self):
        return (type(None), ())

    def _blob_segments(self, P, K):
        lay = _lib.tparams_layout(P, K)
        off = dict(zip(self.names, self.offsets))
        T = P * P + 1
        seg = [(off["cls_token"], lay["cls"], 1, 32, 32, 0), (off["pos_embed"], lay["pos"], 1, T * 32, T * 32, 0)]
        for l, lo in enumerate(lay["layers"]):
            b = f"blocks.{l}."
            seg += [(off[b + "attn.qkv.weight"], lo["wqkv"], 96, 32, 40, 1),
                    (off[b + "attn.proj.weight"], lo["wproj"], 32, 32, 40, 1),
                    (off[b + "mlp.fc1.weight"], lo["wfc1"], 128, 32, 40, 1),
                    (off[b + "mlp.fc2.weight"], lo["wfc2"], 32, 128, 136, 1)]
            for key, name, n in (("ln1_g", "norm1.weight", 32), ("ln1_b", "norm1.bias", 32), ("bqkv", "attn.qkv.bias", 96),
                                 ("bproj", "attn.proj.bias", 32), ("ln2_g", "norm2.weight", 32), ("ln2_b", "norm2.bias", 32),
                                 ("bfc1", "mlp.fc1.bias", 128), ("bfc2", "mlp.fc2.bias", 32)):
                seg.append((off[b + name], lo[key], 1, n, n, 0))
        seg += [(off["norm.weight"], lay["lnf_g"], 1, 32, 32, 0), (off["norm.bias"], lay["lnf_b"], 1, 32, 32, 0),
                (off["head.weight"], lay["whead"], 1, K * 32, K * 32, 0), (off["head.bias"], lay["bhead"], 1, K, K, 0)]
        return torch.tensor(seg, dtype=torch.int64, device=self.device)

    def valid(self) -> bool:
        """Parameters still views into the flat buffer and BN buffers where we recorded them
        (``.to()`` / ``load_state_dict(assign=True)`` replace storages)."""
        base = self.flat.data_ptr()
        if any(p.data_ptr() != base + 4 * off for p, off in zip(self.params, self.offsets)):
            return False
        return self._bn_ptrs == [b.running_mean.data_ptr() for b in _bn_layers(self.model)]

    def workspace(self, n: int) -> torch.Tensor:
        ws = self.ws.get(n)
        if ws is None:
            L = _lib.lib()
            need = L.vc_train_workspace_bytes(ctypes.byref(self.struct), n)
            if need <= 0:
                raise RuntimeError("this configuration is not supported by the training kernels "
                                   "(patch_size <= 11 or 13 / 15, n_classes <= 64)")
            if len(self.ws) >= 4:
                evicted = next(iter(self.ws))
                self.ws.pop(evicted)
                for cb in self.on_evict:          # CUDA graphs hold raw pointers into the workspace they captured
                    cb(evicted)
            ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            _lib.check(L.vc_train_workspace_init(ctypes.byref(self.struct), n, ws.data_ptr(), ws.numel(),
                                                 torch.cuda.current_stream().cuda_stream), "vc_train_workspace_init")
            self.ws[n] = ws
        return ws

    # ---- one batch --------------------------------------------------------------------------------
    def forward_patches(self, hsi, lidar):
        n = hsi.shape[0]
        logits = torch.empty(n, self.model.num_classes, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            ws = self.workspace(n)
            hs = (ctypes.c_int64 * 4)(*hsi.stride())
            ls = (ctypes.c_int64 * 4)(*lidar.stride())
            _lib.check(_lib.lib().vc_train_forward(ctypes.byref(self.struct), hsi.data_ptr(), hs, lidar.data_ptr(), ls, n,
                                                   ws.data_ptr(), ws.numel(), logits.data_ptr(),
                                                   torch.cuda.current_stream().cuda_stream), "vc_train_forward")
        self.pending = n
        self.generation += 1
        self.model._pack = None      # running statistics changed: eval-mode packing is stale
        return logits

    def forward_gather(self, img1, img2, gt, xy, ops=None):
        n = xy.shape[0]
        H, W, _ = img1.shape
        logits = torch.empty(n, self.model.num_classes, dtype=torch.float32, device=self.device)
        labels = torch.empty(n, dtype=torch.int64, device=self.device) if gt is not None else None
        with torch.cuda.device(self.device):
            ws = self.workspace(n)
            _lib.check(_lib.lib().vc_train_forward_gather(
                ctypes.byref(self.struct), img1.data_ptr(), img2.data_ptr(), 0 if gt is None else gt.data_ptr(),
                0 if gt is None else gt.element_size(), H, W, xy.data_ptr(), 0 if ops is None else ops.data_ptr(), n,
                ws.data_ptr(), ws.numel(),
                logits.data_ptr(), 0 if labels is None else labels.data_ptr(),
                torch.cuda.current_stream().cuda_stream), "vc_train_forward_gather")
        self.pending = n
        self.generation += 1
        self.model._pack = None
        return logits, labels

    def backward(self, dlogits, generation=None):
        n = dlogits.shape[0]
        if self.pending != n:
            raise RuntimeError("backward() without a matching training forward on this model")
        if generation is not None and generation != self.generation:
            raise RuntimeError("backward() of a forward whose activations were overwritten by a later training-mode "
                               "forward of this model (one forward per backward: the workspace holds one batch)")
        with torch.cuda.device(self.device):
            ws = self.workspace(n)
            _lib.check(_lib.lib().vc_train_backward(ctypes.byref(self.struct), dlogits.data_ptr(), n, ws.data_ptr(),
                                                    ws.numel(), torch.cuda.current_stream().cuda_stream),
                       "vc_train_backward")
        self.pending = None

    def grad_views(self, flat):
        return tuple(flat[off:off + p.numel()].view(p.shape) for p, off in zip(self.params, self.offsets))


def train_state(model) -> TrainState:
    st = getattr(model, "_train_state", None)
    if st is None or not st.valid():
        st = TrainState(model)
        model._train_state = st
    if st.struct.dropout != float(model.dropout):       # model.dropout may be changed between steps
        st.struct.dropout = float(model.dropout)
    return st


class _TrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, state, hsi, lidar, *params):
        ctx.state = state
        out = state.forward_patches(hsi, lidar)
        ctx.generation = state.generation
        return out

    @staticmethod
    def backward(ctx, dlogits):
        st = ctx.state
        st.backward(dlogits.contiguous().float(), ctx.generation)
        return (None, None, None) + st.grad_views(st.grads.clone())


def train_forward(model, hsi, lidar):
    """Training-mode forward of ``ViTCNN`` (called from ``ViTCNN.forward``)."""
    st = train_state(model)
    if hsi.shape[0] == 0:
        return torch.empty(0, model.num_classes, dtype=torch.float32, device=hsi.device)
    if not torch.is_grad_enabled():
        return st.forward_patches(hsi, lidar)
    return _TrainFn.apply(st, hsi, lidar, *st.params)


# ----------------------------------------------------------------------------------------------------
_CE_SCRATCH = {}


def ce_loss(logits, labels, weight=None, grad_scale=1.0, want_grad=True):
    """nn.CrossEntropyLoss(weight) forward (+ gradient): returns (loss[2] = mean loss, weight sum;
    dlogits or None)."""
    n, K = logits.shape
    out = torch.empty(2, dtype=torch.float32, device=logits.device)
    d = torch.empty_like(logits) if want_grad else None
    scratch = _CE_SCRATCH.get(logits.device)
    if scratch is None:
        scratch = _CE_SCRATCH[logits.device] = torch.zeros(2, dtype=torch.float64, device=logits.device)
    with torch.cuda.device(logits.device):
        _lib.check(_lib.lib().vc_ce_loss(logits.data_ptr(), labels.data_ptr(), 0 if weight is None else weight.data_ptr(),
                                         n, K, float(grad_scale), out.data_ptr(), 0 if d is None else d.data_ptr(),
                                         scratch.data_ptr(), torch.cuda.current_stream().cuda_stream), "vc_ce_loss")
    return out, d


def adam_step(p, g, m, v, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, grad_scale=1.0):
    with torch.cuda.device(p.device):
        _lib.check(_lib.lib().vc_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, betas[0],
                                           betas[1], eps, weight_decay, step, grad_scale,
                                           torch.cuda.current_stream().cuda_stream), "vc_adam_step")


class Trainer:
    """Data-parallel training step on device-resident rasters: gather -> forward -> weighted CE ->
    backward -> all-reduce(sum) of the flat gradient bucket over NCCL -> Adam(grad / world).
    One process per GPU; replicas start identical (same seed / broadcast state_dict).

    ``use_graph=True`` captures the whole step (every kernel of this library, the collective and
    the optimiser) in one CUDA graph after ``graph_warmup`` eager steps: the step is ~65 short
    launches, so at small per-GPU batches the CPU launch rate, not the GPU, sets the pace."""

    def __init__(self, model, lr=1e-3, weights=None, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, process_group=None,
                 use_graph=False, graph_warmup=3):
        import torch.distributed as dist
        self.model = model.train()
        self.state = train_state(model)
        dev = self.state.device
        self.weights = None if weights is None else weights.to(dev, torch.float32).contiguous()
        self.m = torch.zeros_like(self.state.flat)
        self.v = torch.zeros_like(self.state.flat)
        self.hyper = torch.tensor([lr, betas[0], betas[1], eps, weight_decay, 0.0, 0.0, 0.0], dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.lr = lr
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.group = process_group
        self.world = self.dist.get_world_size(process_group) if self.dist else 1
        self.use_graph, self.graph_warmup = use_graph, graph_warmup
        self._graphs = {}       # batch size -> (graph, static xy, static loss)
        # (weak: the state must not keep the trainer -- and with it CUDA graphs that captured NCCL kernels -- alive in a
        # reference cycle; a communicator cannot be destroyed while such a graph exists)
        me = weakref.ref(self)
        self.state.on_evict.append(lambda n, me=me: me() is not None and me()._graphs.pop(n, None))
        self.launches_per_step = 0
        self._eager_steps = 0

    @property
    def t(self) -> int:
        return int(self.step_count.item())

    def set_lr(self, lr: float) -> None:
        self.lr = float(lr)
        self.hyper[0:1].fill_(self.lr)

    def step_lr(self, epoch: int, step_size: int = 30, gamma: float = 0.9, base_lr: float = None) -> float:
        """StepLR(step_size, gamma) as the reference schedules Adam (model_utils.py:498): call once
        per epoch with the number of finished epochs."""
        if base_lr is None:
            base_lr = getattr(self, "_base_lr", self.lr)
        self._base_lr = base_lr
        self.set_lr(base_lr * gamma ** (epoch // step_size))
        return self.lr

    def close(self) -> None:
        """Drop the captured CUDA graphs (they hold NCCL kernels: ``destroy_process_group`` waits for them to be gone)."""
        self._graphs.clear()

    def allreduce_ms(self, iters: int = 20) -> float:
        """Device time of the gradient all-reduce alone (one flat fp32 bucket, NCCL), CUDA events, mean of `iters`
        back-to-back calls after two warm-ups; 0 for a single rank.  The bucket is scratch here (call between steps)."""
        if self.world <= 1:
            return 0.0
        g = self.state.grads
        for _ in range(2):
            self.dist.all_reduce(g, op=self.dist.ReduceOp.SUM, group=self.group)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            self.dist.all_reduce(g, op=self.dist.ReduceOp.SUM, group=self.group)
        e1.record()
        torch.cuda.synchronize()
        g.zero_()
        return e0.elapsed_time(e1) / iters

    def _step_impl(self, img1, img2, gt, xy, ops=None):
        st = self.state
        logits, labels = st.forward_gather(img1, img2, gt, xy, ops)
        loss, dlogits = ce_loss(logits, labels, self.weights)
        st.backward(dlogits)
        if self.world > 1:
            self.dist.all_reduce(st.grads, op=self.dist.ReduceOp.SUM, group=self.group)
        with torch.cuda.device(st.device):
            _lib.check(_lib.lib().vc_adam_step_dev(st.flat.data_ptr(), st.grads.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                                   st.flat.numel(), self.hyper.data_ptr(), self.step_count.data_ptr(),
                                                   1.0 / self.world, torch.cuda.current_stream().cuda_stream),
                       "vc_adam_step_dev")
        return loss

    def step(self, img1, img2, gt, xy, ops=None, validate=False):
        """One optimisation step on patches centred at xy (int32 [n,2], device); ``ops`` (uint8 [n],
        device, optional) = flip / rot90 augmentation code per sample, drawn by the host.  Returns the
        loss tensor [2] (mean loss of this rank, weight sum) without synchronising.  The centres must keep
        their P x P windows inside the raster (MultiModalX.indices do); ``validate=True`` checks that with one
        device->host read and raises ``ValueError`` (the kernels themselves only guarantee memory safety)."""
        if validate:
            from .ops import validate_xy
            validate_xy(xy, img1.shape[0], img1.shape[1], self.model.patch_size, True)
        st = self.state
        if not st.valid():
            raise RuntimeError("model parameters were moved after the Trainer was built")
        self.model._pack = None
        n = xy.shape[0]
        if not self.use_graph or ops is not None:
            return self._step_impl(img1, img2, gt, xy, ops)
        st = train_state(self.model)                      # picks up a changed model.dropout
        entry = self._graphs.get(n)
        if entry is not None and entry[3] == (img1.data_ptr(), img2.data_ptr(), gt.data_ptr(), st.struct.dropout):
            graph, xy_s, loss_s, _ = entry
            xy_s.copy_(xy, non_blocking=True)
            graph.replay()
            return loss_s
        if self._eager_steps < self.graph_warmup:
            self._eager_steps += 1
            return self._step_impl(img1, img2, gt, xy)
        xy_s = xy.clone()
        st.workspace(n)                                   # allocated and initialised outside the capture
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        l0 = _lib.lib().vc_launch_count()
        with torch.cuda.graph(graph):
            loss_s = self._step_impl(img1, img2, gt, xy_s)
        self.launches_per_step = int(_lib.lib().vc_launch_count() - l0)   # kernels of this library in one replay
        self._graphs[n] = (graph, xy_s, loss_s, (img1.data_ptr(), img2.data_ptr(), gt.data_ptr(), st.struct.dropout))
        graph.replay()      # capturing records the step without running it
        return loss_s


def train(savename, run, bands, net, optimizer, criterion, data_loader, epoch, scheduler=None, display_iter=100,
          device=torch.device("cpu"), display=None, val_loader=None, supervision="full"):
    """The reference's training loop (model_utils.py:854-1045) with its signature (main.py:478 calls it
    positionally) and its results: iteration order, loss, optimiser and scheduler stepping, validation after
    every epoch **in training mode** (the reference never switches the mode before ``val()``, :908 / :991),
    the running-mean loss plotted to ``display`` every ``display_iter`` iterations (:940-974), the best
    weights kept with the reference's rule ``abs(metric) >= best`` where ``metric = -val_acc``, or the
    epoch's mean loss without a validation loader (:1015-1017; ties go to the LATER epoch, and without
    validation the rule keeps the HIGHEST mean loss, as upstream), checkpoints through ``save_model``
    every ``save_epoch`` epochs among the best ones and at the last epoch (:1018-1043).  ``bands`` is
    unused upstream too.  Returns the best ``state_dict`` (a deep copy)."""
    import copy

    from .model_utils import save_model, val
    from .utils import camel_to_snake
    best_val_acc = 0.0
    if criterion is None:
        raise Exception("Missing criterion. You must specify a loss function.")
    net.to(device)
    save_epoch = 16 if epoch == 128 else (epoch // 20 if epoch > 20 else 1)
    recent = []                       # losses[max(0, iter_-100) : iter_+1] of the reference: slot 0 is never written
    mean_losses = {}
    iter_ = 1
    loss_win, val_win = None, None
    val_accuracies = []
    model_name = camel_to_snake(str(net.__class__.__name__))
    for e in range(1, epoch + 1):
        net.train()
        avg_loss = 0.0
        for batch_idx, (data, data2, target) in enumerate(data_loader):
            data, data2, target = data.to(device), data2.to(device), target.to(device)
            optimizer.zero_grad()
            if supervision == "full":
                output = net(data, data2)
                loss = criterion(output, target)
            else:
                raise ValueError('supervision mode "{}" is unknown.'.format(supervision))
            loss.backward()
            optimizer.step()
            value = loss.item()
            avg_loss += value
            recent.append(value)
            if len(recent) > 101:
                recent.pop(0)
            window = recent if iter_ > 100 else [0.0] + recent      # the zero-initialised slot 0 (:900, :938)
            mean_losses[iter_] = float(np.mean(np.asarray(window, dtype=np.float64)))
            if display_iter and iter_ % display_iter == 0:
                update = None if loss_win is None else "append"
                loss_win = display.line(
                    X=np.arange(iter_ - display_iter, iter_),
                    Y=np.array([mean_losses.get(k, 0.0) for k in range(iter_ - display_iter, iter_)]),
                    win=loss_win, update=update,
                    opts={"title": "Training loss run:{}".format(run), "xlabel": "Iterations", "ylabel": "Loss"})
                if len(val_accuracies) > 0:
                    val_win = display.line(
                        Y=np.array(val_accuracies), X=np.arange(len(val_accuracies)), win=val_win,
                        opts={"title": "Validation accuracy run:{}".format(run), "xlabel": "Epochs", "ylabel": "Accuracy"})
            iter_ += 1
        avg_loss /= len(data_loader)
        if val_loader is not None:
            val_acc = val(net, val_loader, device=device, supervision=supervision)
            val_accuracies.append(val_acc)
            metric = -val_acc
        else:
            metric = avg_loss
        if isinstance(scheduler, torch.optim.lr_scheduler.ReduceLROnPlateau):
            scheduler.step(metric)
        elif scheduler is not None:
            scheduler.step()
        if abs(metric) >= best_val_acc:
            best_val_acc = abs(metric)
            best_model_wts = copy.deepcopy(net.state_dict())
            if e % save_epoch == 0:
                save_model(savename, net, model_name, data_loader.dataset.name, train_state="train", type="best_epoch",
                           run=run, epoch=e, metric=abs(metric))
        if e == epoch:
            save_model(savename, net, model_name, data_loader.dataset.name, train_state="train", type="final_epoch",
                       run=run, epoch=e, metric=abs(metric))
    return best_model_wts
